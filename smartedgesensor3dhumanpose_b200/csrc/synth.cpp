// synth.cpp — host driver of the synthetic frame generator (synth.h), multi-threaded over frames.
#include "synth.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

extern "C" int ses3d_synth_frames(int32_t n_cams, const ses3d_camera* cams, const ses3d_synth_config* cfg,
                                  int64_t first_frame, int32_t n_frames, ses3d_person2d* persons, int32_t* n_persons,
                                  int32_t* gt_id, float* gt_joints) {
  using namespace ses3d_synth;
  if (!cams || !cfg || !persons || !n_persons || n_cams < 1 || n_frames < 0) return SES3D_E_INVALID;
  if (cfg->n_people < 0 || cfg->n_people > SES3D_SYNTH_MAX_PEOPLE || cfg->p_max < 1) return SES3D_E_INVALID;
  const int C = n_cams, PM = cfg->p_max;
  std::memset(persons, 0, sizeof(ses3d_person2d) * (size_t)n_frames * C * PM);
  if (gt_id) std::fill(gt_id, gt_id + (size_t)n_frames * C * PM, -1);
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int n_threads = (int)std::min<unsigned>(hw, (unsigned)std::max(1, n_frames / 64));
  auto work = [&](int a, int b) {
    for (int f = a; f < b; ++f) {
      Scene sc;
      const uint64_t frame = (uint64_t)(first_frame + f);
      make_scene(*cfg, frame, sc);
      if (gt_joints)
        for (int p = 0; p < cfg->n_people; ++p)
          for (int k = 0; k < 17; ++k) {
            double X[3];
            world_joint(sc, p, k, X);
            float* o = gt_joints + (((size_t)f * cfg->n_people + p) * 17 + k) * 3;
            o[0] = (float)X[0]; o[1] = (float)X[1]; o[2] = (float)X[2];
          }
      for (int c = 0; c < C; ++c) {
        const size_t base = ((size_t)f * C + c) * PM;
        n_persons[(size_t)f * C + c] = make_camera_view(*cfg, cams[c], c, frame, sc, persons + base,
                                                        gt_id ? gt_id + base : nullptr);
      }
    }
  };
  if (n_threads <= 1) work(0, n_frames);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
      pool.emplace_back(work, (int)((int64_t)n_frames * t / n_threads), (int)((int64_t)n_frames * (t + 1) / n_threads));
    for (auto& th : pool) th.join();
  }
  return SES3D_OK;
}
