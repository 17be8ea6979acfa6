// synth_device.cu — device variant of the synthetic frame generator (synth.h): one thread per
// (frame, camera). Compiled with -fmad=false so that it is bit-identical to the host generator.
// Used to feed batches that are too large to ship over PCIe (config 5: 1e7 frames).
#include <cuda_runtime.h>

#include "synth.h"

namespace {

__global__ void k_synth(int n_cams, const ses3d_camera* __restrict__ cams, ses3d_synth_config cfg, long long first_frame,
                        int n_frames, ses3d_person2d* persons, int32_t* n_persons, int32_t* gt_id) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_frames * n_cams) return;
  const int f = (int)(t / n_cams), c = (int)(t % n_cams);
  ses3d_synth::Scene sc;
  const uint64_t frame = (uint64_t)(first_frame + f);
  ses3d_synth::make_scene(cfg, frame, sc);
  const size_t base = ((size_t)f * n_cams + c) * cfg.p_max;
  n_persons[(size_t)f * n_cams + c] =
      ses3d_synth::make_camera_view(cfg, cams[c], c, frame, sc, persons + base, gt_id ? gt_id + base : nullptr);
}

}  // namespace

extern "C" int ses3d_synth_frames_device(int32_t n_cams, const ses3d_camera* cams, const ses3d_synth_config* cfg,
                                         int64_t first_frame, int32_t n_frames, ses3d_person2d* persons,
                                         int32_t* n_persons, int32_t* gt_id, void* stream) {
  if (!cams || !cfg || !persons || !n_persons || n_cams < 1 || n_frames < 0) return SES3D_E_INVALID;
  if (cfg->n_people < 0 || cfg->n_people > SES3D_SYNTH_MAX_PEOPLE || cfg->p_max < 1) return SES3D_E_INVALID;
  if (n_frames == 0) return SES3D_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ses3d_camera* d_cams = nullptr;
  if (cudaMalloc(&d_cams, sizeof(ses3d_camera) * n_cams) != cudaSuccess) return SES3D_E_CUDA;
  cudaError_t e = cudaMemcpyAsync(d_cams, cams, sizeof(ses3d_camera) * n_cams, cudaMemcpyHostToDevice, st);
  const size_t slots = (size_t)n_frames * n_cams * cfg->p_max;
  if (e == cudaSuccess) e = cudaMemsetAsync(persons, 0, slots * sizeof(ses3d_person2d), st);
  if (e == cudaSuccess && gt_id) e = cudaMemsetAsync(gt_id, 0xFF, slots * sizeof(int32_t), st);
  if (e == cudaSuccess) {
    const long long total = (long long)n_frames * n_cams;
    k_synth<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(n_cams, d_cams, *cfg, (long long)first_frame, n_frames,
                                                             persons, n_persons, gt_id);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_cams);
  return e == cudaSuccess ? SES3D_OK : SES3D_E_CUDA;
}
