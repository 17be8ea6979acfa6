// hostsim.cpp — TEST-ONLY serial instantiation of the device algorithms.
//
// The per-frame algorithms of the CUDA library (csrc/*_core.h) are __host__ __device__
// templates over a "team". This file instantiates them with SerialTeam and the library's own
// K0 tables so that the GPU-less container can check the *device logic* (not just the oracle)
// against the oracle. It is built by tests/hostsim/build.py into tests/hostsim/libhostsim.so,
// is never loaded by the product package, and is not a CPU fallback: libses3d.so has none.
#include <algorithm>
#include <cstring>
#include <vector>

#include "assoc_core.h"
#include "fin_core.h"
#include "host_setup.h"
#include "markers_core.h"
#include "prior_core.h"
#include "reproj_core.h"
#include "tri_core.h"

using namespace ses3d;

namespace {
struct Sim {
  HostTables host;
  Tables tb;
};
}  // namespace

static bool g_big_rig_path = false;   // association as the kernel runs it for rigs that do not fit shared memory

extern "C" {

void hostsim_set_big_rig_path(int32_t on) { g_big_rig_path = on != 0; }

void* hostsim_create(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* prm) {
  Sim* s = new Sim;
  if (!build_host_tables(n_cams, cams, *prm, &s->host)) { delete s; return nullptr; }
  s->tb.n_cams = n_cams;
  s->tb.camf = s->host.camf.data();
  s->tb.camd = s->host.camd.data();
  s->tb.F = s->host.F.data();
  s->tb.f_row = s->host.f_row.data();
  s->tb.model = s->host.model;
  s->tb.prm = *prm;
  s->tb.exact_mode = 3;
  return s;
}
void hostsim_destroy(void* h) { delete static_cast<Sim*>(h); }

void hostsim_get_tables(void* h, float* P, float* F) {
  Sim* s = static_cast<Sim*>(h);
  for (int i = 0; i < s->host.n_cams; ++i) std::memcpy(P + (size_t)i * 12, s->host.camf[i].P, 48);
  std::memcpy(F, s->host.F.data(), s->host.F.size() * 4);
}

int hostsim_triangulate_batch(void* h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                              const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out, int32_t* n_out,
                              int32_t* hyp_of, int32_t* n_hyp_out, int32_t* n_hung_out) {
  Sim* s = static_cast<Sim*>(h);
  const Tables& tb = s->tb;
  const int C = tb.n_cams;
  const bool big = g_big_rig_path;   // keypoints in "global scratch"
  std::vector<unsigned char> wsa(pair_ws_bytes(C, p_max, !big, 1) + 64), wsr(round_ws_bytes(C, p_max, h_max) + 64),
      wsf(fin_ws_bytes(h_max) + 64), meta_buf(frame_meta_bytes(C, p_max) + 64);
  std::vector<float> nk_scratch(big ? (size_t)C * p_max * NKP * 2 : 0);
  const bool f64 = tb.prm.precision == SES3D_PRECISION_FP64;
  std::vector<unsigned char> wst((f64 ? tri_ws_bytes<double>(C) : tri_ws_bytes<float>(C)) + 64);
  std::vector<int8_t> hyp_det((size_t)h_max * C);
  std::vector<float> far_scratch(FAR_COV_STRIDE);   // serial team: one private solve copy
  std::vector<double> pair_table(assoc_pair_table_entries(C, p_max));
  std::vector<ses3d_person_cov> tmp(h_max);
  std::vector<int32_t> keep(h_max);
  int32_t overflow = 0;
  SerialTeam tm;
  for (int f = 0; f < n_frames; ++f) {
    const ses3d_person2d* pf = persons + (size_t)f * C * p_max;
    // K2a: pair table + compact detection list; K2b: camera rounds (the two kernels of the association)
    Arena a1(wsa.data());
    AssocWs pws;
    pair_ws_layout(a1, C, p_max, !big, &pws, 1);   // with the line buffer: dense frames take the tiled pair pass
    if (big) pws.nk = nk_scratch.data();
    pws.E = pair_table.data();
    const FrameMeta meta = frame_meta_at(meta_buf.data(), C, p_max);
    pairs_frame(tm, tb, p_max, pf, n_persons + (size_t)f * C, pws, meta);
    Arena a1b(wsr.data());
    AssocWs aws;
    round_ws_layout(a1b, C, p_max, h_max, &aws);
    aws.voff = meta.voff; aws.vslot = meta.vslot; aws.pscore = meta.pscore;
    aws.E = pair_table.data();
    int32_t n_hyp = 0, n_hung = 0;
    rounds_frame(tm, tb, p_max, h_max, n_persons + (size_t)f * C, *meta.n_valid, aws, hyp_det.data(), &n_hyp, &n_hung,
                 &overflow);
    if (n_hyp_out) n_hyp_out[f] = n_hyp;
    if (n_hung_out) n_hung_out[f] = n_hung;
    if (hyp_of) {
      int32_t* ho = hyp_of + (size_t)f * C * p_max;
      for (int i = 0; i < C * p_max; ++i) ho[i] = -1;
      for (int hh = 0; hh < n_hyp; ++hh)
        for (int c = 0; c < C; ++c)
          if (hyp_det[hh * C + c] >= 0) ho[c * p_max + hyp_det[hh * C + c]] = hh;
    }
    for (int hh = 0; hh < h_max; ++hh) {
      keep[hh] = 0;
      if (hh >= n_hyp) continue;
      Arena a2(wst.data());
      if (f64) {
        TriWs<double> tws;
        tri_ws_layout<double>(a2, C, &tws);
        triangulate_hypothesis<double>(tm, tb, p_max, pf, hyp_det.data() + (size_t)hh * C, tws, &tmp[hh], &keep[hh]);
      } else {
        TriWs<float> tws;
        tri_ws_layout<float>(a2, C, &tws);
        tws.far_scratch = far_scratch.data();
        triangulate_hypothesis<float>(tm, tb, p_max, pf, hyp_det.data() + (size_t)hh * C, tws, &tmp[hh], &keep[hh]);
      }
    }
    Arena a3(wsf.data());
    FinWs fws;
    fin_ws_layout(a3, h_max, &fws);
    finalize_frame(tm, tb, h_max, n_hyp, tmp.data(), keep.data(), fws, out + (size_t)f * h_max, n_out + f);
  }
  return overflow ? SES3D_E_CAPACITY : SES3D_OK;
}

int hostsim_reproject_batch(void* h, int32_t n_frames, int32_t h_max, int32_t cam_tile, const ses3d_person_cov* p3d,
                            const int32_t* n_p3d, ses3d_person2d* out, int32_t* n_out) {
  Sim* s = static_cast<Sim*>(h);
  const Tables& tb = s->tb;
  const int C = tb.n_cams;
  // cam_tile > 0 forces person batches of that size (several passes over the sigma-point staging); 0 = one batch
  const int s_cap = reproj_s_cap(C, h_max, cam_tile > 0 ? cam_tile : h_max);
  std::vector<unsigned char> wsr(reproj_ws_bytes(C, 1, s_cap) + 64);
  SerialTeam tm;
  for (int f = 0; f < n_frames; ++f) {
    Arena a(wsr.data());
    ReprojWs ws;
    reproj_ws_layout(a, C, 1, s_cap, &ws);
    reproject_frame(tm, tb, h_max, p3d + (size_t)f * h_max, n_p3d[f], ws, out + (size_t)f * C * h_max,
                    n_out + (size_t)f * C);
  }
  return 0;
}

// ---- visualisation content (markers_core.h), same array shapes as ses3d_markers_batch
void hostsim_markers(void* h, int32_t n_frames, int32_t h_max, const ses3d_person_cov* p3d, const int32_t* n_p3d,
                     int32_t style, ses3d_ellipsoid* ell, double* seg, int32_t* n_seg, int8_t* seg_slot) {
  Sim* s = static_cast<Sim*>(h);
  for (int f = 0; f < n_frames; ++f)
    for (int p = 0; p < h_max; ++p) {
      const size_t unit = (size_t)f * h_max + p;
      const bool live = p < n_p3d[f];
      for (int k = 0; k < NFUS; ++k) {
        ses3d_ellipsoid e = {0, 0, 0, 0, 0, 0, 0};
        if (live && p3d[unit].keypoints[k].score > 0.0f) covariance_ellipsoid(p3d[unit].keypoints[k].cov, &e);
        ell[unit * NFUS + k] = e;
      }
      n_seg[unit] = live ? skeleton_segments(s->tb.model, style, p3d[unit], seg + unit * MARKER_MAX_SEGMENTS * 6,
                                             seg_slot + unit * MARKER_MAX_SEGMENTS) : 0;
    }
}

// ---- pose_prior (prior_core.h) with a serial team: same array shapes as ses3d_prior_run --------------------------
struct PriorSim {
  PriorTables pt;
  int n_seq, max_tracks;
  std::vector<PriorSeqState> states;
  std::vector<PriorTrack> tracks;
  std::vector<uint8_t> order;
};

void* hostsim_prior_create(const ses3d_prior_params* prm, int32_t n_sequences, int32_t max_tracks) {
  PriorSim* s = new PriorSim;
  s->pt.prm = *prm;
  s->pt.limb_sigma_factor = prm->normalize_by_height ? 2.0 : 1.0;
  s->pt.st = &prior_static_host();
  s->n_seq = n_sequences;
  s->max_tracks = max_tracks;
  s->states.resize(n_sequences);
  s->tracks.resize((size_t)n_sequences * max_tracks);
  s->order.assign((size_t)n_sequences * max_tracks, 0);
  std::memset(s->tracks.data(), 0, s->tracks.size() * sizeof(PriorTrack));
  for (auto& st : s->states) prior_state_reset(*prm, &st, false);
  return s;
}
void hostsim_prior_destroy(void* h) { delete static_cast<PriorSim*>(h); }

int hostsim_prior_run(void* h, int32_t n_sequences, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons,
                      const int32_t* n_persons, const int64_t* stamp_ns, int32_t n_cams, const float* fb_delay,
                      ses3d_person_cov* fused, ses3d_person_cov* pred, int32_t* n_out, float* pred_delay,
                      int32_t* track_of, int32_t group) {
  PriorSim* s = static_cast<PriorSim*>(h);
  if (n_sequences > s->n_seq) return SES3D_E_INVALID;
  size_t pers = 0, trans = 0;
  prior_ws_bytes(h_max, s->max_tracks, &pers, &trans);
  std::vector<unsigned char> wsb(pers + 64), fitb(std::max(prior_fit_ws_bytes(group), trans) + 64);
  SerialTeam tm;
  int rc = 0;
  for (int q = 0; q < n_sequences; ++q) {
    for (int f = 0; f < n_frames; ++f) {
      const size_t i = (size_t)q * n_frames + f;
      Arena ar(wsb.data()), tr(fitb.data());   // the tracker's big arrays share memory with the fit workspace
      PriorWs ws;
      prior_ws_layout(ar, tr, h_max, s->max_tracks, &ws);
      prior_frame(tm, s->pt, s->max_tracks, h_max, group, &s->states[q], s->tracks.data() + (size_t)q * s->max_tracks,
                  s->order.data() + (size_t)q * s->max_tracks, ws, fitb.data(), 0, stamp_ns[i], n_cams,
                  fb_delay ? fb_delay + i * n_cams : nullptr, n_persons[i], persons + i * h_max, fused + i * h_max,
                  pred + i * h_max, n_out + i, pred_delay ? pred_delay + i : nullptr,
                  track_of ? track_of + i * h_max : nullptr);
    }
    if (s->states[q].overflow) rc = SES3D_E_CAPACITY;
  }
  return rc;
}

int hostsim_prior_get_tracks(void* h, int32_t sequence, int32_t* ids, int32_t* num_obs) {
  PriorSim* s = static_cast<PriorSim*>(h);
  const PriorSeqState& st = s->states[sequence];
  for (int i = 0; i < st.n_tracks; ++i) {
    const PriorTrack& t = s->tracks[(size_t)sequence * s->max_tracks + s->order[(size_t)sequence * s->max_tracks + i]];
    if (ids) ids[i] = t.id;
    if (num_obs) num_obs[i] = t.num_obs;
  }
  return st.n_tracks;
}

}  // extern "C"
