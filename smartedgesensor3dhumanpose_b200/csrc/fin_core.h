// fin_core.h — per-frame compaction and merge of the triangulated skeletons (kernel K4 "finalize").
//
// Replaces the tail of triangulate_persons: collecting the kept persons in hypothesis order
// (the reference's `omp critical` append S3D:977-981 taken in its serial, deterministic order)
// and the pairwise merge of skeletons closer than merge_dist_thresh (S3D:984-996 with
// calc_3D_dist S3D:392-408, merge_persons S3D:410-423, mergeKeypointCovariance S3D:264-271).
#pragma once
#include "common.h"
#include "team.h"

namespace ses3d {

constexpr int FIN_MAX_WARPS = 16;   // per-warp term buffers are provisioned for teams of up to 16 warps

struct FinWs {
  int* list;    // [h_cap] kept hypothesis indices, in order
  int* scal;    // [2]
  double* D;    // [h_cap*h_cap] pairwise distances of the kept persons
  double* term; // [FIN_MAX_WARPS][32] per-joint distance terms of the pair a warp is working on (-1 = joint not shared)
};

template <class A>
SES_HD void fin_ws_layout(A& ar, int h_cap, FinWs* ws) {
  double* D = ar.template take<double>((size_t)h_cap * h_cap);
  double* term = ar.template take<double>((size_t)FIN_MAX_WARPS * 32);
  int* list = ar.template take<int>(h_cap);
  int* scal = ar.template take<int>(2);
  if (ws) { ws->D = D; ws->list = list; ws->scal = scal; ws->term = term; }
}
inline size_t fin_ws_bytes(int h_cap) {
  ArenaSizer s;
  fin_ws_layout(s, h_cap, nullptr);
  return (s.used + 15) / 16 * 16;
}

SES_HD double dist3d(const ses3d_person_cov& a, const ses3d_person_cov& b) {  // calc_3D_dist S3D:392-408
  int n = 0;
  double d = 0;
  for (int s = 0; s < NFUS; ++s) {
    const ses3d_keypoint_cov &p = a.keypoints[s], &q = b.keypoints[s];
    if (p.score > 0 && q.score > 0) {
      const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      d += sqrt(dx * dx + dy * dy + dz * dz);
      ++n;
    }
  }
  return n > 0 ? d / n : MAX_COSTS;
}

SES_HD void merge_into(ses3d_person_cov& a, const ses3d_person_cov& b) {  // merge_persons S3D:410-423
  for (int s = 0; s < NFUS; ++s) {
    ses3d_keypoint_cov& p = a.keypoints[s];
    const ses3d_keypoint_cov& q = b.keypoints[s];
    const double total = static_cast<double>(p.score + q.score);
    if (total > 0.0) {
      p.x = ((double)p.score * p.x + (double)q.score * q.x) / total;
      p.y = ((double)p.score * p.y + (double)q.score * q.y) / total;
      p.z = ((double)p.score * p.z + (double)q.score * q.z) / total;
      p.score = p.score > q.score ? p.score : q.score;
      for (int i = 0; i < 6; ++i) p.cov[i] = (p.cov[i] + q.cov[i]) / 2.0;
    }
  }
}

// tmp [h_cap]: records written by triangulate_hypothesis (modified in place by merges);
// keep [h_cap]; out [h_cap]; *n_out. *overflow is left untouched unless h_cap is exceeded (cannot happen:
// kept persons <= hypotheses <= h_cap).
template <class Team>
SES_HD void finalize_frame(Team& tm, const Tables& tb, int h_cap, int n_hyp, ses3d_person_cov* tmp,
                           const int32_t* keep, const FinWs& ws, ses3d_person_cov* out, int32_t* n_out) {
  tm.single([&] {
    int n = 0;
    for (int h = 0; h < n_hyp && h < h_cap; ++h)
      if (keep[h]) ws.list[n++] = h;
    ws.scal[0] = n;
  });
  const int n0 = ws.scal[0];
  // calc_3D_dist of every pair (S3D:392-408), one warp per pair: the 21 per-joint distances (an FP64 square root each)
  // in parallel, then summed by the leader in joint order - the reference's sequential sum, bit for bit. (One thread
  // per pair left ~10 threads of the CTA grinding through 21 dependent square roots while the rest waited at the
  // barrier: 43 % of k_finproj's barrier stalls.)
  const int slots = tm.n_warps();   // item e runs on warp e % n_warps; the launchers keep teams at <= FIN_MAX_WARPS warps
  tm.per_warp(n0 * (n0 - 1) / 2, [&](auto& wt, int e) {
    int i = 0, rem = e;   // e -> (i, j), i < j, rows of the upper triangle in order
    while (rem >= n0 - 1 - i) { rem -= n0 - 1 - i; ++i; }
    const int j = i + 1 + rem;
    const ses3d_person_cov &a = tmp[ws.list[i]], &b = tmp[ws.list[j]];
    double* term = ws.term + (size_t)(e % slots) * 32;
    wt.pfor(NFUS, [&](int s) {
      const ses3d_keypoint_cov &p = a.keypoints[s], &q = b.keypoints[s];
      double t = -1.0;
      if (p.score > 0 && q.score > 0) {
        const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        t = sqrt(dx * dx + dy * dy + dz * dz);
      }
      term[s] = t;
    });
    wt.single([&] {
      int n = 0;
      double d = 0;
      for (int s = 0; s < NFUS; ++s)
        if (!(term[s] == -1.0)) { d += term[s]; ++n; }   // -1 = joint not shared; a NaN distance is summed like the reference does
      ws.D[i * n0 + j] = n > 0 ? d / n : MAX_COSTS;
    });
  });
  tm.single([&] {
    // positions hold original indices into list/D; erased entries are compacted away
    int n = n0;
    int* pos = ws.list;  // pos[i] = hypothesis index; orig index recovered through D row bookkeeping below
    // D is indexed by original positions; keep a parallel array of original positions in scal-free storage:
    // reuse the lower triangle of D (never written above) as int storage is avoided for clarity: recompute
    // the original position by searching is O(n); n is tiny.
    // Simpler: carry original positions in the upper 16 bits of list entries.
    for (int i = 0; i < n; ++i) pos[i] = pos[i] | (i << 16);
    for (int i = 0; i < n; ++i) {
      bool dirty = false;
      for (int j = i + 1; j < n;) {
        const int hi = pos[i] & 0xFFFF, hj = pos[j] & 0xFFFF, oi = pos[i] >> 16, oj = pos[j] >> 16;
        const double d = dirty ? dist3d(tmp[hi], tmp[hj]) : ws.D[oi * n0 + oj];
        if (d < tb.prm.merge_dist_thresh) {
          merge_into(tmp[hi], tmp[hj]);
          dirty = true;
          for (int t = j; t + 1 < n; ++t) pos[t] = pos[t + 1];
          --n;
        } else {
          ++j;
        }
      }
    }
    for (int i = 0; i < n; ++i) pos[i] &= 0xFFFF;
    ws.scal[1] = n;
  });
  const int n = ws.scal[1];
  const int words = (int)(sizeof(ses3d_person_cov) / 8);
  tm.pfor(n * words, [&](int e) {
    const int i = e / words, w = e % words;
    reinterpret_cast<uint64_t*>(out + i)[w] = reinterpret_cast<const uint64_t*>(tmp + ws.list[i])[w];
  });
  tm.single([&] { *n_out = n; });
}

}  // namespace ses3d
