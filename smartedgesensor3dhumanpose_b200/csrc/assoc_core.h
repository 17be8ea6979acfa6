// assoc_core.h — cross-view association of one frame (kernel K2 "associate").
//
// Replaces, for one frame: normalize_keypoints (S3D:312-333), calcCost (S3D:335-390), the
// Tanke-Gall iterative greedy matching loop of triangulate_persons (S3D:528-674) and the
// Munkres solver HungarianAlgorithm::assignmentoptimal (Hungarian.cpp:60-397, called S3D:630).
//
// Association indices must be bit-exact with the reference, so this file keeps the
// reference's float/double split and evaluation order and MUST be compiled without FMA
// contraction (nvcc -fmad=false; g++ -ffp-contract=off).
//
// B200 mapping: two kernels, because the two halves of the association want opposite things from an SM.
//  (1) pair table (pairs_frame, kernel K2a, one CTA per frame): calcCost only ever combines a detection of an
//      earlier camera (a hypothesis observation) with a detection of a later camera, and every earlier valid
//      detection belongs to exactly one hypothesis, so the set of (observation, detection) pairs the reference
//      evaluates over all camera rounds is exactly "all cross-camera pairs of valid detections". Their mean epipolar
//      distances E[a][b] (S3D:353-368) are computed up front in ONE flat parallel pass - all the floating-point work
//      of the association, no dependency on the matching - into an L2-resident table. Throughput-bound: wide CTAs,
//      the frame's normalised keypoints in shared memory (13 KB at 16 x 6).
//  (2) camera rounds (rounds_frame, kernel K2b, one WARP per frame): the sequential Tanke-Gall rounds only gather
//      table entries - a cost-matrix entry is the in-order mean of <= n_obs lookups - then the ambiguity test, the
//      warp-cooperative Munkres and the hypothesis update: integer work on a few KB of shared memory, latency-bound.
//      With the keypoints gone a warp needs ~3 KB, so dozens of frames are in flight per SM and hide each other's
//      latency (in round 1 both phases shared one CTA: two of three warps idled through the rounds while the CTA
//      kept 17 KB of shared memory).
//  The compact list of valid detections (FrameMeta) travels from (1) to (2) through global memory.
#pragma once
#include "common.h"
#include "team.h"

namespace ses3d {

enum { SC_N_HYP = 0, SC_N_DET, SC_N_HUNG, SC_OVERFLOW, SC_CURSOR, SC_N_VALID, SC_AMBIG, SC_COUNT };

// Per-frame hand-over from the pair kernel to the rounds kernel (global memory):
//   int32 n_valid | uint16 voff[C+1] | uint16 vslot[n_max] | (pad to 4) | float pscore[n_max]      n_max = C * p_max
struct FrameMeta {
  int32_t* n_valid;
  uint16_t* voff;     // [C+1] first compact index of each camera
  uint16_t* vslot;    // [n_max] compact index -> slot (cam * p_max + det)
  float* pscore;      // [n_max] Person2D.score, by compact index
};
SES_HD size_t frame_meta_bytes(int C, int p_max) {
  size_t b = 4 + 2 * (size_t)(C + 1) + 2 * (size_t)C * p_max;
  b = (b + 3) / 4 * 4;
  return (b + 4 * (size_t)C * p_max + 15) / 16 * 16;
}
SES_HD FrameMeta frame_meta_at(void* base, int C, int p_max) {
  unsigned char* p = static_cast<unsigned char*>(base);
  FrameMeta m;
  m.n_valid = reinterpret_cast<int32_t*>(p);
  m.voff = reinterpret_cast<uint16_t*>(p + 4);
  m.vslot = m.voff + (C + 1);
  size_t b = 4 + 2 * (size_t)(C + 1) + 2 * (size_t)C * p_max;
  b = (b + 3) / 4 * 4;
  m.pscore = reinterpret_cast<float*>(p + b);
  return m;
}

struct AssocWs {
  // ---- pair kernel
  float* nk;          // [C*p_max][17][2] normalised keypoints x, y (shared memory, or global scratch for big rigs)
  uint32_t* kmask;    // [C*p_max] bit k: keypoint k has score > threshold (the strict test of calcCost, S3D:354)
  uint8_t* valid;     // [C*p_max] more than 8 valid keypoints (S3D:579,599), by slot
  uint32_t* pstart;   // [C*p_max+1] cross-camera pairs before compact detection b (b pairs with every earlier camera's)
  float* lines;       // [n_warps][2][PAIR_TILE][17][4] epipolar lines (x, y, z, norm) of a camera-pair tile; nullptr =
                      // flat pair pass only
  // ---- both
  double* E;          // [n(n-1)/2], n = C*p_max: pair table in global memory, index b(b-1)/2 + a for a < b
                      //   (compact indices of valid detections, camera-major); -1 = no joint in common
  float* pscore;      // [C*p_max] Person2D.score, by compact index          (FrameMeta)
  uint16_t* vslot;    // [C*p_max] compact index -> slot (cam * p_max + det)  (FrameMeta)
  uint16_t* voff;     // [C+1] first compact index of each camera             (FrameMeta)
  int* scal;          // [SC_COUNT]
  // ---- rounds kernel
  uint8_t* hyp_nobs;  // [h_cap]
  uint16_t* hyp_obs;  // [h_cap][C] observation list (compact indices), in camera order
  double* cost;       // [h_cap*p_max] column-major n_hyp x n_det (S3D:611)
  double* dist;       // Munkres working copy
  uint8_t *mask, *star, *prime, *nstar;  // [h_cap*p_max]
  uint8_t *cov_r, *cov_c, *handled;      // [h_cap], [p_max], [p_max]
  int* assignment;    // [h_cap]
};

// Shared-memory workspace of the pair kernel; nk_inside = false keeps the keypoints in global scratch
// (rigs whose frame does not fit in shared memory). vslot / voff / pscore are staged here and copied to FrameMeta.
constexpr int PAIR_TILE = 20;   // detections per side of a camera-pair tile (tiled pair pass)
SES_HD size_t pair_tile_floats() { return (size_t)2 * PAIR_TILE * NKP * 4; }

// tile_warps > 0 reserves the per-warp line buffers of the tiled pair pass
template <class A>
SES_HD void pair_ws_layout(A& ar, int C, int p_max, bool nk_inside, AssocWs* ws, int tile_warps = 0) {
  float* lines = tile_warps > 0 ? ar.template take<float>((size_t)tile_warps * pair_tile_floats()) : nullptr;
  if (ws) ws->lines = lines;
  float* nk = nk_inside ? ar.template take<float>((size_t)C * p_max * NKP * 2) : nullptr;
  uint32_t* kmask = ar.template take<uint32_t>((size_t)C * p_max);
  uint32_t* pstart = ar.template take<uint32_t>((size_t)C * p_max + 1);
  int* scal = ar.template take<int>(SC_COUNT);
  uint16_t* vslot = ar.template take<uint16_t>((size_t)C * p_max);
  uint16_t* voff = ar.template take<uint16_t>(C + 1);
  uint8_t* valid = ar.template take<uint8_t>((size_t)C * p_max);
  if (ws) {
    if (nk_inside) ws->nk = nk;
    ws->kmask = kmask; ws->pstart = pstart; ws->scal = scal; ws->vslot = vslot; ws->voff = voff; ws->valid = valid;
  }
}
inline size_t pair_ws_bytes(int C, int p_max, bool nk_inside, int tile_warps = 0) {
  ArenaSizer s;
  pair_ws_layout(s, C, p_max, nk_inside, nullptr, tile_warps);
  return (s.used + 15) / 16 * 16;
}

// Shared-memory workspace of one frame in the rounds kernel (per warp). The compact detection list is read from
// FrameMeta in global memory (L1-resident, a few hundred bytes).
template <class A>
SES_HD void round_ws_layout(A& ar, int C, int p_max, int h_cap, AssocWs* ws) {
  double* cost = ar.template take<double>((size_t)h_cap * p_max);
  double* dist = ar.template take<double>((size_t)h_cap * p_max);
  int* assignment = ar.template take<int>(h_cap);
  int* scal = ar.template take<int>(SC_COUNT);
  uint16_t* hyp_obs = ar.template take<uint16_t>((size_t)h_cap * C);
  uint8_t* hyp_nobs = ar.template take<uint8_t>(h_cap);
  uint8_t* mask = ar.template take<uint8_t>((size_t)h_cap * p_max);
  uint8_t* star = ar.template take<uint8_t>((size_t)h_cap * p_max);
  uint8_t* prime = ar.template take<uint8_t>((size_t)h_cap * p_max);
  uint8_t* nstar = ar.template take<uint8_t>((size_t)h_cap * p_max);
  uint8_t* cov_r = ar.template take<uint8_t>(h_cap);
  uint8_t* cov_c = ar.template take<uint8_t>(p_max);
  uint8_t* handled = ar.template take<uint8_t>(p_max);
  if (ws) {
    ws->cost = cost; ws->dist = dist; ws->assignment = assignment; ws->scal = scal; ws->hyp_obs = hyp_obs;
    ws->hyp_nobs = hyp_nobs; ws->mask = mask; ws->star = star; ws->prime = prime; ws->nstar = nstar;
    ws->cov_r = cov_r; ws->cov_c = cov_c; ws->handled = handled;
  }
}
inline size_t round_ws_bytes(int C, int p_max, int h_cap) {
  ArenaSizer s;
  round_ws_layout(s, C, p_max, h_cap, nullptr);
  return (s.used + 15) / 16 * 16;
}

// entries of the per-frame pair table
SES_HD size_t assoc_pair_table_entries(int C, int p_max) {
  const size_t n = (size_t)C * p_max;
  return n * (n - 1) / 2 + 1;
}

// Symmetric point-to-epipolar-line distance d1 + d2 in float (S3D:355-362):
// l1 = F (x1,y1,1), l2 = F^T (x2,y2,1), d1 = |p2.l1| / sqrt(l1x^2+l1y^2), d2 = |p1.l2| / sqrt(l2x^2+l2y^2).
SES_HD float epipolar_symmetric(const float* F, float x1, float y1, float x2, float y2) {
  const float l1x = sum3(F[0] * x1, F[1] * y1, F[2] * 1.0f);
  const float l1y = sum3(F[3] * x1, F[4] * y1, F[5] * 1.0f);
  const float l1z = sum3(F[6] * x1, F[7] * y1, F[8] * 1.0f);
  const float l2x = sum3(F[0] * x2, F[3] * y2, F[6] * 1.0f);
  const float l2y = sum3(F[1] * x2, F[4] * y2, F[7] * 1.0f);
  const float l2z = sum3(F[2] * x2, F[5] * y2, F[8] * 1.0f);
  const float d1 = ses_abs(sum3(x2 * l1x, y2 * l1y, 1.0f * l1z)) / ses_sqrt(l1x * l1x + l1y * l1y);
  const float d2 = ses_abs(sum3(x1 * l2x, y1 * l2y, 1.0f * l2z)) / ses_sqrt(l2x * l2x + l2y * l2y);
  return d1 + d2;
}

// Munkres on the column-major n_r x n_c matrix `in` (HungarianAlgorithm::assignmentoptimal, Hungarian.cpp:60-397).
// The reference's mutual recursion step2a/2b/3/4/5 is unrolled into a state machine; scan orders, the
// |x| < DBL_EPSILON zero test and the +-h updates are kept so ties resolve identically. The solver is run
// cooperatively by a warp-sized team (lanes over rows / columns / entries):
// every search of the reference ("first zero in this scan order") becomes a ballot + find-first-set over
// the same index order, so primes, stars and covers - and therefore ties - resolve identically.
#ifndef SES_COLD_MUNKRES
#define SES_COLD_MUNKRES 0   // 1: the solver as an out-of-line function - measured slower on B200 (K2a + K2b 0.91 -> 1.11 ms
                             // per 16384 hall frames, dense ring 1.47 -> 1.61); A/B switch for scripts/build_variants.py
#endif
#if SES_COLD_MUNKRES
#define SES_MUNKRES_FN SES_HDN
#else
#define SES_MUNKRES_FN SES_HD
#endif
template <class WT>
SES_MUNKRES_FN void munkres_coop(WT& tm, const AssocWs& ws, const double* in, int n_r, int n_c, int* assignment) {
  const int n_e = n_r * n_c;
  double* dist = ws.dist;
  uint8_t *star = ws.star, *prime = ws.prime, *nstar = ws.nstar, *cov_r = ws.cov_r, *cov_c = ws.cov_c;
  tm.pfor(n_e, [&](int i) { dist[i] = in[i]; star[i] = 0; prime[i] = 0; nstar[i] = 0; });
  tm.pfor(n_r, [&](int r) { cov_r[r] = 0; assignment[r] = -1; });
  tm.pfor(n_c, [&](int c) { cov_c[c] = 0; });
  auto is_zero = [&](int r, int c) { return fabs(dist[r + n_r * c]) < DBL_EPSILON; };
  int min_dim;
  if (n_r <= n_c) {  // Hungarian.cpp:95-131
    min_dim = n_r;
    tm.pfor(n_r, [&](int r) {
      double mn = dist[r];
      for (int c = 1; c < n_c; ++c) { const double v = dist[r + n_r * c]; if (v < mn) mn = v; }
      for (int c = 0; c < n_c; ++c) dist[r + n_r * c] -= mn;
    });
    for (int r = 0; r < n_r; ++r) {
      const int c = tm.first(n_c, [&](int cc) { return is_zero(r, cc) && !cov_c[cc]; });
      if (c < n_c) tm.single([&] { star[r + n_r * c] = 1; cov_c[c] = 1; });
    }
  } else {  // Hungarian.cpp:132-170
    min_dim = n_c;
    tm.pfor(n_c, [&](int c) {
      double mn = dist[n_r * c];
      for (int r = 1; r < n_r; ++r) { const double v = dist[r + n_r * c]; if (v < mn) mn = v; }
      for (int r = 0; r < n_r; ++r) dist[r + n_r * c] -= mn;
    });
    for (int c = 0; c < n_c; ++c) {
      const int r = tm.first(n_r, [&](int rr) { return is_zero(rr, c) && !cov_r[rr]; });
      if (r < n_r) tm.single([&] { star[r + n_r * c] = 1; cov_c[c] = 1; cov_r[r] = 1; });
    }
    tm.pfor(n_r, [&](int r) { cov_r[r] = 0; });
  }
  enum { S2A, S2B, S3, S4, S5, DONE };
  int st = S2B, row4 = 0, col4 = 0;
  // Every pass through S4 adds a star and every S5 uncovers a zero, so a finite matrix needs far fewer than
  // 4 (n_r + n_c)^2 + 64 transitions; the cap only guards against non-finite costs (NaN keypoints), for which
  // the reference's own behaviour is undefined - the kernel must never spin.
  int budget = 4 * (n_r + n_c) * (n_r + n_c) + 64;
  while (st != DONE && --budget > 0) {
    if (st == S2A) {  // Hungarian.cpp:222-242
      tm.pfor(n_c, [&](int c) {
        for (int r = 0; r < n_r; ++r)
          if (star[r + n_r * c]) { cov_c[c] = 1; break; }
      });
      st = S2B;
    } else if (st == S2B) {  // Hungarian.cpp:245-266
      int n = 0;
      for (int c = 0; c < n_c; ++c) n += cov_c[c] ? 1 : 0;
      st = (n == min_dim) ? DONE : S3;
    } else if (st == S3) {  // Hungarian.cpp:269-309: column-outer scan, rows searched by ballot
      bool zeros = true, to4 = false;
      while (zeros && !to4) {
        zeros = false;
        for (int c = 0; c < n_c && !to4; ++c) {
          if (cov_c[c]) continue;
          const int r = tm.first(n_r, [&](int rr) { return !cov_r[rr] && is_zero(rr, c); });
          if (r == n_r) continue;
          const int sc = tm.first(n_c, [&](int cc) { return star[r + n_r * cc] != 0; });
          tm.single([&] {
            prime[r + n_r * c] = 1;
            if (sc < n_c) { cov_r[r] = 1; cov_c[sc] = 0; }
          });
          if (sc == n_c) { row4 = r; col4 = c; to4 = true; }
          else zeros = true;
        }
      }
      st = to4 ? S4 : S5;
    } else if (st == S4) {  // Hungarian.cpp:312-363: the alternating path is a short serial chain
      tm.pfor(n_e, [&](int i) { nstar[i] = star[i]; });
      tm.single([&] {
        nstar[row4 + n_r * col4] = 1;
        int sc = col4, sr = 0;
        for (sr = 0; sr < n_r; ++sr)
          if (star[sr + n_r * sc]) break;
        while (sr < n_r) {
          nstar[sr + n_r * sc] = 0;
          int pc = 0;
          for (; pc < n_c; ++pc)
            if (prime[sr + n_r * pc]) break;
          nstar[sr + n_r * pc] = 1;
          sc = pc;
          for (sr = 0; sr < n_r; ++sr)
            if (star[sr + n_r * sc]) break;
        }
      });
      tm.pfor(n_e, [&](int i) { prime[i] = 0; star[i] = nstar[i]; });
      tm.pfor(n_r, [&](int r) { cov_r[r] = 0; });
      st = S2A;
    } else {  // S5, Hungarian.cpp:366-397
      const double h = tm.min(n_e, [&](int e) {
        const int r = e % n_r, c = e / n_r;
        return (!cov_r[r] && !cov_c[c]) ? dist[e] : DBL_MAX;
      });
      tm.pfor(n_e, [&](int e) {  // per entry the same sequence as the reference: + h (covered row), then - h (uncovered column)
        const int r = e % n_r, c = e / n_r;
        double v = dist[e];
        if (cov_r[r]) v += h;
        if (!cov_c[c]) v -= h;
        dist[e] = v;
      });
      st = S3;
    }
  }
  tm.pfor(n_r, [&](int r) {  // buildassignmentvector (Hungarian.cpp:190-205)
    for (int c = 0; c < n_c; ++c)
      if (star[r + n_r * c]) { assignment[r] = c; break; }
  });
}

// K2a, one frame. persons [C][p_max], n_persons [C]. Outputs: the pair table ws.E and the compact detection list
// (meta). ws = pair_ws_layout.
// Dense frame = on average >= 4 valid detections per camera: its pair table is built by camera-pair tiles (needs the
// line buffers ws.lines). defer_dense: this call has no line buffers and leaves the pair table of dense frames to a
// second call that has them (GPU: k_pairs with a small workspace for the usual sparse frames, k_pairs_dense with the
// line buffers for the rest; both run pairs_frame, a frame's table is written by exactly one of them).
SES_HD bool pairs_frame_is_dense(int n_valid, int C) { return n_valid >= 4 * C; }

// part / n_parts: the frame's pair list may be cut into n_parts slices handled by different teams (crowd rigs: one frame
// has ~10^6 pairs and there are too few frames to fill the GPU). Every slice repeats the cheap set-up; slice 0 alone
// writes the frame's meta record.
template <class Team>
SES_HD void pairs_frame(Team& tm, const Tables& tb, int p_max, const ses3d_person2d* persons, const int32_t* n_persons,
                        const AssocWs& ws, const FrameMeta& meta, bool defer_dense = false, int part = 0, int n_parts = 1) {
  const int C = tb.n_cams;
  const float thr = tb.prm.triangulation_threshold;
  auto np = [&](int c) { const int n = n_persons[c]; return n < 0 ? 0 : (n > p_max ? p_max : n); };

  // normalize_keypoints for every detection of the frame (S3D:312-333)
  tm.pfor(C * p_max * NKP, [&](int i) {
    const int k = i % NKP, cd = i / NKP, c = cd / p_max, d = cd % p_max;
    float x = 0.f, y = 0.f;
    if (d < np(c)) {
      const ses3d_keypoint2d& kp = persons[cd].keypoints[k];
      const CamF& cm = tb.camf[c];
      if (kp.score >= thr) {
        x = (kp.x - cm.cx) / cm.fx;
        y = (kp.y - cm.cy) / cm.fy;
      }
    }
    // one store of the final value: when the keypoints live in global scratch, other slices of the same frame write
    // the identical values concurrently and may already be reading
    float* o = ws.nk + (size_t)i * 2;
    o[0] = x; o[1] = y;
  });
  tm.pfor(C * p_max, [&](int cd) {
    const int c = cd / p_max, d = cd % p_max;
    int n_valid = 0;
    uint32_t strict = 0;
    if (d < np(c))
      for (int k = 0; k < NKP; ++k) {
        const float sc = persons[cd].keypoints[k].score;
        n_valid += sc >= thr ? 1 : 0;                 // S3D:321
        strict |= sc > thr ? (1u << k) : 0u;          // S3D:354 (normalised conf = score when >= thr, else -1)
      }
    ws.valid[cd] = n_valid > NKP / 2 ? 1 : 0;
    ws.kmask[cd] = strict;
  });
  // compact, camera-major list of the valid detections; pstart[b] = cross-camera pairs (a, b'), b' < b
  tm.single([&] {
    int n = 0;
    uint32_t pairs = 0;
    for (int c = 0; c < C; ++c) {
      ws.voff[c] = (uint16_t)n;
      const int before = n;   // every detection of an earlier camera pairs with each detection of camera c
      for (int d = 0; d < np(c); ++d)
        if (ws.valid[c * p_max + d]) {
          ws.vslot[n] = (uint16_t)(c * p_max + d);
          ws.pstart[n] = pairs;
          pairs += (uint32_t)before;
          ++n;
        }
    }
    ws.voff[C] = (uint16_t)n;
    ws.pstart[n] = pairs;
    ws.scal[SC_N_VALID] = n;
    if (part == 0) *meta.n_valid = n;
  });
  const int n_valid = ws.scal[SC_N_VALID];
  if (part == 0) {
    tm.pfor(n_valid, [&](int a) { meta.vslot[a] = ws.vslot[a]; meta.pscore[a] = persons[ws.vslot[a]].score; });
    tm.pfor(C + 1, [&](int c) { meta.voff[c] = ws.voff[c]; });
    // same-camera pairs are never evaluated (and never looked up); give their table entries a defined value so that
    // the triangle [0, n_valid(n_valid-1)/2) can be copied as a block (low-latency k_rounds stages it in shared memory)
    tm.pfor(n_valid, [&](int b) {
      const int c = ws.vslot[b] / p_max;
      for (int a = ws.voff[c]; a < b; ++a) ws.E[(size_t)b * (b - 1) / 2 + a] = -1.0;
    });
  }

  // pair table: mean symmetric epipolar distance of every cross-camera pair of valid detections, the inner loop of
  // calcCost (S3D:347-368), joints in ascending order, float distances summed in double.
  //
  // Dense frames (crowds: on average >= 4 valid detections per camera): camera-pair tiles, one warp each. The epipolar
  // line of a keypoint in the other camera (l1 = F p1 resp. l2 = F^T p2) and its norm depend on one detection and the
  // camera pair only, so a tile of na x nb pairs computes its (na + nb) x 17 lines once into the warp's buffer and a
  // pair costs two dot products and two divisions per joint instead of two 3x3 products, two square roots and two
  // divisions - the same operations on the same values, hence the same bits. Sparse frames (the hall rig sees ~2
  // detections per camera: a line would be used twice) keep the flat pass below.
  if (defer_dense && pairs_frame_is_dense(n_valid, C)) return;
  if (ws.lines && pairs_frame_is_dense(n_valid, C)) {
    tm.per_warp(C * (C - 1) / 2, [&](auto& wt, int t) {
      int ca = 0;   // t = f_row[ca] + cb - ca - 1 (get_fundamental_idx order)
      while (ca + 2 < C && tb.f_row[ca + 1] <= t) ++ca;
      const int cb = t - tb.f_row[ca] + ca + 1;
      const int a_beg = ws.voff[ca], na = ws.voff[ca + 1] - a_beg, b_beg = ws.voff[cb], nb = ws.voff[cb + 1] - b_beg;
      if (na == 0 || nb == 0) return;
      const float* F = tb.F + (size_t)t * 9;
      float* LA = ws.lines + (size_t)(t % tm.n_warps()) * pair_tile_floats();   // [PAIR_TILE][17][4], detections of ca
      float* LB = LA + (size_t)PAIR_TILE * NKP * 4;                              // detections of cb
      for (int ia = 0; ia < na; ia += PAIR_TILE)
        for (int ib = 0; ib < nb; ib += PAIR_TILE) {
          const int ta = (na - ia) < PAIR_TILE ? (na - ia) : PAIR_TILE, tb_n = (nb - ib) < PAIR_TILE ? (nb - ib) : PAIR_TILE;
          wt.pfor((ta + tb_n) * NKP, [&](int i) {
            const int k = i % NKP, q = i / NKP;
            const bool side_a = q < ta;
            const int slot = side_a ? ws.vslot[a_beg + ia + q] : ws.vslot[b_beg + ib + (q - ta)];
            if (!((ws.kmask[slot] >> k) & 1u)) return;
            const float x = ws.nk[((size_t)slot * NKP + k) * 2], y = ws.nk[((size_t)slot * NKP + k) * 2 + 1];
            float lx, ly, lz;
            if (side_a) {   // l1 = F (x1, y1, 1)
              lx = sum3(F[0] * x, F[1] * y, F[2] * 1.0f); ly = sum3(F[3] * x, F[4] * y, F[5] * 1.0f);
              lz = sum3(F[6] * x, F[7] * y, F[8] * 1.0f);
            } else {        // l2 = F^T (x2, y2, 1)
              lx = sum3(F[0] * x, F[3] * y, F[6] * 1.0f); ly = sum3(F[1] * x, F[4] * y, F[7] * 1.0f);
              lz = sum3(F[2] * x, F[5] * y, F[8] * 1.0f);
            }
            float* o = (side_a ? LA + (size_t)(q * NKP + k) * 4 : LB + (size_t)((q - ta) * NKP + k) * 4);
            o[0] = lx; o[1] = ly; o[2] = lz; o[3] = ses_sqrt(lx * lx + ly * ly);
          });
          wt.pfor(ta * tb_n, [&](int i) {
            const int qa = i / tb_n, qb = i % tb_n;
            const int a = a_beg + ia + qa, b = b_beg + ib + qb;
            const int sa = ws.vslot[a], sb = ws.vslot[b];
            const float* hk = ws.nk + ((size_t)sa * NKP) * 2;
            const float* dk = ws.nk + ((size_t)sb * NKP) * 2;
            uint32_t m = ws.kmask[sa] & ws.kmask[sb];
            double cost = 0.;
            int n_joints = 0;
            while (m) {
              const int k = ses_ctz(m);
              m &= m - 1;
              const float* l1 = LA + (size_t)(qa * NKP + k) * 4;
              const float* l2 = LB + (size_t)(qb * NKP + k) * 4;
              const float d1 = ses_abs(sum3(dk[2 * k] * l1[0], dk[2 * k + 1] * l1[1], 1.0f * l1[2])) / l1[3];
              const float d2 = ses_abs(sum3(hk[2 * k] * l2[0], hk[2 * k + 1] * l2[1], 1.0f * l2[2])) / l2[3];
              cost += static_cast<double>(d1 + d2);
              ++n_joints;
            }
            ws.E[(size_t)b * (b - 1) / 2 + a] = n_joints > 0 ? cost / n_joints : -1.0;
          });
        }
    });
    return;
  }
  const long long n_pairs = (long long)ws.pstart[n_valid];
  const int e_lo = (int)(n_pairs * part / n_parts), e_hi = (int)(n_pairs * (part + 1) / n_parts);
  tm.pfor(e_hi - e_lo, [&](int e_rel) {
    const int e = e_lo + e_rel;
    // e -> (a, b): b = the detection whose pair range contains e (binary search), a = offset inside it
    int lo = 0, hi = n_valid;   // invariant: pstart[lo] <= e < pstart[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ws.pstart[mid] <= (uint32_t)e) lo = mid; else hi = mid;
    }
    const int b = lo, a = e - (int)ws.pstart[b];
    const int sa = ws.vslot[a], sb = ws.vslot[b];
    const int ca = sa / p_max, cb = sb / p_max;
    const float* F = tb.F + (size_t)fundamental_idx(tb, ca, cb) * 9;
    const float* hk = ws.nk + ((size_t)sa * NKP) * 2;
    const float* dk = ws.nk + ((size_t)sb * NKP) * 2;
    uint32_t m = ws.kmask[sa] & ws.kmask[sb];
    double cost = 0.;
    int n_joints = 0;
    while (m) {
      const int k = ses_ctz(m);
      m &= m - 1;
      cost += static_cast<double>(epipolar_symmetric(F, hk[2 * k], hk[2 * k + 1], dk[2 * k], dk[2 * k + 1]));
      ++n_joints;
    }
    ws.E[(size_t)b * (b - 1) / 2 + a] = n_joints > 0 ? cost / n_joints : -1.0;
  });
}

// K2b, one frame: the Tanke-Gall camera rounds over the pair table. ws = round_ws_layout + E + the FrameMeta
// pointers (pscore, vslot, voff). Outputs: hyp_det [h_cap][C] (detection slot of hypothesis h in camera c or -1),
// n_hyp, n_hungarian, overflow flag (h_cap exceeded).
template <class Team>
SES_HD void rounds_frame(Team& tm, const Tables& tb, int p_max, int h_cap, const int32_t* n_persons, int n_valid,
                         const AssocWs& ws, int8_t* hyp_det, int32_t* n_hyp_out, int32_t* n_hung_out,
                         int32_t* overflow_out) {
  const int C = tb.n_cams;
  const double max_epi = tb.prm.max_epipolar_error;
  auto np = [&](int c) { const int n = n_persons[c]; return n < 0 ? 0 : (n > p_max ? p_max : n); };
  (void)n_valid;

  auto add_hyp = [&](int a) {  // push_back of a one-observation hypothesis
    const int h = ws.scal[SC_N_HYP];
    if (h >= h_cap) { ws.scal[SC_OVERFLOW] = 1; return; }
    ws.hyp_obs[(size_t)h * C] = (uint16_t)a;
    ws.hyp_nobs[h] = 1;
    ws.scal[SC_N_HYP] = h + 1;
  };

  // cameras with detections, seed hypotheses (S3D:538-586)
  tm.single([&] {
    int n_with = 0;
    for (int c = 0; c < C; ++c) n_with += np(c) > 0 ? 1 : 0;
    ws.scal[SC_N_HYP] = 0; ws.scal[SC_N_HUNG] = 0; ws.scal[SC_OVERFLOW] = 0; ws.scal[SC_N_DET] = 0;
    int c = C;
    if (n_with >= 2) {
      for (c = 0; c < C; ++c) {
        if (np(c) == 0) continue;
        for (int a = ws.voff[c]; a < ws.voff[c + 1]; ++a) add_hyp(a);
        if (ws.scal[SC_N_HYP] > 0) { ++c; break; }
      }
    }
    ws.scal[SC_CURSOR] = c;
  });

  // camera rounds (S3D:588-674)
  for (int cam = ws.scal[SC_CURSOR]; cam < C; ++cam) {
    const int b0 = ws.voff[cam], n_det = ws.voff[cam + 1] - b0, n_hyp = ws.scal[SC_N_HYP];
    if (n_det == 0) continue;  // covers "no person" and "no valid person" (S3D:539-541, 608-609)

    // cost matrix entry = outer part of calcCost (S3D:367-389) over table lookups, observations in order
    tm.pfor(n_hyp * n_det, [&](int e) {
      const int h = e % n_hyp, b = b0 + e / n_hyp;
      const int n_obs = ws.hyp_nobs[h];
      double total = 0., tmp_veto = 0.;
      int n_used = 0;
      const double tolerance = 1.0 - 1.0 / (2 * n_obs), veto_delta = 1.0 / n_obs;
      const double* Eb = ws.E + (size_t)b * (b - 1) / 2;
      for (int o = 0; o < n_obs; ++o) {
        const int a = ws.hyp_obs[(size_t)h * C + o];
        const double cost = Eb[a];
        if (cost >= 0.0) {  // the pair shared at least one joint
          total += cost;
          ++n_used;
          if (cost > max_epi && (ws.pscore[a] > 0.5f || n_obs == 1)) tmp_veto += veto_delta;
        }
      }
      bool veto = tmp_veto > tolerance;
      double c;
      if (n_used > 0) c = total / n_used;
      else { veto = true; c = MAX_COSTS; }
      ws.cost[e] = c;  // column-major: h + n_hyp * d
      ws.mask[e] = (!veto && c < max_epi) ? 1 : 0;
    });

    // provisional assignment = the last passing detection per hypothesis (S3D:616-626); the Munkres solve is
    // needed when any row or column of the mask has more than one hit (S3D:628). One thread per row / column;
    // the flag write is the same value from every writer.
    const bool ambiguous = tm.count(n_hyp + n_det, [&](int i) {
      int cnt = 0;
      if (i < n_hyp) {
        int last = -1;
        for (int d = 0; d < n_det; ++d)
          if (ws.mask[i + n_hyp * d]) { last = d; ++cnt; }
        ws.assignment[i] = last;
      } else {
        const int d = i - n_hyp;
        ws.handled[d] = 0;   // (reset for the update below)
        for (int h = 0; h < n_hyp; ++h) cnt += ws.mask[h + n_hyp * d];
      }
      return cnt > 1;
    }) > 0;
    tm.sync();   // assignment / handled are read by other threads next
    if (ambiguous) {  // S3D:628-634: full Munkres on the cost matrix
      tm.warp0([&](auto& wt) {
        const AssocWs wc = ws;   // a copy: an out-of-line solver must not make the workspace struct addressable
        munkres_coop(wt, wc, wc.cost, n_hyp, n_det, wc.assignment);
      });
    }
    // Apply the assignment (S3D:637-673). The assignment is a matching (one detection per hypothesis at most and vice
    // versa), so the per-hypothesis updates are independent; the new one-observation hypotheses keep the reference's
    // push_back order - first the assigned-but-vetoed detections in hypothesis order, then the unassigned detections
    // in detection order - through two stream compactions.
    tm.pfor(n_hyp, [&](int h) {
      const int d = ws.assignment[h];
      if (d < 0) return;
      ws.handled[d] = 1;
      if (ws.mask[h + n_hyp * d]) {
        const int o = ws.hyp_nobs[h];
        ws.hyp_obs[(size_t)h * C + o] = (uint16_t)(b0 + d);
        ws.hyp_nobs[h] = (uint8_t)(o + 1);
      }
    });
    auto new_hyp = [&](int idx, int a) {   // add_hyp at a known position; beyond h_cap it only counts (overflow below)
      if (idx >= h_cap) return;
      ws.hyp_obs[(size_t)idx * C] = (uint16_t)a;
      ws.hyp_nobs[idx] = 1;
    };
    const int n_vetoed = tm.compact(n_hyp, [&](int h) { const int d = ws.assignment[h]; return d >= 0 && !ws.mask[h + n_hyp * d]; },
                                    [&](int h, int pos) { new_hyp(n_hyp + pos, b0 + ws.assignment[h]); });
    const int n_free = tm.compact(n_det, [&](int d) { return !ws.handled[d]; },
                                  [&](int d, int pos) { new_hyp(n_hyp + n_vetoed + pos, b0 + d); });
    tm.single([&] {
      int total = n_hyp + n_vetoed + n_free;
      if (total > h_cap) { ws.scal[SC_OVERFLOW] = 1; total = h_cap; }
      ws.scal[SC_N_HYP] = total;
      if (ambiguous) ++ws.scal[SC_N_HUNG];
    });
  }

  // export the hypothesis table
  const int n_hyp = ws.scal[SC_N_HYP];
  tm.pfor(h_cap * C, [&](int i) { hyp_det[i] = -1; });
  tm.pfor(n_hyp * C, [&](int i) {
    const int h = i / C, o = i % C;
    if (o < ws.hyp_nobs[h]) {
      const int slot = ws.vslot[ws.hyp_obs[(size_t)h * C + o]];
      hyp_det[h * C + slot / p_max] = (int8_t)(slot % p_max);
    }
  });
  tm.single([&] {
    *n_hyp_out = n_hyp;
    if (n_hung_out) *n_hung_out = ws.scal[SC_N_HUNG];
    if (ws.scal[SC_OVERFLOW]) *overflow_out = 1;
  });
}

}  // namespace ses3d
