// reproj_core.h — semantic-feedback reprojection of one frame (kernel K6 "reproject").
//
// Replaces fusedSkeletonCallback (REP:139-235) with draw_sigma_points (REP:62-75): per joint
// with score > 0 the 3x3 covariance is Cholesky-factored (Eigen llt), 7 sigma points are
// pushed through T_cam<-base and image_geometry::project3dToPixel, giving a pixel mean and
// 2x2 covariance per camera; joints whose mean falls outside the image are skipped, the
// bounding box is the min/max of the accepted means, and a person is emitted for a camera
// only if it has at least one accepted joint. FP64 compute, FP32 store, like the reference.
//
// B200 mapping: one CTA per frame, ONE WARP PER CAMERA inside it. Per batch of <= s_cap persons the whole CTA factors
// the covariances and stores the sigma points once (one thread per (person, joint)); after that single CTA barrier
// every warp handles its camera on its own - cull pass, exact projection of the survivors, bounding boxes, per-camera
// compaction and the coalesced record copy - with warp-level synchronisation only. (The first version ran every one of
// these steps CTA-wide: ~25 CTA barriers per frame, 39 % of all stall samples were barrier waits,
// profiles/r02_ncu_full_summary.json.)
#pragma once
#include "common.h"
#include "team.h"

namespace ses3d {

struct ReprojWs {
  double* S;              // [s_cap*17][21] sigma points of the current person batch (7 points x xyz)
  float* sscore;          // [s_cap*17] 3-D score (0 = joint absent)
  float* ctr;             // [s_cap*17][4] single-precision joint centre x, y, z and the sigma-point radius
  int* nslot;             // [C] records emitted so far per camera (push_back order = person order, REP:229)
  // per warp slot (n_slots of each):
  ses3d_person2d* stage;  // [n_slots][s_cap] the camera's Person2D records of the current person batch
  uint8_t* vflag;         // [n_slots][s_cap*17] joint accepted
  uint16_t* list;         // [n_slots][s_cap*17] (person, joint) items that need the exact projection
  int* cnt;               // [n_slots][2] length of list; output slot of the person being emitted
  int s_cap;              // persons per batch
  int n_cams;
  int n_slots;            // warps of the team (1 for warp / serial teams)
};

template <class A>
SES_HD void reproj_ws_layout(A& ar, int n_cams, int n_slots, int s_cap, ReprojWs* ws) {
  double* S = ar.template take<double>((size_t)s_cap * NKP * 21);
  ses3d_person2d* stage = ar.template take<ses3d_person2d>((size_t)n_slots * s_cap);
  float* sscore = ar.template take<float>((size_t)s_cap * NKP);
  float* ctr = ar.template take<float>((size_t)s_cap * NKP * 4);
  int* nslot = ar.template take<int>((size_t)n_cams);
  int* cnt = ar.template take<int>((size_t)n_slots * 2);
  uint16_t* list = ar.template take<uint16_t>((size_t)n_slots * s_cap * NKP);
  uint8_t* vflag = ar.template take<uint8_t>((size_t)n_slots * s_cap * NKP);
  if (ws) {
    ws->S = S; ws->stage = stage; ws->sscore = sscore; ws->vflag = vflag; ws->s_cap = s_cap; ws->nslot = nslot;
    ws->ctr = ctr; ws->cnt = cnt; ws->list = list; ws->n_cams = n_cams; ws->n_slots = n_slots;
  }
}
inline size_t reproj_ws_bytes(int n_cams, int n_slots, int s_cap) {
  ArenaSizer s;
  reproj_ws_layout(s, n_cams, n_slots, s_cap, nullptr);
  return (s.used + 15) / 16 * 16;
}
// persons per batch: bounded by the frame capacity
inline int reproj_s_cap(int n_cams, int h_max, int want) {
  (void)n_cams;
  const int s = want < h_max ? want : h_max;
  return s < 1 ? 1 : s;
}

// Conservative single-precision pre-test of REP:207-208. A person is in view of only a few cameras, so most
// (joint, camera) pairs end in "mean pixel outside the image -> skip". With the joint centre x0 at depth Z0 and all
// seven sigma points within radius r of it (r = sqrt(3.5) x the largest column norm of the Cholesky factor), every
// sigma point has depth >= Z0 - r and its pixel lies within B = (f (Z0 + |X0|) + |T|) r / (Z0 (Z0 - r)) of the centre's
// pixel, and so does their weighted mean (weights are positive and sum to one). If the centre's pixel is more than
// B + 16 px outside the image, the exact FP64 mean is outside as well and the fourteen IEEE divisions of the exact
// path are skipped. Conditions that keep single precision trustworthy (else the exact path runs): Z0 - r >= 0.5 m,
// |x0| < 100 m, centre within 1e4 px of the principal point -> float evaluation error < 1.1 px (coordinate rounding
// 2.4e-5 m, f dX / Z <= 0.05 px, |u - cx| dZ / Z <= 0.5 px), well inside the 16 px margin. NaN / Inf fail every test.
SES_HD bool reproj_certainly_outside(const float* c4, const CamF& cf, const CamD& cm) {
  const float sx = c4[0], sy = c4[1], sz = c4[2], r = c4[3];
  const float X = cf.P[0] * sx + cf.P[1] * sy + cf.P[2] * sz + cf.P[3];
  const float Y = cf.P[4] * sx + cf.P[5] * sy + cf.P[6] * sz + cf.P[7];
  const float Z = cf.P[8] * sx + cf.P[9] * sy + cf.P[10] * sz + cf.P[11];
  const float zr = Z - r;
  if (!(zr >= 0.5f) || !(ses_abs(sx) < 100.f) || !(ses_abs(sy) < 100.f) || !(ses_abs(sz) < 100.f)) return false;
  const float iz = 1.0f / Z;
  const float du = (cf.fx * X + (float)cm.Tx) * iz, dv = (cf.fy * Y + (float)cm.Ty) * iz;
  if (!(ses_abs(du) < 1e4f) || !(ses_abs(dv) < 1e4f)) return false;
  const float k = r / zr * iz * 1.001f;                      // r / (Z0 (Z0 - r)), rounded up
  const float bu = (cf.fx * (Z + ses_abs(X)) + ses_abs((float)cm.Tx)) * k + 16.f;
  const float bv = (cf.fy * (Z + ses_abs(Y)) + ses_abs((float)cm.Ty)) * k + 16.f;
  const float u = du + cf.cx, v = dv + cf.cy;
  return u < -bu || u > (float)cm.width + bu || v < -bv || v > (float)cm.height + bv;
}

// persons3d [n_p] (n_p <= h_max); out [C][h_max]; n_out [C]. ws.n_slots must equal the number of warps of the team.
template <class Team>
SES_HD void reproject_frame(Team& tm, const Tables& tb, int h_max, const ses3d_person_cov* persons3d, int n_p,
                            const ReprojWs& ws, ses3d_person2d* out, int32_t* n_out) {
  const int C = tb.n_cams;
  const int words = (int)(sizeof(ses3d_person2d) / 4);  // 107
  if (n_p > h_max) n_p = h_max;
  if (n_p < 0) n_p = 0;
  tm.pfor(C, [&](int c) { ws.nslot[c] = 0; });
  for (int p0 = 0; p0 < n_p; p0 += ws.s_cap) {
    const int np_b = (n_p - p0) < ws.s_cap ? (n_p - p0) : ws.s_cap;
    // one thread per (person, joint): Cholesky of the 3x3 covariance and the 7 sigma points (REP:62-75, 184-190)
    tm.pfor(np_b * NKP, [&](int e) {
      const int p = p0 + e / NKP, k = e % NKP;
      const ses3d_keypoint_cov& kp = persons3d[p].keypoints[tb.model.fusion_idx[k]];
      const bool present = kp.score > 0.0f;  // REP:181
      ws.sscore[e] = present ? kp.score : 0.f;
      if (!present) return;
      // lower Cholesky of [[c0 c1 c2][c1 c3 c4][c2 c4 c5]] as cov.llt().matrixL() evaluates it (REP:72, 184-187):
      // Eigen's unblocked LLT stops at the first non-positive pivot and leaves the rest of the lower triangle as it
      // is at that moment; matrixL() is read without checking info(). So a zero or indefinite covariance gives
      // finite sigma points (all equal to the mean for cov = 0), not NaN. A NaN pivot fails `x <= 0` and
      // propagates through sqrt like in Eigen.
      double l00 = kp.cov[0], l10 = kp.cov[1], l20 = kp.cov[2], l11 = kp.cov[3], l21 = kp.cov[4], l22 = kp.cov[5];
      do {
        if (l00 <= 0.0) break;
        l00 = sqrt(l00);
        l10 /= l00; l20 /= l00;
        double x = l11 - l10 * l10;
        if (x <= 0.0) break;
        l11 = x = sqrt(x);
        l21 -= l20 * l10;
        l21 /= x;
        x = l22 - (l20 * l20 + l21 * l21);
        if (x <= 0.0) break;
        l22 = sqrt(x);
      } while (false);
      const double sp = sqrt(3.0 + 0.5);  // sqrt(DIM + kappa) REP:63,68
      // samples: mean, mean - sp*L e_j (j=0..2), mean + sp*L e_j (REP:68-72)
      const double col[3][3] = {{l00, l10, l20}, {0.0, l11, l21}, {0.0, 0.0, l22}};
      double* S = ws.S + (size_t)e * 21;
      {  // single-precision centre and sigma-point radius for the "certainly outside" pre-test
        const double n0 = l00 * l00 + l10 * l10 + l20 * l20, n1 = l11 * l11 + l21 * l21, n2 = l22 * l22;
        const double nm = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
        float* c4 = ws.ctr + (size_t)e * 4;
        c4[0] = (float)kp.x; c4[1] = (float)kp.y; c4[2] = (float)kp.z;
        c4[3] = (float)(sp * sqrt(nm)) * 1.001f + 1e-6f;   // NaN entries disable the pre-test
      }
      S[0] = kp.x; S[1] = kp.y; S[2] = kp.z;
      for (int j = 0; j < 3; ++j) {
        S[(1 + j) * 3 + 0] = (col[j][0] * -sp) + kp.x; S[(1 + j) * 3 + 1] = (col[j][1] * -sp) + kp.y;
        S[(1 + j) * 3 + 2] = (col[j][2] * -sp) + kp.z;
        S[(4 + j) * 3 + 0] = (col[j][0] * sp) + kp.x; S[(4 + j) * 3 + 1] = (col[j][1] * sp) + kp.y;
        S[(4 + j) * 3 + 2] = (col[j][2] * sp) + kp.z;
      }
    });
    // one warp per camera from here on (REP:193-230); the barrier at the end of per_warp also protects S
    tm.per_warp(C, [&](auto& wt, int c) {
      const int slot = c % ws.n_slots;   // == the warp's index: items are dealt round-robin
      ses3d_person2d* stage = ws.stage + (size_t)slot * ws.s_cap;
      uint8_t* vflag = ws.vflag + (size_t)slot * ws.s_cap * NKP;
      uint16_t* list = ws.list + (size_t)slot * ws.s_cap * NKP;
      int* cnt = ws.cnt + slot * 2;
      const CamD& cm = tb.camd[c];
      wt.single([&] { cnt[0] = 0; });
      // cheap single-precision cull, then the exact projection of the survivors with the lanes full
      wt.pfor(np_b * NKP, [&](int e) {
        vflag[e] = 0;
        if (!(ws.sscore[e] > 0.0f)) return;
        if (reproj_certainly_outside(ws.ctr + (size_t)e * 4, tb.camf[c], cm)) return;
        list[team_append(cnt)] = (uint16_t)e;
      });
      const int n_surv = cnt[0];
      wt.sync();   // every lane has read the count before the leader resets it for the warp's next camera
      if (n_surv == 0) return;   // nobody of this batch is anywhere near this camera's image (the common case)
      wt.pfor(np_b * words, [&](int e) { reinterpret_cast<uint32_t*>(stage)[e] = 0u; });
      wt.pfor(n_surv, [&](int li) {
        const int e = list[li];
        const int pl = e / NKP, k = e % NKP;
        const float score = ws.sscore[e];
        const double wden = 2.0 * (3 + 0.5);
        const double w0 = 2 * 0.5 / wden, wi = 1.0 / wden;  // REP:65-66
        const double* S = ws.S + (size_t)e * 21;
        double u[7], v[7];
        for (int s = 0; s < 7; ++s) {
          const double sx = S[s * 3], sy = S[s * 3 + 1], sz = S[s * 3 + 2];
          const double X = cm.P[0] * sx + cm.P[1] * sy + cm.P[2] * sz + cm.P[3];
          const double Y = cm.P[4] * sx + cm.P[5] * sy + cm.P[6] * sz + cm.P[7];
          const double Z = cm.P[8] * sx + cm.P[9] * sy + cm.P[10] * sz + cm.P[11];
          u[s] = (cm.fx * X + cm.Tx) / Z + cm.cx;  // project3dToPixel (image_geometry)
          v[s] = (cm.fy * Y + cm.Ty) / Z + cm.cy;
        }
        double mu = 0, mv = 0;
        for (int s = 0; s < 7; ++s) { const double w = s == 0 ? w0 : wi; mu += u[s] * w; mv += v[s] * w; }
        double cxx = 0, cxy = 0, cyy = 0;
        for (int s = 0; s < 7; ++s) {
          const double w = s == 0 ? w0 : wi;
          const double du = u[s] - mu, dv = v[s] - mv;
          cxx += du * w * du; cxy += du * w * dv; cyy += dv * w * dv;
        }
        if (!(mu < 0 || mu > cm.width || mv < 0 || mv > cm.height)) {  // REP:207-208
          ses3d_keypoint2d& o = stage[pl].keypoints[k];
          o.x = static_cast<float>(mu); o.y = static_cast<float>(mv); o.score = score;
          o.cov[0] = static_cast<float>(cxx); o.cov[1] = static_cast<float>(cxy); o.cov[2] = static_cast<float>(cyy);
          vflag[e] = 1;
        }
      });
      // bbox + emitted flag per person (REP:150,161-162,218-230), one lane per person; emitted persons go out in
      // person order (push_back, REP:229)
      wt.pfor(np_b, [&](int pl) {
        ses3d_person2d& ps = stage[pl];
        float x0 = (float)cm.width, y0 = (float)cm.height, x1 = 0.f, y1 = 0.f;
        int n_valid = 0;
        for (int k = 0; k < NKP; ++k)
          if (vflag[pl * NKP + k]) {
            const float x = ps.keypoints[k].x, y = ps.keypoints[k].y;
            x0 = x < x0 ? x : x0; y0 = y < y0 ? y : y0; x1 = x > x1 ? x : x1; y1 = y > y1 ? y : y1;
            ++n_valid;
          }
        ps.score = 1.0f;  // REP:175
        ps.bbox[0] = x0; ps.bbox[1] = y0; ps.bbox[2] = x1; ps.bbox[3] = y1;
        list[pl] = n_valid > 0 ? 1 : 0;   // the work list is spent: reuse it for the emitted flags
      });
      wt.single([&] {
        int s = ws.nslot[c];
        for (int pl = 0; pl < np_b; ++pl) list[pl] = list[pl] ? (uint16_t)(s++) : (uint16_t)0xFFFF;
        ws.nslot[c] = s;
      });
      for (int pl = 0; pl < np_b; ++pl) {
        const int s = list[pl];
        if (s == 0xFFFF) continue;
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + (size_t)c * h_max + s);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(stage + pl);
        wt.pfor(words, [&](int w) { dst[w] = src[w]; });
      }
    });
  }
  tm.pfor(C, [&](int c) { n_out[c] = ws.nslot[c]; });
}

}  // namespace ses3d
