// tri_core.h — triangulation of one person hypothesis (kernel K3 "triangulate").
//
// Replaces the body of the per-hypothesis loop of triangulate_persons (S3D:681-975):
// per-joint view gathering (S3D:718-738), weighted DLT + reprojection error (S3D:746 ->
// 440-465, 425-438), the 3-view epipolar and >=4-view leave-one-out outlier rejection
// (S3D:748-838), score down-weighting (S3D:840-844), the unscented-transform covariance
// (S3D:846-847 -> 471-523), limb-length covariance inflation (S3D:861-883) and the
// root-distance / feet-height plausibility tests (S3D:923-973). An optional
// Levenberg-Marquardt refinement (not in the reference) sits behind
// params.lm_refine.
//
// B200 mapping: one CTA per (frame, hypothesis). The 17 joints x (4n+1) sigma-point solves
// of the covariance - ~98 % of the solves - are flattened into one index space and spread
// over all threads; each thread owns a complete 4x4 eigen-solve in registers. Each sigma
// point differs from the base system in one view only, so its normal matrix is the base
// Gram plus a rank-<=4 update (two rows removed, two added) instead of a rebuild.
#pragma once
#include "common.h"
#include "geom.h"
#include "team.h"

namespace ses3d {

template <class T>
struct ViewKp {  // one normalised keypoint of one observation
  T x, y, conf, cxx, cxy, cyy;
};

template <class T>
struct TriWs {
  uint8_t* obs_cam;   // [C]
  uint8_t* obs_det;   // [C]
  int* scal;          // [4]: n_obs, keep, total_samples
  ViewKp<T>* vw;      // [C][17]
  uint8_t* vlist;     // [17][C] observation indices used by joint k
  int* jn;            // [17] number of views (0 = joint not triangulated)
  int* jflag;         // [17] 1 = leave-one-out pending
  T* jX;              // [17][3]
  double* jerr;       // [17]
  float* jscore;      // [17]
  double* G0;         // [17][10] unweighted base Gram of the final view set
  int* soff;          // [18] sample offsets
  T* Y;               // [17*(4C+1)][3] transformed sigma points   (aliases the LOO buffers)
  T* looX;            // [17][C][3]
  double* looErr;     // [17][C]
  ses3d_keypoint_cov* kp;  // [21] the output skeleton
};

template <class T, class A>
SES_HD void tri_ws_layout(A& ar, int C, TriWs<T>* ws) {
  double* jerr = ar.template take<double>(NKP);
  double* G0 = ar.template take<double>(NKP * 10);
  ses3d_keypoint_cov* kp = ar.template take<ses3d_keypoint_cov>(NFUS);
  // union { Y ; looX + looErr }
  const size_t y_bytes = (size_t)NKP * (4 * C + 1) * 3 * sizeof(T);
  const size_t loo_bytes = (size_t)NKP * C * (8 + 3 * sizeof(T));
  const size_t u_bytes = (y_bytes > loo_bytes ? y_bytes : loo_bytes);
  double* u = ar.template take<double>((u_bytes + 7) / 8);
  ViewKp<T>* vw = ar.template take<ViewKp<T>>((size_t)C * NKP);
  T* jX = ar.template take<T>(NKP * 3);
  float* jscore = ar.template take<float>(NKP);
  int* jn = ar.template take<int>(NKP);
  int* jflag = ar.template take<int>(NKP);
  int* soff = ar.template take<int>(NKP + 1);
  int* scal = ar.template take<int>(4);
  uint8_t* vlist = ar.template take<uint8_t>((size_t)NKP * C);
  uint8_t* obs_cam = ar.template take<uint8_t>(C);
  uint8_t* obs_det = ar.template take<uint8_t>(C);
  if (ws) {
    ws->jerr = jerr; ws->G0 = G0; ws->kp = kp; ws->Y = reinterpret_cast<T*>(u);
    ws->looErr = u; ws->looX = reinterpret_cast<T*>(u + (size_t)NKP * C);
    ws->vw = vw; ws->jX = jX; ws->jscore = jscore; ws->jn = jn; ws->jflag = jflag; ws->soff = soff;
    ws->scal = scal; ws->vlist = vlist; ws->obs_cam = obs_cam; ws->obs_det = obs_det;
  }
}

template <class T>
inline size_t tri_ws_bytes(int C) {
  ArenaSizer s;
  tri_ws_layout<T>(s, C, nullptr);
  return (s.used + 15) / 16 * 16;
}

template <class T> struct CamSel;
template <> struct CamSel<float> {
  static SES_HD const float* P(const Tables& tb, int c) { return tb.camf[c].P; }
};
template <> struct CamSel<double> {
  static SES_HD const double* P(const Tables& tb, int c) { return tb.camd[c].P; }
};

// normalize_keypoints (S3D:312-333) of one raw keypoint, in T. conf = -1 when below threshold.
SES_HD void normalize_kp(const Tables& tb, int cam, const ses3d_keypoint2d& kp, ViewKp<float>& o) {
  const CamF& cm = tb.camf[cam];
  o.x = 0.f; o.y = 0.f; o.conf = -1.f; o.cxx = 0.f; o.cxy = 0.f; o.cyy = 0.f;
  if (kp.score >= tb.prm.triangulation_threshold) {
    o.x = (kp.x - cm.cx) / cm.fx;
    o.y = (kp.y - cm.cy) / cm.fy;
    o.conf = kp.score;
    o.cxx = kp.cov[0] / (cm.fx * cm.fx);
    o.cxy = kp.cov[1] / (cm.fx * cm.fy);
    o.cyy = kp.cov[2] / (cm.fy * cm.fy);
  }
}
SES_HD void normalize_kp(const Tables& tb, int cam, const ses3d_keypoint2d& kp, ViewKp<double>& o) {
  const CamD& cm = tb.camd[cam];
  o.x = 0.; o.y = 0.; o.conf = -1.; o.cxx = 0.; o.cxy = 0.; o.cyy = 0.;
  if (kp.score >= tb.prm.triangulation_threshold) {
    o.x = ((double)kp.x - cm.cx) / cm.fx;
    o.y = ((double)kp.y - cm.cy) / cm.fy;
    o.conf = (double)kp.score;
    o.cxx = (double)kp.cov[0] / (cm.fx * cm.fx);
    o.cxy = (double)kp.cov[1] / (cm.fx * cm.fy);
    o.cyy = (double)kp.cov[2] / (cm.fy * cm.fy);
  }
}

// Weighted DLT of joint k over the views in list[0..n) skipping index `skip` (-1 = none):
// triangulate(..., weight_by_conf=true, &err)  S3D:440-465 + calcReprojectionError S3D:425-438.
template <class T>
SES_HD void solve_weighted(const Tables& tb, const TriWs<T>& ws, int C, int k, const uint8_t* list, int n, int skip,
                           T X[3], double* err) {
  double G[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    if (i == skip) continue;
    const int o = list[i];
    const ViewKp<T>& v = ws.vw[o * NKP + k];
    const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
    T r[4];
    dlt_row<T>(P, 0, v.x, v.conf, true, r); gram_add<T>(G, r, 1.0);
    dlt_row<T>(P, 1, v.y, v.conf, true, r); gram_add<T>(G, r, 1.0);
  }
  T e[4];
  smallest_eigvec4<T>(G, e);
  X[0] = e[0] / e[3]; X[1] = e[1] / e[3]; X[2] = e[2] / e[3];
  if (err) {
    double avg = 0., norm = 0.;
    for (int i = 0; i < n; ++i) {
      if (i == skip) continue;
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T r = reproj_residual<T>(CamSel<T>::P(tb, ws.obs_cam[o]), X, v.x, v.y);
      avg += static_cast<double>(v.conf * r);
      norm += static_cast<double>(v.conf);
    }
    *err = avg / norm;
  }
}

// LM refinement of sum conf^2 * ||hnorm(P X~) - x||^2 (self-specified, not in the reference)
template <class T>
SES_HD void lm_refine_joint(const Tables& tb, const TriWs<T>& ws, int k, const uint8_t* list, int n, T X[3]) {
  auto cost_at = [&](const T* Y) {
    T f = 0;
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      const T a = P[0] * Y[0] + P[1] * Y[1] + P[2] * Y[2] + P[3];
      const T b = P[4] * Y[0] + P[5] * Y[1] + P[6] * Y[2] + P[7];
      const T c = P[8] * Y[0] + P[9] * Y[1] + P[10] * Y[2] + P[11];
      const T rx = v.conf * (a / c - v.x), ry = v.conf * (b / c - v.y);
      f += rx * rx + ry * ry;
    }
    return f;
  };
  T lambda = T(1e-3);
  T f0 = cost_at(X);
  for (int it = 0; it < tb.prm.lm_max_iters; ++it) {
    T H0 = 0, H1 = 0, H2 = 0, H3 = 0, H4 = 0, H5 = 0, g0 = 0, g1 = 0, g2 = 0;
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      const T a = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
      const T b = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
      const T c = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
      const T ic = T(1) / c, u = a * ic, vv = b * ic;
      const T rx = v.conf * (u - v.x), ry = v.conf * (vv - v.y);
      const T jx0 = v.conf * ic * (P[0] - u * P[8]), jx1 = v.conf * ic * (P[1] - u * P[9]),
              jx2 = v.conf * ic * (P[2] - u * P[10]);
      const T jy0 = v.conf * ic * (P[4] - vv * P[8]), jy1 = v.conf * ic * (P[5] - vv * P[9]),
              jy2 = v.conf * ic * (P[6] - vv * P[10]);
      H0 += jx0 * jx0 + jy0 * jy0; H1 += jx0 * jx1 + jy0 * jy1; H2 += jx0 * jx2 + jy0 * jy2;
      H3 += jx1 * jx1 + jy1 * jy1; H4 += jx1 * jx2 + jy1 * jy2; H5 += jx2 * jx2 + jy2 * jy2;
      g0 += jx0 * rx + jy0 * ry; g1 += jx1 * rx + jy1 * ry; g2 += jx2 * rx + jy2 * ry;
    }
    const T a00 = H0 * (T(1) + lambda), a11 = H3 * (T(1) + lambda), a22 = H5 * (T(1) + lambda);
    const T a01 = H1, a02 = H2, a12 = H4;
    const T c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const T det = a00 * c00 + a01 * c01 + a02 * c02;
    if (!(ses_abs(det) > T(0))) break;
    const T c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const T id = T(1) / det;
    const T d0 = -(c00 * g0 + c01 * g1 + c02 * g2) * id, d1 = -(c01 * g0 + c11 * g1 + c12 * g2) * id,
            d2 = -(c02 * g0 + c12 * g1 + c22 * g2) * id;
    const T Y[3] = {X[0] + d0, X[1] + d1, X[2] + d2};
    const T f1 = cost_at(Y);
    if (f1 < f0) {
      X[0] = Y[0]; X[1] = Y[1]; X[2] = Y[2];
      f0 = f1;
      lambda *= T(0.1);
      if (d0 * d0 + d1 * d1 + d2 * d2 < T(1e-14)) break;
    } else {
      lambda *= T(10);
    }
  }
}

SES_HD double joint_dist(const ses3d_keypoint_cov& a, const ses3d_keypoint_cov& b) {  // calcJointDist S3D:467-469
  return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
SES_HD void add_cov(ses3d_keypoint_cov& kp, double sigma) {  // addToKeypointCovariance S3D:273-277
  kp.cov[0] += sigma * sigma; kp.cov[3] += sigma * sigma; kp.cov[5] += sigma * sigma;
}
SES_HD void zero_kp(ses3d_keypoint_cov& kp) {
  kp.x = 0; kp.y = 0; kp.z = 0; kp.score = 0; kp.pad_ = 0;
  for (int i = 0; i < 6; ++i) kp.cov[i] = 0;
}

// One hypothesis. hyp_det_row [C]: detection slot per camera (-1 = not observed).
// Writes *out (the PersonCov record) and *keep (1 if the person passes S3D:968).
template <class T, class Team>
SES_HD void triangulate_hypothesis(Team& tm, const Tables& tb, int p_max, const ses3d_person2d* persons,
                                   const int8_t* hyp_det_row, const TriWs<T>& ws, ses3d_person_cov* out,
                                   int32_t* keep) {
  const int C = tb.n_cams;
  const double max_reproj = tb.prm.reproj_error_max_acceptable;
  const float thr = tb.prm.triangulation_threshold;

  tm.single([&] {
    int n = 0;
    for (int c = 0; c < C; ++c)
      if (hyp_det_row[c] >= 0) { ws.obs_cam[n] = (uint8_t)c; ws.obs_det[n] = (uint8_t)hyp_det_row[c]; ++n; }
    ws.scal[0] = n;
  });
  const int n_obs = ws.scal[0];
  if (n_obs < 2) {  // S3D:684: hypotheses with a single observation are not triangulated
    tm.single([&] { *keep = 0; });
    return;
  }

  // normalised keypoints + covariances of the hypothesis' observations
  tm.pfor(n_obs * NKP, [&](int i) {
    const int o = i / NKP, k = i % NKP;
    const int cam = ws.obs_cam[o];
    normalize_kp(tb, cam, persons[cam * p_max + ws.obs_det[o]].keypoints[k], ws.vw[i]);
  });
  tm.pfor(NFUS, [&](int s) { zero_kp(ws.kp[s]); });

  // per joint: gather views, weighted DLT, 3-view epipolar rejection (S3D:718-792)
  tm.pfor(NKP, [&](int k) {
    uint8_t* list = ws.vlist + k * C;
    int n = 0;
    float avg_score = 0;
    for (int o = 0; o < n_obs; ++o) {
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      if ((float)v.conf >= thr) { list[n++] = (uint8_t)o; avg_score += (float)v.conf; }
    }
    ws.jflag[k] = 0;
    if (n < 2) { ws.jn[k] = 0; return; }
    avg_score /= n;
    T X[3];
    double err;
    solve_weighted<T>(tb, ws, C, k, list, n, -1, X, &err);
    if (err > max_reproj && n == 3) {
      int best = -1;
      float best_dist = static_cast<float>(err * err);
      for (int i = 0; i < 3; ++i) {
        const int oa = list[i == 0 ? 1 : 0], ob = list[i == 2 ? 1 : 2];
        const ViewKp<T>& a = ws.vw[oa * NKP + k];
        const ViewKp<T>& b = ws.vw[ob * NKP + k];
        const float* F = tb.F + (size_t)fundamental_idx(tb, ws.obs_cam[oa], ws.obs_cam[ob]) * 9;
        const float x1 = (float)a.x, y1 = (float)a.y, x2 = (float)b.x, y2 = (float)b.y;
        const float l1x = sum3(F[0] * x1, F[1] * y1, F[2] * 1.0f);
        const float l1y = sum3(F[3] * x1, F[4] * y1, F[5] * 1.0f);
        const float l1z = sum3(F[6] * x1, F[7] * y1, F[8] * 1.0f);
        const float l2x = sum3(F[0] * x2, F[3] * y2, F[6] * 1.0f);
        const float l2y = sum3(F[1] * x2, F[4] * y2, F[7] * 1.0f);
        const float l2z = sum3(F[2] * x2, F[5] * y2, F[8] * 1.0f);
        const float n1 = sum3(x2 * l1x, y2 * l1y, 1.0f * l1z);
        const float n2 = sum3(x1 * l2x, y1 * l2y, 1.0f * l2z);
        const float d = n1 * n1 / (l1x * l1x + l1y * l1y) + n2 * n2 / (l2x * l2x + l2y * l2y);
        if (d < best_dist) { best_dist = d; best = i; }
      }
      if (best != -1) {
        for (int i = best; i < 2; ++i) list[i] = list[i + 1];
        n = 2;
        solve_weighted<T>(tb, ws, C, k, list, n, -1, X, &err);
        avg_score = ((float)ws.vw[list[0] * NKP + k].conf + (float)ws.vw[list[1] * NKP + k].conf) / 2.0f;
      }
    } else if (err > max_reproj && n >= 4) {
      ws.jflag[k] = 1;
    }
    ws.jn[k] = n;
    ws.jX[k * 3] = X[0]; ws.jX[k * 3 + 1] = X[1]; ws.jX[k * 3 + 2] = X[2];
    ws.jerr[k] = err;
    ws.jscore[k] = avg_score;
  });

  // leave-one-out solves for joints with a large error, one (joint, left-out view) per thread (S3D:799-810)
  tm.pfor(NKP * C, [&](int i) {
    const int k = i / C, v = i % C;
    if (!ws.jflag[k] || v >= ws.jn[k]) return;
    T X[3];
    double e;
    solve_weighted<T>(tb, ws, C, k, ws.vlist + k * C, ws.jn[k], v, X, &e);
    ws.looX[i * 3] = X[0]; ws.looX[i * 3 + 1] = X[1]; ws.looX[i * 3 + 2] = X[2];
    ws.looErr[i] = e;
  });

  // select (S3D:811-837), optional LM, down-weight (S3D:840-844), base Gram for the sigma points
  tm.pfor(NKP, [&](int k) {
    int n = ws.jn[k];
    if (n < 2) return;
    uint8_t* list = ws.vlist + k * C;
    double err = ws.jerr[k];
    float avg_score = ws.jscore[k];
    T X[3] = {ws.jX[k * 3], ws.jX[k * 3 + 1], ws.jX[k * 3 + 2]};
    if (ws.jflag[k]) {
      double best_err = err;
      int best = -1;
      float best_score = avg_score;
      for (int i = 0; i < n; ++i) {
        const double e_sub = ws.looErr[k * C + i];
        if (best_err > e_sub && e_sub < 0.9 * err) {
          best_err = e_sub; best = i;
          float tmp = 0.f;
          for (int j = 0; j < n; ++j)
            if (j != i) tmp += (float)ws.vw[list[j] * NKP + k].conf;
          best_score = tmp / (float)(n - 1);
        }
      }
      if (best != -1) {
        X[0] = ws.looX[(k * C + best) * 3]; X[1] = ws.looX[(k * C + best) * 3 + 1]; X[2] = ws.looX[(k * C + best) * 3 + 2];
        for (int i = best; i < n - 1; ++i) list[i] = list[i + 1];
        --n;
        err = best_err;
        avg_score = best_score;
      }
    }
    if (tb.prm.lm_refine) {
      lm_refine_joint<T>(tb, ws, k, list, n, X);
      double avg = 0., norm = 0.;
      for (int i = 0; i < n; ++i) {
        const ViewKp<T>& v = ws.vw[list[i] * NKP + k];
        const T r = reproj_residual<T>(CamSel<T>::P(tb, ws.obs_cam[list[i]]), X, v.x, v.y);
        avg += static_cast<double>(v.conf * r);
        norm += static_cast<double>(v.conf);
      }
      err = avg / norm;
    }
    if (err > max_reproj) avg_score = (float)((double)avg_score * (max_reproj / err));
    double G[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      T r[4];
      dlt_row<T>(P, 0, v.x, T(1), false, r); gram_add<T>(G, r, 1.0);
      dlt_row<T>(P, 1, v.y, T(1), false, r); gram_add<T>(G, r, 1.0);
    }
    for (int i = 0; i < 10; ++i) ws.G0[k * 10 + i] = G[i];
    ws.jn[k] = n;
    ws.jX[k * 3] = X[0]; ws.jX[k * 3 + 1] = X[1]; ws.jX[k * 3 + 2] = X[2];
    ws.jerr[k] = err;
    ws.jscore[k] = avg_score;
  });

  tm.single([&] {
    int off = 0;
    for (int k = 0; k < NKP; ++k) { ws.soff[k] = off; off += ws.jn[k] >= 2 ? 4 * ws.jn[k] + 1 : 0; }
    ws.soff[NKP] = off;
  });
  const int n_samples_total = ws.soff[NKP];

  // unscented sigma points (S3D:471-506): all joints x (4n+1) samples in one index space
  tm.pfor(n_samples_total, [&](int i) {
    int k = 0;
    while (ws.soff[k + 1] <= i) ++k;
    const int s = i - ws.soff[k];
    const int n = ws.jn[k];
    double G[10];
    for (int j = 0; j < 10; ++j) G[j] = ws.G0[k * 10 + j];
    if (s > 0) {
      const int vi = (s - 1) >> 2, m = (s - 1) & 3;
      const int o = ws.vlist[k * C + vi];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      const T b = ses_sqrt(T(2 * n) + T(0.5));
      const T l11 = ses_sqrt(v.cxx);         // mod_samples S3D:471-487: 2x2 Cholesky
      const T l21 = v.cxy / l11;
      const T l22 = ses_sqrt(v.cyy - l21 * l21);
      T nx = v.x, ny = v.y;
      if (m == 0) { nx = v.x - l11 * b; ny = v.y - l21 * b; }
      else if (m == 1) { ny = v.y - l22 * b; }
      else if (m == 2) { nx = v.x + l11 * b; ny = v.y + l21 * b; }
      else { ny = v.y + l22 * b; }
      T r[4];
      if ((m & 1) == 0) {
        dlt_row<T>(P, 0, v.x, T(1), false, r); gram_add<T>(G, r, -1.0);
        dlt_row<T>(P, 0, nx, T(1), false, r); gram_add<T>(G, r, 1.0);
      }
      dlt_row<T>(P, 1, v.y, T(1), false, r); gram_add<T>(G, r, -1.0);
      dlt_row<T>(P, 1, ny, T(1), false, r); gram_add<T>(G, r, 1.0);
    }
    T e[4];
    smallest_eigvec4<T>(G, e);
    ws.Y[i * 3] = e[0] / e[3]; ws.Y[i * 3 + 1] = e[1] / e[3]; ws.Y[i * 3 + 2] = e[2] / e[3];
  });

  // covariance about the weighted-DLT point (S3D:521-522) and the output keypoint (S3D:849-857)
  tm.pfor(NKP, [&](int k) {
    const int n = ws.jn[k];
    if (n < 2) return;
    const T wden = T(2) * (T(2 * n) + T(0.5));
    const T w0 = (T(2) * T(0.5)) / wden, wi = T(1) / wden;
    const T m0 = ws.jX[k * 3], m1 = ws.jX[k * 3 + 1], m2 = ws.jX[k * 3 + 2];
    T c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
    const int base = ws.soff[k], ns = 4 * n + 1;
    for (int s = 0; s < ns; ++s) {
      const T w = s == 0 ? w0 : wi;
      const T d0 = ws.Y[(base + s) * 3] - m0, d1 = ws.Y[(base + s) * 3 + 1] - m1, d2 = ws.Y[(base + s) * 3 + 2] - m2;
      c00 += (d0 * w) * d0; c01 += (d0 * w) * d1; c02 += (d0 * w) * d2;
      c11 += (d1 * w) * d1; c12 += (d1 * w) * d2; c22 += (d2 * w) * d2;
    }
    ses3d_keypoint_cov& o = ws.kp[tb.model.fusion_idx[k]];
    o.x = (double)m0; o.y = (double)m1; o.z = (double)m2;
    o.score = ws.jscore[k];
    o.cov[0] = (double)c00; o.cov[1] = (double)c01; o.cov[2] = (double)c02;
    o.cov[3] = (double)c11; o.cov[4] = (double)c12; o.cov[5] = (double)c22;
  });

  // skeleton plausibility (S3D:861-973), a few hundred flops: team leader
  tm.single([&] {
    const SkeletonModel& M = tb.model;
    int num_valid = 0;
    for (int k = 0; k < NKP; ++k) num_valid += ws.jn[k] >= 2 ? 1 : 0;
    for (int k = 0; k < NKP; ++k) {
      ses3d_keypoint_cov& kp = ws.kp[M.fusion_idx[k]];
      if (kp.score <= 0) continue;
      const int parent = M.parent[k];
      if (parent >= 0) {
        const ses3d_keypoint_cov& pk = ws.kp[M.fusion_idx[parent]];
        if (pk.score > 0 && M.limb_len[k] > 0) {
          add_cov(kp, tb.prm.limb_cov_offset_sigma * (joint_dist(kp, pk) - M.limb_len[k]) / M.limb_sigma[k]);
        } else if (tb.prm.pose_method == SES3D_POSE_SIMPLE && k == 6 /*RShoulder S3D:83*/) {
          ses3d_keypoint_cov& ls = ws.kp[M.fusion_idx[5 /*LShoulder S3D:86*/]];
          if (ls.score > 0) {
            const double d = joint_dist(kp, ls);
            add_cov(kp, tb.prm.limb_cov_offset_sigma * (d - 0.35) / 0.15);  // shoulderDist, shoulderSigma S3D:103
            add_cov(ls, tb.prm.limb_cov_offset_sigma * (d - 0.35) / 0.15);
          }
        }
      }
    }
    ses3d_keypoint_cov root;
    zero_kp(root);
    const ses3d_keypoint_cov* K = ws.kp;
    if (K[SES3D_FBP_MIDHIP].score > 0) root = K[SES3D_FBP_MIDHIP];
    else if (K[SES3D_FBP_LHIP].score > 0 && K[SES3D_FBP_RHIP].score > 0) {
      root.x = (K[SES3D_FBP_LHIP].x + K[SES3D_FBP_RHIP].x) / 2.;
      root.y = (K[SES3D_FBP_LHIP].y + K[SES3D_FBP_RHIP].y) / 2.;
      root.z = (K[SES3D_FBP_LHIP].z + K[SES3D_FBP_RHIP].z) / 2.;
      root.score = (K[SES3D_FBP_LHIP].score + K[SES3D_FBP_RHIP].score) / 2.f;
    }
    if (root.score > 0) {
      for (int s = 0; s < NFUS; ++s) {
        ses3d_keypoint_cov& kp = ws.kp[s];
        if (kp.score > 0) {
          if (joint_dist(root, kp) > tb.prm.max_joint_dist_to_root) { zero_kp(kp); --num_valid; }
        } else {
          zero_kp(kp);
          --num_valid;
        }
      }
    }
    double feet = 0.0;
    if (K[SES3D_FBP_LANKLE].score > 0 && K[SES3D_FBP_RANKLE].score > 0)
      feet = (K[SES3D_FBP_LANKLE].z + K[SES3D_FBP_RANKLE].z) / 2.0;
    else if (K[SES3D_FBP_LANKLE].score > 0) feet = K[SES3D_FBP_LANKLE].z;
    else if (K[SES3D_FBP_RANKLE].score > 0) feet = K[SES3D_FBP_RANKLE].z;
    if (fabs(feet) > 0.50) num_valid = 0;
    ws.scal[1] = num_valid > tb.prm.min_num_valid_keypoints ? 1 : 0;
  });

  // write the PersonCov record (header + 21 keypoints + zero bbox)
  const int kept = ws.scal[1];
  if (kept) {
    uint64_t* dst = reinterpret_cast<uint64_t*>(out);
    const uint64_t* src = reinterpret_cast<const uint64_t*>(ws.kp);
    tm.pfor((int)(sizeof(ses3d_person_cov) / 8), [&](int i) {
      uint64_t v = 0;  // id = 0, score = 0, bbox_center / bbox_size = 0 (never set by the reference)
      if (i >= 1 && i < 1 + NFUS * 10) v = src[i - 1];
      dst[i] = v;
    });
  }
  tm.single([&] { *keep = kept; });
}

}  // namespace ses3d
