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
// B200 mapping: one warp per (frame, hypothesis), several warps per CTA, no CTA-wide barrier.
// ~95 % of the solves are the 4n+1 sigma-point triangulations of the covariance; all joints'
// sigma points are flattened into one index space so the 32 lanes stay full, and each lane
// owns a complete 4x4 eigen-solve in registers. A sigma point differs from the base system in
// one view only, so its normal matrix is written in the eigenbasis of the base system
// (diag(lambda) + a rank-<=4 update): Jacobi then starts almost diagonal (warm start) and
// converges in 1-3 sweeps instead of 5-7.
#pragma once
#include "common.h"
#include "geom.h"
#include "team.h"

#ifndef SES_COLD_PATHS
#define SES_COLD_PATHS 0   // 1: rare paths as out-of-line functions - measured +0.05 ms / 16384 frames on B200 (the call
                           // ABI and the lost cross-call scheduling cost more than the instruction fetch saved); A/B
                           // switch for scripts/build_variants.py
#endif
#if SES_COLD_PATHS
#define SES_COLD_FN SES_HDN
#else
#define SES_COLD_FN SES_HD
#endif

namespace ses3d {

template <class T>
struct ViewKp {  // one normalised keypoint of one observation (conf = -1: below threshold)
  T x, y, conf;
};

constexpr int Y_CHUNK = 256;  // sigma points solved per pass (bounds the shared-memory staging)
constexpr int TRI_PHASES = 6;  // team.phase() calls on every path through triangulate_hypothesis
constexpr int TRI_BUCKETS = 8; // K3 work lists by observation count: 2, 3, ..., 8, >= 9

template <class T>
struct TriWs {
  uint8_t* obs_cam;   // [C]
  uint8_t* obs_det;   // [C]
  int* scal;          // [4]: n_obs, keep
  ViewKp<T>* vw;      // [C][17]
  uint8_t* vlist;     // [17][C] observation indices used by joint k
  int* jn;            // [17] number of views (0 = joint not triangulated)
  int* jflag;         // [17] 1 = leave-one-out pending
  T* jX;              // [17][3]
  double* jerr;       // [17]
  float* jscore;      // [17]
  T* defl;            // [17][14] deflated base system: v[4], a0, b0[3], C0[6]
  T* cov;             // [17][6]  running covariance sums
  int* soff;          // [18] sample offsets (sample 0 of each joint is the base solve, already known)
  uint8_t* kmap;      // [17*4*C] sample index -> joint
  T* Y;               // [Y_CHUNK][3] transformed sigma points of the current pass  (aliases the LOO scratch)
  T* looX;            // [loo_cap][3]
  double* looErr;     // [loo_cap]
  int loo_cap;
  ses3d_keypoint_cov* kp;  // [21] the output skeleton
  float* far_scratch;      // [team size][FAR_COV_STRIDE] private solve copies for exact_far_covariance (not carved
                           // from the arena: global memory on the GPU; nullptr disables the exact covariance)
};

template <class T, class A>
SES_HD void tri_ws_layout(A& ar, int C, TriWs<T>* ws) {
  double* jerr = ar.template take<double>(NKP);
  // union { Y[Y_CHUNK][3] ; looErr[loo_cap] + looX[loo_cap][3] ; kp[21] (written after the last sigma-point pass) }
  size_t u_bytes = (size_t)Y_CHUNK * 3 * sizeof(T);
  if (u_bytes < NFUS * sizeof(ses3d_keypoint_cov)) u_bytes = NFUS * sizeof(ses3d_keypoint_cov);
  const int loo_cap = (int)(u_bytes / (8 + 3 * sizeof(T)));
  double* u = ar.template take<double>((u_bytes + 7) / 8);
  ViewKp<T>* vw = ar.template take<ViewKp<T>>((size_t)C * NKP);
  T* jX = ar.template take<T>(NKP * 3);
  T* defl = ar.template take<T>(NKP * 14);
  T* cov = ar.template take<T>(NKP * 6);
  float* jscore = ar.template take<float>(NKP);
  int* jn = ar.template take<int>(NKP);
  int* jflag = ar.template take<int>(NKP);
  int* soff = ar.template take<int>(NKP + 1);
  int* scal = ar.template take<int>(4);
  uint8_t* kmap = reinterpret_cast<uint8_t*>(ar.template take<uint32_t>((size_t)NKP * C));   // filled word-wise
  uint8_t* vlist = ar.template take<uint8_t>((size_t)NKP * C);
  uint8_t* obs_cam = ar.template take<uint8_t>(C);
  uint8_t* obs_det = ar.template take<uint8_t>(C);
  if (ws) {
    ws->far_scratch = nullptr;
    ws->jerr = jerr; ws->kp = reinterpret_cast<ses3d_keypoint_cov*>(u); ws->Y = reinterpret_cast<T*>(u);
    ws->looErr = u; ws->looX = reinterpret_cast<T*>(u + loo_cap); ws->loo_cap = loo_cap;
    ws->vw = vw; ws->jX = jX; ws->defl = defl; ws->cov = cov; ws->jscore = jscore; ws->jn = jn;
    ws->jflag = jflag; ws->soff = soff; ws->scal = scal; ws->vlist = vlist; ws->kmap = kmap; ws->obs_cam = obs_cam; ws->obs_det = obs_det;
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
  o.x = 0.f; o.y = 0.f; o.conf = -1.f;
  if (kp.score >= tb.prm.triangulation_threshold) {
    o.x = (kp.x - cm.cx) / cm.fx;
    o.y = (kp.y - cm.cy) / cm.fy;
    o.conf = kp.score;
  }
}
SES_HD void normalize_kp(const Tables& tb, int cam, const ses3d_keypoint2d& kp, ViewKp<double>& o) {
  const CamD& cm = tb.camd[cam];
  o.x = 0.; o.y = 0.; o.conf = -1.;
  if (kp.score >= tb.prm.triangulation_threshold) {
    o.x = ((double)kp.x - cm.cx) / cm.fx;
    o.y = ((double)kp.y - cm.cy) / cm.fy;
    o.conf = (double)kp.score;
  }
}
// normalised 2x2 covariance (S3D:324-327) and its Cholesky factor (mod_samples S3D:473-475)
SES_HD void cholesky_cov(const Tables& tb, int cam, const ses3d_keypoint2d& kp, float& l11, float& l21, float& l22) {
  const CamF& cm = tb.camf[cam];
#if defined(__CUDA_ARCH__)
  // covariance-only path (tolerance-checked): reciprocal scaling and SFU rsqrt instead of 4 divides + 2 sqrt;
  // a zero variance still yields NaN like the reference's 0/0 (S3D:473-475)
  const float cxx = kp.cov[0] * cm.inv_fx2, cxy = kp.cov[1] * cm.inv_fxfy, cyy = kp.cov[2] * cm.inv_fy2;
  const float r11 = rsqrtf(cxx);
  l11 = cxx * r11;
  l21 = cxy * r11;
  const float t = cyy - l21 * l21;
  l22 = t * rsqrtf(t);
#else
  const float cxx = kp.cov[0] / (cm.fx * cm.fx), cxy = kp.cov[1] / (cm.fx * cm.fy), cyy = kp.cov[2] / (cm.fy * cm.fy);
  l11 = ses_sqrt(cxx);
  l21 = cxy / l11;
  l22 = ses_sqrt(cyy - l21 * l21);
#endif
}
SES_HD void cholesky_cov(const Tables& tb, int cam, const ses3d_keypoint2d& kp, double& l11, double& l21, double& l22) {
  const CamD& cm = tb.camd[cam];
#if defined(__CUDA_ARCH__)
  // covariance-only path (tolerance-checked): refined reciprocals instead of 4 IEEE divides + 2 square roots;
  // a zero variance still yields NaN like the reference's 0/0 (S3D:473-475)
  const double ifx = ses_rcp(cm.fx), ify = ses_rcp(cm.fy);
  const double cxx = (double)kp.cov[0] * (ifx * ifx), cxy = (double)kp.cov[1] * (ifx * ify), cyy = (double)kp.cov[2] * (ify * ify);
  const double r11 = ses_rsqrt(cxx);
  l11 = cxx * r11;
  l21 = cxy * r11;
  const double t = cyy - l21 * l21;
  l22 = t * ses_rsqrt(t);
#else
  const double cxx = (double)kp.cov[0] / (cm.fx * cm.fx), cxy = (double)kp.cov[1] / (cm.fx * cm.fy),
               cyy = (double)kp.cov[2] / (cm.fy * cm.fy);
  l11 = sqrt(cxx);
  l21 = cxy / l11;
  l22 = sqrt(cyy - l21 * l21);
#endif
}

// confidence-weighted mean reprojection error, calcReprojectionError S3D:425-438
template <class T>
SES_HD double reproj_error(const Tables& tb, const TriWs<T>& ws, int k, const uint8_t* list, int n, int skip,
                           const T X[3]) {
  double avg = 0., norm = 0.;
  for (int i = 0; i < n; ++i) {
    if (i == skip) continue;
    const int o = list[i];
    const ViewKp<T>& v = ws.vw[o * NKP + k];
    const T r = reproj_residual<T>(CamSel<T>::P(tb, ws.obs_cam[o]), X, v.x, v.y);
    avg += static_cast<double>(v.conf * r);
    norm += static_cast<double>(v.conf);
  }
  return avg / norm;
}

// Weighted DLT of joint k over the views in list[0..n) skipping index `skip` (-1 = none):
// triangulate(..., weight_by_conf=true, &err)  S3D:440-465.
template <class T>
SES_HD void solve_weighted(const Tables& tb, const TriWs<T>& ws, int k, const uint8_t* list, int n, int skip, T X[3],
                           double* err) {
  double G[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    if (i == skip) continue;
    const int o = list[i];
    const ViewKp<T>& v = ws.vw[o * NKP + k];
    const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
    T r[4];
    dlt_row_fast<T>(P, 0, v.x, r);
    r[0] *= v.conf; r[1] *= v.conf; r[2] *= v.conf; r[3] *= v.conf;   // weight_by_conf, S3D:450-453
    gram_add<T>(G, r, 1.0);
    dlt_row_fast<T>(P, 1, v.y, r);
    r[0] *= v.conf; r[1] *= v.conf; r[2] *= v.conf; r[3] *= v.conf;
    gram_add<T>(G, r, 1.0);
  }
  T e[4];
  smallest_eigvec4_fast<T>(G, e);
  const T ie3 = ses_rcp(e[3]);   // hnormalized(), S3D:459
  X[0] = e[0] * ie3; X[1] = e[1] * ie3; X[2] = e[2] * ie3;
  *err = reproj_error<T>(tb, ws, k, list, n, skip, X);
}

// COLD CALLS. The rarely taken paths (Jacobi fallback, 3-view re-solve, leave-one-out, LM, exact re-solves) can be
// built as out-of-line functions (SES_COLD_PATHS / SES_COLD_JACOBI = 1) so that their code does not sit inside the hot
// instruction stream (the kernel is fetch-bound, see kernels_tri.cu). An object whose address is handed to such a
// function becomes addressable and is demoted from registers to local memory on EVERY path, so call sites pass copies
// (of the workspace struct, of small arrays). Measured on B200 (profiles/r02_tri_build_variants.txt): inlined 1.427 ms,
// out of line 1.484 ms per 16 384 frames - the default is inlined.
//
// The same solve for the rare call sites (3-view re-solve, leave-one-out): one shared out-of-line copy.
template <class T>
struct ColdSolve { T x, y, z; double err; };
template <class T>
SES_COLD_FN ColdSolve<T> solve_weighted_cold(const Tables& tb, const TriWs<T>& ws, int k, const uint8_t* list, int n, int skip) {
  T X[3];
  ColdSolve<T> r;
  solve_weighted<T>(tb, ws, k, list, n, skip, X, &r.err);
  r.x = X[0]; r.y = X[1]; r.z = X[2];
  return r;   // by value: the caller's X / err stay in registers
}

// ---- far points: exact re-solve in the oracle's operation order ----------------------------------------------
// The DLT returns a homogeneous vector v and the joint is X = v_xyz / v_w (hnormalized, S3D:459). For a point at
// distance |X| from the rig origin v_w ~ 1/|X|, so a rounding error dv of the unit vector moves X by ~ dv |X|^2: two
// nearly parallel rays that "meet" hundreds of metres away (mismatched views, gross outliers) land wherever the last
// bits of the particular float SVD put them. Measured (scripts/soak_hostsim.py): inside 15 m the Gram / inverse
// iteration path is within 5e-5 m of the oracle, beyond 50 m it can be 1e-2 m off - and so is any other float SVD
// (the oracle's two SVD variants differ by the same amount there). Such joints are therefore re-solved exactly as
// the reference's data flow is restated in the oracle: rows of A built without FMA contraction, one-sided Jacobi
// on the 2n x 4 matrix with sequential column sums, same rotation formulas, same column selection - bit for bit,
// so they carry no tolerance at all. A holds 2n x 4 floats in the warp's sigma-point staging area (idle at this point).
constexpr float FAR_POINT_R2 = 20.0f * 20.0f;

// One DLT solve in the oracle's arithmetic. build_row(r, row) fills row r of A (already normalised / weighted, built
// with the non-contracting x-ops); B [rows][4] and W [4][4] live in shared memory; every thread of the team evaluates
// the same column sums sequentially (read-only), so all threads take the same decisions and end with the same X.
template <class Team, class RowFn>
SES_HD void exact_dlt(Team& tm, float* B, int rows, RowFn&& build_row, float X[3]) {
  float* W = B + 4 * rows;
  tm.pfor(rows, [&](int r) {
    float row[4];
    build_row(r, row);
    for (int c = 0; c < 4; ++c) B[r * 4 + c] = row[c];
  });
  tm.pfor(16, [&](int i) { W[i] = (i >> 2) == (i & 3) ? 1.f : 0.f; });
  const float eps = 1.1920929e-7f;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        float alpha = 0.f, beta = 0.f, gamma = 0.f;
        for (int r = 0; r < rows; ++r) {
          const float bp = B[r * 4 + p], bq = B[r * 4 + q];
          alpha = xadd(alpha, xmul(bp, bp));
          beta = xadd(beta, xmul(bq, bq));
          gamma = xadd(gamma, xmul(bp, bq));
        }
        tm.sync();   // all reads of B done before anybody rotates its rows
        if (ses_abs(gamma) <= xmul(eps, xsqrt(xmul(alpha, beta))) || gamma == 0.f) continue;
        rotated = true;
        const float zeta = xdiv(xsub(beta, alpha), xmul(2.f, gamma));
        const float t = xdiv(zeta >= 0.f ? 1.f : -1.f, xadd(ses_abs(zeta), xsqrt(xadd(1.f, xmul(zeta, zeta)))));
        const float c = xdiv(1.f, xsqrt(xadd(1.f, xmul(t, t))));
        const float sn = xmul(c, t);
        tm.pfor(rows + 4, [&](int r) {   // rows of A and of W are rotated alike
          float* row = B + r * 4;
          const float bp = row[p], bq = row[q];
          row[p] = xsub(xmul(c, bp), xmul(sn, bq));
          row[q] = xadd(xmul(sn, bp), xmul(c, bq));
        });
      }
    if (!rotated) break;
  }
  int best = 0;
  float best_s = FLT_MAX;
  for (int c = 0; c < 4; ++c) {
    float sq = 0.f;
    for (int r = 0; r < rows; ++r) sq = xadd(sq, xmul(B[r * 4 + c], B[r * 4 + c]));
    if (sq < best_s) { best_s = sq; best = c; }
  }
  const float w = W[12 + best];
  X[0] = xdiv(W[best], w); X[1] = xdiv(W[4 + best], w); X[2] = xdiv(W[8 + best], w);
  tm.sync();   // B may be rebuilt by the next solve
}

// row r of A for view list[r / 2] with the keypoint at (x, y): triangulate(), S3D:444-454
SES_HD void exact_row(const float* P, int half, float m, float conf, bool weighted, float row[4]) {
  for (int c = 0; c < 4; ++c) row[c] = xsub(xmul(m, P[8 + c]), P[half * 4 + c]);
  const float z = xadd(xadd(xmul(row[0], row[0]), xmul(row[1], row[1])), xadd(xmul(row[2], row[2]), xmul(row[3], row[3])));
  if (z > 0.f) {
    const float nrm = xsqrt(z);
    for (int c = 0; c < 4; ++c) row[c] = xdiv(row[c], nrm);
  }
  if (weighted)
    for (int c = 0; c < 4; ++c) row[c] = xmul(row[c], conf);
}

// calcReprojectionError (S3D:425-438) in the oracle's arithmetic (no contraction, IEEE division / square root)
SES_HD double exact_reproj_error(const Tables& tb, const TriWs<float>& ws, int k, const uint8_t* list, int n,
                                 const float X[3]) {
  double avg = 0., norm = 0.;
  for (int i = 0; i < n; ++i) {
    const int o = list[i];
    const ViewKp<float>& v = ws.vw[o * NKP + k];
    const float* P = tb.camf[ws.obs_cam[o]].P;
    const float a = xadd(xadd(xmul(P[0], X[0]), xmul(P[1], X[1])), xadd(xmul(P[2], X[2]), xmul(P[3], 1.f)));
    const float b = xadd(xadd(xmul(P[4], X[0]), xmul(P[5], X[1])), xadd(xmul(P[6], X[2]), xmul(P[7], 1.f)));
    const float cc = xadd(xadd(xmul(P[8], X[0]), xmul(P[9], X[1])), xadd(xmul(P[10], X[2]), xmul(P[11], 1.f)));
    const float dx = xsub(xdiv(a, cc), v.x), dy = xsub(xdiv(b, cc), v.y);
    const float e = xsqrt(xadd(xmul(dx, dx), xmul(dy, dy)));
    avg += static_cast<double>(xmul(v.conf, e));
    norm += static_cast<double>(v.conf);
  }
  return avg / norm;
}

template <class Team>
SES_COLD_FN void exact_weighted_resolve(Team& tm, const Tables& tb, const TriWs<float>& ws, int k, const uint8_t* list, int n) {
  float X[3];
  exact_dlt(tm, ws.Y, 2 * n, [&](int r, float* row) {
    const int o = list[r >> 1], half = r & 1;
    const ViewKp<float>& v = ws.vw[o * NKP + k];
    exact_row(tb.camf[ws.obs_cam[o]].P, half, half ? v.y : v.x, v.conf, true, row);
  }, X);
  const double e_avg = exact_reproj_error(tb, ws, k, list, n, X);
  tm.single([&] {
    ws.jX[k * 3] = X[0]; ws.jX[k * 3 + 1] = X[1]; ws.jX[k * 3 + 2] = X[2];
    ws.jerr[k] = e_avg;
  });
}
template <class Team>
SES_HD void exact_weighted_resolve(Team&, const Tables&, const TriWs<double>&, int, const uint8_t*, int) {}

// The same solve run by ONE thread on its own copy of A (rows x 4 followed by the 4 x 4 rotation accumulator): the
// oracle's onesided_jacobi statement for statement. Used where many independent solves exist at once (the far joints
// of a hypothesis, the sigma points of a far joint), one per lane. Element e of a lane's copy lives at B[e * ld]: the
// lanes' copies are interleaved (ld = team size), so a warp-wide access to "element e" is one coalesced line.
SES_HDN void exact_dlt_serial(float* B, int ld, int rows, float X[3]) {
  float* W = B + (size_t)4 * rows * ld;
  for (int i = 0; i < 16; ++i) W[i * ld] = (i >> 2) == (i & 3) ? 1.f : 0.f;
  const float eps = 1.1920929e-7f;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        float alpha = 0.f, beta = 0.f, gamma = 0.f;
        for (int r = 0; r < rows; ++r) {
          const float bp = B[(r * 4 + p) * ld], bq = B[(r * 4 + q) * ld];
          alpha = xadd(alpha, xmul(bp, bp));
          beta = xadd(beta, xmul(bq, bq));
          gamma = xadd(gamma, xmul(bp, bq));
        }
        if (ses_abs(gamma) <= xmul(eps, xsqrt(xmul(alpha, beta))) || gamma == 0.f) continue;
        rotated = true;
        const float zeta = xdiv(xsub(beta, alpha), xmul(2.f, gamma));
        const float t = xdiv(zeta >= 0.f ? 1.f : -1.f, xadd(ses_abs(zeta), xsqrt(xadd(1.f, xmul(zeta, zeta)))));
        const float c = xdiv(1.f, xsqrt(xadd(1.f, xmul(t, t))));
        const float sn = xmul(c, t);
        for (int r = 0; r < rows + 4; ++r) {
          float* row = B + (size_t)r * 4 * ld;
          const float bp = row[p * ld], bq = row[q * ld];
          row[p * ld] = xsub(xmul(c, bp), xmul(sn, bq));
          row[q * ld] = xadd(xmul(sn, bp), xmul(c, bq));
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  float best_s = FLT_MAX;
  for (int c = 0; c < 4; ++c) {
    float sq = 0.f;
    for (int r = 0; r < rows; ++r) sq = xadd(sq, xmul(B[(r * 4 + c) * ld], B[(r * 4 + c) * ld]));
    if (sq < best_s) { best_s = sq; best = c; }
  }
  const float w = W[(12 + best) * ld];
  X[0] = xdiv(W[best * ld], w); X[1] = xdiv(W[(4 + best) * ld], w); X[2] = xdiv(W[(8 + best) * ld], w);
}

constexpr int FAR_COV_MAX_VIEWS = 16;
constexpr int FAR_COV_STRIDE = FAR_COV_MAX_VIEWS * 8 + 16;   // floats per private solve copy

// Exact weighted re-solve of every joint whose jflag has bit 0 set and that fits a private copy (n <= FAR_COV_MAX_VIEWS):
// the solves are independent, so each lane takes one joint - a garbage hypothesis with 17 far joints costs one pass
// instead of 17 cooperative solves. Joints with more views go through the cooperative solver (exact_weighted_resolve).
template <class Team>
SES_COLD_FN void exact_weighted_resolve_lanes(Team& tm, const Tables& tb, const TriWs<float>& ws, int C, float* scratch) {
  const int ld = tm.size();
  tm.pfor(NKP, [&](int k) {
    const int n = ws.jn[k];
    if (!(ws.jflag[k] & 1) || n > FAR_COV_MAX_VIEWS) return;
    const uint8_t* list = ws.vlist + k * C;
    float* B = scratch + (k % ld);
    for (int r = 0; r < 2 * n; ++r) {
      const int o = list[r >> 1], half = r & 1;
      const ViewKp<float>& v = ws.vw[o * NKP + k];
      float row[4];
      exact_row(tb.camf[ws.obs_cam[o]].P, half, half ? v.y : v.x, v.conf, true, row);
      for (int c = 0; c < 4; ++c) B[(r * 4 + c) * ld] = row[c];
    }
    float X[3];
    exact_dlt_serial(B, ld, 2 * n, X);
    ws.jX[k * 3] = X[0]; ws.jX[k * 3 + 1] = X[1]; ws.jX[k * 3 + 2] = X[2];
    ws.jerr[k] = exact_reproj_error(tb, ws, k, list, n, X);
  });
}
template <class Team>
SES_HD void exact_weighted_resolve_lanes(Team&, const Tables&, const TriWs<double>&, int, float*) {}

// Unscented covariance of the far joints (jflag bit 1) in the oracle's arithmetic (calc_covariance S3D:508-523 with
// draw_sigma_points S3D:489-506 and mod_samples S3D:471-487). For a point hundreds of metres away the sigma points
// straddle the pole of X = v_xyz / v_w, so anything but the same arithmetic gives an unrelated (equally meaningless)
// covariance. The 4n + 1 solves of ALL far joints of the hypothesis form one index space: one solve per lane on a
// private copy of A in `scratch` ([FAR_COV_STRIDE][team size] floats, global memory on the GPU), Y_CHUNK solves per
// pass; the transformed points are staged in the (idle) sigma-point buffer and folded per joint in sample order like
// the reference does.
template <class Team>
SES_COLD_FN void exact_far_covariances(Team& tm, const Tables& tb, int p_max, const ses3d_person2d* persons,
                                   const TriWs<float>& ws, int C, float* scratch) {
  const int ld = tm.size();
  for (int k0 = 0; k0 < NKP;) {
    tm.single([&] {  // batch [k0,k1): far joints whose 4n+1 solves fit into the staging buffer; soff = staging offsets
      int k1 = k0, used = 0;
      while (k1 < NKP) {
        const int need = (ws.jflag[k1] & 2) ? 4 * ws.jn[k1] + 1 : 0;
        if (used + need > Y_CHUNK && used > 0) break;
        ws.soff[k1] = used;
        used += need;
        ++k1;
      }
      ws.scal[2] = k1;
      ws.scal[3] = used;
    });
    const int k1 = ws.scal[2], used = ws.scal[3];
    if (used > 0) {
      tm.pfor(used, [&](int i) {
        int k = k0;
        while (k < k1 - 1 && !((ws.jflag[k] & 2) && i < ws.soff[k] + 4 * ws.jn[k] + 1)) ++k;
        const int s = i - ws.soff[k], n = ws.jn[k];
        const uint8_t* list = ws.vlist + k * C;
        const float dimk = xadd((float)(2 * n), 0.5f);          // T(dim) + kappa
        const float b = xsqrt(dimk);
        const int vi = s > 0 ? (s - 1) >> 2 : -1, m = (s - 1) & 3;
        float px = 0.f, py = 0.f;   // the perturbed keypoint of view vi
        if (vi >= 0) {
          const int o = list[vi], cam = ws.obs_cam[o];
          const ses3d_keypoint2d& kp = persons[cam * p_max + ws.obs_det[o]].keypoints[k];
          const CamF& cm = tb.camf[cam];
          const ViewKp<float>& v = ws.vw[o * NKP + k];
          const float cxx = xdiv(kp.cov[0], xmul(cm.fx, cm.fx)), cxy = xdiv(kp.cov[1], xmul(cm.fx, cm.fy)),
                      cyy = xdiv(kp.cov[2], xmul(cm.fy, cm.fy));            // normalize_keypoints S3D:324-327
          const float l11 = xsqrt(cxx), l21 = xdiv(cxy, l11);
          const float l22 = xsqrt(xsub(cyy, xmul(l21, l21)));
          const float dx1 = xmul(l11, b), dy1 = xmul(l21, b), dy2 = xmul(l22, b);
          px = v.x; py = v.y;
          if (m == 0) { px = xsub(v.x, dx1); py = xsub(v.y, dy1); }
          else if (m == 1) { py = xsub(v.y, dy2); }
          else if (m == 2) { px = xadd(v.x, dx1); py = xadd(v.y, dy1); }
          else { py = xadd(v.y, dy2); }
        }
        float* B = scratch + (i % ld);
        for (int r = 0; r < 2 * n; ++r) {
          const int vr = r >> 1, o = list[vr], half = r & 1;
          const ViewKp<float>& v = ws.vw[o * NKP + k];
          const float coord = vr == vi ? (half ? py : px) : (half ? v.y : v.x);
          float row[4];
          exact_row(tb.camf[ws.obs_cam[o]].P, half, coord, 1.f, false, row);
          for (int c = 0; c < 4; ++c) B[(r * 4 + c) * ld] = row[c];
        }
        exact_dlt_serial(B, ld, 2 * n, ws.Y + i * 3);
      });
      tm.pfor(k1 - k0, [&](int kk) {
        const int k = k0 + kk;
        if (!(ws.jflag[k] & 2)) return;
        const int n = ws.jn[k], n_samples = 4 * n + 1;
        const float dimk = xadd((float)(2 * n), 0.5f);
        const float wden = xmul(2.f, dimk);
        const float w0 = xdiv(xmul(2.f, 0.5f), wden), wi = xdiv(1.f, wden);
        const float* Yk = ws.Y + ws.soff[k] * 3;
        const float m0 = ws.jX[k * 3], m1 = ws.jX[k * 3 + 1], m2 = ws.jX[k * 3 + 2];
        float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f;
        for (int s = 0; s < n_samples; ++s) {
          const float w = s == 0 ? w0 : wi;
          const float d0 = xsub(Yk[s * 3], m0), d1 = xsub(Yk[s * 3 + 1], m1), d2 = xsub(Yk[s * 3 + 2], m2);
          c00 = xadd(c00, xmul(xmul(d0, w), d0)); c01 = xadd(c01, xmul(xmul(d0, w), d1)); c02 = xadd(c02, xmul(xmul(d0, w), d2));
          c11 = xadd(c11, xmul(xmul(d1, w), d1)); c12 = xadd(c12, xmul(xmul(d1, w), d2)); c22 = xadd(c22, xmul(xmul(d2, w), d2));
        }
        float* cv = ws.cov + k * 6;
        cv[0] = c00; cv[1] = c01; cv[2] = c02; cv[3] = c11; cv[4] = c12; cv[5] = c22;
      });
    }
    k0 = k1;
  }
}
template <class Team>
SES_HD void exact_far_covariances(Team&, const Tables&, int, const ses3d_person2d*, const TriWs<double>&, int, float*) {}

// LM refinement of sum conf^2 * ||hnorm(P X~) - x||^2 (self-specified, not in the reference)
template <class T>
SES_COLD_FN void lm_refine_joint(const Tables& tb, const TriWs<T>& ws, int k, const uint8_t* list, int n, T X[3]) {
  auto cost_at = [&](const T* Y) {
    T f = 0;
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      const T a = P[0] * Y[0] + P[1] * Y[1] + P[2] * Y[2] + P[3];
      const T b = P[4] * Y[0] + P[5] * Y[1] + P[6] * Y[2] + P[7];
      const T c = P[8] * Y[0] + P[9] * Y[1] + P[10] * Y[2] + P[11];
      const T ic = ses_rcp(c);
      const T rx = v.conf * (a * ic - v.x), ry = v.conf * (b * ic - v.y);
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
      const T ic = ses_rcp(c), u = a * ic, vv = b * ic;
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
//
// live = false: this member of a lockstep team has no hypothesis (padding slot); it only passes the phase barriers - every
// member executes the same barrier call sites. team_ws: workspaces of all members when a team of teams spreads the two
// per-joint phases over all its threads (team.shared_pfor passes the member index); nullptr = every member runs its own
// joints (the measured optimum, see team.h).
template <class T, class Team>
SES_HD void triangulate_hypothesis(Team& tm, const Tables& tb, int p_max, const ses3d_person2d* persons,
                                   const int8_t* hyp_det_row, const TriWs<T>& ws, ses3d_person_cov* out,
                                   int32_t* keep, const TriWs<T>* team_ws = nullptr, bool live = true) {
  const int C = tb.n_cams;
  const double max_reproj = tb.prm.reproj_error_max_acceptable;
  const float thr = tb.prm.triangulation_threshold;
  if (!team_ws) team_ws = &ws;

  tm.phase();   // 1 of TRI_PHASES: gather + normalise
  int n_obs_own = 0;
  if (live) {
    // observation list in camera order: stream compaction of the hypothesis' row of the association table
    n_obs_own = tm.compact(C, [&](int c) { return hyp_det_row[c] >= 0; },
                           [&](int c, int pos) { ws.obs_cam[pos] = (uint8_t)c; ws.obs_det[pos] = (uint8_t)hyp_det_row[c]; });
    if (n_obs_own < 2) {  // S3D:684: hypotheses with a single observation are not triangulated
      tm.single([&] { *keep = 0; });
      n_obs_own = 0;
      live = false;
    }
  }
  if (live) {
    // normalised keypoints of the hypothesis' observations
    tm.pfor(n_obs_own * NKP, [&](int i) {
      const int o = i / NKP, k = i % NKP;
      const int cam = ws.obs_cam[o];
      normalize_kp(tb, cam, persons[cam * p_max + ws.obs_det[o]].keypoints[k], ws.vw[i]);
    });
  }

  // per joint: gather views, weighted DLT, 3-view epipolar rejection (S3D:718-792)
  // 2: weighted solves (shared phase; the explicit capture list keeps member-specific state out of the body)
  tm.shared_pfor(NKP, n_obs_own, [&tb, team_ws, C, thr, max_reproj](int member, int n_obs, int k) {
    const TriWs<T>& ws = team_ws[member];
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
    solve_weighted<T>(tb, ws, k, list, n, -1, X, &err);
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
        const TriWs<T> wc = ws;   // out-of-line calls get copies: ws must not become addressable (see COLD CALLS)
        const ColdSolve<T> cs = solve_weighted_cold<T>(tb, wc, k, list, n, -1);
        X[0] = cs.x; X[1] = cs.y; X[2] = cs.z; err = cs.err;
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

  tm.phase();   // 3: the rare paths (leave-one-out, exact re-solves)
  const bool exact_on = sizeof(T) == 4 && (tb.exact_mode & 1);
  const int cap_n = (int)((size_t)Y_CHUNK * 3 - 16) / 8;   // views the cooperative solver can stage
  if (live) {
  // leave-one-out (S3D:793-838), rare: joints with a large error are handled in batches that fit
  // the scratch; one (joint, left-out view) solve per thread, then the reference's sequential selection
  const bool any_loo = tm.first(NKP, [&](int k) { return ws.jflag[k] != 0; }) < NKP;   // usually none: skip the batching
  for (int k0 = any_loo ? 0 : NKP; k0 < NKP;) {
    tm.single([&] {  // batch [k0,k1): flagged joints whose n solves fit into loo_cap; soff = scratch offsets
      int k1 = k0, used = 0;
      while (k1 < NKP) {
        const int need = ws.jflag[k1] ? ws.jn[k1] : 0;
        if (used + need > ws.loo_cap && used > 0) break;
        ws.soff[k1] = used;
        used += need;
        ++k1;
      }
      ws.scal[2] = k1;
      ws.scal[3] = used;
    });
    const int k1 = ws.scal[2], used = ws.scal[3];
    if (used > 0) {
      tm.pfor(used, [&](int i) {
        int k = k0;
        while (k < NKP - 1 && !(ws.jflag[k] && i < ws.soff[k] + ws.jn[k])) ++k;
        const TriWs<T> wc = ws;
        const ColdSolve<T> cs = solve_weighted_cold<T>(tb, wc, k, wc.vlist + k * C, wc.jn[k], i - wc.soff[k]);
        ws.looX[i * 3] = cs.x; ws.looX[i * 3 + 1] = cs.y; ws.looX[i * 3 + 2] = cs.z;
        ws.looErr[i] = cs.err;
      });
      tm.pfor(k1 - k0, [&](int kk) {
        const int k = k0 + kk;
        if (!ws.jflag[k]) return;
        const int base = ws.soff[k];
        const int n = ws.jn[k];
        uint8_t* list = ws.vlist + k * C;
        const double err = ws.jerr[k];
        double best_err = err;
        int best = -1;
        float best_score = ws.jscore[k];
        for (int i = 0; i < n; ++i) {
          const double e_sub = ws.looErr[base + i];
          if (best_err > e_sub && e_sub < 0.9 * err) {
            best_err = e_sub; best = i;
            float tmp = 0.f;
            for (int j = 0; j < n; ++j)
              if (j != i) tmp += (float)ws.vw[list[j] * NKP + k].conf;
            best_score = tmp / (float)(n - 1);
          }
        }
        if (best != -1) {
          ws.jX[k * 3] = ws.looX[(base + best) * 3]; ws.jX[k * 3 + 1] = ws.looX[(base + best) * 3 + 1];
          ws.jX[k * 3 + 2] = ws.looX[(base + best) * 3 + 2];
          for (int i = best; i < n - 1; ++i) list[i] = list[i + 1];
          ws.jn[k] = n - 1;
          ws.jerr[k] = best_err;
          ws.jscore[k] = best_score;
        }
      });
    }
    k0 = k1;
  }

  // far points and high-residual joints (FP32 mode): exact re-solve of the final view set, see exact_weighted_resolve.
  // jflag (the leave-one-out marks are spent) now holds bit 0 = re-solve exactly, bit 1 = far joint (its covariance is
  // solved exactly as well). Triggers: the approximate position lies beyond the far-point radius, or the residual is
  // above the acceptance threshold (a gross outlier left in the view set: the large smallest singular value narrows
  // the gap to the next one, which amplifies rounding the same way, and the residual scales the published score,
  // S3D:840-844 - solved exactly, both match the oracle to the last bit).
  tm.pfor(NKP, [&](int k) {
    int fl = 0;
    const int n = ws.jn[k];
    if (exact_on && n >= 2 && n <= cap_n) {
      const T x = ws.jX[k * 3], y = ws.jX[k * 3 + 1], z = ws.jX[k * 3 + 2];
      const bool far = x * x + y * y + z * z > T(FAR_POINT_R2);
      if (far || ws.jerr[k] > max_reproj) fl = far ? 3 : 1;
    }
    ws.jflag[k] = fl;
  });
  if (exact_on && tm.first(NKP, [&](int k) { return ws.jflag[k] != 0; }) < NKP) {
    // Most far joints belong to garbage hypotheses (two detections of different people matched) and never reach the
    // output. Exactness is only owed to what is published, so two conservative filters run on the approximate
    // positions first:
    //  (a) the hypothesis is certainly dropped by the plausibility count (S3D:923-968): even when every joint whose
    //      root distance is not clearly beyond the limit is counted as kept, num_valid <= min_num_valid_keypoints.
    //      "Clearly" = by more than a margin that bounds the difference between the approximate and the exact position
    //      (1e-4 relative to the distances involved - an order above anything measured - plus 1 mm);
    //  (b) the root lies inside the far-point radius and the joint more than 3 m beyond it: the joint is reset by the
    //      root-distance rule (S3D:937-953) whatever its last digits are.
    // LM refinement moves the points after this step, so the filters are off when it is enabled.
    tm.single([&] {
      bool root_near = false, dropped = false;
      if (!tb.prm.lm_refine) {
        auto joint_of = [&](int slot) {
          for (int k = 0; k < NKP; ++k)
            if (tb.model.fusion_idx[k] == slot) return ws.jn[k] >= 2 ? k : -1;
          return -1;
        };
        int T_cnt = 0;
        for (int k = 0; k < NKP; ++k) T_cnt += ws.jn[k] >= 2 ? 1 : 0;
        const int k_mid = joint_of(SES3D_FBP_MIDHIP), k_lh = joint_of(SES3D_FBP_LHIP), k_rh = joint_of(SES3D_FBP_RHIP);
        bool have_root = false;
        double rx = 0, ry = 0, rz = 0;
        if (k_mid >= 0) {
          have_root = true;
          rx = ws.jX[k_mid * 3]; ry = ws.jX[k_mid * 3 + 1]; rz = ws.jX[k_mid * 3 + 2];
        } else if (k_lh >= 0 && k_rh >= 0) {
          have_root = true;
          rx = 0.5 * ((double)ws.jX[k_lh * 3] + (double)ws.jX[k_rh * 3]);
          ry = 0.5 * ((double)ws.jX[k_lh * 3 + 1] + (double)ws.jX[k_rh * 3 + 1]);
          rz = 0.5 * ((double)ws.jX[k_lh * 3 + 2] + (double)ws.jX[k_rh * 3 + 2]);
        }
        if (!have_root) {
          dropped = T_cnt <= tb.prm.min_num_valid_keypoints;   // num_valid = T when the root loop is skipped
        } else {
          const double rn = sqrt(rx * rx + ry * ry + rz * rz);
          int f_certain = 0;   // joints certainly further than max_joint_dist_to_root from the root
          for (int k = 0; k < NKP; ++k) {
            if (ws.jn[k] < 2) continue;
            const double x = ws.jX[k * 3], y = ws.jX[k * 3 + 1], z = ws.jX[k * 3 + 2];
            const double d = sqrt((x - rx) * (x - rx) + (y - ry) * (y - ry) + (z - rz) * (z - rz));
            const double margin = 1e-4 * (rn + sqrt(x * x + y * y + z * z)) + 1e-3;
            if (d > tb.prm.max_joint_dist_to_root + margin) ++f_certain;   // NaN distances count as kept
          }
          dropped = 2 * T_cnt - NFUS - f_certain <= tb.prm.min_num_valid_keypoints;   // S3D:937-953, see SURVEY a10
          root_near = rn * rn <= (double)FAR_POINT_R2 && tb.prm.max_joint_dist_to_root <= 2.5;
        }
      }
      const T r_skip = T(23.0 * 23.0);
      const bool far_cov = ws.far_scratch && (tb.exact_mode & 2);
      for (int k = 0; k < NKP; ++k) {
        int fl = ws.jflag[k];
        if (!fl) continue;
        const T x = ws.jX[k * 3], y = ws.jX[k * 3 + 1], z = ws.jX[k * 3 + 2];
        const T r2 = x * x + y * y + z * z;
        if (dropped) fl = 0;                                           // (a) nothing of this hypothesis is published
        else if (root_near && r2 > r_skip && r2 < T(1e30)) fl = 0;     // (b) will be reset by the root-distance rule
        else if (!(far_cov && ws.jn[k] <= FAR_COV_MAX_VIEWS)) fl &= 1;
        ws.jflag[k] = fl;
      }
    });
    const TriWs<T> wc = ws;
    if (wc.far_scratch) exact_weighted_resolve_lanes(tm, tb, wc, C, wc.far_scratch);
    for (int k = 0; k < NKP; ++k)   // joints the lane-parallel pass cannot hold (or no private scratch): cooperative
      if ((ws.jflag[k] & 1) && (!ws.far_scratch || ws.jn[k] > FAR_COV_MAX_VIEWS))
        exact_weighted_resolve(tm, tb, wc, k, wc.vlist + k * C, wc.jn[k]);
  }

  }   // live (phase 3)
  // 4: base systems (shared phase): optional LM, down-weight (S3D:840-844), then the unweighted base system of the final
  // view set: its full eigen-decomposition is the warm-start basis and its smallest eigenvector is sigma point 0
  tm.shared_pfor(NKP, n_obs_own, [&tb, team_ws, C, max_reproj](int member, int, int k) {
    const TriWs<T>& ws = team_ws[member];
    const int n = ws.jn[k];
    for (int i = 0; i < 6; ++i) ws.cov[k * 6 + i] = T(0);
    if (n < 2) return;
    const uint8_t* list = ws.vlist + k * C;
    double err = ws.jerr[k];
    float avg_score = ws.jscore[k];
    T X[3] = {ws.jX[k * 3], ws.jX[k * 3 + 1], ws.jX[k * 3 + 2]};
    if (tb.prm.lm_refine) {
      T Xr[3] = {X[0], X[1], X[2]};   // a copy: the out-of-line call must not make X addressable
      const TriWs<T> wc = ws;
      lm_refine_joint<T>(tb, wc, k, list, n, Xr);
      X[0] = Xr[0]; X[1] = Xr[1]; X[2] = Xr[2];
      err = reproj_error<T>(tb, ws, k, list, n, -1, X);
      ws.jX[k * 3] = X[0]; ws.jX[k * 3 + 1] = X[1]; ws.jX[k * 3 + 2] = X[2];
    }
    if (err > max_reproj) avg_score = (float)((double)avg_score * (max_reproj / err));
    ws.jscore[k] = avg_score;
    if (ws.jflag[k] & 2) return;   // far joint: covariance by exact_far_covariances below
    double G[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      T r[4];
      dlt_row_fast<T>(P, 0, v.x, r); gram_add<T>(G, r, 1.0);
      dlt_row_fast<T>(P, 1, v.y, r); gram_add<T>(G, r, 1.0);
    }
    T e[4];
    smallest_eigvec4_fast<T>(G, e);
    // base system in the deflated basis [e | Q]: a0 = e^T G e, b0 = Q^T G e (~0), C0 = Q^T G Q
    T a0 = 0, b0[3] = {0, 0, 0}, C0[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const int o = list[i];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, ws.obs_cam[o]);
      for (int which = 0; which < 2; ++which) {
        T r[4], sr, p[3];
        dlt_row_fast<T>(P, which, which == 0 ? v.x : v.y, r);
        deflate_project<T>(e, r, sr, p);
        a0 += sr * sr;
        b0[0] += sr * p[0]; b0[1] += sr * p[1]; b0[2] += sr * p[2];
        C0[0] += p[0] * p[0]; C0[1] += p[0] * p[1]; C0[2] += p[0] * p[2];
        C0[3] += p[1] * p[1]; C0[4] += p[1] * p[2]; C0[5] += p[2] * p[2];
      }
    }
    T* df = ws.defl + k * 14;
    df[0] = e[0]; df[1] = e[1]; df[2] = e[2]; df[3] = e[3]; df[4] = a0;
    df[5] = b0[0]; df[6] = b0[1]; df[7] = b0[2];
    for (int i = 0; i < 6; ++i) df[8 + i] = C0[i];
    // sigma point 0 (unperturbed, unweighted) contributes w0 * (y0 - m)(y0 - m)^T   (S3D:521-522)
    const T wden = T(2) * (T(2 * n) + T(0.5));
    const T w0 = (T(2) * T(0.5)) / wden;
    const T ie3 = ses_rcp(e[3]);
    const T d0 = e[0] * ie3 - X[0], d1 = e[1] * ie3 - X[1], d2 = e[2] * ie3 - X[2];
    T* cv = ws.cov + k * 6;
    cv[0] = (d0 * w0) * d0; cv[1] = (d0 * w0) * d1; cv[2] = (d0 * w0) * d2;
    cv[3] = (d1 * w0) * d1; cv[4] = (d1 * w0) * d2; cv[5] = (d2 * w0) * d2;
  });

  tm.phase();   // 5: sigma points
  if (live) {
  static_assert(NKP <= 32, "team.scan handles at most 32 items");
  tm.scan(NKP, [&](int k) { return (ws.jn[k] >= 2 && !(ws.jflag[k] & 2)) ? 4 * ws.jn[k] : 0; }, ws.soff);
  const int n_samples_total = ws.soff[NKP];
  tm.pfor(NKP, [&](int k) {   // a joint owns 4n consecutive samples: four map entries per store
    uint32_t* km = reinterpret_cast<uint32_t*>(ws.kmap);
    for (int w = ws.soff[k] >> 2; w < (ws.soff[k + 1] >> 2); ++w) km[w] = (uint32_t)k * 0x01010101u;
  });

  // unscented sigma points 1..4n (S3D:471-506) of all joints in one index space, Y_CHUNK per pass
  for (int s0 = 0; s0 < n_samples_total; s0 += Y_CHUNK) {
    const int cnt = (n_samples_total - s0) < Y_CHUNK ? (n_samples_total - s0) : Y_CHUNK;
    tm.pfor(cnt, [&](int ii) {
      const int i = s0 + ii;
      const int k = ws.kmap[i];
      const int s = i - ws.soff[k];
      const int n = ws.jn[k];
      const int vi = s >> 2, m = s & 3;
      const int o = ws.vlist[k * C + vi];
      const int cam = ws.obs_cam[o];
      const ViewKp<T>& v = ws.vw[o * NKP + k];
      const T* P = CamSel<T>::P(tb, cam);
      const T* df = ws.defl + k * 14;
      T l11, l21, l22;
      cholesky_cov(tb, cam, persons[cam * p_max + ws.obs_det[o]].keypoints[k], l11, l21, l22);
      const T b = ses_sqrt(T(2 * n) + T(0.5));  // sqrt(dim + kappa), S3D:500
      T nx = v.x, ny = v.y;                    // mod_samples S3D:481-486
      if (m == 0) { nx = v.x - l11 * b; ny = v.y - l21 * b; }
      else if (m == 1) { ny = v.y - l22 * b; }
      else if (m == 2) { nx = v.x + l11 * b; ny = v.y + l21 * b; }
      else { ny = v.y + l22 * b; }
      // perturbed system in the deflated basis: base +/- the rows of the one view that moved
      T a = df[4], bb[3] = {df[5], df[6], df[7]}, Cm[6] = {df[8], df[9], df[10], df[11], df[12], df[13]};
      auto update = [&](int which, T coord, T sign) {
        T r[4], sr, p[3];
        dlt_row_fast<T>(P, which, coord, r);
        deflate_project<T>(df, r, sr, p);
        const T ss = sign * sr;
        a += ss * sr;
        bb[0] += ss * p[0]; bb[1] += ss * p[1]; bb[2] += ss * p[2];
        const T q0 = sign * p[0], q1 = sign * p[1], q2 = sign * p[2];
        Cm[0] += q0 * p[0]; Cm[1] += q0 * p[1]; Cm[2] += q0 * p[2];
        Cm[3] += q1 * p[1]; Cm[4] += q1 * p[2]; Cm[5] += q2 * p[2];
      };
      if ((m & 1) == 0) { update(0, v.x, T(-1)); update(0, nx, T(1)); }
      update(1, v.y, T(-1));
      update(1, ny, T(1));
      T e[4], xs[3];
      if (secular_smallest<T>(a, bb, Cm, xs)) {
        deflate_expand<T>(df, xs, e);
      } else {  // rare: rebuild the full normal matrix of this sigma point and use Jacobi
        double G[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i2 = 0; i2 < n; ++i2) {
          const int o2 = ws.vlist[k * C + i2];
          const ViewKp<T>& v2 = ws.vw[o2 * NKP + k];
          const T* P2 = CamSel<T>::P(tb, ws.obs_cam[o2]);
          T r[4];
          dlt_row_fast<T>(P2, 0, i2 == vi ? nx : v2.x, r); gram_add<T>(G, r, 1.0);
          dlt_row_fast<T>(P2, 1, i2 == vi ? ny : v2.y, r); gram_add<T>(G, r, 1.0);
        }
        T ec[4];
        smallest_eigvec4<T>(G, ec);
        e[0] = ec[0]; e[1] = ec[1]; e[2] = ec[2]; e[3] = ec[3];
      }
      const T inv = ses_rcp(e[3]);
      ws.Y[ii * 3] = e[0] * inv; ws.Y[ii * 3 + 1] = e[1] * inv; ws.Y[ii * 3 + 2] = e[2] * inv;
    });
    // covariance about the weighted-DLT point, samples in the reference's order (S3D:521-522)
    tm.pfor(NKP, [&](int k) {
      const int n = ws.jn[k];
      if (n < 2 || (ws.jflag[k] & 2)) return;
      int a = ws.soff[k], b = ws.soff[k + 1];
      a = a > s0 ? a : s0;
      b = b < s0 + cnt ? b : s0 + cnt;
      if (a >= b) return;
      const T wi = T(1) / (T(2) * (T(2 * n) + T(0.5)));
      const T m0 = ws.jX[k * 3], m1 = ws.jX[k * 3 + 1], m2 = ws.jX[k * 3 + 2];
      T* cv = ws.cov + k * 6;
      T c00 = cv[0], c01 = cv[1], c02 = cv[2], c11 = cv[3], c12 = cv[4], c22 = cv[5];
      for (int i = a; i < b; ++i) {
        const T d0 = ws.Y[(i - s0) * 3] - m0, d1 = ws.Y[(i - s0) * 3 + 1] - m1, d2 = ws.Y[(i - s0) * 3 + 2] - m2;
        c00 += (d0 * wi) * d0; c01 += (d0 * wi) * d1; c02 += (d0 * wi) * d2;
        c11 += (d1 * wi) * d1; c12 += (d1 * wi) * d2; c22 += (d2 * wi) * d2;
      }
      cv[0] = c00; cv[1] = c01; cv[2] = c02; cv[3] = c11; cv[4] = c12; cv[5] = c22;
    });
  }

  }   // live (phase 5)
  tm.phase();   // 6: far covariances, skeleton assembly, plausibility, output
  if (!live) return;
  if (exact_on && ws.far_scratch && tm.first(NKP, [&](int k) { return (ws.jflag[k] & 2) != 0; }) < NKP)
  {
    const TriWs<T> wc = ws;
    exact_far_covariances(tm, tb, p_max, persons, wc, C, wc.far_scratch);
  }

  // output keypoints (S3D:849-857); the record shares storage with the sigma-point staging, which is done
  tm.pfor(NFUS, [&](int s) { zero_kp(ws.kp[s]); });
  tm.pfor(NKP, [&](int k) {
    if (ws.jn[k] < 2) return;
    ses3d_keypoint_cov& o = ws.kp[tb.model.fusion_idx[k]];
    o.x = (double)ws.jX[k * 3]; o.y = (double)ws.jX[k * 3 + 1]; o.z = (double)ws.jX[k * 3 + 2];
    o.score = ws.jscore[k];
    for (int i = 0; i < 6; ++i) o.cov[i] = (double)ws.cov[k * 6 + i];
  });

  // limb-length covariance inflation (S3D:861-883): each joint reads positions only, so the joints
  // are independent except for the shoulder pair, which the RShoulder item updates alone
  tm.pfor(NKP, [&](int k) {
    const SkeletonModel& M = tb.model;
    ses3d_keypoint_cov& kp = ws.kp[M.fusion_idx[k]];
    if (kp.score <= 0) return;
    const int parent = M.parent[k];
    if (parent < 0) return;
    const ses3d_keypoint_cov& pk = ws.kp[M.fusion_idx[parent]];
    if (pk.score > 0 && M.limb_len[k] > 0) {
      add_cov(kp, tb.prm.limb_cov_offset_sigma * (joint_dist(kp, pk) - M.limb_len[k]) / M.limb_sigma[k]);
    } else if (tb.prm.pose_method == SES3D_POSE_SIMPLE && k == 6 /*RShoulder S3D:83*/) {
      ses3d_keypoint_cov& ls = ws.kp[M.fusion_idx[5 /*LShoulder S3D:86: parent Nose, limbLength -1: no own update*/]];
      if (ls.score > 0) {
        const double d = joint_dist(kp, ls);
        add_cov(kp, tb.prm.limb_cov_offset_sigma * (d - 0.35) / 0.15);  // shoulderDist, shoulderSigma S3D:103
        add_cov(ls, tb.prm.limb_cov_offset_sigma * (d - 0.35) / 0.15);
      }
    }
  });

  // root distance / feet height / keep decision (S3D:923-973)
  tm.single([&] {
    const ses3d_keypoint_cov* K = ws.kp;
    ses3d_keypoint_cov root;
    zero_kp(root);
    if (K[SES3D_FBP_MIDHIP].score > 0) root = K[SES3D_FBP_MIDHIP];
    else if (K[SES3D_FBP_LHIP].score > 0 && K[SES3D_FBP_RHIP].score > 0) {
      root.x = (K[SES3D_FBP_LHIP].x + K[SES3D_FBP_RHIP].x) / 2.;
      root.y = (K[SES3D_FBP_LHIP].y + K[SES3D_FBP_RHIP].y) / 2.;
      root.z = (K[SES3D_FBP_LHIP].z + K[SES3D_FBP_RHIP].z) / 2.;
      root.score = (K[SES3D_FBP_LHIP].score + K[SES3D_FBP_RHIP].score) / 2.f;
    }
    // root is stashed for the parallel pass below
    ws.jerr[0] = root.x; ws.jerr[1] = root.y; ws.jerr[2] = root.z; ws.jerr[3] = (double)root.score;
  });
  tm.pfor(NFUS, [&](int s) {  // far-from-root joints are reset (S3D:937-953)
    if (!(ws.jerr[3] > 0)) return;
    ses3d_keypoint_cov& kp = ws.kp[s];
    ses3d_keypoint_cov root;
    root.x = ws.jerr[0]; root.y = ws.jerr[1]; root.z = ws.jerr[2];
    if (kp.score > 0) {
      if (joint_dist(root, kp) > tb.prm.max_joint_dist_to_root) zero_kp(kp);
    } else {
      zero_kp(kp);
    }
  });
  // every slot that is empty after the reset decrements the counter (S3D:940-951): joints reset by the distance test
  // and the never-filled slots alike; joints that were triangulated with a non-positive score were counted once and
  // are removed here as "empty" exactly like the reference does (kp.score > 0 is its only test)
  int num_valid_all = tm.count(NKP, [&](int k) { return ws.jn[k] >= 2; });
  if (ws.jerr[3] > 0) num_valid_all -= tm.count(NFUS, [&](int s) { return !(ws.kp[s].score > 0); });
  tm.single([&] {
    const ses3d_keypoint_cov* K = ws.kp;
    int num_valid = num_valid_all;
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
