// geom.h — per-thread geometric primitives of the triangulation path (host+device).
//
// DLT (S3D:440-465): the reference stacks two unit-normalised rows per view into A (2n x 4)
// and takes the right singular vector of the smallest singular value (Eigen JacobiSVD).
// Here every solve is a 4x4 symmetric eigenproblem on the normal matrix G = A^T A held in
// registers: rows are built in T exactly as the reference builds them, their outer products
// are accumulated in double (so G is A^T A to ~1e-16 and rank-2 updates for the unscented
// sigma points do not lose digits), and the smallest eigenvector comes from cyclic Jacobi
// rotations in T.
#pragma once
#include "common.h"

namespace ses3d {

// order of the 10 unique entries of a symmetric 4x4: 00 01 02 03 11 12 13 22 23 33
template <class T>
SES_HD void dlt_row(const T* P, int which /*0: x-row, 1: y-row*/, T m, T weight, bool weighted, T r[4]) {
  // x * P.row(2) - P.row(which), normalised, optionally scaled by the confidence (S3D:446-453)
  r[0] = m * P[8] - P[which * 4 + 0];
  r[1] = m * P[9] - P[which * 4 + 1];
  r[2] = m * P[10] - P[which * 4 + 2];
  r[3] = m * P[11] - P[which * 4 + 3];
  const T z = sum4(r[0] * r[0], r[1] * r[1], r[2] * r[2], r[3] * r[3]);
  if (z > T(0)) {
    const T nrm = ses_sqrt(z);
    r[0] /= nrm; r[1] /= nrm; r[2] /= nrm; r[3] /= nrm;
  }
  if (weighted) { r[0] *= weight; r[1] *= weight; r[2] *= weight; r[3] *= weight; }
}

// Unweighted row for the sigma-point solves: same row, normalised with one SFU rsqrt instead of
// sqrt + 4 IEEE divides on the GPU float path (<= 2 ulp per entry; positions are tolerance-checked).
template <class T>
SES_HD void dlt_row_fast(const T* P, int which, T m, T r[4]) {
  dlt_row<T>(P, which, m, T(1), false, r);
}
#if defined(__CUDA_ARCH__)
template <>
SES_HD void dlt_row_fast<float>(const float* P, int which, float m, float r[4]) {
  r[0] = fmaf(m, P[8], -P[which * 4 + 0]);
  r[1] = fmaf(m, P[9], -P[which * 4 + 1]);
  r[2] = fmaf(m, P[10], -P[which * 4 + 2]);
  r[3] = fmaf(m, P[11], -P[which * 4 + 3]);
  const float z = (r[0] * r[0] + r[1] * r[1]) + (r[2] * r[2] + r[3] * r[3]);
  const float inv = z > 0.f ? rsqrtf(z) : 1.f;
  r[0] *= inv; r[1] *= inv; r[2] *= inv; r[3] *= inv;
}
#endif

template <class T>
SES_HD void gram_add(double G[10], const T r[4], double sign) {
  const double a = (double)r[0], b = (double)r[1], c = (double)r[2], d = (double)r[3];
  const double sa = sign * a, sb = sign * b, sc = sign * c, sd = sign * d;
  G[0] += sa * a; G[1] += sa * b; G[2] += sa * c; G[3] += sa * d;
  G[4] += sb * b; G[5] += sb * c; G[6] += sb * d;
  G[7] += sc * c; G[8] += sc * d;
  G[9] += sd * d;
}

// ---- 4x4 symmetric eigen-solver: cyclic Jacobi held entirely in registers ----------------------
template <class T>
struct Sym4V {  // matrix (10 unique entries) + accumulated rotations (eigenvector columns)
  T a00, a01, a02, a03, a11, a12, a13, a22, a23, a33;
  T v00, v01, v02, v03, v10, v11, v12, v13, v20, v21, v22, v23, v30, v31, v32, v33;
};

// tan / cos / sin of the Jacobi rotation that annihilates apq (Rutishauser). The float version
// uses the SFU approximations (rsqrt, fast divide): joint positions are tolerance-checked, and
// any |error| in the angle only slows convergence, it does not bias the fixed point.
SES_HD void jacobi_angle(double app, double aqq, double apq, double& t, double& c, double& s) {
  const double theta = (aqq - app) / (2.0 * apq);
  t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  c = 1.0 / sqrt(t * t + 1.0);
  s = t * c;
}
SES_HD void jacobi_angle(float app, float aqq, float apq, float& t, float& c, float& s) {
#if defined(__CUDA_ARCH__)
  const float d = aqq - app;
  const float h2 = fmaf(d, d, 4.0f * apq * apq);
  const float h = h2 * rsqrtf(h2);
  t = __fdividef(2.0f * apq, d + copysignf(h, d));
  c = rsqrtf(fmaf(t, t, 1.0f));
  s = t * c;
#else
  const float theta = (aqq - app) / (2.0f * apq);
  t = (theta >= 0.0f ? 1.0f : -1.0f) / (fabsf(theta) + sqrtf(theta * theta + 1.0f));
  c = 1.0f / sqrtf(t * t + 1.0f);
  s = t * c;
#endif
}

template <class T>
SES_HD void jacobi4(Sym4V<T>& m) {
  const T eps = sizeof(T) == 4 ? T(1.1920929e-7) : T(2.220446049250313e-16);
  const T thr2 = (eps * T(0.125)) * (eps * T(0.125));
  const int max_sweeps = sizeof(T) == 4 ? 10 : 14;
// rotate the (p,q) plane; the two other indices r,s couple through arp/arq and asp/asq
#define SES_ROT(app, aqq, apq, arp, arq, asp, asq, v0p, v0q, v1p, v1q, v2p, v2q, v3p, v3q)      \
  do {                                                                                          \
    const T apq_ = (apq);                                                                       \
    if (apq_ * apq_ > thr2 * ses_abs((app) * (aqq)) && apq_ != T(0)) {                          \
      rotated = true;                                                                           \
      T t, c, s;                                                                                \
      jacobi_angle((app), (aqq), apq_, t, c, s);                                                \
      (app) -= t * apq_;                                                                        \
      (aqq) += t * apq_;                                                                        \
      (apq) = T(0);                                                                             \
      T x_, y_;                                                                                 \
      x_ = (arp); y_ = (arq); (arp) = c * x_ - s * y_; (arq) = s * x_ + c * y_;                 \
      x_ = (asp); y_ = (asq); (asp) = c * x_ - s * y_; (asq) = s * x_ + c * y_;                 \
      x_ = (v0p); y_ = (v0q); (v0p) = c * x_ - s * y_; (v0q) = s * x_ + c * y_;                 \
      x_ = (v1p); y_ = (v1q); (v1p) = c * x_ - s * y_; (v1q) = s * x_ + c * y_;                 \
      x_ = (v2p); y_ = (v2q); (v2p) = c * x_ - s * y_; (v2q) = s * x_ + c * y_;                 \
      x_ = (v3p); y_ = (v3q); (v3p) = c * x_ - s * y_; (v3q) = s * x_ + c * y_;                 \
    }                                                                                           \
  } while (0)
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    bool rotated = false;
    SES_ROT(m.a00, m.a11, m.a01, m.a02, m.a12, m.a03, m.a13, m.v00, m.v01, m.v10, m.v11, m.v20, m.v21, m.v30, m.v31);
    SES_ROT(m.a00, m.a22, m.a02, m.a01, m.a12, m.a03, m.a23, m.v00, m.v02, m.v10, m.v12, m.v20, m.v22, m.v30, m.v32);
    SES_ROT(m.a00, m.a33, m.a03, m.a01, m.a13, m.a02, m.a23, m.v00, m.v03, m.v10, m.v13, m.v20, m.v23, m.v30, m.v33);
    SES_ROT(m.a11, m.a22, m.a12, m.a01, m.a02, m.a13, m.a23, m.v01, m.v02, m.v11, m.v12, m.v21, m.v22, m.v31, m.v32);
    SES_ROT(m.a11, m.a33, m.a13, m.a01, m.a03, m.a12, m.a23, m.v01, m.v03, m.v11, m.v13, m.v21, m.v23, m.v31, m.v33);
    SES_ROT(m.a22, m.a33, m.a23, m.a02, m.a03, m.a12, m.a13, m.v02, m.v03, m.v12, m.v13, m.v22, m.v23, m.v32, m.v33);
    if (!rotated) break;
  }
#undef SES_ROT
}

template <class T>
SES_HD void sym4_load(Sym4V<T>& m, const double g[10]) {
  m.a00 = (T)g[0]; m.a01 = (T)g[1]; m.a02 = (T)g[2]; m.a03 = (T)g[3]; m.a11 = (T)g[4]; m.a12 = (T)g[5];
  m.a13 = (T)g[6]; m.a22 = (T)g[7]; m.a23 = (T)g[8]; m.a33 = (T)g[9];
  m.v00 = 1; m.v01 = 0; m.v02 = 0; m.v03 = 0; m.v10 = 0; m.v11 = 1; m.v12 = 0; m.v13 = 0;
  m.v20 = 0; m.v21 = 0; m.v22 = 1; m.v23 = 0; m.v30 = 0; m.v31 = 0; m.v32 = 0; m.v33 = 1;
}

// NaN inputs must give NaN outputs like the reference's SVD does (e.g. a zero 2-D covariance makes the
// sigma points NaN, S3D:473-475): Jacobi's comparisons would otherwise silently skip every rotation.
SES_HD double nan_poison(const double g[10]) {
  return (((g[0] + g[1]) + (g[2] + g[3])) + ((g[4] + g[5]) + (g[6] + g[7])) + (g[8] + g[9])) * 0.0;
}

template <class T>
SES_HD void sym4_smallest(const Sym4V<T>& m, T v[4]) {
  T best = m.a00;
  v[0] = m.v00; v[1] = m.v10; v[2] = m.v20; v[3] = m.v30;
  if (m.a11 < best) { best = m.a11; v[0] = m.v01; v[1] = m.v11; v[2] = m.v21; v[3] = m.v31; }
  if (m.a22 < best) { best = m.a22; v[0] = m.v02; v[1] = m.v12; v[2] = m.v22; v[3] = m.v32; }
  if (m.a33 < best) { best = m.a33; v[0] = m.v03; v[1] = m.v13; v[2] = m.v23; v[3] = m.v33; }
}

// Eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix g (cold start).
template <class T>
SES_HD void smallest_eigvec4(const double g[10], T v[4]) {
  Sym4V<T> m;
  sym4_load(m, g);
  jacobi4(m);
  sym4_smallest(m, v);
  const T poison = (T)nan_poison(g);
  v[0] += poison; v[1] += poison; v[2] += poison; v[3] += poison;
}

// Full decomposition g = V diag(lam) V^T; V row-major, column c = eigenvector c. Also returns
// the smallest eigenvector (for free). Used once per joint as the basis for the warm solves.
template <class T>
SES_HD void eig4_full(const double g[10], T lam[4], T V[16], T v[4]) {
  Sym4V<T> m;
  sym4_load(m, g);
  jacobi4(m);
  lam[0] = m.a00; lam[1] = m.a11; lam[2] = m.a22; lam[3] = m.a33;
  V[0] = m.v00; V[1] = m.v01; V[2] = m.v02; V[3] = m.v03; V[4] = m.v10; V[5] = m.v11; V[6] = m.v12; V[7] = m.v13;
  V[8] = m.v20; V[9] = m.v21; V[10] = m.v22; V[11] = m.v23; V[12] = m.v30; V[13] = m.v31; V[14] = m.v32; V[15] = m.v33;
  sym4_smallest(m, v);
  const T poison = (T)nan_poison(g);
  v[0] += poison; v[1] += poison; v[2] += poison; v[3] += poison;
  lam[0] += poison;
}

// q = V^T r in double (V, r exact in double): the row expressed in the eigenbasis of the base system.
template <class T>
SES_HD void to_eigenbasis(const T V[16], const T r[4], double q[4]) {
  for (int c = 0; c < 4; ++c)
    q[c] = (double)V[c] * (double)r[0] + (double)V[4 + c] * (double)r[1] + (double)V[8 + c] * (double)r[2] +
           (double)V[12 + c] * (double)r[3];
}
SES_HD void gram_add_d(double G[10], const double q[4], double sign) {
  const double sa = sign * q[0], sb = sign * q[1], sc = sign * q[2], sd = sign * q[3];
  G[0] += sa * q[0]; G[1] += sa * q[1]; G[2] += sa * q[2]; G[3] += sa * q[3];
  G[4] += sb * q[1]; G[5] += sb * q[2]; G[6] += sb * q[3];
  G[7] += sc * q[2]; G[8] += sc * q[3];
  G[9] += sd * q[3];
}

// Warm solve: gp is the perturbed normal matrix expressed in the eigenbasis V0 of the base
// system (diag(lam0) + small symmetric update), so Jacobi starts almost diagonal and needs
// 1-3 sweeps instead of 5-7. Returns the smallest eigenvector in the original basis.
template <class T>
SES_HD void smallest_eigvec4_warm(const double gp[10], const T V0[16], T v[4]) {
  Sym4V<T> m;
  sym4_load(m, gp);
  jacobi4(m);
  T w[4];
  sym4_smallest(m, w);
  const T poison = (T)nan_poison(gp);
  for (int r = 0; r < 4; ++r)
    v[r] = V0[r * 4] * w[0] + V0[r * 4 + 1] * w[1] + V0[r * 4 + 2] * w[2] + V0[r * 4 + 3] * w[3] + poison;
}

// projection residual of one view, S3D:430-433
template <class T>
SES_HD T reproj_residual(const T* P, const T X[3], T x, T y) {
  const T a = sum4(P[0] * X[0], P[1] * X[1], P[2] * X[2], P[3] * T(1));
  const T b = sum4(P[4] * X[0], P[5] * X[1], P[6] * X[2], P[7] * T(1));
  const T c = sum4(P[8] * X[0], P[9] * X[1], P[10] * X[2], P[11] * T(1));
  const T dx = a / c - x, dy = b / c - y;
  return ses_sqrt(dx * dx + dy * dy);
}

}  // namespace ses3d
