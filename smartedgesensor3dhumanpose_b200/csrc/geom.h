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

#ifndef SES_UNROLL1
#define SES_UNROLL1 1   // 1: keep the short solver iterations as loops: -0.02 ms / 16384 frames (A/B: scripts/build_variants.py)
#endif

namespace ses3d {

SES_HD double ses_rsqrt(double x);
SES_HD double ses_rcp(double x);

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
#if defined(__CUDA_ARCH__)
    if (sizeof(T) == 8) {   // FP64 mode (tolerance-checked): one refined reciprocal square root instead of sqrt + 4 divides
      const T inv = (T)ses_rsqrt((double)z);
      r[0] *= inv; r[1] *= inv; r[2] *= inv; r[3] *= inv;
    } else
#endif
    {
      const T nrm = ses_sqrt(z);
      r[0] /= nrm; r[1] /= nrm; r[2] /= nrm; r[3] /= nrm;
    }
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

// Eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix g (cold start): the rarely taken fallback of every
// fast solve. SES_COLD_JACOBI selects whether it is a shared out-of-line function (smaller hot code) or inlined.
#ifndef SES_COLD_JACOBI
#define SES_COLD_JACOBI 0   // 1 = out of line: measured +0.01 .. 0.03 ms (call ABI costs more than the fetch it saves)
#endif
#if SES_COLD_JACOBI
#define SES_JACOBI_FN SES_HDN
#else
#define SES_JACOBI_FN SES_HD
#endif
template <class T>
SES_JACOBI_FN void smallest_eigvec4(const double g[10], T v[4]) {
  Sym4V<T> m;
  sym4_load(m, g);
  jacobi4(m);
  sym4_smallest(m, v);
  const T poison = (T)nan_poison(g);
  v[0] += poison; v[1] += poison; v[2] += poison; v[3] += poison;
}

// ---- fast path: inverse iteration + deflated secular solve ----------------------------------------
// The DLT normal matrix has one tiny eigenvalue (the squared residual) well separated from the other
// three, so (a) inverse iteration finds its eigenvector in 2-3 LDL^T solves instead of 5-7 Jacobi
// sweeps, and (b) a sigma point - the base system plus a rank-<=4 update - is solved in the basis
// [v | Q] (v = base eigenvector, Q = Householder complement) where the matrix is [[a, b^T], [b, C]]
// with small b: the eigenvector is (1, x), (C - lambda I) x = -b, lambda the smallest root of the
// secular equation a - lambda - b^T (C - lambda I)^-1 b = 0 (Newton from lambda = a converges
// monotonically because the function is concave and decreasing left of C's spectrum). Every routine
// reports failure (non-positive pivot, no convergence, NaN) and the caller falls back to Jacobi.
SES_HD float ses_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __fdividef(1.0f, x);
#else
  return 1.0f / x;
#endif
}
// FP64 mode on the GPU: an IEEE double division / square root expands to ~20 instructions with a long dependency chain.
// A single-precision SFU seed refined by two Newton steps in double is accurate to ~1 ulp (seed error 2e-7 ->
// 6e-14 -> 1e-16) at a third of the cost; arguments outside the comfortable float range take the exact path.
SES_HD double ses_rcp(double x) {
#if defined(__CUDA_ARCH__)
  const double ax = fabs(x);
  if (ax > 1e-290 && ax < 1e290) {
    // the double-precision SFU seed (MUFU.RCP64H, ~20 bits) instead of a round trip through float: the two F2F
    // conversions were 11 % of the FP64 kernel's instructions
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * (2.0 - x * y);
    y = y * (2.0 - x * y);
    return y;
  }
#endif
  return 1.0 / x;
}
SES_HD float ses_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
SES_HD double ses_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  if (x > 1e-290 && x < 1e290) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // MUFU.RSQ64H seed, refined below
    y = y * (1.5 - 0.5 * x * y * y);
    y = y * (1.5 - 0.5 * x * y * y);
    return y;
  }
#endif
  return 1.0 / sqrt(x);
}

// Smallest eigenvector (unit length) of the symmetric PSD 4x4 matrix G by inverse iteration.
template <class T>
SES_HD bool invit4(const double G[10], T v[4]) {
  const T g00 = (T)G[0], g01 = (T)G[1], g02 = (T)G[2], g03 = (T)G[3], g11 = (T)G[4], g12 = (T)G[5], g13 = (T)G[6],
          g22 = (T)G[7], g23 = (T)G[8], g33 = (T)G[9];
  const T rel = sizeof(T) == 4 ? T(1e-5) : T(1e-11);     // leading pivots must stay clear of rounding noise
  const T tol2 = sizeof(T) == 4 ? T(4e-12) : T(1e-26);   // squared change of the unit iterate
  if (!(g00 > T(0))) return false;
  const T i0 = ses_rcp(g00);
  const T l10 = g01 * i0, l20 = g02 * i0, l30 = g03 * i0;
  const T d1 = g11 - l10 * g01;
  if (!(d1 > rel * g11)) return false;
  const T i1 = ses_rcp(d1);
  const T t21 = g12 - l20 * g01, t31 = g13 - l30 * g01;
  const T l21 = t21 * i1, l31 = t31 * i1;
  const T d2 = g22 - l20 * g02 - l21 * t21;
  if (!(d2 > rel * g22)) return false;
  const T i2 = ses_rcp(d2);
  const T t32 = g23 - l30 * g02 - l31 * t21;
  const T l32 = t32 * i2;
  T d3 = g33 - l30 * g03 - l31 * t31 - l32 * t32;   // ~ the tiny eigenvalue; may round to <= 0, which is fine
  const T floor3 = g33 * (sizeof(T) == 4 ? T(1e-15) : T(1e-30));
  if (ses_abs(d3) < floor3) d3 = floor3;
  const T i3 = ses_rcp(d3);
  T x0 = 0, x1 = 0, x2 = 0, x3 = 1;
#if SES_UNROLL1
#pragma unroll 1
#endif
  for (int it = 0; it < 6; ++it) {
    // L z = x, z /= d, L^T y = z
    T z0 = x0;
    T z1 = x1 - l10 * z0;
    T z2 = x2 - l20 * z0 - l21 * z1;
    T z3 = x3 - l30 * z0 - l31 * z1 - l32 * z2;
    z0 *= i0; z1 *= i1; z2 *= i2; z3 *= i3;
    const T y3 = z3;
    const T y2 = z2 - l32 * y3;
    const T y1 = z1 - l21 * y2 - l31 * y3;
    const T y0 = z0 - l10 * y1 - l20 * y2 - l30 * y3;
    const T n2 = (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
    if (!(n2 > T(0)) || !(n2 < T(1e30))) return false;
    T inv = ses_rsqrt(n2);
    if ((y0 * x0 + y1 * x1) + (y2 * x2 + y3 * x3) < T(0)) inv = -inv;
    const T w0 = y0 * inv, w1 = y1 * inv, w2 = y2 * inv, w3 = y3 * inv;
    const T e0 = w0 - x0, e1 = w1 - x1, e2 = w2 - x2, e3 = w3 - x3;
    x0 = w0; x1 = w1; x2 = w2; x3 = w3;
    if ((e0 * e0 + e1 * e1) + (e2 * e2 + e3 * e3) < tol2) {
      v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3;
      return true;
    }
  }
  return false;
}

// Smallest eigenvector: inverse iteration, Jacobi when it declines. The out-of-line fallback works on copies: handing
// it the caller's G and v would make those arrays addressable and push them out of registers on the hot path too.
template <class T>
SES_HD void smallest_eigvec4_fast(const double G[10], T v[4]) {
  if (!invit4<T>(G, v)) {
    double Gc[10];
    T vc[4];
    for (int i = 0; i < 10; ++i) Gc[i] = G[i];
    smallest_eigvec4<T>(Gc, vc);
    v[0] = vc[0]; v[1] = vc[1]; v[2] = vc[2]; v[3] = vc[3];
  }
}

// Coordinates of a row r in the basis [v | Q]: s = r.v and p = Q^T r, with Q the Householder
// complement of the unit vector v built around the 4th axis (u = v + sign(v3) e3, never cancels).
template <class T>
SES_HD void deflate_project(const T v[4], const T r[4], T& s, T p[3]) {
  const T sg = v[3] >= T(0) ? T(1) : T(-1);
  const T kappa = ses_rcp(T(1) + ses_abs(v[3]));
  s = (r[0] * v[0] + r[1] * v[1]) + (r[2] * v[2] + r[3] * v[3]);
  const T c = kappa * (s + sg * r[3]);
  p[0] = r[0] - c * v[0]; p[1] = r[1] - c * v[1]; p[2] = r[2] - c * v[2];
}

// w = v + Q x  (back to the original basis)
template <class T>
SES_HD void deflate_expand(const T v[4], const T x[3], T w[4]) {
  const T sg = v[3] >= T(0) ? T(1) : T(-1);
  const T kappa = ses_rcp(T(1) + ses_abs(v[3]));
  const T c = kappa * (x[0] * v[0] + x[1] * v[1] + x[2] * v[2]);
  w[0] = v[0] + x[0] - c * v[0];
  w[1] = v[1] + x[1] - c * v[1];
  w[2] = v[2] + x[2] - c * v[2];
  w[3] = v[3] - c * (v[3] + sg);
}

// Smallest eigenpair of [[a, b^T], [b, C]] (C: 00 01 02 11 12 22) as (1, x); see the block comment above.
template <class T>
SES_HD bool secular_smallest(T a, const T b[3], const T C[6], T x[3]) {
  const T conv = sizeof(T) == 4 ? T(2e-6) : T(1e-13);
  T lam = a;
#if SES_UNROLL1
#pragma unroll 1
#endif
  for (int it = 0; it < 5; ++it) {
    const T m00 = C[0] - lam;
    if (!(m00 > T(0))) return false;
    const T i0 = ses_rcp(m00);
    const T l10 = C[1] * i0, l20 = C[2] * i0;
    const T m11 = C[3] - lam - l10 * C[1];
    if (!(m11 > T(0))) return false;
    const T i1 = ses_rcp(m11);
    const T t21 = C[4] - l20 * C[1];
    const T l21 = t21 * i1;
    const T m22 = C[5] - lam - l20 * C[2] - l21 * t21;
    if (!(m22 > T(0))) return false;
    const T i2 = ses_rcp(m22);
    T z0 = b[0];
    T z1 = b[1] - l10 * z0;
    T z2 = b[2] - l20 * z0 - l21 * z1;
    z0 *= i0; z1 *= i1; z2 *= i2;
    const T y2 = z2;
    const T y1 = z1 - l21 * y2;
    const T y0 = z0 - l10 * y1 - l20 * y2;
    x[0] = -y0; x[1] = -y1; x[2] = -y2;
    const T f = a - lam - (b[0] * y0 + b[1] * y1 + b[2] * y2);
    const T dl = f * ses_rcp(T(1) + (y0 * y0 + y1 * y1 + y2 * y2));
    // x changes by ~ |y| / m_min * |dl|: stop once that is below the working precision
    const T mmin = m00 < m11 ? (m00 < m22 ? m00 : m22) : (m11 < m22 ? m11 : m22);
    if (ses_abs(dl) <= conv * mmin) return true;
    lam += dl;
  }
  return false;
}

// projection residual of one view, S3D:430-433
template <class T>
SES_HD T reproj_residual(const T* P, const T X[3], T x, T y) {
  const T a = sum4(P[0] * X[0], P[1] * X[1], P[2] * X[2], P[3] * T(1));
  const T b = sum4(P[4] * X[0], P[5] * X[1], P[6] * X[2], P[7] * T(1));
  const T c = sum4(P[8] * X[0], P[9] * X[1], P[10] * X[2], P[11] * T(1));
  const T ic = T(1) / c;   // one IEEE reciprocal instead of two divides (<= 1 ulp on the projection)
  const T dx = a * ic - x, dy = b * ic - y;
  return ses_sqrt(dx * dx + dy * dy);
}

}  // namespace ses3d
