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

template <class T>
SES_HD void gram_add(double G[10], const T r[4], double sign) {
  const double a = (double)r[0], b = (double)r[1], c = (double)r[2], d = (double)r[3];
  const double sa = sign * a, sb = sign * b, sc = sign * c, sd = sign * d;
  G[0] += sa * a; G[1] += sa * b; G[2] += sa * c; G[3] += sa * d;
  G[4] += sb * b; G[5] += sb * c; G[6] += sb * d;
  G[7] += sc * c; G[8] += sc * d;
  G[9] += sd * d;
}

// Eigenvector of the smallest eigenvalue of the symmetric 4x4 matrix g (10 unique entries),
// cyclic Jacobi with the rotations of Rutishauser's formulation. Everything stays in registers.
template <class T>
SES_HD void smallest_eigvec4(const double g[10], T v[4]) {
  T a00 = (T)g[0], a01 = (T)g[1], a02 = (T)g[2], a03 = (T)g[3], a11 = (T)g[4], a12 = (T)g[5], a13 = (T)g[6],
    a22 = (T)g[7], a23 = (T)g[8], a33 = (T)g[9];
  T v00 = 1, v01 = 0, v02 = 0, v03 = 0, v10 = 0, v11 = 1, v12 = 0, v13 = 0, v20 = 0, v21 = 0, v22 = 1, v23 = 0,
    v30 = 0, v31 = 0, v32 = 0, v33 = 1;
  const T eps = sizeof(T) == 4 ? T(1.1920929e-7) : T(2.220446049250313e-16);
  const int max_sweeps = sizeof(T) == 4 ? 10 : 14;

// rotate the (p,q) plane; r,s are the two other indices. app/aqq/apq diagonal block,
// arp/arq and asp/asq the coupled off-diagonal entries, v?p/v?q eigenvector columns.
#define SES_ROT(app, aqq, apq, arp, arq, asp, asq, v0p, v0q, v1p, v1q, v2p, v2q, v3p, v3q)      \
  do {                                                                                          \
    const T apq_ = (apq);                                                                       \
    if (ses_abs(apq_) > eps * T(0.125) * ses_sqrt(ses_abs((app) * (aqq))) && apq_ != T(0)) {    \
      rotated = true;                                                                           \
      const T theta = ((aqq) - (app)) / (T(2) * apq_);                                          \
      const T t = (theta >= T(0) ? T(1) : T(-1)) / (ses_abs(theta) + ses_sqrt(theta * theta + T(1))); \
      const T c = T(1) / ses_sqrt(t * t + T(1));                                                \
      const T s = t * c;                                                                        \
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
    SES_ROT(a00, a11, a01, a02, a12, a03, a13, v00, v01, v10, v11, v20, v21, v30, v31);  // (0,1): others 2,3
    SES_ROT(a00, a22, a02, a01, a12, a03, a23, v00, v02, v10, v12, v20, v22, v30, v32);  // (0,2): others 1,3
    SES_ROT(a00, a33, a03, a01, a13, a02, a23, v00, v03, v10, v13, v20, v23, v30, v33);  // (0,3): others 1,2
    SES_ROT(a11, a22, a12, a01, a02, a13, a23, v01, v02, v11, v12, v21, v22, v31, v32);  // (1,2): others 0,3
    SES_ROT(a11, a33, a13, a01, a03, a12, a23, v01, v03, v11, v13, v21, v23, v31, v33);  // (1,3): others 0,2
    SES_ROT(a22, a33, a23, a02, a03, a12, a13, v02, v03, v12, v13, v22, v23, v32, v33);  // (2,3): others 0,1
    if (!rotated) break;
  }
#undef SES_ROT
  // smallest diagonal entry -> its eigenvector column
  T best = a00;
  v[0] = v00; v[1] = v10; v[2] = v20; v[3] = v30;
  if (a11 < best) { best = a11; v[0] = v01; v[1] = v11; v[2] = v21; v[3] = v31; }
  if (a22 < best) { best = a22; v[0] = v02; v[1] = v12; v[2] = v22; v[3] = v32; }
  if (a33 < best) { best = a33; v[0] = v03; v[1] = v13; v[2] = v23; v[3] = v33; }
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
