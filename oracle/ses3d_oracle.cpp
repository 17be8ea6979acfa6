// ses3d_oracle.cpp — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
//
// A dependency-free restatement of the reference's per-frame multi-view
// geometry path, used as the checker for the CUDA library and as the timed
// CPU baseline. Nothing in the product (smartedgesensor3dhumanpose_b200/, the
// C-ABI library) may include, link or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Reference spans restated (S3D = skeleton_3d/src/skeleton_3d_triang_mult_node.cpp,
// REP = pose_reprojection/src/skeleton_reproj_mult_node.cpp, HUN = skeleton_3d/src/Hungarian.cpp):
//   set-up tables            S3D:230-253, 1187-1211
//   normalize_keypoints      S3D:312-333
//   calcCost                 S3D:335-390
//   association loop         S3D:528-674
//   Munkres                  HUN:60-397 (restated iteratively; the verbatim file can be
//                            swapped in through oracle_use_ref_hungarian(), see oracle/Makefile)
//   triangulate / reproj err S3D:425-465
//   outlier rejection        S3D:745-844
//   UT covariance            S3D:471-523
//   skeleton plausibility    S3D:861-973
//   merge                    S3D:392-423, 984-996
//   reprojection             REP:62-75, 139-235
//
// PARITY STATUS: the reference has no tests or golden vectors, and its node cannot be
// compiled here (ROS, Eigen, image_geometry, tf2 absent). Pinned against the real
// reference: the Munkres solver (verbatim Hungarian.cpp built into oracle/_ref). Everything
// that goes through Eigen (JacobiSVD, fixed-size products, llt) or image_geometry is a
// restatement of the published algorithm: PARITY UNPINNED at those third-party boundaries.
//   * Eigen::JacobiSVD (S3D:237,456) -> one-sided Jacobi SVD below (same singular vectors up
//     to rounding).
//   * Eigen fixed-size 3-vector reductions (S3D:357-361, 766-771) evaluate as
//     x0 + (x1 + x2) (Redux.h halves the range); 4-vector as (x0+x1)+(x2+x3). Stated from
//     knowledge of Eigen 3.3; not verifiable offline.
//   * image_geometry::project3dToPixel -> u = (fx*X + Tx)/Z + cx, v = (fy*Y + Ty)/Z + cy.
//
// The float/double split and the evaluation order of the reference are kept. Build with
// -ffp-contract=off and without -ffast-math / -march=native (reference = x86-64 baseline).

#include "ses3d.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// constants / skeleton tables (S3D:43-64, 81-149; REP:47-53)
// ----------------------------------------------------------------------------
const double MAX_COSTS = 1e6;  // S3D:43
const int NKP = SES3D_NUM_KEYPOINTS;
const int NFUS = SES3D_NUM_FUSION_KEYPOINTS;

struct SkeletonModel {
  int parent[17];
  double limb_len[17];
  double limb_sigma[17];
  int fusion_idx[17];
};

// EdgeTPU_BodyParts_Simple S3D:81-104, g_kp2kpFusion_idx_simple S3D:139-142
const SkeletonModel kSimple = {
    {-1, 0, 0, 1, 2, 0, 0, 5, 6, 7, 8, 5, 6, 11, 12, 13, 14},
    {-1, 0.05, 0.05, 0.10, 0.10, -1, -1, 0.28, 0.28, 0.25, 0.25, 0.50, 0.50, 0.45, 0.45, 0.446, 0.446},
    {-1, 0.05, 0.05, 0.05, 0.05, -1, -1, 0.10, 0.10, 0.10, 0.10, 0.15, 0.15, 0.10, 0.10, 0.10, 0.10},
    {0, 16, 15, 18, 17, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11}};
const double kShoulderDist = 0.35, kShoulderSigma = 0.15;  // S3D:103
const int kSimpleRShoulder = 6, kSimpleLShoulder = 5;      // S3D:83,86

// EdgeTPU_BodyParts_H36M S3D:111-133, g_kp2kpFusion_idx_h36m S3D:143-145
const SkeletonModel kH36M = {
    {-1, 0, 0, 2, 3, 2, 2, 5, 6, 7, 8, 4, 4, 11, 12, 13, 14},
    {-1, 0.115, 0.116, 0.255, 0.238, 0.149, 0.149, 0.28, 0.28, 0.25, 0.25, 0.134, 0.134, 0.449, 0.449, 0.446, 0.446},
    {-1, 0.07, 0.07, 0.15, 0.15, 0.10, 0.10, 0.15, 0.15, 0.15, 0.15, 0.10, 0.10, 0.20, 0.20, 0.20, 0.20},
    {0, 19, 1, 20, 8, 5, 2, 6, 3, 7, 4, 12, 9, 13, 10, 14, 11}};

// ----------------------------------------------------------------------------
// small linear algebra
// ----------------------------------------------------------------------------

// One-sided (Hestenes) Jacobi SVD of a rows x cols matrix B (row-major, cols <= 4,
// rows >= cols): on return the columns of B are U*diag(sigma) and W (cols x cols,
// row-major) holds the right singular vectors. Stands in for Eigen::JacobiSVD.
template <class T>
void onesided_jacobi(T* B, int rows, int cols, T* W) {
  for (int i = 0; i < cols; ++i)
    for (int j = 0; j < cols; ++j) W[i * cols + j] = (i == j) ? T(1) : T(0);
  const T eps = std::numeric_limits<T>::epsilon();
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < cols - 1; ++p) {
      for (int q = p + 1; q < cols; ++q) {
        T alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < rows; ++r) {
          const T bp = B[r * cols + p], bq = B[r * cols + q];
          alpha += bp * bp;
          beta += bq * bq;
          gamma += bp * bq;
        }
        if (std::fabs(gamma) <= eps * std::sqrt(alpha * beta) || gamma == T(0)) continue;
        rotated = true;
        const T zeta = (beta - alpha) / (T(2) * gamma);
        const T t = (zeta >= T(0) ? T(1) : T(-1)) / (std::fabs(zeta) + std::sqrt(T(1) + zeta * zeta));
        const T c = T(1) / std::sqrt(T(1) + t * t);
        const T s = c * t;
        for (int r = 0; r < rows; ++r) {
          const T bp = B[r * cols + p], bq = B[r * cols + q];
          B[r * cols + p] = c * bp - s * bq;
          B[r * cols + q] = s * bp + c * bq;
        }
        for (int r = 0; r < cols; ++r) {
          const T wp = W[r * cols + p], wq = W[r * cols + q];
          W[r * cols + p] = c * wp - s * wq;
          W[r * cols + q] = s * wp + c * wq;
        }
      }
    }
    if (!rotated) break;
  }
}

// ----------------------------------------------------------------------------
// Second SVD variant: restatement of Eigen 3.3 JacobiSVD<Matrix<T,-1,4>, ColPivHouseholderQRPreconditioner>
// ::compute(A, ComputeThinV) as triangulate() calls it (S3D:456). Eigen is absent from this image, so this follows
// the published algorithm of Eigen 3.3.x (Ubuntu 18.04 / ROS melodic ships 3.3.4):
//   JacobiSVD.h   compute(): scale by max|a_ij|; rows != cols -> R-SVD step (column-pivoting Householder QR, work
//                 matrix = upper-triangular R, V = column permutation); two-sided Jacobi sweeps p = 1..3, q = 0..p-1
//                 with threshold max(min_normal, 2 eps * maxDiagEntry); real_2x2_jacobi_svd; singular values = |diag|,
//                 sorted descending with column swaps of V.
//   ColPivHouseholderQR.h computeInPlace(): pivot = first column of largest updated norm, LAPACK xGEQPF norm down-date.
//   Householder.h makeHouseholder() / applyHouseholderOnTheLeft(), Jacobi.h makeJacobi() / rotation products.
// What cannot be restated offline: Eigen's vectorised reduction orders for dynamic-size columns (SSE packets,
// alignment dependent) - sums over a column run sequentially here. The variant therefore tracks the reference's
// algorithm (pivoting, rotation order, thresholds), not its last bit. It is used to QUANTIFY how far the primary
// oracle (one-sided Hestenes Jacobi above) can sit from the reference's arithmetic at S3D:456: see DESIGN.md 2 and
// tests/test_oracle.py::test_svd_variants_agree.
// ----------------------------------------------------------------------------
template <class T>
struct JRot { T c, s; };

template <class T>
inline JRot<T> jrot_mul(const JRot<T>& a, const JRot<T>& b) {   // JacobiRotation::operator*
  return {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}

// JacobiRotation::makeJacobi(x, y, z) for the symmetric 2x2 [[x, y], [y, z]]
template <class T>
inline JRot<T> make_jacobi(T x, T y, T z) {
  const T deno = T(2) * std::fabs(y);
  if (deno < std::numeric_limits<T>::min()) return {T(1), T(0)};
  const T tau = (x - z) / deno;
  const T w = std::sqrt(tau * tau + T(1));
  const T t = tau > T(0) ? T(1) / (tau + w) : T(1) / (tau - w);
  const T sign_t = t > T(0) ? T(1) : T(-1);
  const T n = T(1) / std::sqrt(t * t + T(1));
  return {n, -sign_t * (y / std::fabs(y)) * std::fabs(t) * n};
}

// A: rows x 4 row-major (rows >= 4). V: 4x4 row-major, columns ordered by descending singular value; sv[4].
template <class T>
void eigen_jacobi_svd_thinV(const T* A, int rows, T* V, T* sv) {
  const int n = 4;
  const T tiny = std::numeric_limits<T>::min();
  T scale = T(0);
  for (int i = 0; i < rows * n; ++i) { const T a = std::fabs(A[i]); if (a > scale) scale = a; }   // cwiseAbs().maxCoeff()
  if (scale == T(0)) scale = T(1);
  T W[16];  // work matrix, row-major W[r*4+c]
  for (int i = 0; i < 16; ++i) V[i] = T(0);
  if (rows != n) {
    // ---- R-SVD step: ColPivHouseholderQR of A / scale
    std::vector<T> M((size_t)rows * n);   // column-major M[c*rows + r]
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < n; ++c) M[(size_t)c * rows + r] = A[(size_t)r * n + c] / scale;
    T norm_upd[4], norm_dir[4];
    int transp[4];
    for (int k = 0; k < n; ++k) {
      T s2 = T(0);
      for (int r = 0; r < rows; ++r) s2 += M[(size_t)k * rows + r] * M[(size_t)k * rows + r];
      norm_dir[k] = norm_upd[k] = std::sqrt(s2);
    }
    const T downdate_thr = std::sqrt(std::numeric_limits<T>::epsilon());
    for (int k = 0; k < n; ++k) {
      int big = k;
      for (int j = k + 1; j < n; ++j)
        if (norm_upd[j] > norm_upd[big]) big = j;      // first maximum
      transp[k] = big;
      if (big != k) {
        for (int r = 0; r < rows; ++r) std::swap(M[(size_t)k * rows + r], M[(size_t)big * rows + r]);
        std::swap(norm_upd[k], norm_upd[big]);
        std::swap(norm_dir[k], norm_dir[big]);
      }
      // makeHouseholderInPlace on M(k.., k)
      T tail2 = T(0);
      for (int r = k + 1; r < rows; ++r) tail2 += M[(size_t)k * rows + r] * M[(size_t)k * rows + r];
      const T c0 = M[(size_t)k * rows + k];
      T tau, beta;
      if (tail2 <= tiny) {
        tau = T(0);
        beta = c0;
        for (int r = k + 1; r < rows; ++r) M[(size_t)k * rows + r] = T(0);
      } else {
        beta = std::sqrt(c0 * c0 + tail2);
        if (c0 >= T(0)) beta = -beta;
        for (int r = k + 1; r < rows; ++r) M[(size_t)k * rows + r] /= (c0 - beta);
        tau = (beta - c0) / beta;
      }
      M[(size_t)k * rows + k] = beta;
      // applyHouseholderOnTheLeft on the bottom-right corner (rows-k) x (n-k-1)
      if (rows - k == 1) {
        for (int j = k + 1; j < n; ++j) M[(size_t)j * rows + k] *= T(1) - tau;
      } else if (tau != T(0)) {
        for (int j = k + 1; j < n; ++j) {
          T tmp = T(0);
          for (int r = k + 1; r < rows; ++r) tmp += M[(size_t)k * rows + r] * M[(size_t)j * rows + r];
          tmp += M[(size_t)j * rows + k];
          M[(size_t)j * rows + k] -= tau * tmp;
          for (int r = k + 1; r < rows; ++r) M[(size_t)j * rows + r] -= (tau * M[(size_t)k * rows + r]) * tmp;
        }
      }
      // norm down-date (LAPACK xGEQPF)
      for (int j = k + 1; j < n; ++j) {
        if (norm_upd[j] != T(0)) {
          T temp = std::fabs(M[(size_t)j * rows + k]) / norm_upd[j];
          temp = (T(1) + temp) * (T(1) - temp);
          temp = temp < T(0) ? T(0) : temp;
          const T q = norm_upd[j] / norm_dir[j];
          const T temp2 = temp * (q * q);
          if (temp2 <= downdate_thr) {
            T s2 = T(0);
            for (int r = k + 1; r < rows; ++r) s2 += M[(size_t)j * rows + r] * M[(size_t)j * rows + r];
            norm_dir[j] = norm_upd[j] = std::sqrt(s2);
          } else {
            norm_upd[j] *= std::sqrt(temp);
          }
        }
      }
    }
    int perm[4] = {0, 1, 2, 3};
    for (int k = 0; k < n; ++k) std::swap(perm[k], perm[transp[k]]);   // applyTranspositionOnTheRight
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) W[r * n + c] = c >= r ? M[(size_t)c * rows + r] : T(0);   // triangularView<Upper>
    for (int j = 0; j < n; ++j) V[perm[j] * n + j] = T(1);   // m_matrixV = colsPermutation
  } else {
    for (int i = 0; i < 16; ++i) W[i] = A[i] / scale;
    for (int i = 0; i < n; ++i) V[i * n + i] = T(1);
  }
  // ---- two-sided Jacobi sweeps
  const T precision = T(2) * std::numeric_limits<T>::epsilon();
  T max_diag = T(0);
  for (int i = 0; i < n; ++i) max_diag = std::max(max_diag, std::fabs(W[i * n + i]));
  bool finished = false;
  int guard = 0;
  while (!finished && ++guard < 1000) {
    finished = true;
    for (int p = 1; p < n; ++p)
      for (int q = 0; q < p; ++q) {
        const T threshold = std::max(tiny, precision * max_diag);
        if (std::fabs(W[p * n + q]) > threshold || std::fabs(W[q * n + p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd
          T m00 = W[p * n + p], m01 = W[p * n + q], m10 = W[q * n + p], m11 = W[q * n + q];
          JRot<T> rot1;
          const T t = m00 + m11, d = m10 - m01;
          if (std::fabs(d) < tiny) {
            rot1 = {T(1), T(0)};
          } else {
            const T u = t / d;
            const T tmp = std::sqrt(T(1) + u * u);
            rot1 = {u / tmp, T(1) / tmp};
          }
          {  // m.applyOnTheLeft(0, 1, rot1)
            const T x0 = m00, y0 = m10, x1 = m01, y1 = m11;
            m00 = rot1.c * x0 + rot1.s * y0; m10 = -rot1.s * x0 + rot1.c * y0;
            m01 = rot1.c * x1 + rot1.s * y1; m11 = -rot1.s * x1 + rot1.c * y1;
          }
          const JRot<T> j_right = make_jacobi<T>(m00, m01, m11);
          const JRot<T> j_left = jrot_mul<T>(rot1, JRot<T>{j_right.c, -j_right.s});
          for (int i = 0; i < n; ++i) {  // applyOnTheLeft(p, q, j_left): rows p, q
            const T x = W[p * n + i], y = W[q * n + i];
            W[p * n + i] = j_left.c * x + j_left.s * y;
            W[q * n + i] = -j_left.s * x + j_left.c * y;
          }
          for (int i = 0; i < n; ++i) {  // applyOnTheRight(p, q, j_right): columns p, q with the transposed rotation
            const T x = W[i * n + p], y = W[i * n + q];
            W[i * n + p] = j_right.c * x - j_right.s * y;
            W[i * n + q] = j_right.s * x + j_right.c * y;
          }
          for (int i = 0; i < n; ++i) {
            const T x = V[i * n + p], y = V[i * n + q];
            V[i * n + p] = j_right.c * x - j_right.s * y;
            V[i * n + q] = j_right.s * x + j_right.c * y;
          }
          max_diag = std::max(max_diag, std::max(std::fabs(W[p * n + p]), std::fabs(W[q * n + q])));
        }
      }
  }
  for (int i = 0; i < n; ++i) sv[i] = std::fabs(W[i * n + i]) * scale;
  for (int i = 0; i < n; ++i) {   // descending order, column swaps of V
    int pos = i;
    for (int j = i + 1; j < n; ++j)
      if (sv[j] > sv[pos]) pos = j;
    if (sv[pos] == T(0)) break;
    if (pos != i) {
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < n; ++r) std::swap(V[r * n + i], V[r * n + pos]);
    }
  }
}

// pseudo_inv34d S3D:236-240: pinv of a 3x4 (row-major M[12]) -> 4x3 (row-major out[12]).
void pseudo_inv34(const double* M, double* out) {
  // SVD of M^T (4x3): M^T = Q diag(s) W^T  =>  M = W diag(s) Q^T,  pinv(M) = Q diag(1/s) W^T
  double B[12], W[9];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 3; ++c) B[r * 3 + c] = M[c * 4 + r];
  onesided_jacobi<double>(B, 4, 3, W);
  double s[3], smax = 0;
  for (int c = 0; c < 3; ++c) {
    double n = 0;
    for (int r = 0; r < 4; ++r) n += B[r * 3 + c] * B[r * 3 + c];
    s[c] = std::sqrt(n);
    smax = std::max(smax, s[c]);
  }
  const double tol = std::numeric_limits<double>::epsilon() * 4.0 * smax;  // eps * max(cols, rows) * sigma_0
  for (int i = 0; i < 12; ++i) out[i] = 0.0;
  for (int c = 0; c < 3; ++c) {
    if (!(std::fabs(s[c]) > tol)) continue;
    // Q column c = B column c / s[c];  contribution Q_c * (1/s_c) * W_c^T
    const double inv2 = 1.0 / (s[c] * s[c]);
    for (int r = 0; r < 4; ++r)
      for (int k = 0; k < 3; ++k) out[r * 3 + k] += B[r * 3 + c] * inv2 * W[k * 3 + c];
  }
}

// get_fundamental_idx S3D:242-253 (NUM_CAMERAS is unsigned there; i,j are non-negative here)
int fundamental_idx(int i, int j, int n_cams) {
  if (i >= j) return -1;
  if (i > n_cams - 2 || j > n_cams - 1) return -1;
  int start = 0;
  for (int ii = 0; ii < i; ++ii) start += n_cams - ii - 1;
  return start + j - i - 1;
}

// Eigen fixed-size reduction orders (see header comment)
template <class T>
inline T sum3(T a, T b, T c) { return a + (b + c); }
template <class T>
inline T sum4(T a, T b, T c, T d) { return (a + b) + (c + d); }

struct Tables {
  int n_cams = 0;
  ses3d_params prm;
  const SkeletonModel* model = &kSimple;
  std::vector<double> Pd;  // [C][12] row-major double
  std::vector<float> Pf;   // [C][12] cast to float (camera_matrices S3D:1208-1211)
  std::vector<float> F;    // [C(C-1)/2][9] row-major float (S3D:1195-1204)
  std::vector<ses3d_camera> cams;
  int svd_variant = 0;     // 0: Hestenes one-sided Jacobi (primary), 1: Eigen JacobiSVD restatement
  void* ref_hungarian_lib = nullptr;
  void (*ref_hungarian)(int*, double*, double*, int, int) = nullptr;
};

// S3D:1187-1211
void build_tables(Tables& tb) {
  const int C = tb.n_cams;
  tb.Pd.resize((size_t)C * 12);
  tb.Pf.resize((size_t)C * 12);
  std::vector<double> centres((size_t)C * 4);
  for (int i = 0; i < C; ++i) {
    const double* T = tb.cams[i].T_cam_base;
    for (int k = 0; k < 12; ++k) {
      tb.Pd[(size_t)i * 12 + k] = T[k];
      tb.Pf[(size_t)i * 12 + k] = static_cast<float>(T[k]);
    }
    // camera centre = inverse(T).col(3): general 3x3 inverse of the linear part (Affine3d::inverse)
    const double a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], k = T[10];
    const double A = e * k - f * h, B = -(d * k - f * g), Cc = d * h - e * g;
    const double det = a * A + b * B + c * Cc;
    const double inv[9] = {A / det, -(b * k - c * h) / det, (b * f - c * e) / det,
                           B / det, (a * k - c * g) / det, -(a * f - c * d) / det,
                           Cc / det, -(a * h - b * g) / det, (a * e - b * d) / det};
    const double t[3] = {T[3], T[7], T[11]};
    for (int r = 0; r < 3; ++r)
      centres[(size_t)i * 4 + r] = -(inv[r * 3 + 0] * t[0] + inv[r * 3 + 1] * t[1] + inv[r * 3 + 2] * t[2]);
    centres[(size_t)i * 4 + 3] = 1.0;
  }
  tb.F.assign((size_t)C * (C - 1) / 2 * 9, 0.f);
  size_t idx = 0;
  for (int i = 0; i < C; ++i) {
    double Pinv[12];
    pseudo_inv34(&tb.Pd[(size_t)i * 12], Pinv);
    for (int j = i + 1; j < C; ++j, ++idx) {
      const double* Pj = &tb.Pd[(size_t)j * 12];
      const double* Ci = &centres[(size_t)i * 4];
      double e[3];
      for (int r = 0; r < 3; ++r)
        e[r] = sum4(Pj[r * 4 + 0] * Ci[0], Pj[r * 4 + 1] * Ci[1], Pj[r * 4 + 2] * Ci[2], Pj[r * 4 + 3] * Ci[3]);
      const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};  // cross_prod_matrix S3D:230-234
      double M[12];  // ex * Pj  (3x4)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
          M[r * 4 + c] = sum3(ex[r * 3 + 0] * Pj[0 * 4 + c], ex[r * 3 + 1] * Pj[1 * 4 + c], ex[r * 3 + 2] * Pj[2 * 4 + c]);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          const double v = sum4(M[r * 4 + 0] * Pinv[0 * 3 + c], M[r * 4 + 1] * Pinv[1 * 3 + c],
                                M[r * 4 + 2] * Pinv[2 * 3 + c], M[r * 4 + 3] * Pinv[3 * 3 + c]);
          tb.F[idx * 9 + r * 3 + c] = static_cast<float>(v);
        }
    }
  }
}

// ----------------------------------------------------------------------------
// Munkres, restated from HUN:60-397 as an explicit state machine (the reference
// recurses step2a/2b/3/4/5). Column-major dist[row + nRows*col]; scan orders kept.
// ----------------------------------------------------------------------------
void munkres(int* assignment, double* cost, const double* dist_in, int nR, int nC) {
  const int nE = nR * nC;
  std::vector<double> dist(dist_in, dist_in + nE);
  std::vector<char> star(nE, 0), prime(nE, 0), new_star(nE, 0), cov_c(nC, 0), cov_r(nR, 0);
  *cost = 0;
  for (int r = 0; r < nR; ++r) assignment[r] = -1;
  int min_dim;
  if (nR <= nC) {  // HUN:95-131
    min_dim = nR;
    for (int r = 0; r < nR; ++r) {
      double mn = dist[r];
      for (int c = 1; c < nC; ++c) mn = (dist[r + nR * c] < mn) ? dist[r + nR * c] : mn;
      for (int c = 0; c < nC; ++c) dist[r + nR * c] -= mn;
    }
    for (int r = 0; r < nR; ++r)
      for (int c = 0; c < nC; ++c)
        if (std::fabs(dist[r + nR * c]) < DBL_EPSILON && !cov_c[c]) {
          star[r + nR * c] = 1;
          cov_c[c] = 1;
          break;
        }
  } else {  // HUN:132-170
    min_dim = nC;
    for (int c = 0; c < nC; ++c) {
      double mn = dist[nR * c];
      for (int r = 1; r < nR; ++r) mn = (dist[r + nR * c] < mn) ? dist[r + nR * c] : mn;
      for (int r = 0; r < nR; ++r) dist[r + nR * c] -= mn;
    }
    for (int c = 0; c < nC; ++c)
      for (int r = 0; r < nR; ++r)
        if (std::fabs(dist[r + nR * c]) < DBL_EPSILON && !cov_r[r]) {
          star[r + nR * c] = 1;
          cov_c[c] = 1;
          cov_r[r] = 1;
          break;
        }
    for (int r = 0; r < nR; ++r) cov_r[r] = 0;
  }

  enum { S2A, S2B, S3, S4, S5, DONE } st = S2B;
  int row4 = 0, col4 = 0;
  while (st != DONE) {
    switch (st) {
      case S2A:  // HUN:222-242
        for (int c = 0; c < nC; ++c)
          for (int r = 0; r < nR; ++r)
            if (star[r + nR * c]) { cov_c[c] = 1; break; }
        st = S2B;
        break;
      case S2B: {  // HUN:245-266
        int n = 0;
        for (int c = 0; c < nC; ++c) n += cov_c[c] ? 1 : 0;
        st = (n == min_dim) ? DONE : S3;
        break;
      }
      case S3: {  // HUN:269-309
        bool zeros = true, to4 = false;
        while (zeros && !to4) {
          zeros = false;
          for (int c = 0; c < nC && !to4; ++c) {
            if (cov_c[c]) continue;
            for (int r = 0; r < nR; ++r) {
              if (!cov_r[r] && std::fabs(dist[r + nR * c]) < DBL_EPSILON) {
                prime[r + nR * c] = 1;
                int sc = 0;
                for (; sc < nC; ++sc)
                  if (star[r + nR * sc]) break;
                if (sc == nC) {
                  row4 = r; col4 = c; to4 = true;
                } else {
                  cov_r[r] = 1;
                  cov_c[sc] = 0;
                  zeros = true;
                }
                break;
              }
            }
          }
        }
        st = to4 ? S4 : S5;
        break;
      }
      case S4: {  // HUN:312-363
        new_star = star;
        new_star[row4 + nR * col4] = 1;
        int sc = col4, sr = 0;
        for (sr = 0; sr < nR; ++sr)
          if (star[sr + nR * sc]) break;
        while (sr < nR) {
          new_star[sr + nR * sc] = 0;
          const int pr = sr;
          int pc = 0;
          for (; pc < nC; ++pc)
            if (prime[pr + nR * pc]) break;
          new_star[pr + nR * pc] = 1;
          sc = pc;
          for (sr = 0; sr < nR; ++sr)
            if (star[sr + nR * sc]) break;
        }
        std::fill(prime.begin(), prime.end(), 0);
        star = new_star;
        std::fill(cov_r.begin(), cov_r.end(), 0);
        st = S2A;
        break;
      }
      case S5: {  // HUN:366-397
        double h = DBL_MAX;
        for (int r = 0; r < nR; ++r)
          if (!cov_r[r])
            for (int c = 0; c < nC; ++c)
              if (!cov_c[c] && dist[r + nR * c] < h) h = dist[r + nR * c];
        for (int r = 0; r < nR; ++r)
          if (cov_r[r])
            for (int c = 0; c < nC; ++c) dist[r + nR * c] += h;
        for (int c = 0; c < nC; ++c)
          if (!cov_c[c])
            for (int r = 0; r < nR; ++r) dist[r + nR * c] -= h;
        st = S3;
        break;
      }
      default: break;
    }
  }
  for (int r = 0; r < nR; ++r)  // buildassignmentvector HUN:190-205
    for (int c = 0; c < nC; ++c)
      if (star[r + nR * c]) { assignment[r] = c; break; }
  for (int r = 0; r < nR; ++r)  // computeassignmentcost HUN:208-219
    if (assignment[r] >= 0) *cost += dist_in[r + nR * assignment[r]];
}

// ----------------------------------------------------------------------------
// association (float/double split as in the reference)
// ----------------------------------------------------------------------------
struct NormPerson {      // one detection after normalize_keypoints
  float kp[17][3];       // x, y, conf; (0,0,-1) when below threshold (S3D:575,595)
  float cov[17][3];      // xx, xy, yy in normalised coordinates
  double kpd[17][2];     // FP64 variant only: the same normalisation carried out in double
  double covd[17][3];
  float score;           // Person2D.score
  int cam, det;          // camera id, detection slot in the input
};

// normalize_keypoints S3D:312-333; returns the number of valid keypoints
int normalize_person(const ses3d_person2d& in, const ses3d_camera& cam, float thr, NormPerson& out) {
  const float fx = static_cast<float>(cam.fx), fy = static_cast<float>(cam.fy);
  const float cx = static_cast<float>(cam.cx), cy = static_cast<float>(cam.cy);
  int n_valid = 0;
  for (int k = 0; k < NKP; ++k) {
    out.kp[k][0] = 0.f; out.kp[k][1] = 0.f; out.kp[k][2] = -1.f;
    out.cov[k][0] = out.cov[k][1] = out.cov[k][2] = 0.f;
    out.kpd[k][0] = out.kpd[k][1] = 0.0;
    out.covd[k][0] = out.covd[k][1] = out.covd[k][2] = 0.0;
    const ses3d_keypoint2d& kp = in.keypoints[k];
    if (kp.score >= thr) {
      out.kp[k][0] = (kp.x - cx) / fx;
      out.kp[k][1] = (kp.y - cy) / fy;
      out.kp[k][2] = kp.score;
      out.cov[k][0] = kp.cov[0] / (fx * fx);
      out.cov[k][1] = kp.cov[1] / (fx * fy);
      out.cov[k][2] = kp.cov[2] / (fy * fy);
      out.kpd[k][0] = ((double)kp.x - cam.cx) / cam.fx;
      out.kpd[k][1] = ((double)kp.y - cam.cy) / cam.fy;
      out.covd[k][0] = (double)kp.cov[0] / (cam.fx * cam.fx);
      out.covd[k][1] = (double)kp.cov[1] / (cam.fx * cam.fy);
      out.covd[k][2] = (double)kp.cov[2] / (cam.fy * cam.fy);
      ++n_valid;
    }
  }
  out.score = in.score;
  return n_valid;
}

struct Hypothesis { std::vector<NormPerson> obs; };  // PersonHypothesis S3D:153-159 (obs in camera order)

// symmetric point-to-epipolar-line distance d1 + d2, S3D:355-362
inline float epipolar_symmetric(const float* F, float x1, float y1, float x2, float y2) {
  // l1 = F * (x1,y1,1);  l2 = F^T * (x2,y2,1)
  const float l1x = sum3(F[0] * x1, F[1] * y1, F[2] * 1.0f);
  const float l1y = sum3(F[3] * x1, F[4] * y1, F[5] * 1.0f);
  const float l1z = sum3(F[6] * x1, F[7] * y1, F[8] * 1.0f);
  const float l2x = sum3(F[0] * x2, F[3] * y2, F[6] * 1.0f);
  const float l2y = sum3(F[1] * x2, F[4] * y2, F[7] * 1.0f);
  const float l2z = sum3(F[2] * x2, F[5] * y2, F[8] * 1.0f);
  const float d1 = std::fabs(sum3(x2 * l1x, y2 * l1y, 1.0f * l1z)) / std::sqrt(l1x * l1x + l1y * l1y);
  const float d2 = std::fabs(sum3(x1 * l2x, y1 * l2y, 1.0f * l2z)) / std::sqrt(l2x * l2x + l2y * l2y);
  return d1 + d2;
}

// calcCost S3D:335-390
double calc_cost(const Tables& tb, const Hypothesis& hyp, const NormPerson& det, int det_cam, bool& veto) {
  const float thr = tb.prm.triangulation_threshold;
  const double max_epi = tb.prm.max_epipolar_error;
  double total_cost = 0.;
  int n_obs_used = 0;
  const int n_obs = (int)hyp.obs.size();
  if (n_obs == 0) { veto = true; return MAX_COSTS; }
  veto = false;
  double tmp_veto = 0.0;
  const double tolerance = 1.0 - 1.0 / (2 * n_obs), veto_delta = 1.0 / n_obs;
  for (int o = 0; o < n_obs; ++o) {
    double cost = 0.;
    int n_joints = 0;
    const NormPerson& ob = hyp.obs[o];
    const float* F = &tb.F[(size_t)fundamental_idx(ob.cam, det_cam, tb.n_cams) * 9];
    for (int k = 0; k < NKP; ++k) {
      if (ob.kp[k][2] > thr && det.kp[k][2] > thr) {
        cost += static_cast<double>(epipolar_symmetric(F, ob.kp[k][0], ob.kp[k][1], det.kp[k][0], det.kp[k][1]));
        ++n_joints;
      }
    }
    if (n_joints > 0) {
      cost /= n_joints;
      total_cost += cost;
      ++n_obs_used;
      if (cost > max_epi && (ob.score > 0.5f || n_obs == 1)) tmp_veto += veto_delta;
      else if (cost > 2 * max_epi && (ob.score > 0.5f || n_obs == 1)) tmp_veto += 1;  // unreachable, kept
    }
  }
  if (tmp_veto > tolerance) veto = true;
  if (n_obs_used > 0) return total_cost / n_obs_used;
  veto = true;
  return MAX_COSTS;
}

struct AssocResult {
  std::vector<Hypothesis> H;
  int n_hungarian = 0;
};

// S3D:528-674
void associate(const Tables& tb, int p_max, const ses3d_person2d* persons /*[C][p_max]*/, const int32_t* n_persons,
               AssocResult& res) {
  const int C = tb.n_cams;
  const float thr = tb.prm.triangulation_threshold;
  std::vector<int> cams;  // cameras with >= 1 detection (S3D:538-555)
  for (int i = 0; i < C; ++i)
    if (n_persons[i] > 0) cams.push_back(i);
  res.H.clear();
  res.n_hungarian = 0;
  if ((int)cams.size() < 2) return;  // S3D:557-560

  auto valid_dets = [&](int cam, std::vector<NormPerson>& out) {
    out.clear();
    for (int d = 0; d < n_persons[cam]; ++d) {
      NormPerson np;
      np.cam = cam; np.det = d;
      const int nv = normalize_person(persons[(size_t)cam * p_max + d], tb.cams[cam], thr, np);
      if (nv > NKP / 2) out.push_back(np);  // S3D:579,599
    }
  };

  std::vector<Hypothesis>& H = res.H;
  size_t ci = 0;
  std::vector<NormPerson> dets;
  while (H.empty() && ci < cams.size()) {  // S3D:567-586
    valid_dets(cams[ci], dets);
    for (const NormPerson& np : dets) { Hypothesis h; h.obs.push_back(np); H.push_back(h); }
    ++ci;
  }
  for (; ci < cams.size(); ++ci) {  // S3D:588-674
    const int cam = cams[ci];
    valid_dets(cam, dets);
    const int n_hyp = (int)H.size(), n_det = (int)dets.size();
    if (n_det == 0) continue;
    std::vector<double> Cm((size_t)n_hyp * n_det);  // column-major (S3D:611)
    std::vector<int> assignment(n_hyp, -1);
    std::vector<char> mask((size_t)n_hyp * n_det, 0);
    for (int d = 0; d < n_det; ++d)
      for (int h = 0; h < n_hyp; ++h) {
        bool veto;
        const double c = calc_cost(tb, H[h], dets[d], cam, veto);
        Cm[h + (size_t)n_hyp * d] = c;
        if (!veto && c < tb.prm.max_epipolar_error) { mask[h + (size_t)n_hyp * d] = 1; assignment[h] = d; }
      }
    bool ambiguous = false;  // S3D:628
    for (int d = 0; d < n_det && !ambiguous; ++d) {
      int n = 0;
      for (int h = 0; h < n_hyp; ++h) n += mask[h + (size_t)n_hyp * d];
      ambiguous = n > 1;
    }
    for (int h = 0; h < n_hyp && !ambiguous; ++h) {
      int n = 0;
      for (int d = 0; d < n_det; ++d) n += mask[h + (size_t)n_hyp * d];
      ambiguous = n > 1;
    }
    if (ambiguous) {
      double cost = 0.0;
      ++res.n_hungarian;
      if (tb.ref_hungarian) tb.ref_hungarian(assignment.data(), &cost, Cm.data(), n_hyp, n_det);
      else munkres(assignment.data(), &cost, Cm.data(), n_hyp, n_det);
    }
    std::vector<char> handled(n_det, 0);
    for (int h = 0; h < n_hyp; ++h) {  // S3D:637-660
      const int d = assignment[h];
      if (d < 0) continue;
      handled[d] = 1;
      if (!mask[h + (size_t)n_hyp * d]) { Hypothesis nh; nh.obs.push_back(dets[d]); H.push_back(nh); }
      else H[h].obs.push_back(dets[d]);
    }
    for (int d = 0; d < n_det; ++d)  // S3D:662-673
      if (!handled[d]) { Hypothesis nh; nh.obs.push_back(dets[d]); H.push_back(nh); }
  }
}

// ----------------------------------------------------------------------------
// triangulation (templated: T = float is the reference; T = double the "exact" variant)
// ----------------------------------------------------------------------------
template <class T>
struct View {
  T x, y, conf;   // normalised keypoint
  T cxx, cxy, cyy;
  const T* P;     // 12 entries row-major
  int cam;
};

// calcReprojectionError S3D:425-438
template <class T>
double reprojection_error(const T X[3], const std::vector<View<T>>& v) {
  double avg = 0., norm = 0.;
  for (const View<T>& w : v) {
    const T* P = w.P;
    const T a = sum4(P[0] * X[0], P[1] * X[1], P[2] * X[2], P[3] * T(1));
    const T b = sum4(P[4] * X[0], P[5] * X[1], P[6] * X[2], P[7] * T(1));
    const T c = sum4(P[8] * X[0], P[9] * X[1], P[10] * X[2], P[11] * T(1));
    const T dx = a / c - w.x, dy = b / c - w.y;
    const T err = std::sqrt(dx * dx + dy * dy);
    avg += static_cast<double>(w.conf * err);
    norm += static_cast<double>(w.conf);
  }
  return avg / norm;
}

// triangulate S3D:440-465
// svd_variant 0: one-sided Hestenes Jacobi (primary oracle); 1: the Eigen JacobiSVD restatement above.
// sv_out (optional): the four singular values of A in descending order.
template <class T>
void triangulate(const std::vector<View<T>>& v, bool weight_by_conf, T X[3], double* reproj_error, int svd_variant = 0,
                 T* sv_out = nullptr) {
  const int n = (int)v.size();
  std::vector<T> A((size_t)2 * n * 4);
  for (int i = 0; i < n; ++i) {
    const T* P = v[i].P;
    for (int half = 0; half < 2; ++half) {
      T* row = &A[(size_t)(2 * i + half) * 4];
      const T m = half == 0 ? v[i].x : v[i].y;
      for (int k = 0; k < 4; ++k) row[k] = m * P[8 + k] - P[half * 4 + k];
      const T z = sum4(row[0] * row[0], row[1] * row[1], row[2] * row[2], row[3] * row[3]);
      if (z > T(0)) { const T nrm = std::sqrt(z); for (int k = 0; k < 4; ++k) row[k] /= nrm; }
      if (weight_by_conf) for (int k = 0; k < 4; ++k) row[k] *= v[i].conf;
    }
  }
  T W[16];
  int best = 0;
  if (svd_variant == 1) {
    T sv[4];
    eigen_jacobi_svd_thinV<T>(A.data(), 2 * n, W, sv);
    best = 3;   // matrixV().col(3), S3D:456
    if (sv_out) for (int c = 0; c < 4; ++c) sv_out[c] = sv[c];
  } else {
    onesided_jacobi<T>(A.data(), 2 * n, 4, W);
    T best_s = std::numeric_limits<T>::max();
    T col2[4];
    for (int c = 0; c < 4; ++c) {
      T s = 0;
      for (int r = 0; r < 2 * n; ++r) s += A[(size_t)r * 4 + c] * A[(size_t)r * 4 + c];
      col2[c] = s;
      if (s < best_s) { best_s = s; best = c; }
    }
    if (sv_out) {
      std::sort(col2, col2 + 4, [](T a, T b) { return a > b; });
      for (int c = 0; c < 4; ++c) sv_out[c] = std::sqrt(col2[c]);
    }
  }
  const T w = W[3 * 4 + best];
  X[0] = W[0 * 4 + best] / w;
  X[1] = W[1 * 4 + best] / w;
  X[2] = W[2 * 4 + best] / w;
  if (reproj_error) *reproj_error = reprojection_error<T>(X, v);
}

// Levenberg-Marquardt refinement of the reprojection error. NOT IN THE REFERENCE:
// self-specified; minimises sum_i conf_i^2 * || hnorm(P_i X~) - x_i ||^2 over X, start = DLT point.
template <class T>
void lm_refine(const std::vector<View<T>>& v, int max_iters, T X[3]) {
  auto cost_at = [&](const T* Y) {
    T f = 0;
    for (const View<T>& w : v) {
      const T* P = w.P;
      const T a = P[0] * Y[0] + P[1] * Y[1] + P[2] * Y[2] + P[3];
      const T b = P[4] * Y[0] + P[5] * Y[1] + P[6] * Y[2] + P[7];
      const T c = P[8] * Y[0] + P[9] * Y[1] + P[10] * Y[2] + P[11];
      const T rx = w.conf * (a / c - w.x), ry = w.conf * (b / c - w.y);
      f += rx * rx + ry * ry;
    }
    return f;
  };
  T lambda = T(1e-3);
  T f0 = cost_at(X);
  for (int it = 0; it < max_iters; ++it) {
    T H[6] = {0, 0, 0, 0, 0, 0};  // JtJ: xx xy xz yy yz zz
    T g[3] = {0, 0, 0};           // Jt r
    for (const View<T>& w : v) {
      const T* P = w.P;
      const T a = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
      const T b = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
      const T c = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
      const T ic = T(1) / c, u = a * ic, vv = b * ic;
      const T rx = w.conf * (u - w.x), ry = w.conf * (vv - w.y);
      T jx[3], jy[3];
      for (int k = 0; k < 3; ++k) {
        jx[k] = w.conf * ic * (P[k] - u * P[8 + k]);
        jy[k] = w.conf * ic * (P[4 + k] - vv * P[8 + k]);
      }
      H[0] += jx[0] * jx[0] + jy[0] * jy[0];
      H[1] += jx[0] * jx[1] + jy[0] * jy[1];
      H[2] += jx[0] * jx[2] + jy[0] * jy[2];
      H[3] += jx[1] * jx[1] + jy[1] * jy[1];
      H[4] += jx[1] * jx[2] + jy[1] * jy[2];
      H[5] += jx[2] * jx[2] + jy[2] * jy[2];
      for (int k = 0; k < 3; ++k) g[k] += jx[k] * rx + jy[k] * ry;
    }
    const T a00 = H[0] * (T(1) + lambda), a11 = H[3] * (T(1) + lambda), a22 = H[5] * (T(1) + lambda);
    const T a01 = H[1], a02 = H[2], a12 = H[4];
    const T c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const T det = a00 * c00 + a01 * c01 + a02 * c02;
    if (!(std::fabs(det) > T(0))) break;
    const T c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const T id = T(1) / det;
    const T d[3] = {-(c00 * g[0] + c01 * g[1] + c02 * g[2]) * id,
                    -(c01 * g[0] + c11 * g[1] + c12 * g[2]) * id,
                    -(c02 * g[0] + c12 * g[1] + c22 * g[2]) * id};
    const T Y[3] = {X[0] + d[0], X[1] + d[1], X[2] + d[2]};
    const T f1 = cost_at(Y);
    if (f1 < f0) {
      X[0] = Y[0]; X[1] = Y[1]; X[2] = Y[2];
      f0 = f1;
      lambda *= T(0.1);
      if (d[0] * d[0] + d[1] * d[1] + d[2] * d[2] < T(1e-14)) break;
    } else {
      lambda *= T(10);
    }
  }
}

// calc_covariance S3D:508-523 with draw_sigma_points S3D:489-506 and mod_samples S3D:471-487
template <class T>
void ut_covariance(const T mean[3], const std::vector<View<T>>& v, T cov[9], int svd_variant = 0) {
  const int n = (int)v.size();
  const int dim = 2 * n;
  const T kappa = T(0.5);
  const int n_samples = 2 * dim + 1;
  const T wden = T(2) * (T(dim) + kappa);
  const T w0 = (T(2) * kappa) / wden, wi = T(1) / wden;
  const T b = std::sqrt(T(dim) + kappa);
  std::vector<T> Y((size_t)n_samples * 3);
  std::vector<View<T>> s(v);
  auto solve = [&](int sample) { triangulate<T>(s, false, &Y[(size_t)sample * 3], nullptr, svd_variant); };
  solve(0);
  for (int c = 0; c < n; ++c) {
    const T l11 = std::sqrt(v[c].cxx);
    const T l21 = v[c].cxy / l11;
    const T l22 = std::sqrt(v[c].cyy - l21 * l21);
    const T dx1 = l11 * b, dy1 = l21 * b, dy2 = l22 * b;
    s[c].x = v[c].x - dx1; s[c].y = v[c].y - dy1; solve(4 * c + 1);
    s[c].x = v[c].x;       s[c].y = v[c].y - dy2; solve(4 * c + 2);
    s[c].x = v[c].x + dx1; s[c].y = v[c].y + dy1; solve(4 * c + 3);
    s[c].x = v[c].x;       s[c].y = v[c].y + dy2; solve(4 * c + 4);
    s[c].x = v[c].x;       s[c].y = v[c].y;
  }
  for (int i = 0; i < 9; ++i) cov[i] = 0;
  for (int k = 0; k < n_samples; ++k) {
    const T w = k == 0 ? w0 : wi;
    const T d[3] = {Y[(size_t)k * 3] - mean[0], Y[(size_t)k * 3 + 1] - mean[1], Y[(size_t)k * 3 + 2] - mean[2]};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) cov[i * 3 + j] += (d[i] * w) * d[j];
  }
}

inline double joint_dist(const ses3d_keypoint_cov& a, const ses3d_keypoint_cov& b) {  // calcJointDist S3D:467-469
  return std::sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
inline void add_cov(ses3d_keypoint_cov& kp, double sigma) {  // addToKeypointCovariance S3D:273-277
  kp.cov[0] += sigma * sigma; kp.cov[3] += sigma * sigma; kp.cov[5] += sigma * sigma;
}

// Diagnostics of one frame for the parity tests (not part of the reference's outputs):
//   margin  smallest relative distance of any data-dependent floating-point branch of the frame to its threshold
//           (S3D:748/793 err > 0.05; S3D:775 d < bestDist; S3D:813 best > e_sub && e_sub < 0.9 err; S3D:943 root
//           distance > 2 m; S3D:964 |feet| > 0.5; S3D:988 merge distance < 0.2). Two correct float implementations
//           of the SVD may legitimately take different sides when the margin is at rounding level.
//   cond    largest sigma_1 / sigma_3 of a weighted DLT system of the frame (conditioning of the triangulated point)
struct FrameDiag {
  double margin = 1e300;
  double cond = 0.0;
  void branch(double value, double threshold, double scale) {
    const double m = std::fabs(value - threshold) / (std::fabs(scale) > 0 ? std::fabs(scale) : 1.0);
    if (m < margin) margin = m;
  }
};

// per-hypothesis body of the OpenMP loop, S3D:681-975. Returns true if the person is kept.
template <class T>
bool triangulate_hypothesis(const Tables& tb, const Hypothesis& hyp, ses3d_person_cov& person, FrameDiag* dg = nullptr) {
  const int sv_var = tb.svd_variant;
  const SkeletonModel& M = *tb.model;
  const float thrf = tb.prm.triangulation_threshold;
  const double max_reproj = tb.prm.reproj_error_max_acceptable;
  const int C = tb.n_cams;
  const int n_obs = (int)hyp.obs.size();
  std::memset(&person, 0, sizeof(person));
  if (n_obs < 2) return false;

  // camera matrices in T
  std::vector<T> PT((size_t)C * 12);
  for (size_t i = 0; i < PT.size(); ++i) PT[i] = std::is_same<T, float>::value ? T(tb.Pf[i]) : T(tb.Pd[i]);

  int num_valid = 0;
  for (int k = 0; k < NKP; ++k) {
    std::vector<View<T>> views;
    float avg_score = 0;
    for (int o = 0; o < n_obs; ++o) {
      const NormPerson& ob = hyp.obs[o];
      if (ob.kp[k][2] >= thrf) {  // S3D:725
        View<T> w;
        if (std::is_same<T, float>::value) {
          w.x = T(ob.kp[k][0]); w.y = T(ob.kp[k][1]);
          w.cxx = T(ob.cov[k][0]); w.cxy = T(ob.cov[k][1]); w.cyy = T(ob.cov[k][2]);
        } else {  // FP64 variant: normalisation carried out in double from the raw pixels
          w.x = T(ob.kpd[k][0]); w.y = T(ob.kpd[k][1]);
          w.cxx = T(ob.covd[k][0]); w.cxy = T(ob.covd[k][1]); w.cyy = T(ob.covd[k][2]);
        }
        w.conf = T(ob.kp[k][2]);
        w.P = &PT[(size_t)ob.cam * 12];
        w.cam = ob.cam;
        views.push_back(w);
        avg_score += ob.kp[k][2];
      }
    }
    int n = (int)views.size();
    if (n < 2) continue;
    avg_score /= n;

    double err;
    T X[3];
    T sv[4];
    triangulate<T>(views, true, X, &err, sv_var, sv);  // S3D:746
    if (dg) {
      if (n >= 3) dg->branch(err, max_reproj, max_reproj);
      const double c = (double)sv[0] / (double)sv[2];
      if (c > dg->cond || c != c) dg->cond = c;
    }

    if (err > max_reproj && n == 3) {  // S3D:748-792
      int best = -1;
      float best_dist = static_cast<float>(err * err);
      for (int i = 0; i < n; ++i) {
        const View<T>& a = views[i == 0 ? 1 : 0];
        const View<T>& b = views[i == 2 ? 1 : 2];
        const float* F = &tb.F[(size_t)fundamental_idx(a.cam, b.cam, C) * 9];
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
        if (dg) dg->branch(d, best_dist, best_dist);
        if (d < best_dist) { best_dist = d; best = i; }
      }
      if (best != -1) {
        views.erase(views.begin() + best);
        triangulate<T>(views, true, X, &err, sv_var);
        avg_score = ((float)views[0].conf + (float)views[1].conf) / 2.0f;
        n = 2;
      }
    } else if (err > max_reproj && n >= 4) {  // S3D:793-838
      double best_err = err;
      int best = -1;
      float best_score = avg_score;
      T bestX[3] = {0, 0, 0};
      for (int i = 0; i < n; ++i) {
        std::vector<View<T>> sub(views);
        sub.erase(sub.begin() + i);
        double e_sub;
        T Xs[3];
        triangulate<T>(sub, true, Xs, &e_sub, sv_var);
        if (dg) { dg->branch(e_sub, best_err, err); dg->branch(e_sub, 0.9 * err, err); }
        if (best_err > e_sub && e_sub < 0.9 * err) {
          best_err = e_sub; best = i;
          bestX[0] = Xs[0]; bestX[1] = Xs[1]; bestX[2] = Xs[2];
          float tmp = 0.f;
          for (const View<T>& w : sub) tmp += (float)w.conf;
          best_score = tmp / sub.size();
        }
      }
      if (best != -1) {
        views.erase(views.begin() + best);
        X[0] = bestX[0]; X[1] = bestX[1]; X[2] = bestX[2];
        err = best_err;
        avg_score = best_score;
        n = n - 1;
      }
    }

    if (tb.prm.lm_refine) {  // not in the reference; default off
      lm_refine<T>(views, tb.prm.lm_max_iters, X);
      err = reprojection_error<T>(X, views);
    }

    if (err > max_reproj) avg_score *= (max_reproj / err);  // S3D:840-844

    T cov[9];
    ut_covariance<T>(X, views, cov, sv_var);  // S3D:846-847

    ses3d_keypoint_cov& out = person.keypoints[M.fusion_idx[k]];  // S3D:849-857
    out.x = static_cast<double>(X[0]); out.y = static_cast<double>(X[1]); out.z = static_cast<double>(X[2]);
    out.score = avg_score;
    out.cov[0] = (double)cov[0]; out.cov[1] = (double)cov[1]; out.cov[2] = (double)cov[2];
    out.cov[3] = (double)cov[4]; out.cov[4] = (double)cov[5]; out.cov[5] = (double)cov[8];
    ++num_valid;
  }

  // limb-length covariance inflation S3D:861-883 (marker code 885-921 dropped)
  for (int k = 0; k < NKP; ++k) {
    ses3d_keypoint_cov& kp = person.keypoints[M.fusion_idx[k]];
    if (kp.score <= 0) continue;
    const int parent = M.parent[k];
    if (parent >= 0) {
      const ses3d_keypoint_cov& pk = person.keypoints[M.fusion_idx[parent]];
      if (pk.score > 0 && M.limb_len[k] > 0) {
        const double d = joint_dist(kp, pk);
        add_cov(kp, tb.prm.limb_cov_offset_sigma * (d - M.limb_len[k]) / M.limb_sigma[k]);
      } else if (tb.prm.pose_method == SES3D_POSE_SIMPLE && k == kSimpleRShoulder) {
        ses3d_keypoint_cov& ls = person.keypoints[M.fusion_idx[kSimpleLShoulder]];
        if (ls.score > 0) {
          const double d = joint_dist(kp, ls);
          add_cov(kp, tb.prm.limb_cov_offset_sigma * (d - kShoulderDist) / kShoulderSigma);
          add_cov(ls, tb.prm.limb_cov_offset_sigma * (d - kShoulderDist) / kShoulderSigma);
        }
      }
    }
  }

  // root distance S3D:923-953
  ses3d_keypoint_cov root;
  std::memset(&root, 0, sizeof(root));
  const ses3d_keypoint_cov* K = person.keypoints;
  if (K[SES3D_FBP_MIDHIP].score > 0) root = K[SES3D_FBP_MIDHIP];
  else if (K[SES3D_FBP_LHIP].score > 0 && K[SES3D_FBP_RHIP].score > 0) {
    root.x = (K[SES3D_FBP_LHIP].x + K[SES3D_FBP_RHIP].x) / 2.;
    root.y = (K[SES3D_FBP_LHIP].y + K[SES3D_FBP_RHIP].y) / 2.;
    root.z = (K[SES3D_FBP_LHIP].z + K[SES3D_FBP_RHIP].z) / 2.;
    root.score = (K[SES3D_FBP_LHIP].score + K[SES3D_FBP_RHIP].score) / 2.f;
  }
  if (root.score > 0) {
    for (int s = 0; s < NFUS; ++s) {
      ses3d_keypoint_cov& kp = person.keypoints[s];
      if (kp.score > 0) {
        if (dg) dg->branch(joint_dist(root, kp), tb.prm.max_joint_dist_to_root, tb.prm.max_joint_dist_to_root);
        if (joint_dist(root, kp) > tb.prm.max_joint_dist_to_root) { std::memset(&kp, 0, sizeof(kp)); --num_valid; }
      } else {
        std::memset(&kp, 0, sizeof(kp));
        --num_valid;
      }
    }
  }
  // feet height S3D:955-966
  double feet = 0.0;
  if (K[SES3D_FBP_LANKLE].score > 0 && K[SES3D_FBP_RANKLE].score > 0) feet = (K[SES3D_FBP_LANKLE].z + K[SES3D_FBP_RANKLE].z) / 2.0;
  else if (K[SES3D_FBP_LANKLE].score > 0) feet = K[SES3D_FBP_LANKLE].z;
  else if (K[SES3D_FBP_RANKLE].score > 0) feet = K[SES3D_FBP_RANKLE].z;
  if (dg && feet != 0.0) dg->branch(std::fabs(feet), 0.50, 0.50);
  if (std::fabs(feet) > 0.50) num_valid = 0;
  return num_valid > tb.prm.min_num_valid_keypoints;  // S3D:968
}

// calc_3D_dist S3D:392-408
double dist3d(const ses3d_person_cov& a, const ses3d_person_cov& b) {
  int n = 0;
  double d = 0;
  for (int s = 0; s < NFUS; ++s) {
    const ses3d_keypoint_cov &p = a.keypoints[s], &q = b.keypoints[s];
    if (p.score > 0 && q.score > 0) {
      d += std::sqrt(std::pow(p.x - q.x, 2) + std::pow(p.y - q.y, 2) + std::pow(p.z - q.z, 2));
      ++n;
    }
  }
  return n > 0 ? d / n : MAX_COSTS;
}

// merge_persons S3D:410-423 (+ mergeKeypointCovariance S3D:264-271)
void merge_into(ses3d_person_cov& a, const ses3d_person_cov& b) {
  for (int s = 0; s < NFUS; ++s) {
    ses3d_keypoint_cov& p = a.keypoints[s];
    const ses3d_keypoint_cov& q = b.keypoints[s];
    const double total = static_cast<double>(p.score + q.score);
    if (total > 0.0) {
      p.x = ((double)p.score * p.x + (double)q.score * q.x) / total;
      p.y = ((double)p.score * p.y + (double)q.score * q.y) / total;
      p.z = ((double)p.score * p.z + (double)q.score * q.z) / total;
      p.score = std::max(p.score, q.score);
      for (int i = 0; i < 6; ++i) p.cov[i] = (p.cov[i] + q.cov[i]) / 2.0;
    }
  }
}

// One frame of triangulate_persons (S3D:525-997). In the FP64 variant the normalised keypoints
// handed to the DLT are recomputed in double from the raw pixels; association (indices) is
// identical in both variants. *n_joints accumulates the output joints (score > 0).
template <class T>
int triangulate_frame(const Tables& tb, int p_max, const ses3d_person2d* persons, const int32_t* n_persons, int h_max,
                      ses3d_person_cov* out, int32_t* hyp_of, int32_t* n_hyp, int32_t* n_hung, int64_t* n_joints,
                      FrameDiag* dg = nullptr) {
  AssocResult ar;
  associate(tb, p_max, persons, n_persons, ar);
  if (hyp_of) {
    for (int i = 0; i < tb.n_cams * p_max; ++i) hyp_of[i] = -1;
    for (size_t h = 0; h < ar.H.size(); ++h)
      for (const NormPerson& ob : ar.H[h].obs) hyp_of[ob.cam * p_max + ob.det] = (int32_t)h;
  }
  if (n_hyp) *n_hyp = (int32_t)ar.H.size();
  if (n_hung) *n_hung = ar.n_hungarian;
  std::vector<ses3d_person_cov> kept;
  for (size_t h = 0; h < ar.H.size(); ++h) {  // hypothesis order = the reference built without OpenMP
    ses3d_person_cov person;
    if (triangulate_hypothesis<T>(tb, ar.H[h], person, dg)) kept.push_back(person);
  }
  for (size_t i = 0; i < kept.size(); ++i)  // S3D:984-996
    for (size_t j = i + 1; j < kept.size();) {
      if (dg) dg->branch(dist3d(kept[i], kept[j]), tb.prm.merge_dist_thresh, tb.prm.merge_dist_thresh);
      if (dist3d(kept[i], kept[j]) < tb.prm.merge_dist_thresh) { merge_into(kept[i], kept[j]); kept.erase(kept.begin() + j); }
      else ++j;
    }
  if ((int)kept.size() > h_max) return -1;
  for (size_t i = 0; i < kept.size(); ++i) {
    out[i] = kept[i];
    if (n_joints)
      for (int s = 0; s < NFUS; ++s) *n_joints += kept[i].keypoints[s].score > 0 ? 1 : 0;
  }
  return (int)kept.size();
}

// REP:139-235 for one frame
void reproject_frame(const Tables& tb, int h_max, const ses3d_person_cov* persons, int n_persons,
                     ses3d_person2d* out /*[C][h_max]*/, int32_t* n_out /*[C]*/) {
  const int C = tb.n_cams;
  const SkeletonModel& M = *tb.model;
  for (int i = 0; i < C; ++i) n_out[i] = 0;
  const double kappa = 0.5;
  const double wden = 2.0 * (3 + kappa);
  const double w0 = 2 * kappa / wden, wi = 1.0 / wden;
  const double spread = std::sqrt(3 + kappa);
  std::vector<ses3d_person2d> in_cam(C);
  std::vector<int> n_valid(C);
  std::vector<double> mnx(C), mny(C), mxx(C), mxy(C);
  for (int p = 0; p < n_persons; ++p) {
    for (int i = 0; i < C; ++i) {
      std::memset(&in_cam[i], 0, sizeof(ses3d_person2d));
      in_cam[i].score = 1.0f;  // REP:175
      n_valid[i] = 0;
      mnx[i] = tb.cams[i].width; mny[i] = tb.cams[i].height; mxx[i] = 0; mxy[i] = 0;  // REP:150,161-162
    }
    for (int k = 0; k < NKP; ++k) {
      const ses3d_keypoint_cov& kp = persons[p].keypoints[M.fusion_idx[k]];
      if (kp.score <= 0.0f) continue;  // REP:181
      // llt of the symmetric 3x3 (lower Cholesky) REP:72, 184-187
      const double a00 = kp.cov[0], a10 = kp.cov[1], a20 = kp.cov[2], a11 = kp.cov[3], a21 = kp.cov[4], a22 = kp.cov[5];
      // Eigen 3.3 LLT.h llt_inplace<double, Lower>::unblocked (a 3x3 never takes the blocked path): a non-positive
      // pivot x <= 0 stops the factorisation and the remaining lower-triangle entries keep their current values;
      // cov.llt().matrixL() (REP:72) reads that lower triangle without checking info(). A NaN pivot compares false
      // and continues into sqrt(NaN), as in Eigen.
      double l00 = a00, l10 = a10, l20 = a20, l11 = a11, l21 = a21, l22 = a22;
      do {
        if (l00 <= 0.0) break;
        l00 = std::sqrt(l00);
        l10 /= l00; l20 /= l00;
        double x = l11 - l10 * l10;            // x -= A10.squaredNorm()
        if (x <= 0.0) break;
        l11 = x = std::sqrt(x);
        l21 -= l20 * l10;                      // A21 -= A20 * A10^T
        l21 /= x;
        x = l22 - (l20 * l20 + l21 * l21);
        if (x <= 0.0) break;
        l22 = std::sqrt(x);
      } while (false);
      const double L[3][3] = {{l00, 0, 0}, {l10, l11, 0}, {l20, l21, l22}};
      double S[7][3];  // mean, three minus, three plus (REP:68-72)
      for (int s = 0; s < 7; ++s) {
        double e[3] = {0, 0, 0};
        if (s >= 1 && s <= 3) e[s - 1] = -spread;
        if (s >= 4) e[s - 4] = spread;
        const double mean[3] = {kp.x, kp.y, kp.z};
        for (int r = 0; r < 3; ++r) S[s][r] = (L[r][0] * e[0] + L[r][1] * e[1] + L[r][2] * e[2]) + mean[r];
      }
      for (int i = 0; i < C; ++i) {
        const double* T = tb.cams[i].T_cam_base;
        const ses3d_camera& cm = tb.cams[i];
        double u[7], v[7];
        for (int s = 0; s < 7; ++s) {
          const double X = T[0] * S[s][0] + T[1] * S[s][1] + T[2] * S[s][2] + T[3];
          const double Y = T[4] * S[s][0] + T[5] * S[s][1] + T[6] * S[s][2] + T[7];
          const double Z = T[8] * S[s][0] + T[9] * S[s][1] + T[10] * S[s][2] + T[11];
          u[s] = (cm.fx * X + cm.Tx) / Z + cm.cx;  // image_geometry::project3dToPixel
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
        if (mu < 0 || mu > cm.width || mv < 0 || mv > cm.height) continue;  // REP:207-208
        ++n_valid[i];
        ses3d_keypoint2d& o = in_cam[i].keypoints[k];
        o.x = static_cast<float>(mu); o.y = static_cast<float>(mv); o.score = kp.score;
        o.cov[0] = static_cast<float>(cxx); o.cov[1] = static_cast<float>(cxy); o.cov[2] = static_cast<float>(cyy);
        if (mu < mnx[i]) mnx[i] = mu;
        if (mv < mny[i]) mny[i] = mv;
        if (mu > mxx[i]) mxx[i] = mu;
        if (mv > mxy[i]) mxy[i] = mv;
      }
    }
    for (int i = 0; i < C; ++i)
      if (n_valid[i] > 0) {  // REP:225-230
        in_cam[i].bbox[0] = (float)mnx[i]; in_cam[i].bbox[1] = (float)mny[i];
        in_cam[i].bbox[2] = (float)mxx[i]; in_cam[i].bbox[3] = (float)mxy[i];
        out[(size_t)i * h_max + n_out[i]] = in_cam[i];
        ++n_out[i];
      }
  }
}

template <class F>
void parallel_frames(int n_frames, int n_threads, F&& body) {
  if (n_threads <= 1) { body(0, n_frames); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < n_threads; ++t) {
    const int a = (int)((int64_t)n_frames * t / n_threads), b = (int)((int64_t)n_frames * (t + 1) / n_threads);
    pool.emplace_back([=, &body] { body(a, b); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

// ----------------------------------------------------------------------------
// C interface for ctypes (tests, smoke, bench cpu_baseline)
// ----------------------------------------------------------------------------
extern "C" {

void* oracle_create(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* prm) {
  if (n_cams < 2 || !cams || !prm) return nullptr;
  Tables* tb = new Tables;
  tb->n_cams = n_cams;
  tb->prm = *prm;
  tb->model = prm->pose_method == SES3D_POSE_H36M ? &kH36M : &kSimple;
  tb->cams.assign(cams, cams + n_cams);
  build_tables(*tb);
  return tb;
}

void oracle_destroy(void* h) {
  Tables* tb = static_cast<Tables*>(h);
  if (tb && tb->ref_hungarian_lib) dlclose(tb->ref_hungarian_lib);
  delete tb;
}

// Swap the restated Munkres for the reference's verbatim Hungarian.cpp (oracle/_ref/libref_hungarian.so).
int oracle_use_ref_hungarian(void* h, const char* so_path) {
  Tables* tb = static_cast<Tables*>(h);
  void* lib = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!lib) return -1;
  void* fn = dlsym(lib, "ref_hungarian_assignmentoptimal");
  if (!fn) { dlclose(lib); return -2; }
  tb->ref_hungarian_lib = lib;
  tb->ref_hungarian = reinterpret_cast<void (*)(int*, double*, double*, int, int)>(fn);
  return 0;
}

void oracle_get_tables(void* h, float* P, float* F) {
  Tables* tb = static_cast<Tables*>(h);
  if (P) std::memcpy(P, tb->Pf.data(), tb->Pf.size() * sizeof(float));
  if (F) std::memcpy(F, tb->F.data(), tb->F.size() * sizeof(float));
}

void oracle_munkres(int* assignment, double* cost, const double* dist, int n_rows, int n_cols) {
  munkres(assignment, cost, dist, n_rows, n_cols);
}

void oracle_set_svd_variant(void* h, int32_t variant) { static_cast<Tables*>(h)->svd_variant = variant == 1 ? 1 : 0; }

// double-precision variant flag: precision taken from params at create.
// diag (optional) [n_frames][2]: FrameDiag margin, cond.
int oracle_triangulate_batch_ex(void* h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                                const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out, int32_t* n_out,
                                int32_t* hyp_of, int32_t* n_hyp, int32_t* n_hung, int64_t* n_joints_total,
                                int32_t n_threads, double* diag) {
  Tables* tb = static_cast<Tables*>(h);
  const int C = tb->n_cams;
  std::vector<int64_t> joints((size_t)std::max(1, n_threads), 0);
  std::vector<int> status((size_t)std::max(1, n_threads), 0);
  int tcount = std::max(1, n_threads);
  std::vector<std::thread> pool;
  auto body = [&](int t, int a, int b) {
    for (int f = a; f < b; ++f) {
      const ses3d_person2d* pf = persons + (size_t)f * C * p_max;
      const int32_t* nf = n_persons + (size_t)f * C;
      int32_t* ho = hyp_of ? hyp_of + (size_t)f * C * p_max : nullptr;
      int r;
      FrameDiag dg;
      if (tb->prm.precision == SES3D_PRECISION_FP64)
        r = triangulate_frame<double>(*tb, p_max, pf, nf, h_max, out + (size_t)f * h_max, ho,
                                      n_hyp ? n_hyp + f : nullptr, n_hung ? n_hung + f : nullptr, &joints[t],
                                      diag ? &dg : nullptr);
      else
        r = triangulate_frame<float>(*tb, p_max, pf, nf, h_max, out + (size_t)f * h_max, ho,
                                     n_hyp ? n_hyp + f : nullptr, n_hung ? n_hung + f : nullptr, &joints[t],
                                     diag ? &dg : nullptr);
      if (diag) { diag[(size_t)f * 2] = dg.margin; diag[(size_t)f * 2 + 1] = dg.cond; }
      if (r < 0) { status[t] = SES3D_E_CAPACITY; n_out[f] = 0; }
      else n_out[f] = r;
    }
  };
  if (tcount == 1) body(0, 0, n_frames);
  else {
    for (int t = 0; t < tcount; ++t) {
      const int a = (int)((int64_t)n_frames * t / tcount), b = (int)((int64_t)n_frames * (t + 1) / tcount);
      pool.emplace_back(body, t, a, b);
    }
    for (auto& th : pool) th.join();
  }
  int64_t total = 0;
  int st = 0;
  for (int t = 0; t < tcount; ++t) { total += joints[t]; if (status[t]) st = status[t]; }
  if (n_joints_total) *n_joints_total = total;
  return st;
}

int oracle_triangulate_batch(void* h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                             const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out, int32_t* n_out,
                             int32_t* hyp_of, int32_t* n_hyp, int32_t* n_hung, int64_t* n_joints_total,
                             int32_t n_threads) {
  return oracle_triangulate_batch_ex(h, n_frames, p_max, persons, n_persons, h_max, out, n_out, hyp_of, n_hyp, n_hung,
                                     n_joints_total, n_threads, nullptr);
}

int oracle_reproject_batch(void* h, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons3d,
                           const int32_t* n_persons3d, ses3d_person2d* out, int32_t* n_out, int32_t n_threads) {
  Tables* tb = static_cast<Tables*>(h);
  const int C = tb->n_cams;
  parallel_frames(n_frames, n_threads, [&](int a, int b) {
    for (int f = a; f < b; ++f)
      reproject_frame(*tb, h_max, persons3d + (size_t)f * h_max, n_persons3d[f], out + (size_t)f * C * h_max,
                      n_out + (size_t)f * C);
  });
  return 0;
}

// Single DLT solve exposed for the numpy.linalg.svd cross-check: P [n][12], pts [n][3] (x,y,conf)
void oracle_triangulate_point_v(int32_t n, const double* P, const double* pts, int32_t weighted, int32_t use_double,
                                int32_t svd_variant, double X[3], double* reproj, double sv[4]) {
  if (use_double) {
    std::vector<double> Pd(P, P + (size_t)n * 12);
    std::vector<View<double>> v(n);
    for (int i = 0; i < n; ++i) { v[i] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], 0, 0, 0, &Pd[(size_t)i * 12], i}; }
    double Xd[3];
    triangulate<double>(v, weighted != 0, Xd, reproj, svd_variant, sv);
    X[0] = Xd[0]; X[1] = Xd[1]; X[2] = Xd[2];
  } else {
    std::vector<float> Pf((size_t)n * 12);
    for (size_t i = 0; i < Pf.size(); ++i) Pf[i] = (float)P[i];
    std::vector<View<float>> v(n);
    for (int i = 0; i < n; ++i)
      v[i] = {(float)pts[i * 3], (float)pts[i * 3 + 1], (float)pts[i * 3 + 2], 0, 0, 0, &Pf[(size_t)i * 12], i};
    float Xf[3], svf[4];
    triangulate<float>(v, weighted != 0, Xf, reproj, svd_variant, svf);
    X[0] = Xf[0]; X[1] = Xf[1]; X[2] = Xf[2];
    if (sv) for (int i = 0; i < 4; ++i) sv[i] = svf[i];
  }
}

void oracle_triangulate_point(int32_t n, const double* P, const double* pts, int32_t weighted, int32_t use_double,
                              double X[3], double* reproj) {
  if (use_double) {
    std::vector<double> Pd(P, P + (size_t)n * 12);
    std::vector<View<double>> v(n);
    for (int i = 0; i < n; ++i) { v[i] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], 0, 0, 0, &Pd[(size_t)i * 12], i}; }
    double Xd[3];
    triangulate<double>(v, weighted != 0, Xd, reproj);
    X[0] = Xd[0]; X[1] = Xd[1]; X[2] = Xd[2];
  } else {
    std::vector<float> Pf((size_t)n * 12);
    for (size_t i = 0; i < Pf.size(); ++i) Pf[i] = (float)P[i];
    std::vector<View<float>> v(n);
    for (int i = 0; i < n; ++i)
      v[i] = {(float)pts[i * 3], (float)pts[i * 3 + 1], (float)pts[i * 3 + 2], 0, 0, 0, &Pf[(size_t)i * 12], i};
    float Xf[3];
    triangulate<float>(v, weighted != 0, Xf, reproj);
    X[0] = Xf[0]; X[1] = Xf[1]; X[2] = Xf[2];
  }
}

// LM refinement exposed for the scipy least_squares cross-check (double only)
void oracle_lm_refine(int32_t n, const double* P, const double* pts, int32_t max_iters, double X[3]) {
  std::vector<double> Pd(P, P + (size_t)n * 12);
  std::vector<View<double>> v(n);
  for (int i = 0; i < n; ++i) v[i] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], 0, 0, 0, &Pd[(size_t)i * 12], i};
  lm_refine<double>(v, max_iters, X);
}

// UT covariance exposed for unit checks: cov2d [n][3] (xx,xy,yy)
void oracle_ut_covariance(int32_t n, const double* P, const double* pts, const double* cov2d, const double mean[3],
                          double cov[9]) {
  std::vector<double> Pd(P, P + (size_t)n * 12);
  std::vector<View<double>> v(n);
  for (int i = 0; i < n; ++i)
    v[i] = {pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], cov2d[i * 3], cov2d[i * 3 + 1], cov2d[i * 3 + 2],
            &Pd[(size_t)i * 12], i};
  ut_covariance<double>(mean, v, cov);
}

}  // extern "C"
