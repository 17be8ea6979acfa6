// host_setup.cpp — K0 "setup_cameras": the constant tables the hot path reads.
//
// Replaces the set-up block of skeleton_3d's main() (S3D:1187-1211): projection matrices
// P_i = [R_i|t_i] (base -> camera, K = I because keypoints are normalised), camera centres
// C_i = inverse(T_i).col(3), and for every pair i<j the fundamental (here: essential) matrix
// F_ij = [P_j C_i]_x P_j pinv(P_i) (cross_prod_matrix S3D:230-234, pseudo_inv34d S3D:236-240),
// computed in double and cast to float, stored in get_fundamental_idx order (S3D:242-253).
// Also the skeleton tables selected by pose_method (S3D:81-145, 1101-1112).
//
// Runs once per handle on the host; compiled with -ffp-contract=off so that the float casts
// are reproducible.
#include "host_setup.h"

#include <cmath>
#include <limits>

namespace ses3d {
namespace {

// EdgeTPU_BodyParts_Simple (S3D:81-104) and the fusion-slot map (S3D:139-142)
const SkeletonModel kModelSimple = {
    {-1, 0, 0, 1, 2, 0, 0, 5, 6, 7, 8, 5, 6, 11, 12, 13, 14},
    {-1, 0.05, 0.05, 0.10, 0.10, -1, -1, 0.28, 0.28, 0.25, 0.25, 0.50, 0.50, 0.45, 0.45, 0.446, 0.446},
    {-1, 0.05, 0.05, 0.05, 0.05, -1, -1, 0.10, 0.10, 0.10, 0.10, 0.15, 0.15, 0.10, 0.10, 0.10, 0.10},
    {SES3D_FBP_NOSE, SES3D_FBP_LEYE, SES3D_FBP_REYE, SES3D_FBP_LEAR, SES3D_FBP_REAR, SES3D_FBP_LSHOULDER,
     SES3D_FBP_RSHOULDER, SES3D_FBP_LELBOW, SES3D_FBP_RELBOW, SES3D_FBP_LWRIST, SES3D_FBP_RWRIST, SES3D_FBP_LHIP,
     SES3D_FBP_RHIP, SES3D_FBP_LKNEE, SES3D_FBP_RKNEE, SES3D_FBP_LANKLE, SES3D_FBP_RANKLE}};
// EdgeTPU_BodyParts_H36M (S3D:111-133) and its map (S3D:143-145)
const SkeletonModel kModelH36M = {
    {-1, 0, 0, 2, 3, 2, 2, 5, 6, 7, 8, 4, 4, 11, 12, 13, 14},
    {-1, 0.115, 0.116, 0.255, 0.238, 0.149, 0.149, 0.28, 0.28, 0.25, 0.25, 0.134, 0.134, 0.449, 0.449, 0.446, 0.446},
    {-1, 0.07, 0.07, 0.15, 0.15, 0.10, 0.10, 0.15, 0.15, 0.15, 0.15, 0.10, 0.10, 0.20, 0.20, 0.20, 0.20},
    {SES3D_FBP_NOSE, SES3D_FBP_HEAD, SES3D_FBP_NECK, SES3D_FBP_BELLY, SES3D_FBP_MIDHIP, SES3D_FBP_LSHOULDER,
     SES3D_FBP_RSHOULDER, SES3D_FBP_LELBOW, SES3D_FBP_RELBOW, SES3D_FBP_LWRIST, SES3D_FBP_RWRIST, SES3D_FBP_LHIP,
     SES3D_FBP_RHIP, SES3D_FBP_LKNEE, SES3D_FBP_RKNEE, SES3D_FBP_LANKLE, SES3D_FBP_RANKLE}};

// Moore-Penrose inverse of a row-major 3x4 through a one-sided Jacobi SVD of its transpose:
// M^T = Q S W^T  =>  pinv(M) = Q S^-1 W^T, singular values below eps*4*s_max dropped (S3D:238-239).
void pinv_3x4(const double* M, double* out /*4x3 row-major*/) {
  double B[4][3], W[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 3; ++c) B[r][c] = M[c * 4 + r];
  const double eps = std::numeric_limits<double>::epsilon();
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < 4; ++r) {
          alpha += B[r][p] * B[r][p];
          beta += B[r][q] * B[r][q];
          gamma += B[r][p] * B[r][q];
        }
        if (std::fabs(gamma) <= eps * std::sqrt(alpha * beta) || gamma == 0.0) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t);
        const double s = c * t;
        for (int r = 0; r < 4; ++r) {
          const double bp = B[r][p], bq = B[r][q];
          B[r][p] = c * bp - s * bq;
          B[r][q] = s * bp + c * bq;
        }
        for (int r = 0; r < 3; ++r) {
          const double wp = W[r][p], wq = W[r][q];
          W[r][p] = c * wp - s * wq;
          W[r][q] = s * wp + c * wq;
        }
      }
    if (!rotated) break;
  }
  double sv[3], smax = 0;
  for (int c = 0; c < 3; ++c) {
    double n = 0;
    for (int r = 0; r < 4; ++r) n += B[r][c] * B[r][c];
    sv[c] = std::sqrt(n);
    smax = sv[c] > smax ? sv[c] : smax;
  }
  const double tol = eps * 4.0 * smax;
  for (int i = 0; i < 12; ++i) out[i] = 0.0;
  for (int c = 0; c < 3; ++c) {
    if (!(std::fabs(sv[c]) > tol)) continue;
    const double inv2 = 1.0 / (sv[c] * sv[c]);
    for (int r = 0; r < 4; ++r)
      for (int k = 0; k < 3; ++k) out[r * 3 + k] += B[r][c] * inv2 * W[k][c];
  }
}

}  // namespace

bool build_host_tables(int n_cams, const ses3d_camera* cams, const ses3d_params& prm, HostTables* out) {
  if (n_cams < 2 || n_cams > 255 || !cams || !out) return false;
  const int C = n_cams;
  out->n_cams = C;
  out->model = prm.pose_method == SES3D_POSE_H36M ? kModelH36M : kModelSimple;
  out->camf.resize(C);
  out->camd.resize(C);
  std::vector<double> centre((size_t)C * 4);
  for (int i = 0; i < C; ++i) {
    const double* T = cams[i].T_cam_base;
    CamF& cf = out->camf[i];
    CamD& cd = out->camd[i];
    for (int k = 0; k < 12; ++k) { cd.P[k] = T[k]; cf.P[k] = static_cast<float>(T[k]); }
    cd.fx = cams[i].fx; cd.fy = cams[i].fy; cd.cx = cams[i].cx; cd.cy = cams[i].cy;
    cd.Tx = cams[i].Tx; cd.Ty = cams[i].Ty;
    cd.width = (double)cams[i].width; cd.height = (double)cams[i].height;
    cf.fx = static_cast<float>(cams[i].fx); cf.fy = static_cast<float>(cams[i].fy);   // S3D:314-317
    cf.cx = static_cast<float>(cams[i].cx); cf.cy = static_cast<float>(cams[i].cy);
    cf.inv_fx2 = 1.0f / (cf.fx * cf.fx); cf.inv_fxfy = 1.0f / (cf.fx * cf.fy); cf.inv_fy2 = 1.0f / (cf.fy * cf.fy);
    cf.pad_ = 0.f;
    // camera centre: inverse of the affine transform, linear part inverted by cofactors (S3D:1191)
    const double a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9], k = T[10];
    const double A = e * k - f * h, B = -(d * k - f * g), Cc = d * h - e * g;
    const double det = a * A + b * B + c * Cc;
    if (!(std::fabs(det) > 0.0) || !(cams[i].fx != 0.0) || !(cams[i].fy != 0.0)) return false;
    const double inv[9] = {A / det, -(b * k - c * h) / det, (b * f - c * e) / det,
                           B / det, (a * k - c * g) / det, -(a * f - c * d) / det,
                           Cc / det, -(a * h - b * g) / det, (a * e - b * d) / det};
    const double t[3] = {T[3], T[7], T[11]};
    for (int r = 0; r < 3; ++r)
      centre[(size_t)i * 4 + r] = -(inv[r * 3 + 0] * t[0] + inv[r * 3 + 1] * t[1] + inv[r * 3 + 2] * t[2]);
    centre[(size_t)i * 4 + 3] = 1.0;
  }
  out->f_row.resize(C);
  int start = 0;
  for (int i = 0; i < C; ++i) { out->f_row[i] = start; start += C - i - 1; }
  out->F.assign((size_t)C * (C - 1) / 2 * 9, 0.f);
  for (int i = 0; i < C; ++i) {
    double Pinv[12];
    pinv_3x4(out->camd[i].P, Pinv);
    for (int j = i + 1; j < C; ++j) {
      const double* Pj = out->camd[j].P;
      const double* Ci = &centre[(size_t)i * 4];
      double e[3];  // epipole of camera i in image j: P_j * C_i (S3D:1197)
      for (int r = 0; r < 3; ++r)
        e[r] = sum4(Pj[r * 4 + 0] * Ci[0], Pj[r * 4 + 1] * Ci[1], Pj[r * 4 + 2] * Ci[2], Pj[r * 4 + 3] * Ci[3]);
      const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
      double M[12];  // [e]_x P_j
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
          M[r * 4 + c] = sum3(ex[r * 3 + 0] * Pj[0 * 4 + c], ex[r * 3 + 1] * Pj[1 * 4 + c], ex[r * 3 + 2] * Pj[2 * 4 + c]);
      float* Fo = &out->F[(size_t)(out->f_row[i] + j - i - 1) * 9];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          Fo[r * 3 + c] = static_cast<float>(sum4(M[r * 4 + 0] * Pinv[0 * 3 + c], M[r * 4 + 1] * Pinv[1 * 3 + c],
                                                  M[r * 4 + 2] * Pinv[2 * 3 + c], M[r * 4 + 3] * Pinv[3 * 3 + c]));
    }
  }
  return true;
}

}  // namespace ses3d
