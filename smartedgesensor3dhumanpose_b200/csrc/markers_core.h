// markers_core.h — the numeric content of the reference's rviz markers (visualisation only).
//
//   covariance ellipsoid of a joint   setMarkerPose  S3D:279-310 == PRI:237-254: eigen-decomposition of the 3x3
//       covariance (Eigen::SelfAdjointEigenSolver, eigenvalues ascending), orientation = the eigenvector matrix made
//       right-handed (negated when its determinant is not positive), scale = 2 x 2.7955 x sqrt(eigenvalue)
//   skeleton line list                S3D:898-916 (detector joints, parent table S3D:100/129) and
//                                     addJointToSkeleton PRI:273-382 (fusion slots incl. Neck / MidHip / Belly)
// ROS message assembly (headers, namespaces, colours, lifetimes) stays in the node. Eigen is not under
// /root/reference, and the sign of an eigenvector is arbitrary anyway: the ellipsoid is pinned by its invariants
// (eigenvalues, right-handed orthonormal frame, R diag(lambda) R^T = covariance), not by the quaternion's bits.
#pragma once
#include "common.h"

namespace ses3d {

enum { MARKER_STYLE_SKELETON3D = 0, MARKER_STYLE_POSE_PRIOR = 1 };
constexpr int MARKER_MAX_SEGMENTS = 22;   // 21 joints + the second Belly segment (PRI:317-337)

// cyclic Jacobi on a symmetric 3x3 (c = xx,xy,xz,yy,yz,zz): eigenvalues ascending in l, eigenvectors in the columns of V
SES_HD void eig3_sym(const double c[6], double l[3], double V[9]) {
  double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double dia = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (!(off > 1e-32 * dia)) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = cs * akp - sn * akq;
          a[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = cs * apk - sn * aqk;
          a[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; ++k) {  // V <- V J
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cs * vkp - sn * vkq;
          v[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  int o[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (a[o[j]][o[j]] > a[o[j + 1]][o[j + 1]]) { const int t = o[j]; o[j] = o[j + 1]; o[j + 1] = t; }
  for (int j = 0; j < 3; ++j) {
    l[j] = a[o[j]][o[j]];
    for (int k = 0; k < 3; ++k) V[k * 3 + j] = v[k][o[j]];
  }
}

// Eigen::Quaternion(Matrix3) — the trace-based conversion (row-major R)
SES_HD void quat_from_rotation(const double R[9], double q[4] /* w x y z */) {
  const double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    double s = sqrt(t + 1.0);
    q[0] = 0.5 * s;
    s = 0.5 / s;
    q[1] = (R[7] - R[5]) * s; q[2] = (R[2] - R[6]) * s; q[3] = (R[3] - R[1]) * s;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    q[1 + i] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (R[k * 3 + j] - R[j * 3 + k]) * s;
    q[1 + j] = (R[j * 3 + i] + R[i * 3 + j]) * s;
    q[1 + k] = (R[k * 3 + i] + R[i * 3 + k]) * s;
  }
}

// setMarkerPose: position is the joint itself; returns orientation (w, x, y, z) and scale (x, y, z)
SES_HD void covariance_ellipsoid(const double cov[6], ses3d_ellipsoid* e) {
  double l[3], V[9];
  eig3_sym(cov, l, V);
  const double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) +
                     V[2] * (V[3] * V[7] - V[4] * V[6]);
  if (!(det > 0.0))
    for (int i = 0; i < 9; ++i) V[i] = -1.0 * V[i];   // "Determinant must be +1!"
  double q[4];
  quat_from_rotation(V, q);
  e->qw = q[0]; e->qx = q[1]; e->qy = q[2]; e->qz = q[3];
  e->sx = 2.0 * 2.7955 * sqrt(l[0]); e->sy = 2.0 * 2.7955 * sqrt(l[1]); e->sz = 2.0 * 2.7955 * sqrt(l[2]);
}

// LINE_LIST segments of one skeleton. seg [MARKER_MAX_SEGMENTS][2][3] (start, end), slot_of_seg = fusion slot whose
// colour the segment carries (g_colors index). Returns the number of segments.
SES_HD int skeleton_segments(const SkeletonModel& model, int style, const ses3d_person_cov& person, double* seg,
                             int8_t* slot_of_seg) {
  int n = 0;
  auto put = [&](int from_slot, int to_slot) {   // from_slot < 0: the joint itself (degenerate segment)
    const ses3d_keypoint_cov& b = person.keypoints[to_slot];
    const ses3d_keypoint_cov& a = from_slot >= 0 ? person.keypoints[from_slot] : b;
    double* s = seg + (size_t)n * 6;
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = b.x; s[4] = b.y; s[5] = b.z;
    if (slot_of_seg) slot_of_seg[n] = (int8_t)to_slot;
    ++n;
  };
  auto has = [&](int slot) { return person.keypoints[slot].score > 0.0f; };
  if (style == MARKER_STYLE_SKELETON3D) {   // S3D:861-916: detector joints in order, parent from the body-part table
    for (int k = 0; k < NKP; ++k) {
      const int slot = model.fusion_idx[k];
      if (!has(slot)) continue;
      const int par = model.parent[k];
      put(par >= 0 && has(model.fusion_idx[par]) ? model.fusion_idx[par] : -1, slot);
    }
    return n;
  }
  // addJointToSkeleton PRI:273-382: fusion slots in ascending order; a referenced joint counts only when it was added
  // before, and every referenced slot has a smaller index than the referring one
  auto first_of = [&](int a, int b, int c) { return a >= 0 && has(a) ? a : (b >= 0 && has(b) ? b : (c >= 0 && has(c) ? c : -1)); };
  for (int s = 0; s < NFUS; ++s) {
    if (!has(s)) continue;
    switch (s) {
      case SES3D_FBP_NOSE: put(-1, s); break;
      case SES3D_FBP_HEAD: case SES3D_FBP_REYE: case SES3D_FBP_LEYE: case SES3D_FBP_NECK:
        put(first_of(SES3D_FBP_NOSE, -1, -1), s); break;
      case SES3D_FBP_RELBOW: case SES3D_FBP_RWRIST: case SES3D_FBP_LELBOW: case SES3D_FBP_LWRIST: case SES3D_FBP_RKNEE:
      case SES3D_FBP_RANKLE: case SES3D_FBP_LKNEE: case SES3D_FBP_LANKLE:
        put(first_of(s - 1, -1, -1), s); break;
      case SES3D_FBP_RSHOULDER: case SES3D_FBP_LSHOULDER: case SES3D_FBP_MIDHIP:
        put(first_of(SES3D_FBP_NECK, SES3D_FBP_NOSE, -1), s); break;
      case SES3D_FBP_BELLY:
        put(first_of(SES3D_FBP_NECK, -1, -1), s);
        put(first_of(SES3D_FBP_MIDHIP, -1, -1), s);
        break;
      case SES3D_FBP_RHIP: case SES3D_FBP_LHIP:
        put(first_of(SES3D_FBP_MIDHIP, SES3D_FBP_NECK, s - 7), s); break;
      case SES3D_FBP_REAR: case SES3D_FBP_LEAR:
        put(first_of(s - 2, -1, -1), s); break;
      default: break;
    }
  }
  return n;
}

}  // namespace ses3d
