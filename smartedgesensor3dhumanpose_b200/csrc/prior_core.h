// prior_core.h — the pose_prior stage for one message of one stream (kernel K7 "prior").
//
// Replaces skeletonCallback of pose_prior/src/pose_prior_mult_node.cpp (PRI:505-921) with its helpers:
// TrackingHypothesis::calc_normed_dist / calc_3d_dist (PRI:84-119), UnaryFactor (PRI:126-145),
// remove_old_tracks (PRI:191-211), addBinaryFactors (PRI:384-481), setInitialState (PRI:483-503), and the
// gtsam calls it makes (LevenbergMarquardtOptimizer PRI:746-749, Marginals PRI:760-789; gtsam 4.0.3 is not
// under /root/reference, its published algorithm with default parameters is restated — see
// oracle/pose_prior_oracle.cpp for the statement-by-statement CPU version this is checked against).
//
// B200 mapping. The stage is stateful per stream, so one CTA owns one stream ("sequence") and walks its frames
// in order; the tracker state lives in HBM/L2 between launches. Inside a frame
//  * the track/detection cost matrix is filled one thread per entry, the assignment is solved by the
//    warp-cooperative Munkres of assoc_core.h (same scan order as the reference's Hungarian.cpp);
//  * every detection is fitted by ONE WARP, lane k = skeleton joint k. The factor graph of a skeleton is a
//    forest (one bone per joint to its parent), and a range factor's Hessian block is rank one:
//    H_child,parent = -w w^T with w = (x_c - x_p) / (|x_c - x_p| sigma). The damped normal equations
//    (J^T J + lambda I) delta = -J^T e are therefore solved exactly by leaf-to-root elimination of 3x3 blocks
//    (no fill-in, each Schur complement is the scalar alpha = w^T D^-1 w times w w^T) and a root-to-leaf
//    back-substitution: 6 tree levels up, 6 down, all joints of a level in parallel, instead of a dense 57 x 57
//    Cholesky per LM trial. The marginal covariances follow from the same factorisation by the downward
//    recursion Sigma_c = D_c^-1 + (w^T Sigma_p w) z z^T, z = D_c^-1 w (no dense inverse);
//  * track pruning / merging is integer work by the leader on a pair-distance table computed in parallel.
// Everything is FP64 like gtsam; results are tolerance-checked against the oracle (different elimination order).
#pragma once
#include "assoc_core.h"
#include "common.h"
#include "team.h"

namespace ses3d {

constexpr int PRIOR_NAVG = 3;             // g_n_mov_avg PRI:53
constexpr double PRIOR_MAX_DIST = 1e6;    // MAX_DIST PRI:65
constexpr int PRIOR_MAX_TRACKS = 64;      // slot bitmap width
constexpr int PRIOR_LEVELS = 6;

struct PriorTrack {                       // TrackingHypothesis PRI:68-82
  double prev[NFUS][3];                   // prevEstimate (root-relative, height-normalised)
  double vel[NFUS][PRIOR_NAVG][3];        // velBuffer
  double t_prev, height_prev, root_prev[3];
  uint32_t exists;                        // bit k: prevEstimate holds joint k
  int32_t num_obs, id, pad_;
};

struct PriorSeqState {                    // file-scope state of one node instance
  double t_prev;                          // g_t_prev PRI:58
  double delay_buf[PRIOR_NAVG];           // g_fb_delay_buffer PRI:54
  unsigned long long used;                // bitmap of occupied track slots
  int32_t next_id, frame_nr, n_tracks;    // g_next_id, g_frame_nr PRI:59-60; g_tracks.size()
  int32_t overflow;                       // sticky: a frame needed more than max_tracks tracks
};

struct PriorTables {
  ses3d_prior_params prm;
  double limb_sigma_factor;               // PRI:934-937
};

SES_HD void prior_state_reset(const ses3d_prior_params& prm, PriorSeqState* st, bool keep_t_prev) {  // reset() PRI:182-189
  if (!keep_t_prev) st->t_prev = 0.0;     // static storage; reset() leaves g_t_prev alone
  for (int i = 0; i < PRIOR_NAVG; ++i) st->delay_buf[i] = prm.avg_delay;
  st->used = 0ull;
  st->next_id = 0; st->frame_nr = 0; st->n_tracks = 0; st->overflow = 0;
}

// ---- static skeleton forest (union of the bone tables PRI:384-481, rooted at MidHip) -------------------------
SES_HD int prior_level(int k) {
  const int8_t L[NFUS] = {3, 2, 3, 4, 5, 3, 4, 5, 0, 1, 2, 3, 1, 2, 3, 4, 4, 5, 5, 4, 1};
  return L[k];
}
SES_HD int prior_static_parent(int k) {   // Neck (1): Belly (20) when measured, else MidHip (8) — PRI:464-471
  const int8_t P[NFUS] = {1, 20, 1, 2, 3, 1, 5, 6, -1, 8, 9, 10, 8, 12, 13, 0, 0, 15, 16, 0, 8};
  return P[k];
}
SES_HD int prior_child(int k, int i) {    // i-th potential child of joint k, -1 = none
  const int8_t K[NFUS][4] = {{19, 15, 16, -1}, {0, 2, 5, -1},    {3, -1, -1, -1},  {4, -1, -1, -1},  {-1, -1, -1, -1},
                             {6, -1, -1, -1},  {7, -1, -1, -1},  {-1, -1, -1, -1}, {9, 12, 20, 1},   {10, -1, -1, -1},
                             {11, -1, -1, -1}, {-1, -1, -1, -1}, {13, -1, -1, -1}, {14, -1, -1, -1}, {-1, -1, -1, -1},
                             {17, -1, -1, -1}, {18, -1, -1, -1}, {-1, -1, -1, -1}, {-1, -1, -1, -1}, {-1, -1, -1, -1},
                             {1, -1, -1, -1}};
  return K[k][i];
}
// bone between joint k and its parent: length and sigma (absolute PRI:434-479 / height-normalised PRI:386-431);
// via_midhip selects the Simple-Baselines MidHip<->Neck bone for k = Neck
SES_HD void prior_bone(int k, bool normalised, bool via_midhip, double* len, double* sigma) {
  const double A[NFUS][2] = {{0.20, 0.025},  {0.25534, 0.035}, {0.15, 0.042},  {0.28, 0.045},  {0.25, 0.063},
                             {0.15, 0.042},  {0.28, 0.045},    {0.25, 0.063},  {0.0, 1.0},     {0.134, 0.033},
                             {0.449, 0.051}, {0.446, 0.051},   {0.134, 0.033}, {0.449, 0.051}, {0.446, 0.051},
                             {0.05, 0.035},  {0.05, 0.035},    {0.10, 0.05},   {0.10, 0.05},   {0.11500, 0.035},
                             {0.23846, 0.071}};
  const double N[NFUS][2] = {{0.33, 0.050},  {0.51, 0.05},   {0.262, 0.092}, {0.515, 0.071}, {0.444, 0.084},
                             {0.262, 0.092}, {0.515, 0.071}, {0.444, 0.084}, {0.0, 1.0},     {0.17, 0.062},
                             {0.694, 0.111}, {0.708, 0.097}, {0.17, 0.062},  {0.694, 0.111}, {0.708, 0.097},
                             {0.085, 0.06},  {0.085, 0.06},  {0.167, 0.08},  {0.167, 0.08},  {0.23, 0.05},
                             {0.49, 0.05}};
  if (k == SES3D_FBP_NECK && via_midhip) {
    *len = normalised ? 1.000 : 0.50;
    *sigma = normalised ? 0.02 : 0.071;
    return;
  }
  *len = normalised ? N[k][0] : A[k][0];
  *sigma = normalised ? N[k][1] : A[k][1];
}
SES_HD double prior_vel_sigma(int k) {    // FUSION_BODY_PARTS::vel_sigmas, fusion_body_parts.h:33
  const double V[NFUS] = {2., 1., 1., 2., 3., 1., 2., 3., 1., 1., 2., 3., 1., 2., 3., 2., 2., 2., 2., 2., 1.};
  return V[k];
}

// ---- workspaces ---------------------------------------------------------------------------------------------
struct PriorFitWs {       // one detection's factor graph, one per warp; arrays indexed by joint
  double *m, *x, *xn, *dl, *ub, *gu, *w, *z, *y, *pa;   // [21][3]
  double *R, *W, *Dinv, *Sg;                            // [21][6]  (00,01,02,11,12,22)
  double *e, *alpha, *beta;                             // [21]
  int8_t* par;                                          // [21] parent joint of the bone, -1 none
  uint8_t *msd, *usev;                                  // [21] measured / velocity usable
  int* scal;                                            // [4] {measured mask, factorisation failed, -, -}
};
enum { PF_MASK = 0, PF_FAIL = 1 };

template <class A>
SES_HD void prior_fit_ws_layout(A& ar, PriorFitWs* ws) {
  double* v3[10];
  for (int i = 0; i < 10; ++i) v3[i] = ar.template take<double>(NFUS * 3);
  double* v6[4];
  for (int i = 0; i < 4; ++i) v6[i] = ar.template take<double>(NFUS * 6);
  double* v1[3];
  for (int i = 0; i < 3; ++i) v1[i] = ar.template take<double>(NFUS);
  int* scal = ar.template take<int>(4);
  int8_t* par = ar.template take<int8_t>(NFUS);
  uint8_t* msd = ar.template take<uint8_t>(NFUS);
  uint8_t* usev = ar.template take<uint8_t>(NFUS);
  if (ws) {
    ws->m = v3[0]; ws->x = v3[1]; ws->xn = v3[2]; ws->dl = v3[3]; ws->ub = v3[4]; ws->gu = v3[5]; ws->w = v3[6];
    ws->z = v3[7]; ws->y = v3[8]; ws->pa = v3[9];
    ws->R = v6[0]; ws->W = v6[1]; ws->Dinv = v6[2]; ws->Sg = v6[3];
    ws->e = v1[0]; ws->alpha = v1[1]; ws->beta = v1[2];
    ws->scal = scal; ws->par = par; ws->msd = msd; ws->usev = usev;
  }
}
inline size_t prior_fit_ws_bytes() {
  ArenaSizer s;
  prior_fit_ws_layout(s, nullptr);
  return (s.used + 15) / 16 * 16;
}

struct PriorWs {          // per-stream frame workspace (shared memory of the CTA)
  double *cost, *dist;    // [h_max * max_tracks] column-major n_det x n_trk (PRI:551)
  double* D;              // [max_tracks * max_tracks] pair distances for the merge loop
  double* dscal;          // [4] {t, pred_delta_t}
  uint8_t *star, *prime, *nstar, *cov_r, *cov_c;
  int* assignment;        // [h_max]
  int* slot;              // [h_max] track slot of each detection, -1 = none (capacity)
  int* out_idx;           // [h_max] position in the published list or -1
  uint8_t *isnew, *has;   // [h_max]
  int* scal;              // [4] {n_trk at frame start, n_pub}
};
enum { PW_T = 0, PW_PDT = 1 };
enum { PS_NTRK = 0, PS_NPUB = 1 };

template <class A>
SES_HD void prior_ws_layout(A& ar, int h_max, int max_tracks, PriorWs* ws) {
  double* cost = ar.template take<double>((size_t)h_max * max_tracks);
  double* dist = ar.template take<double>((size_t)h_max * max_tracks);
  double* D = ar.template take<double>((size_t)max_tracks * max_tracks);
  double* dscal = ar.template take<double>(4);
  int* assignment = ar.template take<int>(h_max);
  int* slot = ar.template take<int>(h_max);
  int* out_idx = ar.template take<int>(h_max);
  int* scal = ar.template take<int>(4);
  uint8_t* star = ar.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* prime = ar.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* nstar = ar.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* cov_r = ar.template take<uint8_t>(h_max);
  uint8_t* cov_c = ar.template take<uint8_t>(max_tracks);
  uint8_t* isnew = ar.template take<uint8_t>(h_max);
  uint8_t* has = ar.template take<uint8_t>(h_max);
  if (ws) {
    ws->cost = cost; ws->dist = dist; ws->D = D; ws->dscal = dscal; ws->assignment = assignment; ws->slot = slot;
    ws->out_idx = out_idx; ws->scal = scal; ws->star = star; ws->prime = prime; ws->nstar = nstar; ws->cov_r = cov_r;
    ws->cov_c = cov_c; ws->isnew = isnew; ws->has = has;
  }
}
inline size_t prior_ws_bytes(int h_max, int max_tracks) {
  ArenaSizer s;
  prior_ws_layout(s, h_max, max_tracks, nullptr);
  return (s.used + 15) / 16 * 16;
}

// ---- 3x3 helpers --------------------------------------------------------------------------------------------
// sqrt-information of noiseModel::Gaussian::Covariance(S): diagonal S -> 1/sigma; else upper Cholesky factor of S^-1.
// S, R as (00,01,02,11,12,22); R upper triangular.
SES_HD void prior_sqrt_information(const double S[6], double R[6]) {
  if (S[1] == 0.0 && S[2] == 0.0 && S[4] == 0.0) {
    R[0] = 1.0 / sqrt(S[0]); R[1] = 0.0; R[2] = 0.0; R[3] = 1.0 / sqrt(S[3]); R[4] = 0.0; R[5] = 1.0 / sqrt(S[5]);
    return;
  }
  const double c00 = S[3] * S[5] - S[4] * S[4], c01 = S[4] * S[2] - S[1] * S[5], c02 = S[1] * S[4] - S[3] * S[2];
  const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
  const double id = 1.0 / det;
  const double i00 = c00 * id, i10 = c01 * id, i20 = c02 * id;
  const double i11 = (S[0] * S[5] - S[2] * S[2]) * id, i21 = (S[1] * S[2] - S[0] * S[4]) * id;
  const double i22 = (S[0] * S[3] - S[1] * S[1]) * id;
  const double l00 = sqrt(i00);
  const double l10 = i10 / l00, l20 = i20 / l00;
  const double l11 = sqrt(i11 - l10 * l10);
  const double l21 = (i21 - l20 * l10) / l11;
  const double l22 = sqrt(i22 - l20 * l20 - l21 * l21);
  R[0] = l00; R[1] = l10; R[2] = l20; R[3] = l11; R[4] = l21; R[5] = l22;
}
SES_HD void sym6_mul(const double A[6], const double v[3], double o[3]) {
  o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  o[1] = A[1] * v[0] + A[3] * v[1] + A[4] * v[2];
  o[2] = A[2] * v[0] + A[4] * v[1] + A[5] * v[2];
}
// inverse of a symmetric positive definite 3x3; false when a Cholesky pivot is not positive / not finite
SES_HD bool sym6_inverse_spd(const double D[6], double I[6]) {
  const double p0 = D[0];
  if (!(p0 > 0.0)) return false;
  const double l10 = D[1] / p0, l20 = D[2] / p0;
  const double p1 = D[3] - l10 * D[1];
  if (!(p1 > 0.0)) return false;
  const double t21 = D[4] - l20 * D[1];
  const double l21 = t21 / p1;
  const double p2 = D[5] - l20 * D[2] - l21 * t21;
  if (!(p2 > 0.0) || !(p2 < DBL_MAX)) return false;
  // D = L diag(p) L^T  =>  D^-1 = L^-T diag(1/p) L^-1, L^-1 = [[1,0,0],[-l10,1,0],[l10 l21 - l20, -l21, 1]]
  const double q0 = 1.0 / p0, q1 = 1.0 / p1, q2 = 1.0 / p2;
  const double a = l10 * l21 - l20;
  I[5] = q2;
  I[4] = -l21 * q2;
  I[2] = a * q2;
  I[3] = q1 + l21 * l21 * q2;
  I[1] = -l10 * q1 - l21 * a * q2;
  I[0] = q0 + l10 * l10 * q1 + a * a * q2;
  return true;
}

// ---- tracking helpers ---------------------------------------------------------------------------------------
SES_HD double prior_normed_dist(const PriorTables& pt, const PriorTrack& tr, const ses3d_person_cov& person, double t) {
  const double delta_t = t - tr.t_prev;  // PRI:84-101
  int used = 0;
  double dist = 0;
  for (int k = 0; k < NFUS; ++k) {
    const ses3d_keypoint_cov& kp = person.keypoints[k];
    if (kp.score > pt.prm.min_score && ((tr.exists >> k) & 1u)) {
      const double dx = kp.x - (tr.prev[k][0] * tr.height_prev + tr.root_prev[0]);
      const double dy = kp.y - (tr.prev[k][1] * tr.height_prev + tr.root_prev[1]);
      const double dz = kp.z - (tr.prev[k][2] * tr.height_prev + tr.root_prev[2]);
      dist += sqrt(dx * dx + dy * dy + dz * dz) / (prior_vel_sigma(k) * delta_t);
      ++used;
    }
  }
  return used > 0 ? dist / used : PRIOR_MAX_DIST;
}
SES_HD double prior_track_dist(const PriorTrack& a, const PriorTrack& b) {  // calc_3d_dist PRI:103-119
  int used = 0;
  double dist = 0;
  const uint32_t both = a.exists & b.exists;
  for (int k = 0; k < NFUS; ++k) {
    if (!((both >> k) & 1u)) continue;
    double d2 = 0;
    for (int i = 0; i < 3; ++i) {
      const double d = (a.prev[k][i] * a.height_prev + a.root_prev[i]) - (b.prev[k][i] * b.height_prev + b.root_prev[i]);
      d2 += d * d;
    }
    dist += sqrt(d2);
    ++used;
  }
  return used > 0 ? dist / used : PRIOR_MAX_DIST;
}
SES_HD double prior_stamp_to_sec(int64_t ns) {  // ros::Time::toSec()
  const int64_t sec = ns / 1000000000LL, nsec = ns % 1000000000LL;
  return (double)sec + 1e-9 * (double)nsec;
}

// root / neck of a detection (PRI:631-656) — pure function of the record, evaluated redundantly by every thread
struct PriorRootNeck {
  double root[3], neck[3];
  float root_score, neck_score;
  double height;
};
SES_HD PriorRootNeck prior_root_neck(const PriorTables& pt, const ses3d_person_cov& person) {
  PriorRootNeck rn;
  rn.root[0] = rn.root[1] = rn.root[2] = 0.0;
  rn.neck[0] = rn.neck[1] = rn.neck[2] = 0.0;
  rn.root_score = 0.f; rn.neck_score = 0.f;
  rn.height = 1.0;
  if (pt.prm.pose_method == SES3D_POSE_H36M) {
    const ses3d_keypoint_cov& r = person.keypoints[SES3D_FBP_MIDHIP];
    const ses3d_keypoint_cov& n = person.keypoints[SES3D_FBP_NECK];
    rn.root[0] = r.x; rn.root[1] = r.y; rn.root[2] = r.z; rn.root_score = r.score;
    rn.neck[0] = n.x; rn.neck[1] = n.y; rn.neck[2] = n.z; rn.neck_score = n.score;
  } else {
    const ses3d_keypoint_cov& hl = person.keypoints[SES3D_FBP_LHIP];
    const ses3d_keypoint_cov& hr = person.keypoints[SES3D_FBP_RHIP];
    const ses3d_keypoint_cov& sl = person.keypoints[SES3D_FBP_LSHOULDER];
    const ses3d_keypoint_cov& sr = person.keypoints[SES3D_FBP_RSHOULDER];
    if (hl.score > 0.0f && hr.score > 0.0f) {
      rn.root[0] = (hl.x + hr.x) / 2.0; rn.root[1] = (hl.y + hr.y) / 2.0; rn.root[2] = (hl.z + hr.z) / 2.0;
      rn.root_score = (hl.score + hr.score) / 2.0f;
    }
    if (sl.score > 0.0f && sr.score > 0.0f) {
      rn.neck[0] = (sl.x + sr.x) / 2.0; rn.neck[1] = (sl.y + sr.y) / 2.0; rn.neck[2] = (sl.z + sr.z) / 2.0;
      rn.neck_score = (sl.score + sr.score) / 2.0f;
    }
  }
  if (rn.root_score > pt.prm.min_score && pt.prm.normalize_by_height) {  // PRI:658-668
    if (rn.neck_score > pt.prm.min_score) {
      const double dx = rn.neck[0] - rn.root[0], dy = rn.neck[1] - rn.root[1], dz = rn.neck[2] - rn.root[2];
      rn.height = sqrt(dx * dx + dy * dy + dz * dz);
    } else {
      rn.height = 0.60;
    }
  }
  return rn;
}
// does the detection contribute at least one factor (num_meas > 0, PRI:739-741)?
SES_HD bool prior_has_measurement(const PriorTables& pt, const ses3d_person_cov& person) {
  const PriorRootNeck rn = prior_root_neck(pt, person);
  if (rn.root_score > pt.prm.min_score) return true;
  if (pt.prm.pose_method == SES3D_POSE_SIMPLE && rn.neck_score > pt.prm.min_score) return true;
  for (int k = 0; k < NFUS; ++k)
    if (k != SES3D_FBP_MIDHIP && person.keypoints[k].score > pt.prm.min_score) return true;
  return false;
}

// ---- the skeleton fit of one detection (PRI:587-853), run by a warp-sized team ---------------------------------
// unary + bone error of joint k at the point xs (NonlinearFactorGraph::error summand)
SES_HD double prior_joint_error(const PriorTables& pt, const PriorFitWs& ws, const double* xs, int k) {
  if (!ws.msd[k]) return 0.0;
  const double* R = ws.R + 6 * k;
  const double d0 = xs[3 * k] - ws.m[3 * k], d1 = xs[3 * k + 1] - ws.m[3 * k + 1], d2 = xs[3 * k + 2] - ws.m[3 * k + 2];
  const double w0 = R[0] * d0 + R[1] * d1 + R[2] * d2, w1 = R[3] * d1 + R[4] * d2, w2 = R[5] * d2;
  double err = 0.5 * (w0 * w0 + w1 * w1 + w2 * w2);
  const int p = ws.par[k];
  if (p >= 0) {
    double len, sigma;
    prior_bone(k, pt.prm.normalize_by_height != 0, p == SES3D_FBP_MIDHIP, &len, &sigma);
    const double dx = xs[3 * k] - xs[3 * p], dy = xs[3 * k + 1] - xs[3 * p + 1], dz = xs[3 * k + 2] - xs[3 * p + 2];
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const double e = (r - len) * (1.0 / (sigma * pt.limb_sigma_factor));   // Isotropic::whiten: v * invsigma
    err += 0.5 * (e * e);
  }
  return err;
}

// linearise at ws.x: whitened unary residual ub, its gradient gu = R^T ub, bone direction w and residual e
template <class WT>
SES_HD void prior_linearize(WT& tm, const PriorTables& pt, const PriorFitWs& ws) {
  tm.pfor(NFUS, [&](int k) {
    if (!ws.msd[k]) return;
    const double* R = ws.R + 6 * k;
    const double* x = ws.x + 3 * k;
    const double d0 = x[0] - ws.m[3 * k], d1 = x[1] - ws.m[3 * k + 1], d2 = x[2] - ws.m[3 * k + 2];
    const double u0 = R[0] * d0 + R[1] * d1 + R[2] * d2, u1 = R[3] * d1 + R[4] * d2, u2 = R[5] * d2;
    ws.ub[3 * k] = u0; ws.ub[3 * k + 1] = u1; ws.ub[3 * k + 2] = u2;
    ws.gu[3 * k] = R[0] * u0;
    ws.gu[3 * k + 1] = R[1] * u0 + R[3] * u1;
    ws.gu[3 * k + 2] = R[2] * u0 + R[4] * u1 + R[5] * u2;
    const int p = ws.par[k];
    double w0 = 0, w1 = 0, w2 = 0, e = 0;
    if (p >= 0) {
      double len, sigma;
      prior_bone(k, pt.prm.normalize_by_height != 0, p == SES3D_FBP_MIDHIP, &len, &sigma);
      const double is = 1.0 / (sigma * pt.limb_sigma_factor);
      const double dx = x[0] - ws.x[3 * p], dy = x[1] - ws.x[3 * p + 1], dz = x[2] - ws.x[3 * p + 2];
      const double r = sqrt(dx * dx + dy * dy + dz * dz);
      const double s = is / r;
      w0 = dx * s; w1 = dy * s; w2 = dz * s;
      e = (r - len) * is;
    }
    ws.w[3 * k] = w0; ws.w[3 * k + 1] = w1; ws.w[3 * k + 2] = w2;
    ws.e[k] = e;
  });
}

// leaf-to-root elimination of (J^T J + lambda I): per joint Dinv, z = Dinv w, y = Dinv b, alpha = w.z, beta = w.y.
// Sets ws.scal[PF_FAIL] when a block is not positive definite (gtsam: IndeterminantLinearSystemException).
template <class WT>
SES_HD void prior_eliminate(WT& tm, const PriorFitWs& ws, double lambda) {
  tm.single([&] { ws.scal[PF_FAIL] = 0; });
  for (int L = PRIOR_LEVELS - 1; L >= 0; --L) {
    tm.pfor(NFUS, [&](int k) {
      if (!ws.msd[k] || prior_level(k) != L) return;
      const double* W = ws.W + 6 * k;
      double D[6] = {W[0] + lambda, W[1], W[2], W[3] + lambda, W[4], W[5] + lambda};
      double b[3] = {-ws.gu[3 * k], -ws.gu[3 * k + 1], -ws.gu[3 * k + 2]};
      if (ws.par[k] >= 0) {
        const double* w = ws.w + 3 * k;
        const double e = ws.e[k];
        D[0] += w[0] * w[0]; D[1] += w[0] * w[1]; D[2] += w[0] * w[2];
        D[3] += w[1] * w[1]; D[4] += w[1] * w[2]; D[5] += w[2] * w[2];
        b[0] -= w[0] * e; b[1] -= w[1] * e; b[2] -= w[2] * e;
      }
      for (int i = 0; i < 4; ++i) {
        const int c = prior_child(k, i);
        if (c < 0) break;
        if (!ws.msd[c] || ws.par[c] != k) continue;
        const double* w = ws.w + 3 * c;
        const double f = 1.0 - ws.alpha[c];        // bone term w w^T minus the child's Schur complement alpha w w^T
        const double g = ws.e[c] + ws.beta[c];     // -(-w e) from the bone gradient, + beta w from the elimination
        D[0] += f * w[0] * w[0]; D[1] += f * w[0] * w[1]; D[2] += f * w[0] * w[2];
        D[3] += f * w[1] * w[1]; D[4] += f * w[1] * w[2]; D[5] += f * w[2] * w[2];
        b[0] += g * w[0]; b[1] += g * w[1]; b[2] += g * w[2];
      }
      double* I = ws.Dinv + 6 * k;
      if (!sym6_inverse_spd(D, I)) {
        ws.scal[PF_FAIL] = 1;
        I[0] = I[3] = I[5] = 1.0; I[1] = I[2] = I[4] = 0.0;
      }
      sym6_mul(I, ws.w + 3 * k, ws.z + 3 * k);
      sym6_mul(I, b, ws.y + 3 * k);
      const double* w = ws.w + 3 * k;
      ws.alpha[k] = w[0] * ws.z[3 * k] + w[1] * ws.z[3 * k + 1] + w[2] * ws.z[3 * k + 2];
      ws.beta[k] = w[0] * ws.y[3 * k] + w[1] * ws.y[3 * k + 1] + w[2] * ws.y[3 * k + 2];
    });
  }
}

// root-to-leaf back-substitution: delta_k = y_k + z_k (w_k . delta_parent)
template <class WT>
SES_HD void prior_backsubstitute(WT& tm, const PriorFitWs& ws) {
  for (int L = 0; L < PRIOR_LEVELS; ++L) {
    tm.pfor(NFUS, [&](int k) {
      if (!ws.msd[k] || prior_level(k) != L) return;
      double s = 0.0;
      const int p = ws.par[k];
      if (p >= 0) s = ws.w[3 * k] * ws.dl[3 * p] + ws.w[3 * k + 1] * ws.dl[3 * p + 1] + ws.w[3 * k + 2] * ws.dl[3 * p + 2];
      for (int i = 0; i < 3; ++i) ws.dl[3 * k + i] = ws.y[3 * k + i] + ws.z[3 * k + i] * s;
    });
  }
}

// marginal covariances after prior_eliminate(lambda = 0): Sigma_k = Dinv_k + (w_k^T Sigma_p w_k) z_k z_k^T
template <class WT>
SES_HD void prior_marginals(WT& tm, const PriorFitWs& ws) {
  for (int L = 0; L < PRIOR_LEVELS; ++L) {
    tm.pfor(NFUS, [&](int k) {
      if (!ws.msd[k] || prior_level(k) != L) return;
      const double* I = ws.Dinv + 6 * k;
      double* S = ws.Sg + 6 * k;
      double g = 0.0;
      const int p = ws.par[k];
      if (p >= 0) {
        double t[3];
        sym6_mul(ws.Sg + 6 * p, ws.w + 3 * k, t);
        g = ws.w[3 * k] * t[0] + ws.w[3 * k + 1] * t[1] + ws.w[3 * k + 2] * t[2];
      }
      const double* z = ws.z + 3 * k;
      S[0] = I[0] + g * z[0] * z[0]; S[1] = I[1] + g * z[0] * z[1]; S[2] = I[2] + g * z[0] * z[2];
      S[3] = I[3] + g * z[1] * z[1]; S[4] = I[4] + g * z[1] * z[2]; S[5] = I[5] + g * z[2] * z[2];
    });
  }
}

// GaussianFactorGraph::error(delta) of the undamped linearised system
template <class WT>
SES_HD double prior_linear_error(WT& tm, const PriorFitWs& ws, bool at_zero) {
  return tm.sum(NFUS, [&](int k) -> double {
    if (!ws.msd[k]) return 0.0;
    double u0 = ws.ub[3 * k], u1 = ws.ub[3 * k + 1], u2 = ws.ub[3 * k + 2], e = ws.e[k];
    if (!at_zero) {
      const double* R = ws.R + 6 * k;
      const double* d = ws.dl + 3 * k;
      u0 += R[0] * d[0] + R[1] * d[1] + R[2] * d[2];
      u1 += R[3] * d[1] + R[4] * d[2];
      u2 += R[5] * d[2];
      const int p = ws.par[k];
      if (p >= 0) {
        const double* w = ws.w + 3 * k;
        e += w[0] * (d[0] - ws.dl[3 * p]) + w[1] * (d[1] - ws.dl[3 * p + 1]) + w[2] * (d[2] - ws.dl[3 * p + 2]);
      }
    }
    return 0.5 * (u0 * u0 + u1 * u1 + u2 * u2) + 0.5 * (e * e);
  });
}

// One detection. tr = its track (exclusively owned by this team during the call). fused / pred may be null
// (track not published yet, PRI:845-848). Returns false when the detection has no usable joint (PRI:739-741).
template <class WT>
SES_HD bool prior_fit_person(WT& tm, const PriorTables& pt, const ses3d_person_cov& person, PriorTrack& tr, double t,
                             double t_prev_global, int frame_nr, double pred_delta_t, const PriorFitWs& ws,
                             ses3d_person_cov* fused, ses3d_person_cov* pred) {
  const ses3d_prior_params& q = pt.prm;
  const PriorRootNeck rn = prior_root_neck(pt, person);
  const double height = rn.height;
  const bool simple = q.pose_method == SES3D_POSE_SIMPLE;

  // measurements and their sqrt-information (PRI:658-737)
  tm.pfor(NFUS, [&](int k) {
    bool ms = false;
    double c[6], mk[3] = {0.0, 0.0, 0.0};
    if (k == SES3D_FBP_MIDHIP) {
      if (rn.root_score > q.min_score) {
        ms = true;
        if (simple) {
          const double* a = person.keypoints[SES3D_FBP_LHIP].cov;
          const double* b = person.keypoints[SES3D_FBP_RHIP].cov;
          for (int i = 0; i < 6; ++i) c[i] = (a[i] + b[i]) / 2.0;
        } else {
          for (int i = 0; i < 6; ++i) c[i] = person.keypoints[SES3D_FBP_MIDHIP].cov[i];
        }
        for (int i = 0; i < 6; ++i) c[i] = c[i] / (height * height) / (q.root_sigma_factor * q.root_sigma_factor);
      }
    } else {
      const ses3d_keypoint_cov& kp = person.keypoints[k];
      if (kp.score > q.min_score) {
        ms = true;
        for (int i = 0; i < 6; ++i) c[i] = kp.cov[i] / (height * height);
        mk[0] = (kp.x - rn.root[0]) / height; mk[1] = (kp.y - rn.root[1]) / height; mk[2] = (kp.z - rn.root[2]) / height;
      }
      if (k == SES3D_FBP_NECK && simple && rn.neck_score > q.min_score) {
        ms = true;
        const double* a = person.keypoints[SES3D_FBP_LSHOULDER].cov;
        const double* b = person.keypoints[SES3D_FBP_RSHOULDER].cov;
        for (int i = 0; i < 6; ++i) c[i] = (a[i] + b[i]) / 2.0 / (height * height);
        for (int i = 0; i < 3; ++i) mk[i] = (rn.neck[i] - rn.root[i]) / height;
      }
    }
    ws.msd[k] = ms ? 1 : 0;
    if (ms) {
      double* R = ws.R + 6 * k;
      prior_sqrt_information(c, R);
      double* W = ws.W + 6 * k;                     // R^T R, constant over the LM iterations
      W[0] = R[0] * R[0]; W[1] = R[0] * R[1]; W[2] = R[0] * R[2];
      W[3] = R[1] * R[1] + R[3] * R[3]; W[4] = R[1] * R[2] + R[3] * R[4];
      W[5] = R[2] * R[2] + R[4] * R[4] + R[5] * R[5];
      ws.m[3 * k] = mk[0]; ws.m[3 * k + 1] = mk[1]; ws.m[3 * k + 2] = mk[2];
    }
  });
  tm.single([&] {
    int mask = 0;
    for (int k = 0; k < NFUS; ++k) mask |= ws.msd[k] ? (1 << k) : 0;
    ws.scal[PF_MASK] = mask;
    if (tr.height_prev < 0.0) {  // PRI:699-702
      tr.height_prev = height;
      tr.root_prev[0] = rn.root[0]; tr.root_prev[1] = rn.root[1]; tr.root_prev[2] = rn.root[2];
    }
  });
  const uint32_t mask = (uint32_t)ws.scal[PF_MASK];
  if (mask == 0) return false;

  // setInitialState (PRI:483-503) + bone parents (addBinaryFactors PRI:384-481)
  const uint32_t had = tr.exists;
  const double height_prev = tr.height_prev;
  const double root_prev[3] = {tr.root_prev[0], tr.root_prev[1], tr.root_prev[2]};
  tm.pfor(NFUS, [&](int k) {
    const bool ex = (had >> k) & 1u, ms = (mask >> k) & 1u;
    if (ex && !ms)
      for (int b = 0; b < PRIOR_NAVG; ++b) tr.vel[k][b][0] = tr.vel[k][b][1] = tr.vel[k][b][2] = 0.0;
    ws.usev[k] = (ex && ms) ? 1 : 0;
    int par = -1;
    if (ms) {
      for (int i = 0; i < 3; ++i) {
        const double v = ex ? tr.prev[k][i] : ws.m[3 * k + i];
        ws.x[3 * k + i] = v;
        ws.pa[3 * k + i] = v * height_prev + root_prev[i];   // previous absolute position (velocity, PRI:820-821)
      }
      par = prior_static_parent(k);
      if (k == SES3D_FBP_NECK && !((mask >> SES3D_FBP_BELLY) & 1u)) par = SES3D_FBP_MIDHIP;
      if (par >= 0 && !((mask >> par) & 1u)) par = -1;
    }
    ws.par[k] = (int8_t)par;
  });

  // LevenbergMarquardtOptimizer(graph, prevEstimate).optimize(), default parameters
  double error = tm.sum(NFUS, [&](int k) { return prior_joint_error(pt, ws, ws.x, k); });
  double lambda = q.lm_lambda_initial;
  int iterations = 0;
  if (error > 0.0 && iterations < q.lm_max_iterations) {
    for (int guard = 0; guard < 4 * q.lm_max_iterations + 64; ++guard) {
      const double current_error = error;
      prior_linearize(tm, pt, ws);
      const double old_lin = prior_linear_error(tm, ws, true);
      for (;;) {  // tryLambda until it reports "stop"
        bool step_ok = false, stop_searching = false;
        double new_error = DBL_MAX;
        prior_eliminate(tm, ws, lambda);
        if (!ws.scal[PF_FAIL]) {
          prior_backsubstitute(tm, ws);
          const double new_lin = prior_linear_error(tm, ws, false);
          const double lin_change = old_lin - new_lin;
          if (lin_change >= 0) {
            tm.pfor(NFUS * 3, [&](int i) { if (ws.msd[i / 3]) ws.xn[i] = ws.x[i] + ws.dl[i]; });
            new_error = tm.sum(NFUS, [&](int k) { return prior_joint_error(pt, ws, ws.xn, k); });
            const double cost_change = error - new_error;
            if (lin_change > 1e-20) step_ok = (cost_change / lin_change) > q.lm_min_model_fidelity;
            if (fabs(cost_change) < q.lm_relative_error_tol * error) stop_searching = true;
          }
        }
        tm.sync();
        if (step_ok) {
          tm.pfor(NFUS * 3, [&](int i) { if (ws.msd[i / 3]) ws.x[i] = ws.xn[i]; });
          error = new_error;
          lambda = lambda / q.lm_lambda_factor;
          if (lambda < 0.0) lambda = 0.0;
          ++iterations;
          break;
        } else if (!stop_searching) {
          lambda *= q.lm_lambda_factor;
          if (lambda >= q.lm_lambda_upper_bound) break;
        } else {
          break;
        }
      }
      bool converged;
      if (error <= 0.0) converged = true;
      else {
        const double abs_dec = current_error - error;
        const double rel_dec = abs_dec / current_error;
        converged = (q.lm_relative_error_tol != 0.0 && rel_dec <= q.lm_relative_error_tol) ||
                    abs_dec <= q.lm_absolute_error_tol;
      }
      if (!(iterations < q.lm_max_iterations && !converged && current_error < DBL_MAX && current_error == current_error))
        break;
    }
  }

  // Marginals(graph, result) (PRI:760-767)
  prior_linearize(tm, pt, ws);
  prior_eliminate(tm, ws, 0.0);
  const bool use_marginals = !ws.scal[PF_FAIL];
  if (use_marginals) prior_marginals(tm, ws);

  // outputs (PRI:770-837) and the track update (PRI:839-843)
  if (fused) {
    double* fz = reinterpret_cast<double*>(fused);
    double* pz = reinterpret_cast<double*>(pred);
    tm.pfor((int)(sizeof(ses3d_person_cov) / 8), [&](int i) { fz[i] = 0.0; pz[i] = 0.0; });
  }
  const double dtg = t - t_prev_global;
  const int vslot = frame_nr % PRIOR_NAVG;
  tm.pfor(NFUS, [&](int k) {
    if (!((mask >> k) & 1u)) return;
    double jf[3], jp[3];
    for (int i = 0; i < 3; ++i) jf[i] = ws.x[3 * k + i] * height + rn.root[i];
    for (int i = 0; i < 3; ++i) jp[i] = jf[i];
    if (ws.usev[k]) {
      for (int i = 0; i < 3; ++i) tr.vel[k][vslot][i] = (jf[i] - ws.pa[3 * k + i]) / dtg;
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;
        for (int b = 0; b < PRIOR_NAVG; ++b) acc += tr.vel[k][b][i];
        jp[i] += acc / PRIOR_NAVG * pred_delta_t;
      }
    }
    for (int i = 0; i < 3; ++i) tr.prev[k][i] = ws.x[3 * k + i];
    if (!fused) return;
    float score;
    if (k == SES3D_FBP_MIDHIP) score = rn.root_score;
    else if (k == SES3D_FBP_NECK) score = rn.neck_score;
    else score = person.keypoints[k].score;
    if (!(score > q.min_score)) score = q.min_score;   // std::max(g_min_score, score)
    double cv[6];
    if (use_marginals) {
      for (int i = 0; i < 6; ++i) cv[i] = ws.Sg[6 * k + i] * height * height;
    } else {
      const double d = q.default_res_sigma * q.default_res_sigma;
      cv[0] = d; cv[1] = 0; cv[2] = 0; cv[3] = d; cv[4] = 0; cv[5] = d;
    }
    if (k == SES3D_FBP_MIDHIP)
      for (int i = 0; i < 6; ++i) cv[i] *= (q.root_sigma_factor * q.root_sigma_factor);
    ses3d_keypoint_cov& o = fused->keypoints[k];
    o.x = jf[0]; o.y = jf[1]; o.z = jf[2]; o.score = score;
    for (int i = 0; i < 6; ++i) o.cov[i] = cv[i];
    ses3d_keypoint_cov& po = pred->keypoints[k];
    po.x = jp[0]; po.y = jp[1]; po.z = jp[2]; po.score = score;
    const double pn = q.pred_noise_sigma * q.pred_noise_sigma;
    po.cov[0] = cv[0] + pn; po.cov[1] = cv[1]; po.cov[2] = cv[2]; po.cov[3] = cv[3] + pn; po.cov[4] = cv[4];
    po.cov[5] = cv[5] + pn;
  });
  tm.single([&] {
    tr.t_prev = t;
    tr.exists = mask;
    tr.height_prev = height;
    tr.root_prev[0] = rn.root[0]; tr.root_prev[1] = rn.root[1]; tr.root_prev[2] = rn.root[2];
    ++tr.num_obs;
    if (fused) { fused->id = (uint32_t)tr.id; pred->id = (uint32_t)tr.id; }
  });
  return true;
}

// remove_old_tracks (PRI:191-211): compact the order list, free the slots. Leader only.
SES_HD void prior_remove_old(const PriorTables& pt, PriorSeqState* st, const PriorTrack* tracks, uint8_t* order, double t) {
  int n = 0;
  for (int i = 0; i < st->n_tracks; ++i) {
    const int s = order[i];
    if (t - tracks[s].t_prev > pt.prm.t_max_unobserved) st->used &= ~(1ull << s);
    else order[n++] = (uint8_t)s;
  }
  st->n_tracks = n;
}

// One message of one stream: skeletonCallback PRI:505-921. tm = the stream's team (CTA / serial).
// persons [n_det], fused / pred [h_max]; fit_ws = base of the per-warp fit workspaces (fit_ws_stride bytes apart).
template <class Team>
SES_HD void prior_frame(Team& tm, const PriorTables& pt, int max_tracks, int h_max, PriorSeqState* st,
                        PriorTrack* tracks, uint8_t* order, const PriorWs& ws, unsigned char* fit_ws, size_t fit_ws_stride,
                        int64_t stamp_ns, int n_cams, const float* fb_delay, int n_det_in,
                        const ses3d_person_cov* persons, ses3d_person_cov* fused, ses3d_person_cov* pred,
                        int32_t* n_out, float* pred_delay, int32_t* track_of) {
  const ses3d_prior_params& q = pt.prm;
  const int n_det = n_det_in < 0 ? 0 : (n_det_in > h_max ? h_max : n_det_in);
  tm.single([&] {
    const double t = prior_stamp_to_sec(stamp_ns);
    double curr = 0.0;  // PRI:513-526
    int n_valid = 0;
    for (int c = 0; c < n_cams; ++c) {
      const float d = fb_delay ? fb_delay[c] : -1.0f;
      if (d > 0.0f) { curr += (double)d; ++n_valid; }
    }
    if (n_valid > 0) curr /= n_valid; else curr = q.avg_delay;
    st->delay_buf[st->frame_nr % PRIOR_NAVG] = curr;
    double acc = 0.0;
    for (int i = 0; i < PRIOR_NAVG; ++i) acc += st->delay_buf[i];
    ws.dscal[PW_T] = t;
    ws.dscal[PW_PDT] = acc / PRIOR_NAVG;
    if (pred_delay) *pred_delay = (float)ws.dscal[PW_PDT];
    ws.scal[PS_NTRK] = st->n_tracks;
    ws.scal[PS_NPUB] = 0;
  });
  const double t = ws.dscal[PW_T], pred_delta_t = ws.dscal[PW_PDT];
  const int n_trk = ws.scal[PS_NTRK];
  if (track_of) tm.pfor(h_max, [&](int i) { track_of[i] = -1; });

  if (n_det == 0) {  // PRI:537-546
    tm.single([&] {
      prior_remove_old(pt, st, tracks, order, t);
      st->t_prev = t;
      *n_out = 0;
    });
    return;
  }

  // association of detections to tracks (PRI:548-568)
  if (n_trk > 0) {
    tm.pfor(n_det * n_trk, [&](int e) {
      const int p = e % n_det, tr = e / n_det;
      ws.cost[e] = prior_normed_dist(pt, tracks[order[tr]], persons[p], t);
    });
    tm.warp0([&](auto& w) {
      AssocWs aws;
      aws.dist = ws.dist; aws.star = ws.star; aws.prime = ws.prime; aws.nstar = ws.nstar; aws.cov_r = ws.cov_r;
      aws.cov_c = ws.cov_c;
      munkres_coop(w, aws, ws.cost, n_det, n_trk, ws.assignment);
    });
    tm.pfor(n_det, [&](int p) {
      const int a = ws.assignment[p];
      if (a >= 0 && ws.cost[p + n_det * a] > q.dist_threshold) ws.assignment[p] = -1;
    });
  } else {
    tm.pfor(n_det, [&](int p) { ws.assignment[p] = -1; });
  }

  // new tracks for unassigned detections, in detection order (PRI:570-580)
  tm.single([&] {
    for (int p = 0; p < n_det; ++p) {
      ws.isnew[p] = 0;
      if (ws.assignment[p] >= 0) { ws.slot[p] = order[ws.assignment[p]]; continue; }
      int s = -1;
      for (int i = 0; i < max_tracks; ++i)
        if (!((st->used >> i) & 1ull)) { s = i; break; }
      if (s < 0 || st->n_tracks >= max_tracks) { st->overflow = 1; ws.slot[p] = -1; continue; }
      st->used |= 1ull << s;
      order[st->n_tracks++] = (uint8_t)s;
      PriorTrack& tr = tracks[s];
      tr.t_prev = -DBL_MAX;      // "never observed" (the reference leaves t_prev uninitialised, PRI:79-82)
      tr.height_prev = -1.0;
      tr.root_prev[0] = tr.root_prev[1] = tr.root_prev[2] = 0.0;
      tr.exists = 0u; tr.num_obs = 0; tr.id = st->next_id++;
      ws.slot[p] = s;
      ws.isnew[p] = 1;
    }
  });
  tm.pfor(n_det * NFUS * PRIOR_NAVG * 3, [&](int i) {
    const int p = i / (NFUS * PRIOR_NAVG * 3);
    if (ws.isnew[p]) (&tracks[ws.slot[p]].vel[0][0][0])[i % (NFUS * PRIOR_NAVG * 3)] = 0.0;
  });
  tm.pfor(n_det, [&](int p) { ws.has[p] = prior_has_measurement(pt, persons[p]) ? 1 : 0; });
  tm.single([&] {  // persons are published in detection order once their track has enough observations (PRI:845-848)
    int n_pub = 0;
    for (int p = 0; p < n_det; ++p) {
      const int s = ws.slot[p];
      ws.out_idx[p] = (s >= 0 && ws.has[p] && tracks[s].num_obs + 1 > q.min_num_obs_track) ? n_pub++ : -1;
      if (track_of && s >= 0) track_of[p] = tracks[s].id;
    }
    ws.scal[PS_NPUB] = n_pub;
  });
  const int n_pub = ws.scal[PS_NPUB];
  const double t_prev_global = st->t_prev;
  const int frame_nr = st->frame_nr;

  // the skeleton fits: one warp per detection (PRI:587-853)
  tm.per_warp(n_det, [&](auto& w, int p) {
    const int s = ws.slot[p];
    if (s < 0) return;
    const int wid = w.size() == 1 ? 0 : (tm.rank() / 32);   // serial build: one workspace
    Arena ar(fit_ws + (size_t)wid * fit_ws_stride);
    PriorFitWs fws;
    prior_fit_ws_layout(ar, &fws);
    const int oi = ws.out_idx[p];
    prior_fit_person(w, pt, persons[p], tracks[s], t, t_prev_global, frame_nr, pred_delta_t, fws,
                     oi >= 0 ? fused + oi : nullptr, oi >= 0 ? pred + oi : nullptr);
  });

  // track life cycle: prune, then merge close tracks (PRI:866-903)
  tm.single([&] { prior_remove_old(pt, st, tracks, order, t); });
  const int n = st->n_tracks;
  tm.pfor(n * n, [&](int e) {
    const int i = e / n, j = e % n;
    if (i < j) ws.D[e] = prior_track_dist(tracks[order[i]], tracks[order[j]]);
  });
  tm.single([&] {
    // positions refer to the list at entry; the merge loop erases entries but never modifies a track
    int live[PRIOR_MAX_TRACKS];
    int m = n;
    for (int i = 0; i < n; ++i) live[i] = i;
    for (int i = 0; i < m; ++i) {
      for (int j = i + 1; j < m;) {
        if (ws.D[live[i] * n + live[j]] < q.merge_dist_thresh) {
          const int sj = order[live[j]], si = order[live[i]];
          const uint32_t id_to_remove = (uint32_t)tracks[sj].id, id_keep = (uint32_t)tracks[si].id;
          st->used &= ~(1ull << sj);
          for (int k = j; k + 1 < m; ++k) live[k] = live[k + 1];
          --m;
          for (int k = 0; k < n_pub; ++k)
            if (fused[k].id == id_to_remove) { fused[k].id = id_keep; pred[k].id = id_keep; }
        } else {
          ++j;
        }
      }
    }
    uint8_t tmp[PRIOR_MAX_TRACKS];
    for (int i = 0; i < m; ++i) tmp[i] = order[live[i]];
    for (int i = 0; i < m; ++i) order[i] = tmp[i];
    st->n_tracks = m;
    *n_out = n_pub;
    st->t_prev = t;   // PRI:909-910
    ++st->frame_nr;
  });
}

}  // namespace ses3d
