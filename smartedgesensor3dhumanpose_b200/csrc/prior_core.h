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
//  * the detections of a message are fitted in groups of up to three by one warp each, all detections of a group in
//    lock step. The factor graph of a skeleton is a forest (one bone per joint to its parent), and a range factor's
//    Hessian block is rank one: H_child,parent = -w w^T with w = (x_c - x_p) / (|x_c - x_p| sigma). The damped
//    normal equations (J^T J + lambda I) delta = -J^T e are therefore solved exactly by leaf-to-root elimination of
//    3x3 blocks (no fill-in, each Schur complement is the scalar alpha = w^T D^-1 w times w w^T) and a root-to-leaf
//    back-substitution: 6 tree levels up, 6 down, the (detection, joint-of-level) items of a level in parallel,
//    instead of a dense 57 x 57 Cholesky per LM trial. The marginal covariances follow from the same factorisation by
//    the downward recursion Sigma_c = D_c^-1 + (w^T Sigma_p w) z z^T, z = D_c^-1 w (no dense inverse). Every
//    detection carries its own LM state (lambda, errors, flags) and idles once it has converged;
//  * a message costs four CTA barriers: costs in parallel; Munkres, new tracks and publication slots on the first
//    warp; the fits; pruning, the pair-distance table and the merge loop on the first warp.
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

struct PriorStatic;
struct PriorTables {
  ses3d_prior_params prm;
  double limb_sigma_factor;               // PRI:934-937
  const PriorStatic* st;                  // skeleton tables (shared memory on the GPU)
};

SES_HD void prior_state_reset(const ses3d_prior_params& prm, PriorSeqState* st, bool keep_t_prev) {  // reset() PRI:182-189
  if (!keep_t_prev) st->t_prev = 0.0;     // static storage; reset() leaves g_t_prev alone
  for (int i = 0; i < PRIOR_NAVG; ++i) st->delay_buf[i] = prm.avg_delay;
  st->used = 0ull;
  st->next_id = 0; st->frame_nr = 0; st->n_tracks = 0; st->overflow = 0;
}

// ---- static skeleton forest (union of the bone tables PRI:384-481, rooted at MidHip) -------------------------
// One table object: a static const instance on the host, a __constant__ instance copied into shared memory at
// kernel start on the GPU (per-lane indexed lookups would serialise in the constant cache and, as function-local
// arrays, were rebuilt on the stack at every call).
struct PriorStatic {
  int8_t level[NFUS];        // depth in the forest: children are always deeper than their parent
  int8_t parent[NFUS];       // Neck (1): Belly (20) when measured, else MidHip (8) — PRI:464-471
  int8_t kids[NFUS][4];      // potential children, -1 padded
  double bone[2][NFUS][2];   // [absolute PRI:434-479 | height-normalised PRI:386-431][child joint]{length, sigma}
  double neck_midhip[2][2];  // the Simple-Baselines MidHip<->Neck bone PRI:470-471 / 422-423
  double vel_sigma[NFUS];    // FUSION_BODY_PARTS::vel_sigmas, fusion_body_parts.h:33
  int8_t lvl_joint[6][5];    // joints of every level (at most five), -1 padded
};
#define SES3D_PRIOR_STATIC_INIT                                                                                        \
  {                                                                                                                    \
    {3, 2, 3, 4, 5, 3, 4, 5, 0, 1, 2, 3, 1, 2, 3, 4, 4, 5, 5, 4, 1},                                                   \
    {1, 20, 1, 2, 3, 1, 5, 6, -1, 8, 9, 10, 8, 12, 13, 0, 0, 15, 16, 0, 8},                                            \
    {{19, 15, 16, -1}, {0, 2, 5, -1},    {3, -1, -1, -1},  {4, -1, -1, -1},  {-1, -1, -1, -1}, {6, -1, -1, -1},        \
     {7, -1, -1, -1},  {-1, -1, -1, -1}, {9, 12, 20, 1},   {10, -1, -1, -1}, {11, -1, -1, -1}, {-1, -1, -1, -1},       \
     {13, -1, -1, -1}, {14, -1, -1, -1}, {-1, -1, -1, -1}, {17, -1, -1, -1}, {18, -1, -1, -1}, {-1, -1, -1, -1},       \
     {-1, -1, -1, -1}, {-1, -1, -1, -1}, {1, -1, -1, -1}},                                                             \
    {{{0.20, 0.025},  {0.25534, 0.035}, {0.15, 0.042},  {0.28, 0.045},  {0.25, 0.063},  {0.15, 0.042},  {0.28, 0.045}, \
      {0.25, 0.063},  {0.0, 1.0},       {0.134, 0.033}, {0.449, 0.051}, {0.446, 0.051}, {0.134, 0.033}, {0.449, 0.051},\
      {0.446, 0.051}, {0.05, 0.035},    {0.05, 0.035},  {0.10, 0.05},   {0.10, 0.05},   {0.11500, 0.035},              \
      {0.23846, 0.071}},                                                                                               \
     {{0.33, 0.050},  {0.51, 0.05},   {0.262, 0.092}, {0.515, 0.071}, {0.444, 0.084}, {0.262, 0.092}, {0.515, 0.071},  \
      {0.444, 0.084}, {0.0, 1.0},     {0.17, 0.062},  {0.694, 0.111}, {0.708, 0.097}, {0.17, 0.062},  {0.694, 0.111},  \
      {0.708, 0.097}, {0.085, 0.06},  {0.085, 0.06},  {0.167, 0.08},  {0.167, 0.08},  {0.23, 0.05},   {0.49, 0.05}}},  \
    {{0.50, 0.071}, {1.000, 0.02}},                                                                                    \
    {2., 1., 1., 2., 3., 1., 2., 3., 1., 1., 2., 3., 1., 2., 3., 2., 2., 2., 2., 2., 1.},                              \
    {{8, -1, -1, -1, -1}, {9, 12, 20, -1, -1}, {10, 13, 1, -1, -1}, {11, 14, 0, 2, 5}, {3, 6, 19, 15, 16},             \
     {4, 7, 17, 18, -1}}                                                                                               \
  }
inline const PriorStatic& prior_static_host() {
  static const PriorStatic t = SES3D_PRIOR_STATIC_INIT;
  return t;
}
// bone between joint k and its parent; via_midhip selects the MidHip<->Neck bone for k = Neck
SES_HD void prior_bone(const PriorStatic& T, int k, bool normalised, bool via_midhip, double* len, double* sigma) {
  const double* b = (k == SES3D_FBP_NECK && via_midhip) ? T.neck_midhip[normalised ? 1 : 0] : T.bone[normalised ? 1 : 0][k];
  *len = b[0];
  *sigma = b[1];
}

// ---- workspaces ---------------------------------------------------------------------------------------------
constexpr int PRIOR_GMAX = 6;   // detections one warp fits together: 6 x 5 joints of the widest tree level = 30 lanes
constexpr int PRIOR_LW = 5;     // widest level

struct PriorFitScal {           // LM / bookkeeping state of one detection of the group
  double lambda, error, cur_error, old_lin, height, height_prev;
  double root[3], neck[3], root_prev[3];
  float root_score, neck_score;
  uint32_t mask, had;           // measured joints; joints the track's prevEstimate held
  int32_t iterations, trials;
  uint8_t active, lm, need_lin, fail, accept, use_marginals, fresh, pad_[1];
};

struct PriorFitWs {       // the factor graphs of one group, one workspace per warp; index i = g * 21 + joint
  double *x, *w, *z;                 // [G*21][3]
  double* my;                        // [G*21][6]: measurement m | y = Dinv b, overwritten in place by delta in the
                                     //            back-substitution; the marginal pass ends by writing Sigma over both
  double* W;                         // [G*21][6]  (00,01,02,11,12,22): information of the unary factor; the marginal
                                     //            elimination overwrites it with the block inverse D^-1
  double *e, *alpha, *beta;          // [G*21]
  double *ta, *tb;                   // [G*21] per-joint error terms; alias alpha / beta (never live together)
  int8_t* par;                       // [G*21] parent joint of the bone, -1 none
  uint8_t *msd, *usev;               // [G*21] measured / velocity usable
  PriorFitScal* sc;                  // [G]
};

template <class A>
SES_HD void prior_fit_ws_layout(A& ar, int G, PriorFitWs* ws) {
  const size_t n = (size_t)G * NFUS;
  double* v3[3];
  for (int i = 0; i < 3; ++i) v3[i] = ar.template take<double>(n * 3);
  double* v6[2];
  for (int i = 0; i < 2; ++i) v6[i] = ar.template take<double>(n * 6);
  double* v1[3];
  for (int i = 0; i < 3; ++i) v1[i] = ar.template take<double>(n);
  PriorFitScal* sc = ar.template take<PriorFitScal>(G);
  int8_t* par = ar.template take<int8_t>(n);
  uint8_t* msd = ar.template take<uint8_t>(n);
  uint8_t* usev = ar.template take<uint8_t>(n);
  if (ws) {
    ws->x = v3[0]; ws->w = v3[1]; ws->z = v3[2];
    ws->my = v6[0]; ws->W = v6[1];
    ws->e = v1[0]; ws->alpha = v1[1]; ws->beta = v1[2]; ws->ta = v1[1]; ws->tb = v1[2];
    ws->sc = sc; ws->par = par; ws->msd = msd; ws->usev = usev;
  }
}
inline size_t prior_fit_ws_bytes(int G) {
  ArenaSizer s;
  prior_fit_ws_layout(s, G, nullptr);
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
  uint8_t* has;           // [h_max]
  int* scal;              // [4] {n_trk at frame start, n_pub}
};
enum { PW_T = 0, PW_PDT = 1 };
enum { PS_NTRK = 0, PS_NPUB = 1 };

// The small arrays live for the whole message; the big ones (cost / Munkres tables, pair distances) are only used
// while no fit is running and therefore share memory with the fit workspaces (`tr` arena).
template <class A, class B>
SES_HD void prior_ws_layout(A& ar, B& tr, int h_max, int max_tracks, PriorWs* ws) {
  double* dscal = ar.template take<double>(4);
  int* assignment = ar.template take<int>(h_max);
  int* slot = ar.template take<int>(h_max);
  int* out_idx = ar.template take<int>(h_max);
  int* scal = ar.template take<int>(4);
  uint8_t* has = ar.template take<uint8_t>(h_max);
  double* cost = tr.template take<double>((size_t)h_max * max_tracks);
  double* dist = tr.template take<double>((size_t)h_max * max_tracks);
  double* D = tr.template take<double>((size_t)max_tracks * max_tracks);
  uint8_t* star = tr.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* prime = tr.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* nstar = tr.template take<uint8_t>((size_t)h_max * max_tracks);
  uint8_t* cov_r = tr.template take<uint8_t>(h_max);
  uint8_t* cov_c = tr.template take<uint8_t>(max_tracks);
  if (ws) {
    ws->cost = cost; ws->dist = dist; ws->D = D; ws->dscal = dscal; ws->assignment = assignment; ws->slot = slot;
    ws->out_idx = out_idx; ws->scal = scal; ws->star = star; ws->prime = prime; ws->nstar = nstar; ws->cov_r = cov_r;
    ws->cov_c = cov_c; ws->has = has;
  }
}
inline void prior_ws_bytes(int h_max, int max_tracks, size_t* persistent, size_t* transient) {
  ArenaSizer a, b;
  prior_ws_layout(a, b, h_max, max_tracks, nullptr);
  *persistent = (a.used + 15) / 16 * 16;
  *transient = (b.used + 15) / 16 * 16;
}

// ---- 3x3 helpers --------------------------------------------------------------------------------------------
// information matrix W = R^T R of noiseModel::Gaussian::Covariance(S): diagonal S -> whitening by 1/sigma
// (Diagonal / Isotropic), else Information(S^-1) whose Cholesky factor R satisfies R^T R = S^-1.
// S, W as (00,01,02,11,12,22).
SES_HD void prior_information(const double S[6], double W[6]) {
  if (S[1] == 0.0 && S[2] == 0.0 && S[4] == 0.0) {
    const double r0 = 1.0 / sqrt(S[0]), r1 = 1.0 / sqrt(S[3]), r2 = 1.0 / sqrt(S[5]);
    W[0] = r0 * r0; W[1] = 0.0; W[2] = 0.0; W[3] = r1 * r1; W[4] = 0.0; W[5] = r2 * r2;
    return;
  }
  const double c00 = S[3] * S[5] - S[4] * S[4], c01 = S[4] * S[2] - S[1] * S[5], c02 = S[1] * S[4] - S[3] * S[2];
  const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
  const double id = 1.0 / det;
  W[0] = c00 * id; W[1] = c01 * id; W[2] = c02 * id;
  W[3] = (S[0] * S[5] - S[2] * S[2]) * id;
  W[4] = (S[1] * S[2] - S[0] * S[4]) * id;
  W[5] = (S[0] * S[3] - S[1] * S[1]) * id;
}
SES_HD void sym6_mul(const double A[6], const double v[3], double o[3]) {
  o[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  o[1] = A[1] * v[0] + A[3] * v[1] + A[4] * v[2];
  o[2] = A[2] * v[0] + A[4] * v[1] + A[5] * v[2];
}
// inverse of a symmetric positive definite 3x3 through its LDL^T factorisation (three reciprocals; the adjugate
// formula needs one but loses ~4 digits on the ill-conditioned information blocks of elongated covariances);
// false when a pivot is not positive / not finite
SES_HD bool sym6_inverse_spd(const double D[6], double I[6]) {
  const double p0 = D[0];
  if (!(p0 > 0.0)) return false;
  const double q0 = 1.0 / p0;
  const double l10 = D[1] * q0, l20 = D[2] * q0;
  const double p1 = D[3] - l10 * D[1];
  if (!(p1 > 0.0)) return false;
  const double q1 = 1.0 / p1;
  const double t21 = D[4] - l20 * D[1];
  const double l21 = t21 * q1;
  const double p2 = D[5] - l20 * D[2] - l21 * t21;
  if (!(p2 > 0.0) || !(p2 < DBL_MAX)) return false;
  const double q2 = 1.0 / p2;
  // D = L diag(p) L^T  =>  D^-1 = L^-T diag(1/p) L^-1, L^-1 = [[1,0,0],[-l10,1,0],[l10 l21 - l20, -l21, 1]]
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
      dist += sqrt(dx * dx + dy * dy + dz * dz) / (pt.st->vel_sigma[k] * delta_t);
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

// ---- the skeleton fits of one group of detections (PRI:587-853), run by a warp-sized team -------------------------
// A group is up to PRIOR_GMAX detections of the same message. Dense phases run over (detection, joint) items, the
// tree sweeps over (detection, joint-of-level) items, so the lanes of the warp stay occupied; every detection
// carries its own LM state (lambda, errors, flags) through the same instruction stream and simply idles once
// it has converged.

// bone residual (|x_k - x_p| - len) / sigma and, optionally, the whitened direction w; xk/xp are the end points
SES_HD double prior_bone_residual(const PriorTables& pt, int k, int p, const double* xk, const double* xp, double* w) {
  double len, sigma;
  prior_bone(*pt.st, k, pt.prm.normalize_by_height != 0, p == SES3D_FBP_MIDHIP, &len, &sigma);
  const double is = 1.0 / (sigma * pt.limb_sigma_factor);   // Isotropic::whiten: v * invsigma
  const double dx = xk[0] - xp[0], dy = xk[1] - xp[1], dz = xk[2] - xp[2];
  const double r = sqrt(dx * dx + dy * dy + dz * dz);
  if (w) { const double s = is / r; w[0] = dx * s; w[1] = dy * s; w[2] = dz * s; }
  return (r - len) * is;
}

// leaf-to-root elimination of (J^T J + lambda I) for every detection with `lm` (or all active ones when
// for_marginals, lambda = 0): per joint Dinv, z = Dinv w, y = Dinv b, alpha = w.z, beta = w.y. A block that is not
// positive definite sets the detection's fail flag (gtsam: IndeterminantLinearSystemException).
template <class WT>
SES_HD void prior_eliminate(WT& tm, const PriorTables& pt, int G, const PriorFitWs& ws, bool for_marginals) {
  for (int L = PRIOR_LEVELS - 1; L >= 0; --L) {
    tm.pfor(G * PRIOR_LW, [&](int it) {
      const int g = it / PRIOR_LW, k = pt.st->lvl_joint[L][it % PRIOR_LW];
      if (k < 0) return;
      PriorFitScal& sc = ws.sc[g];
      if (!(for_marginals ? sc.active : sc.lm)) return;
      const int i = g * NFUS + k;
      if (!ws.msd[i]) return;
      const double lambda = for_marginals ? 0.0 : sc.lambda;
      const double* W = ws.W + 6 * i;
      double D[6] = {W[0] + lambda, W[1], W[2], W[3] + lambda, W[4], W[5] + lambda};
      double b[3];
      {  // minus the gradient of the unary factor, W (x - m), recomputed here (9 FMAs) instead of being stored
        const double* x = ws.x + 3 * i;
        const double d[3] = {x[0] - ws.my[6 * i], x[1] - ws.my[6 * i + 1], x[2] - ws.my[6 * i + 2]};
        sym6_mul(W, d, b);
        b[0] = -b[0]; b[1] = -b[1]; b[2] = -b[2];
      }
      const double* w = ws.w + 3 * i;
      if (ws.par[i] >= 0) {
        const double e = ws.e[i];
        D[0] += w[0] * w[0]; D[1] += w[0] * w[1]; D[2] += w[0] * w[2];
        D[3] += w[1] * w[1]; D[4] += w[1] * w[2]; D[5] += w[2] * w[2];
        b[0] -= w[0] * e; b[1] -= w[1] * e; b[2] -= w[2] * e;
      }
      for (int c4 = 0; c4 < 4; ++c4) {
        const int c = pt.st->kids[k][c4];
        if (c < 0) break;
        const int ic = g * NFUS + c;
        if (!ws.msd[ic] || ws.par[ic] != k) continue;
        const double* wc = ws.w + 3 * ic;
        const double f = 1.0 - ws.alpha[ic];        // bone term w w^T minus the child's Schur complement alpha w w^T
        const double q = ws.e[ic] + ws.beta[ic];    // -(-w e) from the bone gradient, + beta w from the elimination
        D[0] += f * wc[0] * wc[0]; D[1] += f * wc[0] * wc[1]; D[2] += f * wc[0] * wc[2];
        D[3] += f * wc[1] * wc[1]; D[4] += f * wc[1] * wc[2]; D[5] += f * wc[2] * wc[2];
        b[0] += q * wc[0]; b[1] += q * wc[1]; b[2] += q * wc[2];
      }
      double I[6];
      if (!sym6_inverse_spd(D, I)) {
        sc.fail = 1;
        I[0] = I[3] = I[5] = 1.0; I[1] = I[2] = I[4] = 0.0;
      }
      if (for_marginals)   // W is dead after this point of the marginal pass: keep D^-1 in its place
        for (int a = 0; a < 6; ++a) ws.W[6 * i + a] = I[a];
      double* z = ws.z + 3 * i;
      double* y = ws.my + 6 * i + 3;
      sym6_mul(I, w, z);
      sym6_mul(I, b, y);
      ws.alpha[i] = w[0] * z[0] + w[1] * z[1] + w[2] * z[2];
      ws.beta[i] = w[0] * y[0] + w[1] * y[1] + w[2] * y[2];
    });
  }
}

// root-to-leaf back-substitution: delta_k = y_k + z_k (w_k . delta_parent)
template <class WT>
SES_HD void prior_backsubstitute(WT& tm, const PriorTables& pt, int G, const PriorFitWs& ws) {
  for (int L = 0; L < PRIOR_LEVELS; ++L) {
    tm.pfor(G * PRIOR_LW, [&](int it) {
      const int g = it / PRIOR_LW, k = pt.st->lvl_joint[L][it % PRIOR_LW];
      if (k < 0 || !ws.sc[g].lm || ws.sc[g].fail) return;
      const int i = g * NFUS + k;
      if (!ws.msd[i]) return;
      double s = 0.0;
      const int p = ws.par[i];
      if (p >= 0) {
        const double* dp = ws.my + 6 * (g * NFUS + p) + 3;
        s = ws.w[3 * i] * dp[0] + ws.w[3 * i + 1] * dp[1] + ws.w[3 * i + 2] * dp[2];
      }
      double* y = ws.my + 6 * i + 3;   // delta overwrites y
      for (int a = 0; a < 3; ++a) y[a] = y[a] + ws.z[3 * i + a] * s;
    });
  }
}

// marginal covariances after prior_eliminate(for_marginals): Sigma_k = Dinv_k + (w_k^T Sigma_p w_k) z_k z_k^T,
// written over m | y (no longer needed)
template <class WT>
SES_HD void prior_marginals(WT& tm, const PriorTables& pt, int G, const PriorFitWs& ws) {
  for (int L = 0; L < PRIOR_LEVELS; ++L) {
    tm.pfor(G * PRIOR_LW, [&](int it) {
      const int g = it / PRIOR_LW, k = pt.st->lvl_joint[L][it % PRIOR_LW];
      if (k < 0 || !ws.sc[g].active || !ws.sc[g].use_marginals) return;
      const int i = g * NFUS + k;
      if (!ws.msd[i]) return;
      const double* I = ws.W + 6 * i;   // D^-1, stored by the marginal elimination
      double* S = ws.my + 6 * i;
      double q = 0.0;
      const int p = ws.par[i];
      if (p >= 0) {
        double t[3];
        sym6_mul(ws.my + 6 * (g * NFUS + p), ws.w + 3 * i, t);
        q = ws.w[3 * i] * t[0] + ws.w[3 * i + 1] * t[1] + ws.w[3 * i + 2] * t[2];
      }
      const double* z = ws.z + 3 * i;
      S[0] = I[0] + q * z[0] * z[0]; S[1] = I[1] + q * z[0] * z[1]; S[2] = I[2] + q * z[0] * z[2];
      S[3] = I[3] + q * z[1] * z[1]; S[4] = I[4] + q * z[1] * z[2]; S[5] = I[5] + q * z[2] * z[2];
    });
  }
}

// linearise at x for the detections selected by `which` (0: lm && need_lin, 1: all active): bone direction w and
// residual e; ta = the joint's share of linear.error(0)
template <class WT>
SES_HD void prior_linearize(WT& tm, const PriorTables& pt, int G, const PriorFitWs& ws, int which) {
  tm.pfor(G * NFUS, [&](int i) {
    const int g = i / NFUS, k = i % NFUS;
    const PriorFitScal& sc = ws.sc[g];
    if (!(which ? sc.active : (sc.lm && sc.need_lin)) || !ws.msd[i]) return;
    const double* x = ws.x + 3 * i;
    const double d[3] = {x[0] - ws.my[6 * i], x[1] - ws.my[6 * i + 1], x[2] - ws.my[6 * i + 2]};
    double gu[3];
    sym6_mul(ws.W + 6 * i, d, gu);
    double err = 0.5 * (d[0] * gu[0] + d[1] * gu[1] + d[2] * gu[2]);
    const int p = ws.par[i];
    double e = 0.0;
    double* w = ws.w + 3 * i;
    w[0] = w[1] = w[2] = 0.0;
    if (p >= 0) {
      e = prior_bone_residual(pt, k, p, x, ws.x + 3 * (g * NFUS + p), w);
      err += 0.5 * (e * e);
    }
    ws.e[i] = e;
    ws.ta[i] = err;
  });
}

SES_HD double prior_sum_terms(const PriorFitWs& ws, const double* t, int g) {
  double s = 0.0;
  for (int k = 0; k < NFUS; ++k)
    if (ws.msd[g * NFUS + k]) s += t[g * NFUS + k];
  return s;
}

// G detections persons[0..G) of one message; slot[g] = track slot (-1: skip), out_idx[g] = position in the
// published list (-1: not published, PRI:845-848). Tracks are exclusively owned by this team during the call.
template <class WT>
SES_HD void prior_fit_group(WT& tm, const PriorTables& pt, int G, const ses3d_person_cov* persons, PriorTrack* tracks,
                            const int* slot, const int* out_idx, ses3d_person_cov* fused, ses3d_person_cov* pred,
                            double t, double t_prev_global, int frame_nr, double pred_delta_t, const PriorFitWs& ws) {
  const ses3d_prior_params& q = pt.prm;
  const bool simple = q.pose_method == SES3D_POSE_SIMPLE;

  // root / neck / height of every detection (PRI:631-668), first-observation initialisation (PRI:699-702)
  tm.pfor(G, [&](int g) {
    PriorFitScal& sc = ws.sc[g];
    sc.active = 0; sc.lm = 0; sc.need_lin = 0; sc.fail = 0; sc.accept = 0; sc.use_marginals = 0; sc.fresh = 0;
    sc.mask = 0u; sc.had = 0u; sc.iterations = 0; sc.trials = 0;
    if (slot[g] < 0) return;
    const PriorRootNeck rn = prior_root_neck(pt, persons[g]);
    for (int a = 0; a < 3; ++a) { sc.root[a] = rn.root[a]; sc.neck[a] = rn.neck[a]; }
    sc.root_score = rn.root_score; sc.neck_score = rn.neck_score; sc.height = rn.height;
    PriorTrack& tr = tracks[slot[g]];
    sc.fresh = tr.height_prev < 0.0 ? 1 : 0;   // first message of the track: its velocity buffer starts at zero
    if (sc.fresh) {
      tr.height_prev = rn.height;
      tr.root_prev[0] = rn.root[0]; tr.root_prev[1] = rn.root[1]; tr.root_prev[2] = rn.root[2];
    }
    sc.height_prev = tr.height_prev;
    for (int a = 0; a < 3; ++a) sc.root_prev[a] = tr.root_prev[a];
    sc.had = tr.exists;
    sc.active = 1;   // provisional: cleared below when the detection has no usable joint
  });

  // measurements and their information matrices (PRI:658-737)
  tm.pfor(G * NFUS, [&](int i) {
    const int g = i / NFUS, k = i % NFUS;
    const PriorFitScal& sc = ws.sc[g];
    bool ms = false;
    double c[6], mk[3] = {0.0, 0.0, 0.0};
    if (sc.active) {
      const ses3d_person_cov& person = persons[g];
      const double height = sc.height;
      if (k == SES3D_FBP_MIDHIP) {
        if (sc.root_score > q.min_score) {
          ms = true;
          if (simple) {
            const double* a = person.keypoints[SES3D_FBP_LHIP].cov;
            const double* b = person.keypoints[SES3D_FBP_RHIP].cov;
            for (int j = 0; j < 6; ++j) c[j] = (a[j] + b[j]) / 2.0;
          } else {
            for (int j = 0; j < 6; ++j) c[j] = person.keypoints[SES3D_FBP_MIDHIP].cov[j];
          }
          for (int j = 0; j < 6; ++j) c[j] = c[j] / (height * height) / (q.root_sigma_factor * q.root_sigma_factor);
        }
      } else {
        const ses3d_keypoint_cov& kp = person.keypoints[k];
        if (kp.score > q.min_score) {
          ms = true;
          for (int j = 0; j < 6; ++j) c[j] = kp.cov[j] / (height * height);
          mk[0] = (kp.x - sc.root[0]) / height; mk[1] = (kp.y - sc.root[1]) / height; mk[2] = (kp.z - sc.root[2]) / height;
        }
        if (k == SES3D_FBP_NECK && simple && sc.neck_score > q.min_score) {
          ms = true;
          const double* a = person.keypoints[SES3D_FBP_LSHOULDER].cov;
          const double* b = person.keypoints[SES3D_FBP_RSHOULDER].cov;
          for (int j = 0; j < 6; ++j) c[j] = (a[j] + b[j]) / 2.0 / (height * height);
          for (int j = 0; j < 3; ++j) mk[j] = (sc.neck[j] - sc.root[j]) / height;
        }
      }
    }
    ws.msd[i] = ms ? 1 : 0;
    if (ms) {
      prior_information(c, ws.W + 6 * i);
      ws.my[6 * i] = mk[0]; ws.my[6 * i + 1] = mk[1]; ws.my[6 * i + 2] = mk[2];
    }
  });
  tm.pfor(G, [&](int g) {
    PriorFitScal& sc = ws.sc[g];
    uint32_t mask = 0;
    for (int k = 0; k < NFUS; ++k) mask |= ws.msd[g * NFUS + k] ? (1u << k) : 0u;
    sc.mask = mask;
    if (mask == 0u) sc.active = 0;   // num_meas == 0: continue (PRI:739-741)
  });

  // setInitialState (PRI:483-503) + bone parents (addBinaryFactors PRI:384-481)
  tm.pfor(G * NFUS, [&](int i) {
    const int g = i / NFUS, k = i % NFUS;
    const PriorFitScal& sc = ws.sc[g];
    ws.usev[i] = 0;
    ws.par[i] = -1;
    if (slot[g] < 0) return;
    PriorTrack& tr = tracks[slot[g]];
    const bool ex = (sc.had >> k) & 1u, ms = (sc.mask >> k) & 1u;
    if (sc.fresh || (sc.active && ex && !ms))   // TrackingHypothesis ctor (PRI:80) / setInitialState (PRI:492)
      for (int b = 0; b < PRIOR_NAVG; ++b) tr.vel[k][b][0] = tr.vel[k][b][1] = tr.vel[k][b][2] = 0.0;
    if (!sc.active) return;
    if (!ms) return;
    ws.usev[i] = ex ? 1 : 0;
    for (int a = 0; a < 3; ++a) ws.x[3 * i + a] = ex ? tr.prev[k][a] : ws.my[6 * i + a];
    int par = pt.st->parent[k];
    if (k == SES3D_FBP_NECK && !((sc.mask >> SES3D_FBP_BELLY) & 1u)) par = SES3D_FBP_MIDHIP;
    if (par >= 0 && !((sc.mask >> par) & 1u)) par = -1;
    ws.par[i] = (int8_t)par;
  });

  // LevenbergMarquardtOptimizer(graph, prevEstimate).optimize(), default parameters, all detections in lock step
  tm.pfor(G * NFUS, [&](int i) {   // graph.error(initial)
    const int g = i / NFUS, k = i % NFUS;
    if (!ws.sc[g].active || !ws.msd[i]) return;
    const double* x = ws.x + 3 * i;
    const double d[3] = {x[0] - ws.my[6 * i], x[1] - ws.my[6 * i + 1], x[2] - ws.my[6 * i + 2]};
    double wd[3];
    sym6_mul(ws.W + 6 * i, d, wd);
    double err = 0.5 * (d[0] * wd[0] + d[1] * wd[1] + d[2] * wd[2]);
    const int p = ws.par[i];
    if (p >= 0) {
      const double e = prior_bone_residual(pt, k, p, x, ws.x + 3 * (g * NFUS + p), nullptr);
      err += 0.5 * (e * e);
    }
    ws.ta[i] = err;
  });
  tm.pfor(G, [&](int g) {
    PriorFitScal& sc = ws.sc[g];
    if (!sc.active) return;
    sc.error = prior_sum_terms(ws, ws.ta, g);
    sc.lambda = q.lm_lambda_initial;
    sc.lm = (sc.error > 0.0 && q.lm_max_iterations > 0) ? 1 : 0;   // errorTol = 0, maxIterations
    sc.need_lin = 1;
  });
  const int trial_cap = 4 * q.lm_max_iterations + 64;
  for (int round = 0; round < trial_cap; ++round) {
    if (tm.first(G, [&](int g) { return ws.sc[g].lm != 0; }) == G) break;
    prior_linearize(tm, pt, G, ws, 0);
    tm.pfor(G, [&](int g) {
      PriorFitScal& sc = ws.sc[g];
      sc.fail = 0;
      if (!sc.lm || !sc.need_lin) return;
      sc.old_lin = prior_sum_terms(ws, ws.ta, g);   // linear.error(0)
      sc.cur_error = sc.error;                      // currentError = error(); iterate();
      sc.need_lin = 0;
    });
    prior_eliminate(tm, pt, G, ws, false);
    prior_backsubstitute(tm, pt, G, ws);
    tm.pfor(G * NFUS, [&](int i) {   // linear.error(delta) and graph.error(x + delta), joint by joint
      const int g = i / NFUS, k = i % NFUS;
      const PriorFitScal& sc = ws.sc[g];
      if (!sc.lm || sc.fail || !ws.msd[i]) return;
      const double* x = ws.x + 3 * i;
      const double* dl = ws.my + 6 * i + 3;
      const double xn[3] = {x[0] + dl[0], x[1] + dl[1], x[2] + dl[2]};
      const double d[3] = {xn[0] - ws.my[6 * i], xn[1] - ws.my[6 * i + 1], xn[2] - ws.my[6 * i + 2]};
      double wd[3];
      sym6_mul(ws.W + 6 * i, d, wd);
      const double unary = 0.5 * (d[0] * wd[0] + d[1] * wd[1] + d[2] * wd[2]);   // linear in x: same in both errors
      double lin = unary, nonlin = unary;
      const int p = ws.par[i];
      if (p >= 0) {
        const int ip = g * NFUS + p;
        const double* xp = ws.x + 3 * ip;
        const double* dp = ws.my + 6 * ip + 3;
        const double* w = ws.w + 3 * i;
        const double el = ws.e[i] + w[0] * (dl[0] - dp[0]) + w[1] * (dl[1] - dp[1]) + w[2] * (dl[2] - dp[2]);
        lin += 0.5 * (el * el);
        const double xpn[3] = {xp[0] + dp[0], xp[1] + dp[1], xp[2] + dp[2]};
        const double en = prior_bone_residual(pt, k, p, xn, xpn, nullptr);
        nonlin += 0.5 * (en * en);
      }
      ws.ta[i] = lin;
      ws.tb[i] = nonlin;
    });
    tm.pfor(G, [&](int g) {   // tryLambda() + the outer loop's convergence test, per detection
      PriorFitScal& sc = ws.sc[g];
      sc.accept = 0;
      if (!sc.lm) return;
      bool step_ok = false, stop_searching = false;
      double new_error = DBL_MAX;
      if (!sc.fail) {
        const double new_lin = prior_sum_terms(ws, ws.ta, g);
        const double lin_change = sc.old_lin - new_lin;
        if (lin_change >= 0) {
          new_error = prior_sum_terms(ws, ws.tb, g);
          const double cost_change = sc.error - new_error;
          if (lin_change > 1e-20) step_ok = (cost_change / lin_change) > q.lm_min_model_fidelity;
          if (fabs(cost_change) < q.lm_relative_error_tol * sc.error) stop_searching = true;
        }
      }
      bool end_iterate = true;
      if (step_ok) {
        sc.accept = 1;
        sc.error = new_error;
        sc.lambda = sc.lambda / q.lm_lambda_factor;
        if (sc.lambda < 0.0) sc.lambda = 0.0;
        ++sc.iterations;
      } else if (!stop_searching) {
        sc.lambda *= q.lm_lambda_factor;
        end_iterate = sc.lambda >= q.lm_lambda_upper_bound;
      }
      if (end_iterate) {
        bool converged;
        if (sc.error <= 0.0) converged = true;
        else {
          const double abs_dec = sc.cur_error - sc.error;
          const double rel_dec = abs_dec / sc.cur_error;
          converged = (q.lm_relative_error_tol != 0.0 && rel_dec <= q.lm_relative_error_tol) ||
                      abs_dec <= q.lm_absolute_error_tol;
        }
        const bool finite = sc.cur_error < DBL_MAX && sc.cur_error == sc.cur_error;
        if (sc.iterations < q.lm_max_iterations && !converged && finite) sc.need_lin = 1;
        else sc.lm = 0;
      }
      if (++sc.trials >= trial_cap) sc.lm = 0;
    });
    tm.pfor(G * NFUS, [&](int i) {
      if (!ws.sc[i / NFUS].accept || !ws.msd[i]) return;
      for (int a = 0; a < 3; ++a) ws.x[3 * i + a] += ws.my[6 * i + 3 + a];
    });
  }

  // Marginals(graph, result) (PRI:760-767)
  tm.pfor(G, [&](int g) { ws.sc[g].fail = 0; ws.sc[g].lm = 0; });
  prior_linearize(tm, pt, G, ws, 1);
  prior_eliminate(tm, pt, G, ws, true);
  tm.pfor(G, [&](int g) { ws.sc[g].use_marginals = (ws.sc[g].active && !ws.sc[g].fail) ? 1 : 0; });
  prior_marginals(tm, pt, G, ws);

  // outputs (PRI:770-837) and the track update (PRI:839-843)
  const int rec_words = (int)(sizeof(ses3d_person_cov) / 8);
  tm.pfor(G * rec_words, [&](int i) {
    const int g = i / rec_words;
    if (!ws.sc[g].active || out_idx[g] < 0) return;
    reinterpret_cast<double*>(fused + out_idx[g])[i % rec_words] = 0.0;
    reinterpret_cast<double*>(pred + out_idx[g])[i % rec_words] = 0.0;
  });
  const double dtg = t - t_prev_global;
  const int vslot = frame_nr % PRIOR_NAVG;
  tm.pfor(G * NFUS, [&](int i) {
    const int g = i / NFUS, k = i % NFUS;
    const PriorFitScal& sc = ws.sc[g];
    if (!sc.active || !ws.msd[i]) return;
    PriorTrack& tr = tracks[slot[g]];
    const double height = sc.height;
    double jf[3], jp[3];
    for (int a = 0; a < 3; ++a) jf[a] = ws.x[3 * i + a] * height + sc.root[a];
    for (int a = 0; a < 3; ++a) jp[a] = jf[a];
    if (ws.usev[i]) {
      for (int a = 0; a < 3; ++a)
        tr.vel[k][vslot][a] = (jf[a] - (tr.prev[k][a] * sc.height_prev + sc.root_prev[a])) / dtg;
      for (int a = 0; a < 3; ++a) {
        double acc = 0.0;
        for (int b = 0; b < PRIOR_NAVG; ++b) acc += tr.vel[k][b][a];
        jp[a] += acc / PRIOR_NAVG * pred_delta_t;
      }
    }
    for (int a = 0; a < 3; ++a) tr.prev[k][a] = ws.x[3 * i + a];
    const int oi = out_idx[g];
    if (oi < 0) return;
    float score;
    if (k == SES3D_FBP_MIDHIP) score = sc.root_score;
    else if (k == SES3D_FBP_NECK) score = sc.neck_score;
    else score = persons[g].keypoints[k].score;
    if (!(score > q.min_score)) score = q.min_score;   // std::max(g_min_score, score)
    double cv[6];
    if (sc.use_marginals) {
      for (int a = 0; a < 6; ++a) cv[a] = ws.my[6 * i + a] * height * height;
    } else {
      const double d = q.default_res_sigma * q.default_res_sigma;
      cv[0] = d; cv[1] = 0; cv[2] = 0; cv[3] = d; cv[4] = 0; cv[5] = d;
    }
    if (k == SES3D_FBP_MIDHIP)
      for (int a = 0; a < 6; ++a) cv[a] *= (q.root_sigma_factor * q.root_sigma_factor);
    ses3d_keypoint_cov& o = fused[oi].keypoints[k];
    o.x = jf[0]; o.y = jf[1]; o.z = jf[2]; o.score = score;
    for (int a = 0; a < 6; ++a) o.cov[a] = cv[a];
    ses3d_keypoint_cov& po = pred[oi].keypoints[k];
    po.x = jp[0]; po.y = jp[1]; po.z = jp[2]; po.score = score;
    const double pn = q.pred_noise_sigma * q.pred_noise_sigma;
    po.cov[0] = cv[0] + pn; po.cov[1] = cv[1]; po.cov[2] = cv[2]; po.cov[3] = cv[3] + pn; po.cov[4] = cv[4];
    po.cov[5] = cv[5] + pn;
  });
  tm.pfor(G, [&](int g) {
    const PriorFitScal& sc = ws.sc[g];
    if (!sc.active) return;
    PriorTrack& tr = tracks[slot[g]];
    tr.t_prev = t;
    tr.exists = sc.mask;
    tr.height_prev = sc.height;
    tr.root_prev[0] = sc.root[0]; tr.root_prev[1] = sc.root[1]; tr.root_prev[2] = sc.root[2];
    ++tr.num_obs;
    if (out_idx[g] >= 0) { fused[out_idx[g]].id = (uint32_t)tr.id; pred[out_idx[g]].id = (uint32_t)tr.id; }
  });
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
// persons [n_det], fused / pred [h_max]; fit_ws = base of the per-warp fit workspaces (fit_ws_stride bytes apart,
// each sized for `group` detections fitted together).
template <class Team>
SES_HD void prior_frame(Team& tm, const PriorTables& pt, int max_tracks, int h_max, int group, PriorSeqState* st,
                        PriorTrack* tracks, uint8_t* order, const PriorWs& ws, unsigned char* fit_ws, size_t fit_ws_stride,
                        int64_t stamp_ns, int n_cams, const float* fb_delay, int n_det_in,
                        const ses3d_person_cov* persons, ses3d_person_cov* fused, ses3d_person_cov* pred,
                        int32_t* n_out, float* pred_delay, int32_t* track_of) {
  const ses3d_prior_params& q = pt.prm;
  const int n_det = n_det_in < 0 ? 0 : (n_det_in > h_max ? h_max : n_det_in);
  // Four barrier-separated phases per message: (1) cost matrix in parallel, (2) everything serial about the
  // tracker on the first warp, (3) the fits, (4) track life cycle on the first warp.
  const double t = prior_stamp_to_sec(stamp_ns);   // pure function of the stamp: every thread evaluates it
  const int n_trk = st->n_tracks;                   // unchanged until phase 2
  const double t_prev_global = st->t_prev;          // unchanged until phase 4
  const int frame_nr = st->frame_nr;

  // ---- phase 1: association costs (PRI:554-559), one thread per (detection, track); the leader also does the
  // feedback-delay moving average (PRI:513-526)
  tm.pfor(n_det * n_trk + h_max + 1, [&](int e) {
    if (e < n_det * n_trk) {
      const int p = e % n_det, tr = e / n_det;
      ws.cost[e] = prior_normed_dist(pt, tracks[order[tr]], persons[p], t);
      return;
    }
    e -= n_det * n_trk;
    if (e < h_max) {
      if (track_of) track_of[e] = -1;
      if (e < n_det) ws.has[e] = prior_has_measurement(pt, persons[e]) ? 1 : 0;
      return;
    }
    double curr = 0.0;
    int n_valid = 0;
    for (int c = 0; c < n_cams; ++c) {
      const float d = fb_delay ? fb_delay[c] : -1.0f;
      if (d > 0.0f) { curr += (double)d; ++n_valid; }
    }
    if (n_valid > 0) curr /= n_valid; else curr = q.avg_delay;
    st->delay_buf[frame_nr % PRIOR_NAVG] = curr;
    double acc = 0.0;
    for (int i = 0; i < PRIOR_NAVG; ++i) acc += st->delay_buf[i];
    ws.dscal[PW_PDT] = acc / PRIOR_NAVG;
    if (pred_delay) *pred_delay = (float)ws.dscal[PW_PDT];
    ws.scal[PS_NPUB] = 0;
  });
  const double pred_delta_t = ws.dscal[PW_PDT];

  if (n_det == 0) {  // PRI:537-546
    tm.single([&] {
      prior_remove_old(pt, st, tracks, order, t);
      st->t_prev = t;
      *n_out = 0;
    });
    return;
  }

  // ---- phase 2 (first warp): assignment (PRI:561-567), new tracks (PRI:570-580), publication slots (PRI:845-848)
  tm.warp0([&](auto& w) {
    if (n_trk > 0) {
      AssocWs aws;
      aws.dist = ws.dist; aws.star = ws.star; aws.prime = ws.prime; aws.nstar = ws.nstar; aws.cov_r = ws.cov_r;
      aws.cov_c = ws.cov_c;
      munkres_coop(w, aws, ws.cost, n_det, n_trk, ws.assignment);
      w.pfor(n_det, [&](int p) {
        const int a = ws.assignment[p];
        if (a >= 0 && ws.cost[p + n_det * a] > q.dist_threshold) ws.assignment[p] = -1;
      });
    } else {
      w.pfor(n_det, [&](int p) { ws.assignment[p] = -1; });
    }
    w.single([&] {
      int n_pub = 0;
      for (int p = 0; p < n_det; ++p) {
        int s = -1;
        if (ws.assignment[p] >= 0) {
          s = order[ws.assignment[p]];
        } else {  // new track, in detection order
          for (int i = 0; i < max_tracks; ++i)
            if (!((st->used >> i) & 1ull)) { s = i; break; }
          if (s < 0 || st->n_tracks >= max_tracks) { st->overflow = 1; s = -1; }
          else {
            st->used |= 1ull << s;
            order[st->n_tracks++] = (uint8_t)s;
            PriorTrack& tr = tracks[s];
            tr.t_prev = -DBL_MAX;      // "never observed" (the reference leaves t_prev uninitialised, PRI:79-82)
            tr.height_prev = -1.0;     // also tells the fit that the velocity buffer has to be zeroed
            tr.root_prev[0] = tr.root_prev[1] = tr.root_prev[2] = 0.0;
            tr.exists = 0u; tr.num_obs = 0; tr.id = st->next_id++;
          }
        }
        ws.slot[p] = s;
        // persons are published in detection order once their track has enough observations
        ws.out_idx[p] = (s >= 0 && ws.has[p] && tracks[s].num_obs + 1 > q.min_num_obs_track) ? n_pub++ : -1;
        if (track_of && s >= 0) track_of[p] = tracks[s].id;
      }
      ws.scal[PS_NPUB] = n_pub;
    });
  });
  const int n_pub = ws.scal[PS_NPUB];

  // ---- phase 3: the skeleton fits, one warp per group of up to `group` detections (PRI:587-853)
  const int n_groups = (n_det + group - 1) / group;
  tm.per_warp(n_groups, [&](auto& w, int gi) {
    const int wid = w.size() == 1 ? 0 : (tm.rank() / 32);   // serial build: one workspace
    Arena ar(fit_ws + (size_t)wid * fit_ws_stride);
    PriorFitWs fws;
    prior_fit_ws_layout(ar, group, &fws);
    const int p0 = gi * group;
    const int G = n_det - p0 < group ? n_det - p0 : group;
    prior_fit_group(w, pt, G, persons + p0, tracks, ws.slot + p0, ws.out_idx + p0, fused, pred, t, t_prev_global,
                    frame_nr, pred_delta_t, fws);
  });

  // ---- phase 4 (first warp): prune (PRI:867), then merge close tracks (PRI:870-903)
  tm.warp0([&](auto& w) {
    w.single([&] { prior_remove_old(pt, st, tracks, order, t); });
    const int n = st->n_tracks;
    w.pfor(n * n, [&](int e) {
      const int i = e / n, j = e % n;
      if (i < j) ws.D[e] = prior_track_dist(tracks[order[i]], tracks[order[j]]);
    });
    w.single([&] {
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
      if (m != n) {
        uint8_t tmp[PRIOR_MAX_TRACKS];
        for (int i = 0; i < m; ++i) tmp[i] = order[live[i]];
        for (int i = 0; i < m; ++i) order[i] = tmp[i];
        st->n_tracks = m;
      }
      *n_out = n_pub;
      st->t_prev = t;   // PRI:909-910
      ++st->frame_nr;
    });
  });
}

}  // namespace ses3d
