// pose_prior_oracle.cpp — CPU restatement of the reference's pose_prior node. TEST INFRASTRUCTURE ONLY
// (imported by tests/, __graft_entry__.smoke() and bench scripts' CPU-baseline legs; never by the product).
//
// Follows pose_prior/src/pose_prior_mult_node.cpp (PRI) statement by statement:
//   TrackingHypothesis            PRI:68-121      (calc_normed_dist :84-101, calc_3d_dist :103-119)
//   UnaryFactor                   PRI:126-145
//   remove_old_tracks             PRI:191-211     (marker output dropped)
//   addBinaryFactors              PRI:384-481
//   setInitialState               PRI:483-503
//   skeletonCallback              PRI:505-921
// and links the Munkres solver of the skeleton oracle (or the reference's verbatim Hungarian.cpp through
// oracle/_ref/libref_hungarian.so; pose_prior/src/Hungarian.cpp is byte-identical to skeleton_3d's).
//
// Third-party arithmetic NOT under /root/reference: gtsam 4.0.3 (README.md:22). Restated from its published
// algorithm, dense and in double:
//   noiseModel::Gaussian::Covariance(S)   -> diagonal S: sigmas = sqrt(diag); else R = chol_upper(S^-1), e_w = R e
//   noiseModel::Isotropic::Sigma(1, s)    -> e_w = e / s
//   RangeFactor<Point3>(a, b, len)        -> e = |x_b - x_a| - len, de/dx_a = -(x_b-x_a)^T/r, de/dx_b = +(x_b-x_a)^T/r
//   graph.error(x)                        -> sum over factors of 0.5 |e_w|^2
//   LevenbergMarquardtOptimizer, default LevenbergMarquardtParams: lambdaInitial 1e-5, lambdaFactor 10,
//     lambdaUpperBound 1e5, lambdaLowerBound 0, useFixedLambdaFactor, no diagonal damping (damped system =
//     J^T J + lambda I), minModelFidelity 1e-3, maxIterations 100, relativeErrorTol = absoluteErrorTol = 1e-5,
//     errorTol 0; inner loop tryLambda(), outer loop NonlinearOptimizer::defaultOptimize() + checkConvergence().
//     One detail differs between gtsam releases and cannot be checked offline: a trial step counts only when the
//     linearised cost change exceeds 1e-20 (4.0.x as remembered) resp. epsilon x the old linearised error (later
//     releases). Both thresholds are only reached at machine-precision convergence, where the relative-tolerance
//     test of the same function ends the search anyway; 1e-20 is used here and in the kernel.
//   Marginals(graph, x).marginalCovariance(k) -> the k-th 3x3 diagonal block of (J^T J)^-1 at x;
//     IndeterminantLinearSystemException when the Cholesky factorisation meets a non-positive pivot
// gtsam is absent from this image and the reference has no tests => PARITY UNPINNED at this boundary; the fit is
// pinned instead by scipy.optimize.least_squares and a dense numpy inverse (tests/test_pose_prior.py).
//
// Two places where the reference has undefined / non-deterministic behaviour and this file picks one:
//  * a track created for a detection without any usable joint keeps an uninitialised t_prev (PRI:79-82, 739-741);
//    here such a track counts as "never observed" and is removed at the end of the frame that created it;
//  * the OpenMP team appends its private result vectors in thread-arrival order (PRI:582-586, 855-860); here
//    persons are emitted in detection order (the reference built without OpenMP).
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <limits>
#include <thread>
#include <vector>

#include "ses3d.h"

extern "C" void oracle_munkres(int* assignment, double* cost, const double* dist, int n_rows, int n_cols);

namespace {

constexpr int NF = SES3D_NUM_FUSION_KEYPOINTS;
constexpr double MAX_DIST = 1e6;  // PRI:65
constexpr int N_MOV_AVG = 3;      // PRI:53
// FUSION_BODY_PARTS::vel_sigmas, fusion_body_parts.h:33
const double kVelSigmas[NF] = {2., 1., 1., 2., 3., 1., 2., 3., 1., 1., 2., 3., 1., 2., 3., 2., 2., 2., 2., 2., 1.};

struct Bone { int a, b; double len, sigma; int only_without_belly; };
// PRI:434-479: absolute bone lengths (norm_height = false)
const Bone kBonesAbs[] = {
    {8, 9, 0.134, 0.033, 0},  {8, 12, 0.134, 0.033, 0}, {9, 10, 0.449, 0.051, 0},   {10, 11, 0.446, 0.051, 0},
    {12, 13, 0.449, 0.051, 0}, {13, 14, 0.446, 0.051, 0}, {1, 0, 0.20, 0.025, 0},     {1, 2, 0.15, 0.042, 0},
    {1, 5, 0.15, 0.042, 0},   {2, 3, 0.28, 0.045, 0},   {3, 4, 0.25, 0.063, 0},     {5, 6, 0.28, 0.045, 0},
    {6, 7, 0.25, 0.063, 0},   {8, 20, 0.23846, 0.071, 0}, {20, 1, 0.25534, 0.035, 0}, {0, 19, 0.11500, 0.035, 0},
    {8, 1, 0.50, 0.071, 1},   {0, 15, 0.05, 0.035, 0},  {0, 16, 0.05, 0.035, 0},    {15, 17, 0.10, 0.05, 0},
    {16, 18, 0.10, 0.05, 0}};
// PRI:386-431: height-normalised bone lengths (norm_height = true)
const Bone kBonesNorm[] = {
    {8, 9, 0.17, 0.062, 0},   {8, 12, 0.17, 0.062, 0},  {9, 10, 0.694, 0.111, 0},  {10, 11, 0.708, 0.097, 0},
    {12, 13, 0.694, 0.111, 0}, {13, 14, 0.708, 0.097, 0}, {1, 0, 0.33, 0.050, 0},    {1, 2, 0.262, 0.092, 0},
    {1, 5, 0.262, 0.092, 0},  {2, 3, 0.515, 0.071, 0},  {3, 4, 0.444, 0.084, 0},   {5, 6, 0.515, 0.071, 0},
    {6, 7, 0.444, 0.084, 0},  {8, 20, 0.49, 0.05, 0},   {20, 1, 0.51, 0.05, 0},    {0, 19, 0.23, 0.05, 0},
    {8, 1, 1.000, 0.02, 1},   {0, 15, 0.085, 0.06, 0},  {0, 16, 0.085, 0.06, 0},   {15, 17, 0.167, 0.08, 0},
    {16, 18, 0.167, 0.08, 0}};
constexpr int N_BONES = 21;

struct Track {  // TrackingHypothesis PRI:68-82
  bool exists[NF];
  double prev[NF][3];               // prevEstimate (root-relative, height-normalised)
  double vel[NF][N_MOV_AVG][3];     // velBuffer
  double t_prev;
  int num_obs, id;
  double height_prev;
  double root_prev[3];
  explicit Track(int id_) : t_prev(-std::numeric_limits<double>::infinity()), num_obs(0), id(id_), height_prev(-1.0) {
    memset(exists, 0, sizeof exists);
    memset(prev, 0, sizeof prev);
    memset(vel, 0, sizeof vel);
    root_prev[0] = root_prev[1] = root_prev[2] = 0.0;
  }
};

struct Prior {
  ses3d_prior_params prm;
  double limb_sigma_factor;
  void* ref_lib = nullptr;
  void (*ref_hungarian)(int*, double*, double*, int, int) = nullptr;
  // file-scope state of the node
  std::vector<Track> tracks;             // g_tracks PRI:123
  double t_prev = 0.0;                   // g_t_prev PRI:58 (static storage: zero)
  int next_id = 0, frame_nr = 0;         // PRI:59-60
  double delay_buf[N_MOV_AVG];           // g_fb_delay_buffer PRI:54
  long long lm_iterations = 0, lm_inner = 0, fits = 0;
  void reset() {                         // PRI:182-189 (g_t_prev is not reset by the reference either)
    tracks.clear();
    for (double& d : delay_buf) d = prm.avg_delay;
    next_id = 0;
    frame_nr = 0;
  }
};

// ---- small dense helpers -------------------------------------------------------------------------------------
inline void sym6_to_mat(const double c[6], double M[9]) {  // PRI:672-674 etc.
  M[0] = c[0]; M[1] = c[1]; M[2] = c[2];
  M[3] = c[1]; M[4] = c[3]; M[5] = c[4];
  M[6] = c[2]; M[7] = c[4]; M[8] = c[5];
}

// sqrt-information R (upper triangular, row-major 3x3) of noiseModel::Gaussian::Covariance(S)
void sqrt_information(const double S[9], double R[9]) {
  for (int i = 0; i < 9; ++i) R[i] = 0.0;
  const bool diagonal = S[1] == 0.0 && S[2] == 0.0 && S[3] == 0.0 && S[5] == 0.0 && S[6] == 0.0 && S[7] == 0.0;
  if (diagonal) {  // Diagonal::Variances / Isotropic: whiten = e / sigma
    R[0] = 1.0 / sqrt(S[0]); R[4] = 1.0 / sqrt(S[4]); R[8] = 1.0 / sqrt(S[8]);
    return;
  }
  // information = S^-1 (cofactors), R = llt(information).matrixU()
  const double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
  const double det = S[0] * c00 + S[1] * c01 + S[2] * c02;
  const double id = 1.0 / det;
  double I[9];
  I[0] = c00 * id; I[1] = (S[2] * S[7] - S[1] * S[8]) * id; I[2] = (S[1] * S[5] - S[2] * S[4]) * id;
  I[3] = c01 * id; I[4] = (S[0] * S[8] - S[2] * S[6]) * id; I[5] = (S[2] * S[3] - S[0] * S[5]) * id;
  I[6] = c02 * id; I[7] = (S[1] * S[6] - S[0] * S[7]) * id; I[8] = (S[0] * S[4] - S[1] * S[3]) * id;
  // lower Cholesky L of I (reads the lower triangle like Eigen::LLT<.., Lower>), R = L^T
  const double l00 = sqrt(I[0]);
  const double l10 = I[3] / l00, l20 = I[6] / l00;
  const double l11 = sqrt(I[4] - l10 * l10);
  const double l21 = (I[7] - l20 * l10) / l11;
  const double l22 = sqrt(I[8] - l20 * l20 - l21 * l21);
  R[0] = l00; R[1] = l10; R[2] = l20; R[4] = l11; R[5] = l21; R[8] = l22;
}

struct Fit {  // one person's factor graph
  int n = 0;                 // variables
  int key[NF];               // variable -> fusion slot, ascending (gtsam Values iterates by key)
  int var_of[NF];            // slot -> variable or -1
  double R[NF][9], m[NF][3]; // unary factor of every variable
  int n_bones = 0;
  int ba[N_BONES], bb[N_BONES];
  double blen[N_BONES], binv_sigma[N_BONES];
};

double graph_error(const Fit& g, const double* x) {  // NonlinearFactorGraph::error
  double total = 0.0;
  for (int v = 0; v < g.n; ++v) {
    const double d0 = x[3 * v] - g.m[v][0], d1 = x[3 * v + 1] - g.m[v][1], d2 = x[3 * v + 2] - g.m[v][2];
    const double* R = g.R[v];
    const double w0 = R[0] * d0 + R[1] * d1 + R[2] * d2, w1 = R[4] * d1 + R[5] * d2, w2 = R[8] * d2;
    total += 0.5 * (w0 * w0 + w1 * w1 + w2 * w2);
  }
  for (int e = 0; e < g.n_bones; ++e) {
    const double* pa = x + 3 * g.ba[e];
    const double* pb = x + 3 * g.bb[e];
    const double dx = pb[0] - pa[0], dy = pb[1] - pa[1], dz = pb[2] - pa[2];
    const double r = sqrt(dx * dx + dy * dy + dz * dz);
    const double w = (r - g.blen[e]) * g.binv_sigma[e];
    total += 0.5 * (w * w);
  }
  return total;
}

// Linearisation at x: H = J^T J (dense, N x N, N = 3n) and gvec = J^T e_w; also keeps what linear.error(delta) needs.
struct Linear {
  int N;
  std::vector<double> H, g;
  // per factor whitened Jacobian rows and residuals
  double ub[NF][3];               // unary: e_w = R (x - m)
  double bu[N_BONES][3], be[N_BONES];  // bone: unit direction / sigma, whitened residual
};

void linearize(const Fit& g, const double* x, Linear& L) {
  const int N = 3 * g.n;
  L.N = N;
  L.H.assign((size_t)N * N, 0.0);
  L.g.assign(N, 0.0);
  for (int v = 0; v < g.n; ++v) {
    const double d[3] = {x[3 * v] - g.m[v][0], x[3 * v + 1] - g.m[v][1], x[3 * v + 2] - g.m[v][2]};
    const double* R = g.R[v];
    double w[3];
    for (int r = 0; r < 3; ++r) w[r] = R[3 * r] * d[0] + R[3 * r + 1] * d[1] + R[3 * r + 2] * d[2];
    for (int r = 0; r < 3; ++r) L.ub[v][r] = w[r];
    for (int i = 0; i < 3; ++i) {
      double gi = 0.0;
      for (int r = 0; r < 3; ++r) gi += R[3 * r + i] * w[r];
      L.g[3 * v + i] += gi;
      for (int j = 0; j < 3; ++j) {
        double h = 0.0;
        for (int r = 0; r < 3; ++r) h += R[3 * r + i] * R[3 * r + j];
        L.H[(size_t)(3 * v + i) * N + 3 * v + j] += h;
      }
    }
  }
  for (int e = 0; e < g.n_bones; ++e) {
    const int a = g.ba[e], b = g.bb[e];
    const double d[3] = {x[3 * b] - x[3 * a], x[3 * b + 1] - x[3 * a + 1], x[3 * b + 2] - x[3 * a + 2]};
    const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double is = g.binv_sigma[e];
    double u[3];
    for (int i = 0; i < 3; ++i) u[i] = d[i] / r * is;   // whitened d e / d x_b; d e / d x_a = -u
    const double w = (r - g.blen[e]) * is;
    for (int i = 0; i < 3; ++i) L.bu[e][i] = u[i];
    L.be[e] = w;
    for (int i = 0; i < 3; ++i) {
      L.g[3 * a + i] -= u[i] * w;
      L.g[3 * b + i] += u[i] * w;
      for (int j = 0; j < 3; ++j) {
        const double h = u[i] * u[j];
        L.H[(size_t)(3 * a + i) * N + 3 * a + j] += h;
        L.H[(size_t)(3 * b + i) * N + 3 * b + j] += h;
        L.H[(size_t)(3 * a + i) * N + 3 * b + j] -= h;
        L.H[(size_t)(3 * b + i) * N + 3 * a + j] -= h;
      }
    }
  }
}

// GaussianFactorGraph::error(delta) of the undamped linearised system: sum 0.5 |J delta + e_w|^2
double linear_error(const Fit& g, const Linear& L, const double* delta) {
  double total = 0.0;
  for (int v = 0; v < g.n; ++v) {
    const double* R = g.R[v];
    double s = 0.0;
    for (int r = 0; r < 3; ++r) {
      const double w = R[3 * r] * delta[3 * v] + R[3 * r + 1] * delta[3 * v + 1] + R[3 * r + 2] * delta[3 * v + 2] + L.ub[v][r];
      s += w * w;
    }
    total += 0.5 * s;
  }
  for (int e = 0; e < g.n_bones; ++e) {
    const int a = g.ba[e], b = g.bb[e];
    double w = L.be[e];
    for (int i = 0; i < 3; ++i) w += L.bu[e][i] * (delta[3 * b + i] - delta[3 * a + i]);
    total += 0.5 * (w * w);
  }
  return total;
}

// In-place lower Cholesky of the N x N matrix A (row-major); false on a non-positive / non-finite pivot
// (gtsam: IndeterminantLinearSystemException)
bool cholesky(std::vector<double>& A, int N) {
  for (int j = 0; j < N; ++j) {
    double d = A[(size_t)j * N + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * N + k] * A[(size_t)j * N + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = sqrt(d);
    A[(size_t)j * N + j] = d;
    for (int i = j + 1; i < N; ++i) {
      double s = A[(size_t)i * N + j];
      for (int k = 0; k < j; ++k) s -= A[(size_t)i * N + k] * A[(size_t)j * N + k];
      A[(size_t)i * N + j] = s / d;
    }
  }
  return true;
}
void chol_solve(const std::vector<double>& Lc, int N, double* b) {  // b <- (L L^T)^-1 b
  for (int i = 0; i < N; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * N + k] * b[k];
    b[i] = s / Lc[(size_t)i * N + i];
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < N; ++k) s -= Lc[(size_t)k * N + i] * b[k];
    b[i] = s / Lc[(size_t)i * N + i];
  }
}

// LevenbergMarquardtOptimizer(graph, initial).optimize() with default parameters (see the header comment).
void lm_optimize(const Prior& P, const Fit& g, double* x, long long* outer, long long* inner) {
  const ses3d_prior_params& q = P.prm;
  const int N = 3 * g.n;
  double error = graph_error(g, x);
  double lambda = q.lm_lambda_initial;
  int iterations = 0;
  if (error <= 0.0) return;                              // errorTol = 0
  if (iterations >= q.lm_max_iterations) return;
  Linear L;
  std::vector<double> A, delta(N), newx(N);
  double current_error;
  do {
    current_error = error;
    // ---- iterate(): linearise once, then try lambdas until tryLambda() says stop
    linearize(g, x, L);
    for (;;) {
      ++*inner;
      A = L.H;
      for (int i = 0; i < N; ++i) A[(size_t)i * N + i] += lambda;   // buildDampedSystem, no diagonal damping
      bool step_ok = false, stop_searching = false;
      double new_error = std::numeric_limits<double>::infinity();
      if (cholesky(A, N)) {
        for (int i = 0; i < N; ++i) delta[i] = -L.g[i];
        chol_solve(A, N, delta.data());
        std::vector<double> zero(N, 0.0);
        const double old_lin = linear_error(g, L, zero.data());
        const double new_lin = linear_error(g, L, delta.data());
        const double lin_change = old_lin - new_lin;
        if (lin_change >= 0) {
          for (int i = 0; i < N; ++i) newx[i] = x[i] + delta[i];   // Values::retract, Point3: x + delta
          new_error = graph_error(g, newx.data());
          const double cost_change = error - new_error;
          if (lin_change > 1e-20) {
            const double fidelity = cost_change / lin_change;
            step_ok = fidelity > q.lm_min_model_fidelity;
          }
          if (fabs(cost_change) < q.lm_relative_error_tol * error) stop_searching = true;
        }
      }
      if (step_ok) {                       // decreaseLambda: new state, iterations + 1
        for (int i = 0; i < N; ++i) x[i] = newx[i];
        error = new_error;
        lambda = std::max(0.0, lambda / q.lm_lambda_factor);
        ++iterations;
        break;
      } else if (!stop_searching) {        // increaseLambda
        lambda *= q.lm_lambda_factor;
        if (lambda >= q.lm_lambda_upper_bound) break;
      } else {
        break;
      }
    }
    ++*outer;
    // ---- checkConvergence(relativeErrorTol, absoluteErrorTol, errorTol, currentError, error())
    bool converged;
    if (error <= 0.0) converged = true;
    else {
      const double abs_dec = current_error - error;
      const double rel_dec = abs_dec / current_error;
      converged = (q.lm_relative_error_tol != 0.0 && rel_dec <= q.lm_relative_error_tol) || abs_dec <= q.lm_absolute_error_tol;
    }
    if (!(iterations < q.lm_max_iterations && !converged && std::isfinite(current_error))) break;
  } while (true);
}

// Marginals(graph, x): joint covariance (J^T J)^-1; false = IndeterminantLinearSystemException
bool marginals(const Fit& g, const double* x, std::vector<double>& Sigma) {
  Linear L;
  linearize(g, x, L);
  const int N = L.N;
  std::vector<double> A = L.H;
  if (!cholesky(A, N)) return false;
  Sigma.assign((size_t)N * N, 0.0);
  std::vector<double> col(N);
  for (int c = 0; c < N; ++c) {
    std::fill(col.begin(), col.end(), 0.0);
    col[c] = 1.0;
    chol_solve(A, N, col.data());
    for (int r = 0; r < N; ++r) Sigma[(size_t)r * N + c] = col[r];
  }
  return true;
}

// ---- TrackingHypothesis methods ---------------------------------------------------------------------------------
double calc_normed_dist(const Prior& P, const Track& tr, const ses3d_person_cov& person, double t) {  // PRI:84-101
  const double delta_t = t - tr.t_prev;
  int used = 0;
  double dist = 0;
  for (int k = 0; k < NF; ++k) {
    const ses3d_keypoint_cov& kp = person.keypoints[k];
    if (kp.score > P.prm.min_score && tr.exists[k]) {
      const double px = tr.prev[k][0] * tr.height_prev + tr.root_prev[0];
      const double py = tr.prev[k][1] * tr.height_prev + tr.root_prev[1];
      const double pz = tr.prev[k][2] * tr.height_prev + tr.root_prev[2];
      const double dx = kp.x - px, dy = kp.y - py, dz = kp.z - pz;
      dist += sqrt(dx * dx + dy * dy + dz * dz) / (kVelSigmas[k] * delta_t);
      ++used;
    }
  }
  return used > 0 ? dist / used : MAX_DIST;
}

double calc_3d_dist(const Track& a, const Track& b) {  // PRI:103-119
  int used = 0;
  double dist = 0;
  for (int k = 0; k < NF; ++k) {
    if (!a.exists[k] || !b.exists[k]) continue;
    double d2 = 0;
    for (int i = 0; i < 3; ++i) {
      const double pa = a.prev[k][i] * a.height_prev + a.root_prev[i];
      const double pb = b.prev[k][i] * b.height_prev + b.root_prev[i];
      d2 += (pa - pb) * (pa - pb);
    }
    dist += sqrt(d2);
    ++used;
  }
  return used > 0 ? dist / used : MAX_DIST;
}

void remove_old_tracks(Prior& P, double t) {  // PRI:191-211
  auto& v = P.tracks;
  v.erase(std::remove_if(v.begin(), v.end(), [&](const Track& tr) { return t - tr.t_prev > P.prm.t_max_unobserved; }),
          v.end());
}

double stamp_to_sec(int64_t ns) {  // ros::Time::toSec(): (double)sec + 1e-9 * (double)nsec
  const int64_t sec = ns / 1000000000LL, nsec = ns % 1000000000LL;
  return (double)sec + 1e-9 * (double)nsec;
}

// One person of skeletonCallback's parallel loop (PRI:587-853). Returns false when num_meas == 0 (PRI:739-741).
bool fuse_person(Prior& P, const ses3d_person_cov& person, Track& tr, double t, double pred_delta_t,
                 ses3d_person_cov* fused, ses3d_person_cov* pred) {
  const ses3d_prior_params& q = P.prm;
  memset(fused, 0, sizeof *fused);
  memset(pred, 0, sizeof *pred);
  fused->id = (uint32_t)tr.id;
  pred->id = (uint32_t)tr.id;

  bool measured[NF] = {false}, use_velocity[NF] = {false};
  double meas[NF][3] = {{0}};
  double Runary[NF][9];
  int num_meas = 0;

  ses3d_keypoint_cov root, neck;   // value-initialised messages: everything zero (PRI:631)
  memset(&root, 0, sizeof root);
  memset(&neck, 0, sizeof neck);
  double height = 1.0;
  const ses3d_keypoint_cov& hl = person.keypoints[SES3D_FBP_LHIP];
  const ses3d_keypoint_cov& hr = person.keypoints[SES3D_FBP_RHIP];
  const ses3d_keypoint_cov& sl = person.keypoints[SES3D_FBP_LSHOULDER];
  const ses3d_keypoint_cov& sr = person.keypoints[SES3D_FBP_RSHOULDER];
  if (q.pose_method == SES3D_POSE_H36M) {  // PRI:633-636
    root = person.keypoints[SES3D_FBP_MIDHIP];
    neck = person.keypoints[SES3D_FBP_NECK];
  } else {  // PRI:637-656
    if (hl.score > 0.0f && hr.score > 0.0f) {
      root.x = (hl.x + hr.x) / 2.0; root.y = (hl.y + hr.y) / 2.0; root.z = (hl.z + hr.z) / 2.0;
      root.score = (hl.score + hr.score) / 2.0f;
    }
    if (sl.score > 0.0f && sr.score > 0.0f) {
      neck.x = (sl.x + sr.x) / 2.0; neck.y = (sl.y + sr.y) / 2.0; neck.z = (sl.z + sr.z) / 2.0;
      neck.score = (sl.score + sr.score) / 2.0f;
    }
  }

  if (root.score > q.min_score) {  // PRI:658-694
    if (q.normalize_by_height) {
      if (neck.score > q.min_score) {
        const double dx = neck.x - root.x, dy = neck.y - root.y, dz = neck.z - root.z;
        height = sqrt(dx * dx + dy * dy + dz * dz);
      } else {
        height = 0.60;
      }
    }
    double rc[9];
    if (q.pose_method == SES3D_POSE_H36M) {
      sym6_to_mat(root.cov, rc);
    } else {
      double a[9], b[9];
      sym6_to_mat(hl.cov, a);
      sym6_to_mat(hr.cov, b);
      for (int i = 0; i < 9; ++i) rc[i] = (a[i] + b[i]) / 2.0;
    }
    for (int i = 0; i < 9; ++i) rc[i] = rc[i] / (height * height) / (q.root_sigma_factor * q.root_sigma_factor);
    sqrt_information(rc, Runary[SES3D_FBP_MIDHIP]);
    meas[SES3D_FBP_MIDHIP][0] = meas[SES3D_FBP_MIDHIP][1] = meas[SES3D_FBP_MIDHIP][2] = 0.0;
    measured[SES3D_FBP_MIDHIP] = true;
    ++num_meas;
  }

  if (tr.height_prev < 0.0) {  // PRI:699-702
    tr.height_prev = height;
    tr.root_prev[0] = root.x; tr.root_prev[1] = root.y; tr.root_prev[2] = root.z;
  }

  for (int k = 0; k < NF; ++k) {  // PRI:704-719
    if (k == SES3D_FBP_MIDHIP) continue;
    const ses3d_keypoint_cov& kp = person.keypoints[k];
    if (kp.score > q.min_score) {
      double c[9];
      sym6_to_mat(kp.cov, c);
      for (int i = 0; i < 9; ++i) c[i] = c[i] / (height * height);
      sqrt_information(c, Runary[k]);
      meas[k][0] = (kp.x - root.x) / height; meas[k][1] = (kp.y - root.y) / height; meas[k][2] = (kp.z - root.z) / height;
      measured[k] = true;
      ++num_meas;
    }
  }

  if (q.pose_method == SES3D_POSE_SIMPLE && neck.score > q.min_score) {  // PRI:721-737
    // (a directly measured Neck slot would make Values::insert throw in the reference; the skeleton_3d stage
    //  never fills that slot in "simple" mode, S3D:139-141 — here the shoulder mean simply replaces it)
    double a[9], b[9], c[9];
    sym6_to_mat(sl.cov, a);
    sym6_to_mat(sr.cov, b);
    for (int i = 0; i < 9; ++i) c[i] = (a[i] + b[i]) / 2.0 / (height * height);
    sqrt_information(c, Runary[SES3D_FBP_NECK]);
    meas[SES3D_FBP_NECK][0] = (neck.x - root.x) / height;
    meas[SES3D_FBP_NECK][1] = (neck.y - root.y) / height;
    meas[SES3D_FBP_NECK][2] = (neck.z - root.z) / height;
    if (!measured[SES3D_FBP_NECK]) ++num_meas;
    measured[SES3D_FBP_NECK] = true;
  }

  if (num_meas == 0) return false;  // PRI:739-741

  // setInitialState, PRI:483-503
  for (int k = 0; k < NF; ++k) {
    if (tr.exists[k] && !measured[k]) {
      tr.exists[k] = false;
      memset(tr.vel[k], 0, sizeof tr.vel[k]);
    }
  }
  for (int k = 0; k < NF; ++k) {
    if (!measured[k]) continue;
    if (!tr.exists[k]) {
      tr.exists[k] = true;
      tr.prev[k][0] = meas[k][0]; tr.prev[k][1] = meas[k][1]; tr.prev[k][2] = meas[k][2];
    } else {
      use_velocity[k] = true;
    }
  }

  // graph: unary factors + addBinaryFactors (PRI:384-481)
  Fit g;
  for (int k = 0; k < NF; ++k) g.var_of[k] = -1;
  for (int k = 0; k < NF; ++k) {
    if (!measured[k]) continue;
    const int v = g.n++;
    g.key[v] = k;
    g.var_of[k] = v;
    memcpy(g.R[v], Runary[k], sizeof g.R[v]);
    memcpy(g.m[v], meas[k], sizeof g.m[v]);
  }
  const Bone* bones = q.normalize_by_height ? kBonesNorm : kBonesAbs;
  for (int e = 0; e < N_BONES; ++e) {
    const Bone& b = bones[e];
    if (!measured[b.a] || !measured[b.b]) continue;
    if (b.only_without_belly && measured[SES3D_FBP_BELLY]) continue;
    const int i = g.n_bones++;
    g.ba[i] = g.var_of[b.a];
    g.bb[i] = g.var_of[b.b];
    g.blen[i] = b.len;
    g.binv_sigma[i] = 1.0 / (b.sigma * P.limb_sigma_factor);
  }

  // LevenbergMarquardtOptimizer optimizer(graph, curr_track.prevEstimate); result = optimizer.optimize(); PRI:746-749
  std::vector<double> x(3 * g.n);
  for (int v = 0; v < g.n; ++v)
    for (int i = 0; i < 3; ++i) x[3 * v + i] = tr.prev[g.key[v]][i];
  ++P.fits;
  lm_optimize(P, g, x.data(), &P.lm_iterations, &P.lm_inner);

  std::vector<double> Sigma;
  const bool use_marginals = marginals(g, x.data(), Sigma);  // PRI:760-767
  const int N = 3 * g.n;

  for (int v = 0; v < g.n; ++v) {  // PRI:770-837
    const int k = g.key[v];
    ses3d_keypoint_cov& out = fused->keypoints[k];
    const double rootv[3] = {root.x, root.y, root.z};
    double jf[3];
    for (int i = 0; i < 3; ++i) jf[i] = x[3 * v + i] * height + rootv[i];
    out.x = jf[0]; out.y = jf[1]; out.z = jf[2];
    if (k == SES3D_FBP_MIDHIP) out.score = std::max(q.min_score, root.score);
    else if (k == SES3D_FBP_NECK) out.score = std::max(q.min_score, neck.score);
    else out.score = std::max(q.min_score, person.keypoints[k].score);

    double cov[9];
    if (use_marginals) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[3 * i + j] = Sigma[(size_t)(3 * v + i) * N + 3 * v + j] * height * height;
    } else {
      for (int i = 0; i < 9; ++i) cov[i] = 0.0;
      cov[0] = cov[4] = cov[8] = q.default_res_sigma * q.default_res_sigma;
    }
    if (k == SES3D_FBP_MIDHIP)
      for (int i = 0; i < 9; ++i) cov[i] *= (q.root_sigma_factor * q.root_sigma_factor);
    out.cov[0] = cov[0]; out.cov[1] = cov[1]; out.cov[2] = cov[2]; out.cov[3] = cov[4]; out.cov[4] = cov[5]; out.cov[5] = cov[8];

    double jp[3] = {jf[0], jf[1], jf[2]};
    if (use_velocity[k]) {  // PRI:819-826
      double* slot = tr.vel[k][P.frame_nr % N_MOV_AVG];
      for (int i = 0; i < 3; ++i)
        slot[i] = ((x[3 * v + i] * height + rootv[i]) - (tr.prev[k][i] * tr.height_prev + tr.root_prev[i])) / (t - P.t_prev);
      for (int i = 0; i < 3; ++i) {
        double acc = 0.0;   // std::accumulate from Zero, in buffer order
        for (int b = 0; b < N_MOV_AVG; ++b) acc += tr.vel[k][b][i];
        jp[i] += acc / N_MOV_AVG * pred_delta_t;
      }
    }
    ses3d_keypoint_cov& po = pred->keypoints[k];
    po.x = jp[0]; po.y = jp[1]; po.z = jp[2];
    po.score = out.score;
    for (int i = 0; i < 6; ++i) po.cov[i] = out.cov[i];
    po.cov[0] += q.pred_noise_sigma * q.pred_noise_sigma;  // addToKeypointCovariance PRI:222-226
    po.cov[3] += q.pred_noise_sigma * q.pred_noise_sigma;
    po.cov[5] += q.pred_noise_sigma * q.pred_noise_sigma;
  }

  // PRI:839-843
  tr.t_prev = t;
  for (int v = 0; v < g.n; ++v)
    for (int i = 0; i < 3; ++i) tr.prev[g.key[v]][i] = x[3 * v + i];
  tr.height_prev = height;
  tr.root_prev[0] = root.x; tr.root_prev[1] = root.y; tr.root_prev[2] = root.z;
  ++tr.num_obs;
  return true;
}

// skeletonCallback, PRI:505-921
int step(Prior& P, int64_t stamp_ns, int n_cams, const float* fb_delay, int n_det, const ses3d_person_cov* persons,
         int h_max, ses3d_person_cov* fused, ses3d_person_cov* pred, int32_t* n_out, float* pred_delay,
         int32_t* track_of) {
  const ses3d_prior_params& q = P.prm;
  const double t = stamp_to_sec(stamp_ns);

  double curr = 0.0;  // PRI:513-526
  int n_valid = 0;
  for (int c = 0; c < n_cams; ++c) {
    const float d = fb_delay ? fb_delay[c] : -1.0f;
    if (d > 0.0f) { curr += (double)d; ++n_valid; }
  }
  if (n_valid > 0) curr /= n_valid; else curr = q.avg_delay;
  P.delay_buf[P.frame_nr % N_MOV_AVG] = curr;
  double acc = 0.0;
  for (int i = 0; i < N_MOV_AVG; ++i) acc += P.delay_buf[i];
  const double pred_delta_t = acc / N_MOV_AVG;
  if (pred_delay) *pred_delay = (float)pred_delta_t;
  *n_out = 0;
  if (track_of) for (int i = 0; i < h_max; ++i) track_of[i] = -1;

  const int n_hyp = (int)P.tracks.size();
  if (n_det == 0) {  // PRI:537-546
    remove_old_tracks(P, t);
    P.t_prev = t;
    return 0;
  }

  std::vector<int> assignment;
  if (n_hyp > 0) {  // PRI:550-568
    std::vector<double> C((size_t)n_det * n_hyp);
    assignment.assign(n_det, -1);
    double cost = 0.0;
    for (int tr = 0; tr < n_hyp; ++tr)
      for (int p = 0; p < n_det; ++p) C[p + (size_t)n_det * tr] = calc_normed_dist(P, P.tracks[tr], persons[p], t);
    if (P.ref_hungarian) P.ref_hungarian(assignment.data(), &cost, C.data(), n_det, n_hyp);
    else oracle_munkres(assignment.data(), &cost, C.data(), n_det, n_hyp);
    for (int i = 0; i < n_det; ++i)
      if (assignment[i] >= 0 && C[i + (size_t)n_det * assignment[i]] > q.dist_threshold) assignment[i] = -1;
  }

  std::vector<int> track_ids(n_det);  // PRI:570-580
  for (int p = 0; p < n_det; ++p) {
    if (!assignment.empty() && assignment[p] >= 0) track_ids[p] = assignment[p];
    else {
      P.tracks.push_back(Track(P.next_id));
      track_ids[p] = (int)P.tracks.size() - 1;
      ++P.next_id;
    }
  }

  int n_pub = 0, rc = 0;
  for (int p = 0; p < n_det; ++p) {  // PRI:587-853
    Track& tr = P.tracks[track_ids[p]];
    if (track_of) track_of[p] = tr.id;
    ses3d_person_cov pf, pp;
    if (!fuse_person(P, persons[p], tr, t, pred_delta_t, &pf, &pp)) continue;
    if (tr.num_obs > q.min_num_obs_track) {  // PRI:845-848
      if (n_pub < h_max) { fused[n_pub] = pf; pred[n_pub] = pp; }
      else rc = SES3D_E_CAPACITY;
      ++n_pub;
    }
  }
  n_pub = std::min(n_pub, h_max);

  remove_old_tracks(P, t);  // PRI:867

  for (size_t i = 0; i < P.tracks.size(); ++i) {  // PRI:870-903
    for (size_t j = i + 1; j < P.tracks.size();) {
      if (calc_3d_dist(P.tracks[i], P.tracks[j]) < q.merge_dist_thresh) {
        const int id_to_remove = P.tracks[j].id;
        P.tracks.erase(P.tracks.begin() + j);
        for (int k = 0; k < n_pub; ++k) {
          if ((int)fused[k].id == id_to_remove) {
            fused[k].id = (uint32_t)P.tracks[i].id;
            pred[k].id = (uint32_t)P.tracks[i].id;
          }
        }
      } else {
        ++j;
      }
    }
  }
  *n_out = n_pub;
  P.t_prev = t;   // PRI:909-910
  ++P.frame_nr;
  return rc;
}

}  // namespace

extern "C" {

void prior_oracle_default_params(ses3d_prior_params* p) {
  memset(p, 0, sizeof *p);
  p->pose_method = SES3D_POSE_SIMPLE;
  p->normalize_by_height = 0;
  p->min_num_obs_track = 10;
  p->lm_max_iterations = 100;
  p->min_score = 0.10f;
  p->pred_noise_sigma = 0.12;
  p->default_res_sigma = 0.10;
  p->avg_delay = 0.10;
  p->root_sigma_factor = 100.0;
  p->t_max_unobserved = 1.0;
  p->dist_threshold = 5.0;
  p->merge_dist_thresh = 0.20;
  p->lm_lambda_initial = 1e-5;
  p->lm_lambda_factor = 10.0;
  p->lm_lambda_upper_bound = 1e5;
  p->lm_relative_error_tol = 1e-5;
  p->lm_absolute_error_tol = 1e-5;
  p->lm_min_model_fidelity = 1e-3;
}

// n_sequences independent trackers behind one handle
struct PriorSet { std::vector<Prior> seq; };

void* prior_oracle_create(const ses3d_prior_params* prm, int32_t n_sequences, const char* ref_hungarian_so) {
  PriorSet* s = new PriorSet;
  s->seq.resize(std::max(1, n_sequences));
  void* lib = nullptr;
  void* fn = nullptr;
  if (ref_hungarian_so) {
    lib = dlopen(ref_hungarian_so, RTLD_NOW | RTLD_LOCAL);
    if (lib) fn = dlsym(lib, "ref_hungarian_assignmentoptimal");
    if (!fn) { delete s; return nullptr; }
  }
  for (Prior& P : s->seq) {
    P.prm = *prm;
    P.limb_sigma_factor = prm->normalize_by_height ? 2.0 : 1.0;  // PRI:934-937
    P.ref_lib = lib;
    P.ref_hungarian = reinterpret_cast<void (*)(int*, double*, double*, int, int)>(fn);
    P.reset();
  }
  return s;
}
void prior_oracle_destroy(void* h) { delete static_cast<PriorSet*>(h); }
void prior_oracle_reset(void* h) { for (Prior& P : static_cast<PriorSet*>(h)->seq) P.reset(); }

// Same array shapes as ses3d_prior_run (sequence-major). n_threads > 1 runs sequences in parallel.
int prior_oracle_run(void* h, int32_t n_sequences, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons,
                     const int32_t* n_persons, const int64_t* stamp_ns, int32_t n_cams, const float* fb_delay,
                     ses3d_person_cov* fused, ses3d_person_cov* pred, int32_t* n_out, float* pred_delay,
                     int32_t* track_of, int32_t n_threads) {
  PriorSet* s = static_cast<PriorSet*>(h);
  if ((int)s->seq.size() < n_sequences) return SES3D_E_INVALID;
  std::atomic<int> next(0), rc_all(0);
  auto work = [&] {
    for (;;) {
      const int q = next.fetch_add(1);
      if (q >= n_sequences) break;
      for (int f = 0; f < n_frames; ++f) {
        const size_t i = (size_t)q * n_frames + f;
        const int rc = step(s->seq[q], stamp_ns[i], n_cams, fb_delay ? fb_delay + i * n_cams : nullptr, n_persons[i],
                            persons + i * h_max, h_max, fused + i * h_max, pred + i * h_max, n_out + i,
                            pred_delay ? pred_delay + i : nullptr, track_of ? track_of + i * h_max : nullptr);
        if (rc != 0) rc_all = rc;
      }
    }
  };
  if (n_threads <= 1) work();
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  return rc_all;
}

int prior_oracle_get_tracks(void* h, int32_t sequence, int32_t* ids, int32_t* num_obs) {
  PriorSet* s = static_cast<PriorSet*>(h);
  if (sequence < 0 || sequence >= (int)s->seq.size()) return SES3D_E_INVALID;
  const Prior& P = s->seq[sequence];
  for (size_t i = 0; i < P.tracks.size(); ++i) {
    if (ids) ids[i] = P.tracks[i].id;
    if (num_obs) num_obs[i] = P.tracks[i].num_obs;
  }
  return (int)P.tracks.size();
}

// {fits, outer LM iterations, inner lambda trials} summed over all sequences
void prior_oracle_stats(void* h, int64_t out[3]) {
  out[0] = out[1] = out[2] = 0;
  for (const Prior& P : static_cast<PriorSet*>(h)->seq) { out[0] += P.fits; out[1] += P.lm_iterations; out[2] += P.lm_inner; }
}

}  // extern "C"
