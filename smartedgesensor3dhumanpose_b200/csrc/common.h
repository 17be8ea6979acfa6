// common.h — shared definitions of the device algorithms (host+device).
//
// Every per-frame algorithm in csrc/*_core.h is written once as a __host__ __device__
// template over a "team" (team.h): on the GPU the team is a warp or a CTA, in the CPU
// test build (tests/hostsim) it is a serial loop. The product never runs the serial
// instantiation; it exists so that the device logic can be checked against the oracle
// in the GPU-less container.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "ses3d.h"

#if defined(__CUDACC__)
#define SES_HD __host__ __device__ __forceinline__
#define SES_HDN __host__ __device__ __noinline__ inline
#else
#define SES_HD inline
#define SES_HDN inline
#endif

namespace ses3d {

constexpr int NKP = SES3D_NUM_KEYPOINTS;          // 17
constexpr int NFUS = SES3D_NUM_FUSION_KEYPOINTS;  // 21
constexpr double MAX_COSTS = 1e6;                 // S3D:43

// Skeleton tables (S3D:81-145), selected once by pose_method.
struct SkeletonModel {
  int parent[17];
  double limb_len[17];
  double limb_sigma[17];
  int fusion_idx[17];
};

// Per-camera constants in the precisions the hot path uses.
struct CamF {     // float view: normalize_keypoints (S3D:314-317) + Matrix34f (S3D:1208-1211)
  float P[12];
  float fx, fy, cx, cy;
  float inv_fx2, inv_fxfy, inv_fy2, pad_;   // 1/(fx*fx), 1/(fx*fy), 1/(fy*fy): covariance scaling of the sigma points
};
struct CamD {     // double view: FP64 triangulation mode and reprojection (REP:152-163, 196-197)
  double P[12];
  double fx, fy, cx, cy, Tx, Ty;
  double width, height;
};

// Read-only tables built by ses3d_create (K0) and replicated on the GPU.
struct Tables {
  int n_cams;
  const CamF* camf;   // [C]
  const CamD* camd;   // [C]
  const float* F;     // [C(C-1)/2][9] row-major, get_fundamental_idx order
  const int* f_row;   // [C]: start index of row i in F (sum_{ii<i} (C-ii-1)), so idx(i,j) = f_row[i] + j-i-1
  SkeletonModel model;
  ses3d_params prm;
  int exact_mode;     // bit 0: exact re-solve of far / high-residual joints, bit 1: exact covariance of far joints
                      // (3 = default; SES3D_TRI_EXACT overrides it for A/B measurements only)
};

SES_HD int fundamental_idx(const Tables& tb, int i, int j) { return tb.f_row[i] + j - i - 1; }  // S3D:242-253, i<j

// Eigen's fixed-size reduction orders: 3 terms a + (b + c); 4 terms (a + b) + (c + d).
template <class T> SES_HD T sum3(T a, T b, T c) { return a + (b + c); }
template <class T> SES_HD T sum4(T a, T b, T c, T d) { return (a + b) + (c + d); }

SES_HD int ses_ctz(uint32_t x) {  // index of the lowest set bit (x != 0)
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
SES_HD float ses_sqrt(float x) { return sqrtf(x); }
SES_HD double ses_sqrt(double x) { return sqrt(x); }
SES_HD float ses_abs(float x) { return fabsf(x); }
SES_HD double ses_abs(double x) { return fabs(x); }

// Single-precision operations that are never contracted into FMAs, for the code paths that must reproduce the
// reference's x86-64 (SSE2, no FMA) arithmetic bit for bit inside translation units compiled with FMA contraction on.
// Host builds are compiled with -ffp-contract=off, where the plain operators already mean this.
#if defined(__CUDA_ARCH__)
SES_HD float xmul(float a, float b) { return __fmul_rn(a, b); }
SES_HD float xadd(float a, float b) { return __fadd_rn(a, b); }
SES_HD float xsub(float a, float b) { return __fsub_rn(a, b); }
SES_HD float xdiv(float a, float b) { return __fdiv_rn(a, b); }
SES_HD float xsqrt(float a) { return __fsqrt_rn(a); }
#else
SES_HD float xmul(float a, float b) { return a * b; }
SES_HD float xadd(float a, float b) { return a + b; }
SES_HD float xsub(float a, float b) { return a - b; }
SES_HD float xdiv(float a, float b) { return a / b; }
SES_HD float xsqrt(float a) { return sqrtf(a); }
#endif

// carve typed arrays out of a byte workspace (shared memory on the GPU)
struct Arena {
  unsigned char* p;
  size_t used;
  SES_HD explicit Arena(void* base) : p(static_cast<unsigned char*>(base)), used(0) {}
  template <class T> SES_HD T* take(size_t n) {
    used = (used + alignof(T) - 1) / alignof(T) * alignof(T);
    T* r = reinterpret_cast<T*>(p + used);
    used += n * sizeof(T);
    return r;
  }
};
// same arithmetic without a buffer: workspace sizing on the host
struct ArenaSizer {
  size_t used = 0;
  template <class T> SES_HD T* take(size_t n) {
    used = (used + alignof(T) - 1) / alignof(T) * alignof(T);
    used += n * sizeof(T);
    return nullptr;
  }
};

}  // namespace ses3d
