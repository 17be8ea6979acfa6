// kernels_tri.cu — K3 triangulate for sm_100a: one CTA per (frame, hypothesis), FP32 or FP64.
// The algorithm lives in tri_core.h. Joint positions are tolerance-checked against the
// oracle (1e-3 m FP32 / 1e-4 m FP64), so FMA contraction stays on here.
#include "launch.h"
#include "tri_core.h"

namespace ses3d {

extern __shared__ __align__(16) unsigned char smem_raw[];

template <class T>
__global__ void __launch_bounds__(128)
k_triangulate(const Tables tb, int n_frames, int p_max, int h_cap, const ses3d_person2d* __restrict__ persons,
              const int8_t* __restrict__ hyp_det, ses3d_person_cov* __restrict__ tmp, int32_t* __restrict__ keep) {
  const size_t fh = blockIdx.x;  // frame * h_cap + hypothesis
  const size_t f = fh / h_cap;
  if (f >= (size_t)n_frames) return;
  Arena ar(smem_raw);
  TriWs<T> ws;
  tri_ws_layout<T>(ar, tb.n_cams, &ws);
  BlockTeam tm;
  triangulate_hypothesis<T>(tm, tb, p_max, persons + f * tb.n_cams * p_max, hyp_det + fh * tb.n_cams, ws, tmp + fh,
                            keep + fh);
}

cudaError_t launch_triangulate(const Tables& tb, LaunchDims d, const ses3d_person2d* persons, const int8_t* hyp_det,
                               ses3d_person_cov* tmp, int32_t* keep, cudaStream_t st) {
  const bool f64 = tb.prm.precision == SES3D_PRECISION_FP64;
  const size_t smem = f64 ? tri_ws_bytes<double>(tb.n_cams) : tri_ws_bytes<float>(tb.n_cams);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  const unsigned grid = (unsigned)((size_t)d.n_frames * d.h_cap);
  cudaError_t e;
  if (f64) {
    e = cudaFuncSetAttribute(k_triangulate<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_triangulate<double><<<grid, 128, smem, st>>>(tb, d.n_frames, d.p_max, d.h_cap, persons, hyp_det, tmp, keep);
  } else {
    e = cudaFuncSetAttribute(k_triangulate<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_triangulate<float><<<grid, 128, smem, st>>>(tb, d.n_frames, d.p_max, d.h_cap, persons, hyp_det, tmp, keep);
  }
  return cudaGetLastError();
}

}  // namespace ses3d
