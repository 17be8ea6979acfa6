// kernels_tri.cu — K3 triangulate for sm_100a: one warp per (frame, hypothesis), 4 warps per CTA, FP32 or FP64.
// The algorithm lives in tri_core.h. Joint positions are tolerance-checked against the
// oracle (1e-3 m FP32 / 1e-4 m FP64), so FMA contraction stays on here.
#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "tri_core.h"

namespace ses3d {

extern __shared__ __align__(16) unsigned char smem_raw[];

template <class T, int kWarpsPerCta>
__global__ void __launch_bounds__(32 * kWarpsPerCta, sizeof(T) == 4 ? 24 / kWarpsPerCta : 1)
k_triangulate(const Tables tb, int p_max, int h_cap, size_t ws_bytes, const ses3d_person2d* __restrict__ persons,
              const int8_t* __restrict__ hyp_det, const uint32_t* __restrict__ work, int32_t* work_count,
              ses3d_person_cov* __restrict__ tmp, int32_t* __restrict__ keep, float* far_scratch, int dynamic) {
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  const int n_work = work_count[0];
  WarpTeam tm;
  Arena ar(smem_raw + (size_t)warp * ws_bytes);
  TriWs<T> ws;
  tri_ws_layout<T>(ar, tb.n_cams, &ws);
  if (far_scratch)
    ws.far_scratch = far_scratch + ((size_t)blockIdx.x * kWarpsPerCta + warp) * 32 * FAR_COV_STRIDE;
  // dynamic work distribution: work_count[1] is the next unclaimed item (zeroed with work_count[0] by K2's launcher).
  // A hypothesis can cost several times the average (far joints are re-solved exactly), so items are handed out one by
  // one instead of in fixed strides - no warp is left holding a queue behind a slow item.
  const int total_warps = (int)gridDim.x * kWarpsPerCta;
  int w_static = (int)blockIdx.x * kWarpsPerCta + warp - total_warps;
  for (;;) {
    int w = 0;
    if (dynamic) {
      if (lane == 0) w = atomicAdd(work_count + 1, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
    } else {
      w = (w_static += total_warps);
    }
    if (w >= n_work) break;
    const size_t fh = work[w];  // frame * h_cap + hypothesis
    const size_t f = fh / h_cap;
    triangulate_hypothesis<T>(tm, tb, p_max, persons + f * tb.n_cams * p_max, hyp_det + fh * tb.n_cams, ws, tmp + fh,
                              keep + fh);
    tm.sync();
  }
}

size_t triangulate_far_scratch_bytes_per_warp();

template <class T, int W>
static cudaError_t launch_tri_impl(LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                                   const int8_t* hyp_det, const uint32_t* work, int32_t* work_count,
                                   ses3d_person_cov* tmp, int32_t* keep, float* far_scratch, size_t far_scratch_bytes,
                                   cudaStream_t st) {
  const size_t ws_bytes = tri_ws_bytes<T>(tb.n_cams);
  const size_t smem = ws_bytes * W;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  // persistent grid: as many CTAs as fit on the chip at the kernel's occupancy (a multiple of the SM count),
  // never more than the work. The occupancy query is cached per (kernel, shared-memory size).
  const void* fn = reinterpret_cast<const void*>(&k_triangulate<T, W>);
  int per_sm = 0;
  for (int i = 0; i < cfg.n_occ; ++i)
    if (cfg.occ[i].fn == fn && cfg.occ[i].smem == smem) per_sm = cfg.occ[i].per_sm;
  if (per_sm == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_triangulate<T, W>, 32 * W, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (cfg.n_occ < 8) cfg.occ[cfg.n_occ++] = {fn, smem, per_sm};
  }
  const size_t units = (size_t)d.n_frames * d.h_cap;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((units + W - 1) / W, (size_t)cfg.n_sm * per_sm));
  if (sizeof(T) != 4 || far_scratch_bytes < (size_t)grid * W * triangulate_far_scratch_bytes_per_warp()) far_scratch = nullptr;
  k_triangulate<T, W><<<grid, 32 * W, smem, st>>>(tb, d.p_max, d.h_cap, ws_bytes, persons, hyp_det, work, work_count, tmp,
                                                  keep, far_scratch, cfg.tri_dynamic);
  return cudaGetLastError();
}

cudaError_t init_kernels_tri(int) {
  const int budget = 200 * 1024;
  cudaError_t e;
#define SES_ATTR(T, W)                                                                                                  \
  if ((e = cudaFuncSetAttribute(k_triangulate<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  SES_ATTR(float, 2) SES_ATTR(float, 4) SES_ATTR(float, 8) SES_ATTR(double, 2) SES_ATTR(double, 4) SES_ATTR(double, 8)
#undef SES_ATTR
  return cudaSuccess;
}

size_t triangulate_far_scratch_bytes_per_warp() { return (size_t)32 * FAR_COV_STRIDE * sizeof(float); }
// upper bound of the warps a launch can have on this device (11 CTAs of 2 warps per SM measured; 32 is the hard limit)
size_t triangulate_far_scratch_bytes(const LaunchCfg& cfg) { return (size_t)cfg.n_sm * 32 * triangulate_far_scratch_bytes_per_warp(); }

cudaError_t launch_triangulate(LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                               const int8_t* hyp_det, const uint32_t* work, int32_t* work_count,
                               ses3d_person_cov* tmp, int32_t* keep, float* far_scratch, size_t far_scratch_bytes,
                               cudaStream_t st) {
  // measured on B200 (hall16 x 6): 2 warps/CTA 0.98 ms, 4: 1.00 ms, 8: 1.10 ms per 8192 frames
  if (tb.prm.precision == SES3D_PRECISION_FP64) {
    if (cfg.tri_warps_f64 == 2) return launch_tri_impl<double, 2>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
    if (cfg.tri_warps_f64 == 8) return launch_tri_impl<double, 8>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
    return launch_tri_impl<double, 4>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
  }
  if (cfg.tri_warps == 2) return launch_tri_impl<float, 2>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
  if (cfg.tri_warps == 8) return launch_tri_impl<float, 8>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
  return launch_tri_impl<float, 4>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep, far_scratch, far_scratch_bytes, st);
}

}  // namespace ses3d
