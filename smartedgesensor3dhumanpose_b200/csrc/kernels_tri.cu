// kernels_tri.cu — K3 triangulate for sm_100a: one warp per (frame, hypothesis), 4 warps per CTA, FP32 or FP64.
// The algorithm lives in tri_core.h. Joint positions are tolerance-checked against the
// oracle (1e-3 m FP32 / 1e-4 m FP64), so FMA contraction stays on here.
#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "tri_core.h"

namespace ses3d {

extern __shared__ __align__(16) unsigned char smem_raw[];

static_assert(TRI_BUCKETS == kTriBuckets, "tri_core.h and launch.h disagree on the number of work sub-lists");

// Work items come in TRI_BUCKETS sub-lists (by observation count, K2b). They are consumed as ONE virtual list, longest
// items first, every sub-list padded to a multiple of `pad` slots so that the kWarpsPerCta items a lockstep CTA claims
// at once come from the same sub-list. Returns the work word of virtual index v, or false for a padding slot.
struct WorkView {
  int n[TRI_BUCKETS];
  int total;   // padded length of the virtual list
  __device__ __forceinline__ void load(const int32_t* work_count, int pad) {
    total = 0;
    for (int b = 0; b < TRI_BUCKETS; ++b) {
      n[b] = work_count[b];
      total += (n[b] + pad - 1) / pad * pad;
    }
  }
  __device__ __forceinline__ bool at(const uint32_t* work, size_t cap, int pad, int v, uint32_t* item) const {
    for (int b = TRI_BUCKETS - 1; b >= 0; --b) {
      const int padded = (n[b] + pad - 1) / pad * pad;
      if (v < padded) {
        if (v >= n[b]) return false;
        *item = work[(size_t)b * cap + v];
        return true;
      }
      v -= padded;
    }
    return false;
  }
};

template <class T, int kWarpsPerCta, bool kLockstep>
// register budget: 24 warps per SM in FP32 mode; 12 in FP64 mode (168 registers: three 4-warp CTAs per SM - at 171 only
// two fit, measured 3.55 -> 3.89 ms on config 3)
__global__ void __launch_bounds__(32 * kWarpsPerCta, sizeof(T) == 4 ? 24 / kWarpsPerCta : (kWarpsPerCta <= 12 ? 12 / kWarpsPerCta : 1))
k_triangulate(const Tables tb, int p_max, int h_cap, size_t work_cap, size_t ws_bytes,
              const ses3d_person2d* __restrict__ persons, const int8_t* __restrict__ hyp_det,
              const uint32_t* __restrict__ work, int32_t* work_count, ses3d_person_cov* __restrict__ tmp,
              int32_t* __restrict__ keep, float* far_scratch, int dynamic) {
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  Arena ar(smem_raw + (size_t)warp * ws_bytes);
  TriWs<T> ws;
  tri_ws_layout<T>(ar, tb.n_cams, &ws);
  if (far_scratch)
    ws.far_scratch = far_scratch + ((size_t)blockIdx.x * kWarpsPerCta + warp) * 32 * FAR_COV_STRIDE;
  WorkView wv;
  int32_t* claim = work_count + TRI_BUCKETS;
  if (kLockstep) {
    // Lockstep CTA: its warps claim kWarpsPerCta neighbouring items (same sub-list = similar cost) and pass the
    // phases of triangulate_hypothesis together (team.phase() = CTA barrier). The body is ~50 KB of mostly
    // straight-line code; warps that drift apart each stream their own part of it through the 32 KB instruction
    // cache (measured: 49 % of all stall samples were instruction fetch), warps that stay together share the lines.
    __shared__ int s_base;
    LockstepWarpTeam tm;
    wv.load(work_count, kWarpsPerCta);
    for (;;) {
      if (threadIdx.x == 0) s_base = atomicAdd(claim, kWarpsPerCta);
      __syncthreads();
      const int base = s_base;
      __syncthreads();
      if (base >= wv.total) break;
      uint32_t item = 0;
      const bool live = wv.at(work, work_cap, kWarpsPerCta, base + warp, &item);   // false: padding slot
      const size_t fh = item;  // frame * h_cap + hypothesis
      const size_t f = fh / h_cap;
      triangulate_hypothesis<T>(tm, tb, p_max, persons + f * tb.n_cams * p_max, hyp_det + fh * tb.n_cams, ws, tmp + fh,
                                keep + fh, nullptr, live);
    }
    return;
  }
  WarpTeam tm;
  wv.load(work_count, 1);
  // dynamic work distribution: a hypothesis can cost several times the average, so items are handed out one by one
  // instead of in fixed strides - no warp is left holding a queue behind a slow item.
  const int total_warps = (int)gridDim.x * kWarpsPerCta;
  int w_static = (int)blockIdx.x * kWarpsPerCta + warp - total_warps;
  for (;;) {
    int w = 0;
    if (dynamic) {
      if (lane == 0) w = atomicAdd(claim, 1);
      w = __shfl_sync(0xffffffffu, w, 0);
    } else {
      w = (w_static += total_warps);
    }
    if (w >= wv.total) break;
    uint32_t item = 0;
    if (!wv.at(work, work_cap, 1, w, &item)) continue;
    const size_t fh = item;  // frame * h_cap + hypothesis
    const size_t f = fh / h_cap;
    triangulate_hypothesis<T>(tm, tb, p_max, persons + f * tb.n_cams * p_max, hyp_det + fh * tb.n_cams, ws, tmp + fh,
                              keep + fh);
    tm.sync();
  }
}

size_t triangulate_far_scratch_bytes_per_warp();

template <class T, int W, bool L>
static cudaError_t launch_tri_impl(LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                                   const int8_t* hyp_det, const uint32_t* work, int32_t* work_count,
                                   ses3d_person_cov* tmp, int32_t* keep, float* far_scratch, size_t far_scratch_bytes,
                                   cudaStream_t st) {
  const size_t ws_bytes = tri_ws_bytes<T>(tb.n_cams);
  const size_t smem = ws_bytes * W;
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  // persistent grid: as many CTAs as fit on the chip at the kernel's occupancy (a multiple of the SM count),
  // never more than the work. The occupancy query is cached per (kernel, shared-memory size).
  const void* fn = reinterpret_cast<const void*>(&k_triangulate<T, W, L>);
  int per_sm = 0;
  for (int i = 0; i < cfg.n_occ; ++i)
    if (cfg.occ[i].fn == fn && cfg.occ[i].smem == smem) per_sm = cfg.occ[i].per_sm;
  if (per_sm == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_triangulate<T, W, L>, 32 * W, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (cfg.n_occ < 8) cfg.occ[cfg.n_occ++] = {fn, smem, per_sm};
  }
  const size_t units = (size_t)d.n_frames * d.h_cap;
  const unsigned grid = (unsigned)std::max<size_t>(1, std::min<size_t>((units + W - 1) / W, (size_t)cfg.n_sm * per_sm));
  if (sizeof(T) != 4 || far_scratch_bytes < (size_t)grid * W * triangulate_far_scratch_bytes_per_warp()) far_scratch = nullptr;
  k_triangulate<T, W, L><<<grid, 32 * W, smem, st>>>(tb, d.p_max, d.h_cap, units, ws_bytes, persons, hyp_det, work,
                                                     work_count, tmp, keep, far_scratch, cfg.tri_dynamic);
  return cudaGetLastError();
}

cudaError_t init_kernels_tri(int) {
  const int budget = 200 * 1024;
  cudaError_t e;
#define SES_ATTR(T, W)                                                                                                  \
  if ((e = cudaFuncSetAttribute(k_triangulate<T, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e; \
  if ((e = cudaFuncSetAttribute(k_triangulate<T, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  SES_ATTR(float, 2) SES_ATTR(float, 4) SES_ATTR(float, 8) SES_ATTR(float, 12) SES_ATTR(float, 16) SES_ATTR(double, 2) SES_ATTR(double, 4) SES_ATTR(double, 8)
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
#define SES_TRI(T, W)                                                                                                    \
  return cfg.tri_lockstep ? launch_tri_impl<T, W, true>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep,       \
                                                        far_scratch, far_scratch_bytes, st)                              \
                          : launch_tri_impl<T, W, false>(cfg, tb, d, persons, hyp_det, work, work_count, tmp, keep,      \
                                                         far_scratch, far_scratch_bytes, st)
  if (tb.prm.precision == SES3D_PRECISION_FP64) {
    if (cfg.tri_warps_f64 == 2) SES_TRI(double, 2);
    if (cfg.tri_warps_f64 == 8) SES_TRI(double, 8);
    SES_TRI(double, 4);
  }
  if (cfg.tri_warps == 2) SES_TRI(float, 2);
  if (cfg.tri_warps == 8) SES_TRI(float, 8);
  if (cfg.tri_warps == 12) SES_TRI(float, 12);
  if (cfg.tri_warps == 16) SES_TRI(float, 16);
  SES_TRI(float, 4);
#undef SES_TRI
}

}  // namespace ses3d
