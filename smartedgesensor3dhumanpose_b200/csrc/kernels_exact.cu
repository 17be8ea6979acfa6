// kernels_exact.cu — K2 associate, K4 finalize, K6 reproject for sm_100a.
//
// Compiled with -fmad=false: association indices must be bit-exact with the reference's
// x86-64 (SSE2, no FMA) arithmetic, and finalize/reproject are FP64 paths whose results
// then match the CPU oracle bit for bit as well. The algorithms live in *_core.h; this
// file only binds a CTA to a frame and carves the shared-memory workspace.
#include <algorithm>
#include <type_traits>
#include <cstdlib>

#include "assoc_core.h"
#include "fin_core.h"
#include "launch.h"
#include "reproj_core.h"

namespace ses3d {

extern __shared__ __align__(16) unsigned char smem_raw[];

// K2a: pair table + compact detection list of one frame per CTA (assoc_core.h::pairs_frame). tile_warps > 0 makes this
// the dense-frame instance: it carries the line buffers of the tiled pair pass, looks only at frames the first
// instance found dense (their detection count is in the frame's meta record) and redoes their set-up itself.
__global__ void __launch_bounds__(256)
k_pairs(const Tables tb, int n_frames, int p_max, const ses3d_person2d* __restrict__ persons,
        const int32_t* __restrict__ n_persons, float* nk_scratch, double* pair_table, unsigned char* meta_base,
        size_t meta_stride, int tile_warps, int defer_dense, int n_parts) {
  const int f = (int)(blockIdx.x / (unsigned)n_parts), part = (int)(blockIdx.x % (unsigned)n_parts);
  if (f >= n_frames) return;
  const int C = tb.n_cams;
  const FrameMeta meta = frame_meta_at(meta_base + (size_t)f * meta_stride, C, p_max);
  if (tile_warps > 0 && !pairs_frame_is_dense(*meta.n_valid, C)) return;
  Arena ar(smem_raw);
  AssocWs ws;
  pair_ws_layout(ar, C, p_max, nk_scratch == nullptr, &ws, tile_warps);
  if (nk_scratch) ws.nk = nk_scratch + (size_t)f * C * p_max * NKP * 2;
  ws.E = pair_table + (size_t)f * assoc_pair_table_entries(C, p_max);
  BlockTeam tm;
  pairs_frame(tm, tb, p_max, persons + (size_t)f * C * p_max, n_persons + (size_t)f * C, ws, meta, defer_dense != 0, part,
              n_parts);
}

// K2b: the sequential camera rounds, one WARP per frame (assoc_core.h::rounds_frame), plus the K3 work list.
// A lockstep variant (the frames of a CTA passing the camera rounds together behind a CTA barrier, to share
// instruction-cache lines: the rounds are ~70 KB of branchy code and a third of the stall samples were instruction
// fetch) was measured on B200 and removed: 0.945 ms against 0.898 ms per 16384 frames (K2a + K2b) - the rounds of
// different frames differ too much in length for the barrier to pay.
// kMode ROUNDS_BLOCK: one CTA per frame (BlockTeam; the cost-matrix gathers spread over the whole CTA, Munkres on its
// first warp) - for rigs whose frames are heavy and few (crowds: a frame's 64 camera rounds take ~15 ms on a single
// warp, and 512 frames do not fill the GPU with one warp each).
enum { ROUNDS_WARP = 0, ROUNDS_BLOCK = 2 };
// per-warp staging area of a low-latency k_rounds launch: pair-table entries, person scores, slots, camera offsets
__host__ __device__ inline size_t rounds_stage_bytes(int C, int p_max, int stage_entries) {
  const size_t n = (size_t)C * p_max;
  return ((size_t)stage_entries * 8 + n * 4 + n * 2 + (size_t)(C + 1) * 2 + 15) / 16 * 16;
}
template <int kMode> struct RoundsTeam { typedef WarpTeam type; };
template <> struct RoundsTeam<ROUNDS_BLOCK> { typedef BlockTeam type; };

template <int kMode>
__global__ void __launch_bounds__(kMode == ROUNDS_BLOCK ? 256 : 128)
k_rounds(const Tables tb, int n_frames, int p_max, int h_cap, size_t ws_bytes, const int32_t* __restrict__ n_persons,
         const double* pair_table, unsigned char* meta_base, size_t meta_stride, int8_t* __restrict__ hyp_det,
         int32_t* __restrict__ n_hyp, int32_t* __restrict__ n_hung, int32_t* overflow, int32_t* hyp_of_dump,
         int32_t* __restrict__ keep, uint32_t* __restrict__ work, int32_t* work_count, int32_t* __restrict__ n_out_zero,
         int stage_entries) {
  const int warp = kMode == ROUNDS_BLOCK ? 0 : (int)(threadIdx.x >> 5);
  const int f = kMode == ROUNDS_BLOCK ? (int)blockIdx.x : (int)blockIdx.x * (int)(blockDim.x >> 5) + warp;
  const int C = tb.n_cams;
  typename RoundsTeam<kMode>::type tm;
  if (f >= n_frames) return;   // padding warp of the last CTA (no CTA-wide barrier follows in warp mode)
  Arena ar(smem_raw + (size_t)warp * ws_bytes);
  AssocWs ws;
  round_ws_layout(ar, C, p_max, h_cap, &ws);
  const FrameMeta meta = frame_meta_at(meta_base + (size_t)f * meta_stride, C, p_max);
  ws.voff = meta.voff; ws.vslot = meta.vslot; ws.pscore = meta.pscore;
  ws.E = const_cast<double*>(pair_table) + (size_t)f * assoc_pair_table_entries(C, p_max);
  if (kMode == ROUNDS_WARP && stage_entries > 0) {
    // Low-latency launches (a handful of frames): every camera round starts with dependent loads of the frame's meta
    // record and pair-table entries, ~1.5 us of L2 round trips per round on an otherwise idle SM. Stage the frame's
    // part of both in shared memory once (coalesced, all loads in flight together).
    const int n_max = C * p_max, nv = *meta.n_valid, need = nv * (nv - 1) / 2, lane = (int)(threadIdx.x & 31u);
    if (need <= stage_entries) {
      unsigned char* sp = smem_raw + (size_t)(blockDim.x >> 5) * ws_bytes + (size_t)warp * rounds_stage_bytes(C, p_max, stage_entries);
      double* Es = reinterpret_cast<double*>(sp);
      float* ps = reinterpret_cast<float*>(Es + stage_entries);
      uint16_t* vs = reinterpret_cast<uint16_t*>(ps + n_max);
      uint16_t* vo = vs + n_max;
      for (int i = lane; i < need; i += 32) Es[i] = ws.E[i];
      for (int i = lane; i < nv; i += 32) { ps[i] = meta.pscore[i]; vs[i] = meta.vslot[i]; }
      for (int i = lane; i <= C; i += 32) vo[i] = meta.voff[i];
      __syncwarp();
      ws.E = Es; ws.pscore = ps; ws.vslot = vs; ws.voff = vo;
    }
  }
  int8_t* hd = hyp_det + (size_t)f * h_cap * C;
  rounds_frame(tm, tb, p_max, h_cap, n_persons + (size_t)f * C, *meta.n_valid, ws, hd, n_hyp + f,
               n_hung ? n_hung + f : nullptr, overflow);
  // work list for K3: every hypothesis with at least two observations (S3D:684); order is irrelevant
  // because results are addressed by (frame, hypothesis)
  tm.pfor(h_cap, [&](int h) { keep[(size_t)f * h_cap + h] = 0; });
  // The list is kept in kTriBuckets sub-lists by observation count (the cost of a hypothesis grows with it): K3 runs
  // the long items first and its lockstep CTAs take items of similar cost.
  const size_t work_cap = (size_t)n_frames * h_cap;
  tm.pfor(ws.scal[SC_N_HYP], [&](int h) {
    const int nobs = ws.hyp_nobs[h];
    if (nobs < 2) return;
    const int b = (nobs < kTriBuckets + 1 ? nobs : kTriBuckets + 1) - 2;
    const int at = atomicAdd(work_count + b, 1);
    work[(size_t)b * work_cap + at] = (uint32_t)((size_t)f * h_cap + h);
  });
  tm.single([&] { if (n_out_zero) n_out_zero[f] = 0; });
  if (hyp_of_dump) {  // [C][p_max] hypothesis index of each detection
    int32_t* ho = hyp_of_dump + (size_t)f * C * p_max;
    tm.pfor(C * p_max, [&](int i) { ho[i] = -1; });
    const int nh = ws.scal[SC_N_HYP];
    tm.pfor(nh * C, [&](int i) {
      const int h = i / C, c = i % C;
      const int d = hd[h * C + c];
      if (d >= 0) ho[c * p_max + d] = h;
    });
  }
}

__global__ void __launch_bounds__(64)
k_finalize(const Tables tb, int n_frames, int h_cap, const int32_t* __restrict__ n_hyp, ses3d_person_cov* tmp,
           const int32_t* __restrict__ keep, ses3d_person_cov* __restrict__ out, int32_t* __restrict__ n_out) {
  const int f = blockIdx.x;
  if (f >= n_frames) return;
  Arena ar(smem_raw);
  FinWs ws;
  fin_ws_layout(ar, h_cap, &ws);
  BlockTeam tm;
  finalize_frame(tm, tb, h_cap, n_hyp[f], tmp + (size_t)f * h_cap, keep + (size_t)f * h_cap, ws,
                 out + (size_t)f * h_cap, n_out + f);
}

// kThreads / kMinBlocks only set the register budget (launch bounds); the launchers pick the instance from the CTA size
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_reproject(const Tables tb, int n_frames, int h_max, int s_cap,
            const ses3d_person_cov* __restrict__ persons3d,
            const int32_t* __restrict__ n_persons3d, ses3d_person2d* __restrict__ out, int32_t* __restrict__ n_out) {
  const int f = blockIdx.x;
  if (f >= n_frames) return;
  Arena ar(smem_raw);
  ReprojWs ws;
  reproj_ws_layout(ar, tb.n_cams, (int)(blockDim.x >> 5), s_cap, &ws);
  BlockTeam tm;
  reproject_frame(tm, tb, h_max, persons3d + (size_t)f * h_max, n_persons3d[f], ws,
                  out + (size_t)f * tb.n_cams * h_max, n_out + (size_t)f * tb.n_cams);
}

// K4 + K6 fused (process calls): the CTA first compacts / merges the frame's skeletons into the PersonCovList
// (finalize_frame), then re-projects that list into every camera (reproject_frame). The two steps use the shared
// memory one after the other (aliased workspaces); the list is read back through L1 / L2, it never waits for DRAM.
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_finproj(const Tables tb, int n_frames, int h_max, int s_cap, const int32_t* __restrict__ n_hyp,
          ses3d_person_cov* tmp, const int32_t* __restrict__ keep, ses3d_person_cov* out3d, int32_t* n_out3d,
          ses3d_person2d* __restrict__ out2d, int32_t* __restrict__ n_out2d) {
  const int f = blockIdx.x;
  if (f >= n_frames) return;
  BlockTeam tm;
  {
    Arena ar(smem_raw);
    FinWs ws;
    fin_ws_layout(ar, h_max, &ws);
    finalize_frame(tm, tb, h_max, n_hyp[f], tmp + (size_t)f * h_max, keep + (size_t)f * h_max, ws,
                   out3d + (size_t)f * h_max, n_out3d + f);
  }
  __threadfence_block();
  tm.sync();
  Arena ar(smem_raw);
  ReprojWs ws;
  reproj_ws_layout(ar, tb.n_cams, (int)(blockDim.x >> 5), s_cap, &ws);
  reproject_frame(tm, tb, h_max, out3d + (size_t)f * h_max, n_out3d[f], ws,
                  out2d + (size_t)f * tb.n_cams * h_max, n_out2d + (size_t)f * tb.n_cams);
}

// Batch of independent assignment problems, one warp each (test / diagnostics entry: lets the tests drive the
// warp-cooperative Munkres with adversarial tied matrices). cost: [n][rows*cols] column-major.
__global__ void __launch_bounds__(32)
k_munkres_batch(int n, int rows, int cols, const double* __restrict__ cost, int32_t* __restrict__ assignment) {
  const int i = blockIdx.x;
  if (i >= n) return;
  Arena ar(smem_raw);
  AssocWs ws;
  round_ws_layout(ar, 1, cols, rows, &ws);
  WarpTeam tm;
  const int n_e = rows * cols;
  tm.pfor(n_e, [&](int e) { ws.cost[e] = cost[(size_t)i * n_e + e]; });
  munkres_coop(tm, ws, ws.cost, rows, cols, ws.assignment);
  tm.pfor(rows, [&](int r) { assignment[(size_t)i * rows + r] = ws.assignment[r]; });
}

static const size_t kSmemBudget = 200 * 1024;   // of the 227 KB a CTA may opt in to
static const size_t kAssocSmemTarget = 64 * 1024;

cudaError_t launch_munkres_batch(int n, int rows, int cols, const double* cost, int32_t* assignment, cudaStream_t st) {
  if (rows < 1 || cols < 1 || rows > 1024 || cols > 127) return cudaErrorInvalidValue;
  const size_t smem = round_ws_bytes(1, cols, rows);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  k_munkres_batch<<<n, 32, smem, st>>>(n, rows, cols, cost, assignment);
  return cudaGetLastError();
}


size_t associate_smem_bytes(int n_cams, int p_max, int h_cap, bool* needs_scratch) {
  (void)h_cap;
  size_t b = pair_ws_bytes(n_cams, p_max, true);
  bool scratch = b > kAssocSmemTarget;
  if (scratch) b = pair_ws_bytes(n_cams, p_max, false);
  if (needs_scratch) *needs_scratch = scratch;
  return b;
}

int associate_launches(const LaunchCfg& cfg, int p_max) { return (cfg.pairs_tiled && p_max >= 4) ? 3 : 2; }
size_t associate_pair_table_bytes(int n_cams, int p_max) { return assoc_pair_table_entries(n_cams, p_max) * sizeof(double); }
size_t associate_meta_bytes(int n_cams, int p_max) { return frame_meta_bytes(n_cams, p_max); }

static int env_int(const char* name, int fallback) {
  const char* v = getenv(name);
  return v ? atoi(v) : fallback;
}

// Once per handle (ses3d_create): SM count, environment overrides, and the opt-in to large dynamic shared memory for
// every kernel of this translation unit. The attribute is a ceiling, not a reservation: occupancy still follows the
// bytes each launch actually asks for.
cudaError_t init_kernels_tri(int device);
cudaError_t init_kernels(LaunchCfg* cfg, int device) {
  cfg->device = device;
  cudaError_t e = cudaDeviceGetAttribute(&cfg->n_sm, cudaDevAttrMultiProcessorCount, device);
  if (e != cudaSuccess) return e;
  if (cfg->n_sm <= 0) cfg->n_sm = 148;
  cfg->assoc_threads = env_int("SES3D_ASSOC_THREADS", 0);
  if (cfg->assoc_threads) cfg->assoc_threads = std::max(32, std::min(256, cfg->assoc_threads / 32 * 32));
  cfg->rounds_warps = std::max(1, std::min(4, env_int("SES3D_ROUNDS_WARPS", 4)));
  cfg->reproj_scap = std::max(1, env_int("SES3D_REPROJ_SCAP", 6));
  cfg->reproj_threads = std::max(0, std::min(512, env_int("SES3D_REPROJ_THREADS", 128) / 32 * 32));   // 0 = one warp per camera, at most 16
  cfg->tri_warps = env_int("SES3D_TRI_WARPS", 4);
  cfg->tri_warps_f64 = env_int("SES3D_TRI_WARPS_F64", 4);
  cfg->tri_dynamic = env_int("SES3D_TRI_DYNAMIC", 1);
  cfg->tri_lockstep = env_int("SES3D_TRI_LOCKSTEP", 1);
  cfg->pairs_tiled = env_int("SES3D_PAIRS_TILED", 0);
  cfg->rounds_block = env_int("SES3D_ROUNDS_BLOCK", -1);
  cfg->pairs_split = env_int("SES3D_PAIRS_SPLIT", 0);
  cfg->latency_frames = env_int("SES3D_LATENCY_FRAMES", -1);
  if (cfg->latency_frames < 0) cfg->latency_frames = cfg->n_sm;
  const int budget = (int)kSmemBudget;
  if ((e = cudaFuncSetAttribute(k_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_rounds<ROUNDS_WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_rounds<ROUNDS_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
#define SES_RP_ATTR(T_, B_)                                                                                                 \
  if ((e = cudaFuncSetAttribute(k_reproject<T_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e; \
  if ((e = cudaFuncSetAttribute(k_finproj<T_, B_>, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  SES_RP_ATTR(128, 5) SES_RP_ATTR(256, 3) SES_RP_ATTR(512, 1)
#undef SES_RP_ATTR
  if ((e = cudaFuncSetAttribute(k_munkres_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, budget)) != cudaSuccess) return e;
  return init_kernels_tri(device);
}

// K2 = K2a + K2b. n_out_zero (nullable): per-frame output count to clear (frames without work items never reach the
// finalize step of a fused pipeline).
cudaError_t launch_associate(const LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                             const int32_t* n_persons, float* nk_scratch, double* pair_table, unsigned char* meta,
                             int8_t* hyp_det, int32_t* n_hyp, int32_t* n_hung, int32_t* overflow, int32_t* hyp_of_dump,
                             int32_t* keep, uint32_t* work, int32_t* work_count, cudaStream_t st) {
  bool scratch;
  const size_t smem = associate_smem_bytes(tb.n_cams, d.p_max, d.h_cap, &scratch);
  if (smem > kSmemBudget) return cudaErrorInvalidConfiguration;
  // the pair pass is pure throughput work: wide CTAs. B200, hall16 x 6 (ms per 16384 frames): see RESULTS.md
  int threads = scratch ? 256 : 128;
  if (d.n_frames <= 296) threads = 256;   // fewer frames than two per SM: latency mode
  if (cfg.assoc_threads) threads = cfg.assoc_threads;
  if (scratch && !nk_scratch) return cudaErrorInvalidValue;
  if (!pair_table || !meta) return cudaErrorInvalidValue;
  // [0, kTriBuckets) item counts per observation-count bucket, [kTriBuckets] K3's claim counter
  cudaError_t e = cudaMemsetAsync(work_count, 0, (kTriBuckets + 1) * sizeof(int32_t), st);
  if (e != cudaSuccess) return e;
  const size_t meta_stride = frame_meta_bytes(tb.n_cams, d.p_max);
  // sparse frames (and every frame of rigs with < 4 slots per camera) in the small-workspace instance, dense frames
  // in a second instance that carries the line buffers of the tiled pair pass
  const int tile_warps = (cfg.pairs_tiled && d.p_max >= 4) ? 8 : 0;
  // big rigs with few frames per launch (crowds): several CTAs share a frame's pair list so that the GPU is full
  int n_parts = 1;
  if (cfg.pairs_split > 0) n_parts = cfg.pairs_split;
  // B200, 64 x 20 crowd, 512 frames (K2a + K2b ms): 1 slice 34.3, 2: 31.2, 4: 27.4, 8: 26.6, 16: 26.6
  else if (scratch && tile_warps == 0) n_parts = std::max(1, std::min(16, (28 * cfg.n_sm + d.n_frames - 1) / d.n_frames));
  // a handful of frames (the per-message calls of the ROS nodes): slices of ~256 pairs, one pair per thread
  else if (tile_warps == 0 && cfg.latency_frames > 0 && d.n_frames * 16 <= cfg.latency_frames) {
    const int n = tb.n_cams * d.p_max;
    n_parts = std::max(1, std::min(16, n * (n - 1) / 2 / 256));
  }
  if (tile_warps > 0) n_parts = 1;   // the dense instance reads the meta record slice 0 of the first instance wrote
  k_pairs<<<d.n_frames * n_parts, threads, smem, st>>>(tb, d.n_frames, d.p_max, persons, n_persons,
                                                       scratch ? nk_scratch : nullptr, pair_table, meta, meta_stride, 0,
                                                       tile_warps > 0, n_parts);
  if (tile_warps > 0) {
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const size_t smem_dense = pair_ws_bytes(tb.n_cams, d.p_max, !scratch, tile_warps);
    if (smem_dense > kSmemBudget) return cudaErrorInvalidConfiguration;
    k_pairs<<<d.n_frames, 32 * tile_warps, smem_dense, st>>>(tb, d.n_frames, d.p_max, persons, n_persons,
                                                             scratch ? nk_scratch : nullptr, pair_table, meta, meta_stride,
                                                             tile_warps, 0, 1);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  const size_t rws = round_ws_bytes(tb.n_cams, d.p_max, d.h_cap);
  int warps = cfg.rounds_warps;
  while (warps > 1 && rws * warps > kSmemBudget) warps >>= 1;
  if (rws * warps > kSmemBudget) return cudaErrorInvalidConfiguration;
#define SES_ROUNDS(MODE, GRID, THREADS, SMEM)                                                                              \
  k_rounds<MODE><<<(GRID), (THREADS), (SMEM), st>>>(tb, d.n_frames, d.p_max, d.h_cap, rws, n_persons, pair_table, meta,      \
                                                    meta_stride, hyp_det, n_hyp, n_hung, overflow, hyp_of_dump, keep, work,  \
                                                    work_count, nullptr, stage_entries)
  // heavy frames (big cost matrices: >= 256 entries) that are too few to fill the GPU one warp each: CTA per frame
  const bool block_mode = cfg.rounds_block == 1 ||
                          (cfg.rounds_block < 0 && d.h_cap * d.p_max >= 256 && d.n_frames < 8 * cfg.n_sm * warps);
  int stage_entries = 0;
  if (!block_mode && cfg.latency_frames > 0 && d.n_frames * 16 <= cfg.latency_frames) {
    // one frame per CTA, the frame's meta record and pair table staged in shared memory (as much as fits)
    warps = 1;
    const size_t room = kSmemBudget - rws;
    const int n = tb.n_cams * d.p_max;
    stage_entries = n * (n - 1) / 2;
    while (stage_entries > 0 && rounds_stage_bytes(tb.n_cams, d.p_max, stage_entries) > room) stage_entries /= 2;
  }
  const size_t stage = stage_entries > 0 ? rounds_stage_bytes(tb.n_cams, d.p_max, stage_entries) : 0;
  if (block_mode) SES_ROUNDS(ROUNDS_BLOCK, d.n_frames, 256, rws);
  else SES_ROUNDS(ROUNDS_WARP, (d.n_frames + warps - 1) / warps, 32 * warps, (rws + stage) * warps);
#undef SES_ROUNDS
  return cudaGetLastError();
}

cudaError_t launch_finalize(const Tables& tb, LaunchDims d, const int32_t* n_hyp, ses3d_person_cov* tmp,
                            const int32_t* keep, ses3d_person_cov* out, int32_t* n_out, cudaStream_t st) {
  const size_t smem = fin_ws_bytes(d.h_cap);
  if (smem > kSmemBudget) return cudaErrorInvalidConfiguration;
  k_finalize<<<d.n_frames, 32, smem, st>>>(tb, d.n_frames, d.h_cap, n_hyp, tmp, keep, out, n_out);
  return cudaGetLastError();
}

// Cameras are dealt round-robin to the CTA's warps, s_cap persons per batch. B200, hall16 x 6, ms per 16384 frames
// (fused with finalize): 4 warps 1.41, 8 warps 1.77, 16 warps (one per camera, 1 CTA / SM at 128 registers) 3.97 -
// the FP64 projection needs ~96 registers, so small CTAs keep more warps resident.
static void reproject_config(const LaunchCfg& cfg, const Tables& tb, int n_frames, int h_max, int* threads, int* s_cap,
                             size_t* smem) {
  int warps = cfg.reproj_threads > 0 ? cfg.reproj_threads / 32 : std::min(tb.n_cams, 16);
  // fewer frames than SMs: occupancy is irrelevant, one warp per camera shortens the frame's critical path
  if (cfg.latency_frames > 0 && n_frames <= cfg.latency_frames) warps = std::min(tb.n_cams, 16);
  warps = std::max(1, std::min(warps, std::min(tb.n_cams, 16)));   // __launch_bounds__(512)
  *threads = 32 * warps;
  *s_cap = reproj_s_cap(tb.n_cams, h_max, cfg.reproj_scap);
  *smem = reproj_ws_bytes(tb.n_cams, warps, *s_cap);
}

cudaError_t launch_finproj(const LaunchCfg& cfg, const Tables& tb, LaunchDims d, const int32_t* n_hyp,
                           ses3d_person_cov* tmp, const int32_t* keep, ses3d_person_cov* out3d, int32_t* n_out3d,
                           ses3d_person2d* out2d, int32_t* n_out2d, cudaStream_t st) {
  int threads, s_cap;
  size_t smem;
  reproject_config(cfg, tb, d.n_frames, d.h_cap, &threads, &s_cap, &smem);
  smem = std::max(smem, fin_ws_bytes(d.h_cap));
  if (smem > kSmemBudget) return cudaErrorInvalidConfiguration;
#define SES_FP(T_, B_) k_finproj<T_, B_><<<d.n_frames, threads, smem, st>>>(tb, d.n_frames, d.h_cap, s_cap, n_hyp, tmp, keep, \
                                                                            out3d, n_out3d, out2d, n_out2d)
  if (threads <= 128) SES_FP(128, 5);
  else if (threads <= 256) SES_FP(256, 3);
  else SES_FP(512, 1);
#undef SES_FP
  return cudaGetLastError();
}

cudaError_t launch_reproject(const LaunchCfg& cfg, const Tables& tb, int n_frames, int h_max,
                             const ses3d_person_cov* persons3d, const int32_t* n_persons3d, ses3d_person2d* out,
                             int32_t* n_out, cudaStream_t st) {
  int threads, s_cap;
  size_t smem;
  reproject_config(cfg, tb, n_frames, h_max, &threads, &s_cap, &smem);
  if (smem > kSmemBudget) return cudaErrorInvalidConfiguration;
  if (threads <= 128) k_reproject<128, 5><<<n_frames, threads, smem, st>>>(tb, n_frames, h_max, s_cap, persons3d, n_persons3d, out, n_out);
  else if (threads <= 256) k_reproject<256, 3><<<n_frames, threads, smem, st>>>(tb, n_frames, h_max, s_cap, persons3d, n_persons3d, out, n_out);
  else k_reproject<512, 1><<<n_frames, threads, smem, st>>>(tb, n_frames, h_max, s_cap, persons3d, n_persons3d, out, n_out);
  return cudaGetLastError();
}

}  // namespace ses3d
