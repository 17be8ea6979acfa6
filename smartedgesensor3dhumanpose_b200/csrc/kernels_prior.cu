// kernels_prior.cu — K7 "prior" for sm_100a: one CTA per message stream, its frames walked in order; inside a
// frame one warp per detection (skeleton fit) and warp 0 for the assignment. The algorithm lives in prior_core.h.
// FP64 throughout (gtsam computes in double); tolerance-checked against the oracle, so FMA contraction stays on.
#include <algorithm>
#include <cstdlib>

#include "launch.h"
#include "prior_core.h"

namespace ses3d {

extern __shared__ __align__(16) unsigned char smem_raw[];

__constant__ PriorStatic c_prior_static = SES3D_PRIOR_STATIC_INIT;
constexpr size_t kStaticBytes = (sizeof(PriorStatic) + 15) / 16 * 16;

__global__ void __launch_bounds__(256)
k_prior(const PriorTables pt_in, int n_seq, int n_frames, int h_max, int max_tracks, int group, size_t ws_bytes,
        size_t fit_bytes,
        PriorSeqState* __restrict__ states, PriorTrack* __restrict__ tracks, uint8_t* __restrict__ order,
        const ses3d_person_cov* __restrict__ persons, const int32_t* __restrict__ n_persons,
        const int64_t* __restrict__ stamp_ns, int n_cams, const float* __restrict__ fb_delay,
        ses3d_person_cov* __restrict__ fused, ses3d_person_cov* __restrict__ pred, int32_t* __restrict__ n_out,
        float* __restrict__ pred_delay, int32_t* __restrict__ track_of) {
  const int s = blockIdx.x;
  if (s >= n_seq) return;
  {  // skeleton tables: constant memory -> shared memory (lanes index them with different joints)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&c_prior_static);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem_raw);
    for (int i = (int)threadIdx.x; i < (int)(sizeof(PriorStatic) / 4); i += (int)blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  PriorTables pt = pt_in;
  pt.st = reinterpret_cast<const PriorStatic*>(smem_raw);
  unsigned char* fit_ws = smem_raw + kStaticBytes + ws_bytes;   // fit workspaces, shared with the tracker's big arrays
  Arena ar(smem_raw + kStaticBytes), tr(fit_ws);
  PriorWs ws;
  prior_ws_layout(ar, tr, h_max, max_tracks, &ws);
  BlockTeam tm;
  PriorSeqState* st = states + s;
  PriorTrack* trk = tracks + (size_t)s * max_tracks;
  uint8_t* ord = order + (size_t)s * max_tracks;
  for (int f = 0; f < n_frames; ++f) {
    const size_t i = (size_t)s * n_frames + f;
    prior_frame(tm, pt, max_tracks, h_max, group, st, trk, ord, ws, fit_ws, fit_bytes, stamp_ns[i], n_cams,
                fb_delay ? fb_delay + i * n_cams : nullptr, n_persons[i], persons + i * h_max, fused + i * h_max,
                pred + i * h_max, n_out + i, pred_delay ? pred_delay + i : nullptr,
                track_of ? track_of + i * h_max : nullptr);
    tm.sync();
  }
}

__global__ void k_prior_reset(const ses3d_prior_params prm, int n_seq, PriorSeqState* states, int keep_t_prev) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seq) prior_state_reset(prm, states + s, keep_t_prev != 0);
}

cudaError_t launch_prior_reset(const ses3d_prior_params& prm, int n_seq, PriorSeqState* states, bool keep_t_prev,
                               cudaStream_t st) {
  k_prior_reset<<<(n_seq + 127) / 128, 128, 0, st>>>(prm, n_seq, states, keep_t_prev ? 1 : 0);
  return cudaGetLastError();
}

// tuning overrides, read once per handle creation (ses3d_prior_create -> init_prior_kernels), never per launch
static int g_prior_group_env = 0, g_prior_warps_env = 0;
cudaError_t init_prior_kernels(int) {
  const char* eg = getenv("SES3D_PRIOR_GROUP");
  const char* ew = getenv("SES3D_PRIOR_WARPS");
  g_prior_group_env = eg ? atoi(eg) : 0;   // re-read at every create: an override does not outlive its variable
  g_prior_warps_env = ew ? atoi(ew) : 0;
  return cudaFuncSetAttribute(k_prior, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t launch_prior(const PriorTables& pt, int n_seq, int n_frames, int h_max, int max_tracks,
                         PriorSeqState* states, PriorTrack* tracks, uint8_t* order, const ses3d_person_cov* persons,
                         const int32_t* n_persons, const int64_t* stamp_ns, int n_cams, const float* fb_delay,
                         ses3d_person_cov* fused, ses3d_person_cov* pred, int32_t* n_out, float* pred_delay,
                         int32_t* track_of, cudaStream_t st) {
  // One warp fits a group of up to `group` detections together (prior_core.h); the groups of a message are spread
  // over the CTA's warps. Measured on B200 (2048 streams x 32 messages x 6 people, h_max 8, ms per launch):
  // group/warps 3/1 23.6, 3/2 19.0, 3/3 22.7, 2/2 22.7, 2/3 19.6, 2/4 22.5, 4/2 24.5 - just enough warps to fit a
  // typical message in one round, and as little shared memory per CTA as possible (occupancy decides).
  int group = std::max(1, std::min(3, h_max));
  if (g_prior_group_env > 0) group = std::max(1, std::min(PRIOR_GMAX, g_prior_group_env));
  // two warps per stream whatever h_max is: a fuller message simply takes more rounds of groups, and the smaller CTA
  // keeps eight streams resident per SM (demo chain, h_max 16, 2048 streams: 40.8 ms with 4 warps -> 36.3 ms with 2)
  int warps = std::max(1, std::min(2, (h_max + group - 1) / group));
  // Few streams (a ROS node tracks ONE; anything that leaves the GPU under-filled): occupancy is irrelevant, the
  // time of a message is what counts - one detection per warp, up to six warps. B200, 32 messages x 6 people per
  // stream, ms per launch (3 x 2 warps -> 1 x 6 warps): 1 stream 3.94 -> 3.30, 64: 4.18 -> 3.58, 256: 4.75 -> 4.19,
  // 512: 5.14 -> 7.58 (so the switch sits at 256 streams).
  if (n_seq <= 256 && g_prior_group_env <= 0 && g_prior_warps_env <= 0) {
    group = 1;
    warps = std::max(1, std::min(6, h_max));
  }
  if (g_prior_warps_env > 0) warps = std::max(1, std::min(8, g_prior_warps_env));
  size_t ws_bytes = 0, transient_bytes = 0;
  prior_ws_bytes(h_max, max_tracks, &ws_bytes, &transient_bytes);
  const size_t fit_bytes = prior_fit_ws_bytes(group);
  static_assert(sizeof(PriorStatic) % 4 == 0, "copied word by word");
  const size_t smem = kStaticBytes + ws_bytes + std::max(fit_bytes * warps, transient_bytes);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  k_prior<<<n_seq, 32 * warps, smem, st>>>(pt, n_seq, n_frames, h_max, max_tracks, group, ws_bytes, fit_bytes, states, tracks,
                                            order, persons, n_persons, stamp_ns, n_cams, fb_delay, fused, pred, n_out,
                                            pred_delay, track_of);
  return cudaGetLastError();
}

}  // namespace ses3d
