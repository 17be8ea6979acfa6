// launch.h — host-side launchers of the four kernels (implemented in kernels_*.cu).
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace ses3d {

constexpr int kTriBuckets = 8;   // == TRI_BUCKETS (tri_core.h): K3 work sub-lists by observation count

struct LaunchDims {
  int n_frames, p_max, h_cap;
};

// Per-handle launch configuration, fixed at ses3d_create: the device's SM count, the tuning overrides from the
// environment (SES3D_*: read once, never on the per-batch path) and a small cache of occupancy queries. The kernels'
// dynamic shared-memory ceilings are raised once per device by init_kernels().
struct LaunchCfg {
  int device = 0;
  int n_sm = 148;
  int assoc_threads = 0;       // 0 = automatic
  int reproj_scap = 6;         // persons per batch of the reprojection kernel
  int reproj_threads = 128;    // 0 = one warp per camera (at most 16 warps)
  int tri_warps = 4, tri_warps_f64 = 4;
  int rounds_warps = 4;        // frames (warps) per CTA of the camera-rounds kernel
  int tri_dynamic = 1;         // K3 hands work items out one by one (0: fixed strides)
  int pairs_tiled = 0;         // K2a: dense frames build their pair table by camera-pair tiles (second k_pairs instance);
                               // measured SLOWER than the flat pass on every rig (crowd 49 vs 30 ms), kept for A/B only
  int latency_frames = 0;      // launches of at most this many frames use the low-latency shapes (default: SM count)
  int pairs_split = 0;         // K2a: CTAs per frame (0 = automatic: > 1 only for big rigs with few frames per launch)
  int rounds_block = -1;       // K2b: CTA per frame instead of warp per frame (-1 = automatic: heavy, few frames)
  int tri_lockstep = 1;        // K3: the warps of a CTA pass the phases of their hypotheses together (I-cache sharing)
  struct OccEntry { const void* fn; size_t smem; int per_sm; };
  OccEntry occ[8] = {};
  int n_occ = 0;
};
cudaError_t init_kernels(LaunchCfg* cfg, int device);
cudaError_t init_prior_kernels(int device);

// K2 = K2a (pair table, one CTA per frame) + K2b (camera rounds, one warp per frame). nk_scratch != nullptr selects
// the global-memory keypoint path; meta = n_frames x associate_meta_bytes() hand-over scratch.
cudaError_t launch_associate(const LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                             const int32_t* n_persons, float* nk_scratch, double* pair_table, unsigned char* meta,
                             int8_t* hyp_det, int32_t* n_hyp, int32_t* n_hung, int32_t* overflow, int32_t* hyp_of_dump,
                             int32_t* keep, uint32_t* work, int32_t* work_count, cudaStream_t st);
size_t associate_smem_bytes(int n_cams, int p_max, int h_cap, bool* needs_scratch);
int associate_launches(const LaunchCfg& cfg, int p_max);    // kernels launch_associate enqueues (2 or 3)
size_t associate_pair_table_bytes(int n_cams, int p_max);   // per frame
size_t associate_meta_bytes(int n_cams, int p_max);         // per frame

// K3: one warp per (frame, hypothesis) work item, persistent grid over the work lists K2 wrote: kTriBuckets sub-lists
// of n_frames * h_cap slots each (by observation count), work_count[b] = items in sub-list b, work_count[kTriBuckets]
// = claim counter (all zeroed by launch_associate). far_scratch: global workspace of
// triangulate_far_scratch_bytes() for the exact covariance of far joints (nullptr / too small: approximate path).
cudaError_t launch_triangulate(LaunchCfg& cfg, const Tables& tb, LaunchDims d, const ses3d_person2d* persons,
                               const int8_t* hyp_det, const uint32_t* work, int32_t* work_count, ses3d_person_cov* tmp,
                               int32_t* keep, float* far_scratch, size_t far_scratch_bytes, cudaStream_t st);
size_t triangulate_far_scratch_bytes(const LaunchCfg& cfg);

// K4: one CTA per frame
cudaError_t launch_finalize(const Tables& tb, LaunchDims d, const int32_t* n_hyp, ses3d_person_cov* tmp,
                            const int32_t* keep, ses3d_person_cov* out, int32_t* n_out, cudaStream_t st);

// K4 + K6 fused for the process calls: one CTA per frame
cudaError_t launch_finproj(const LaunchCfg& cfg, const Tables& tb, LaunchDims d, const int32_t* n_hyp,
                           ses3d_person_cov* tmp, const int32_t* keep, ses3d_person_cov* out3d, int32_t* n_out3d,
                           ses3d_person2d* out2d, int32_t* n_out2d, cudaStream_t st);

// K6: one CTA per frame
cudaError_t launch_reproject(const LaunchCfg& cfg, const Tables& tb, int n_frames, int h_max, const ses3d_person_cov* persons3d,
                             const int32_t* n_persons3d, ses3d_person2d* out, int32_t* n_out, cudaStream_t st);

// diagnostics: a batch of independent assignment problems through the warp-cooperative Munkres
cudaError_t launch_munkres_batch(int n, int rows, int cols, const double* cost, int32_t* assignment, cudaStream_t st);

// ragged <-> padded record movement (kernels_pack.cu)
cudaError_t launch_scan_counts(const int32_t* counts, int n, int cap, long long* offsets, long long* running,
                               cudaStream_t st);
cudaError_t launch_move_records(int direction /*0 pack, 1 unpack*/, int n_units, int cap, int rec_bytes,
                                const int32_t* counts, const long long* offsets, void* strided, void* dense,
                                long long dense_limit, cudaStream_t st);

}  // namespace ses3d

// K7 (kernels_prior.cu): pose_prior, one CTA per message stream
#include "prior_core.h"
namespace ses3d {
cudaError_t launch_prior_reset(const ses3d_prior_params& prm, int n_seq, PriorSeqState* states, bool keep_t_prev,
                               cudaStream_t st);
cudaError_t launch_prior(const PriorTables& pt, int n_seq, int n_frames, int h_max, int max_tracks,
                         PriorSeqState* states, PriorTrack* tracks, uint8_t* order, const ses3d_person_cov* persons,
                         const int32_t* n_persons, const int64_t* stamp_ns, int n_cams, const float* fb_delay,
                         ses3d_person_cov* fused, ses3d_person_cov* pred, int32_t* n_out, float* pred_delay,
                         int32_t* track_of, cudaStream_t st);
}  // namespace ses3d

// K8 (kernels_markers.cu): visualisation content, one warp per (frame, person)
namespace ses3d {
cudaError_t launch_markers(const SkeletonModel& model, int n_frames, int h_max, int style,
                           const ses3d_person_cov* persons3d, const int32_t* n_persons3d, ses3d_ellipsoid* ell,
                           double* seg, int32_t* n_seg, int8_t* seg_slot, cudaStream_t st);
}  // namespace ses3d

// K9 (kernels_overlay.cu): 2-D skeleton overlay images, one CTA per image
namespace ses3d {
cudaError_t launch_overlay(int n_images, int p_max, const ses3d_person2d* persons, const int32_t* n_persons, int width,
                           int height, unsigned char* rgb, cudaStream_t st);
}  // namespace ses3d
