// kernels_markers.cu — K8 "markers" (visualisation): one warp per (frame, person); lanes 0..20 compute the
// covariance ellipsoids of the 21 fusion slots, lane 31 walks the skeleton's LINE_LIST segments. The algorithms live
// in markers_core.h.
#include "launch.h"
#include "markers_core.h"

namespace ses3d {

__global__ void __launch_bounds__(128)
k_markers(const SkeletonModel model, int n_units, int h_max, int style, const ses3d_person_cov* __restrict__ persons3d,
          const int32_t* __restrict__ n_persons3d, ses3d_ellipsoid* __restrict__ ell, double* __restrict__ seg,
          int32_t* __restrict__ n_seg, int8_t* __restrict__ seg_slot) {
  const int unit = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);   // frame * h_max + person
  const int lane = (int)(threadIdx.x & 31u);
  if (unit >= n_units) return;
  const int f = unit / h_max, p = unit % h_max;
  const bool live = p < n_persons3d[f];
  const ses3d_person_cov& person = persons3d[unit];
  if (ell && lane < NFUS) {
    ses3d_ellipsoid e = {0, 0, 0, 0, 0, 0, 0};
    if (live && person.keypoints[lane].score > 0.0f) covariance_ellipsoid(person.keypoints[lane].cov, &e);
    ell[(size_t)unit * NFUS + lane] = e;
  }
  if (seg && lane == 31) {
    int n = 0;
    if (live)
      n = skeleton_segments(model, style, person, seg + (size_t)unit * MARKER_MAX_SEGMENTS * 6,
                            seg_slot ? seg_slot + (size_t)unit * MARKER_MAX_SEGMENTS : nullptr);
    n_seg[unit] = n;
  }
}

cudaError_t launch_markers(const SkeletonModel& model, int n_frames, int h_max, int style,
                           const ses3d_person_cov* persons3d, const int32_t* n_persons3d, ses3d_ellipsoid* ell,
                           double* seg, int32_t* n_seg, int8_t* seg_slot, cudaStream_t st) {
  const long long n_units = (long long)n_frames * h_max;
  if (n_units == 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((n_units * 32 + 127) / 128);
  k_markers<<<blocks, 128, 0, st>>>(model, (int)n_units, h_max, style, persons3d, n_persons3d, ell, seg, n_seg, seg_slot);
  return cudaGetLastError();
}

}  // namespace ses3d
