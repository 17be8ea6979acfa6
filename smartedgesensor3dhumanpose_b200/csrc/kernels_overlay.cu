// kernels_overlay.cu — K9 "overlay" (visualisation): the 2-D skeleton overlay image of
// person_msgs/scripts/pose2D_plot_node.py (draw_humans :18-66, callback_pose :82-91) rasterised on the GPU, one CTA per
// image. The node draws, per person and in this order, a filled circle at every keypoint with score >= 0.25 (:33-47),
// a thick line for every CocoPairs limb whose two joints were drawn, in the colour of the limb's second joint (:50-54),
// and the bounding box grown by 6 px in colours[id % 24] with id = 0 (:57-63), onto a white rgb8 image (:85).
// OpenCV is not available here, so the rasterisation rules are this file's own (integer tests, no anti-aliasing):
//   circle   (x-cx)^2 + (y-cy)^2 <= r^2                       r = max(1, int(w/360)) * 5
//   line     squared distance to the segment <= (t/2)^2         t = max(1, int(w/360)) * 4, round caps
//   box      inside [x1-h, x2+h] x [y1-h, y2+h], outside the box shrunk by h on every side, h = t/2, t = max(1, int(w/360)) * 2
// Later primitives overwrite earlier ones (painter's order = the node's drawing order). Bandwidth-trivial: 0.9 MB per
// 640x480 image written once plus the primitives' bounding boxes.
#include "launch.h"

namespace ses3d {

namespace {

__constant__ unsigned char kCocoColors[24][3] = {
    {255, 0, 0},   {255, 85, 0},  {255, 170, 0}, {255, 255, 0}, {170, 255, 0}, {85, 255, 0},  {0, 255, 0},   {0, 255, 85},
    {0, 255, 170}, {0, 255, 255}, {0, 170, 255}, {0, 85, 255},  {0, 0, 255},   {50, 0, 255},  {100, 0, 255}, {170, 0, 255},
    {255, 0, 255}, {255, 150, 0}, {85, 170, 0},  {42, 128, 85}, {0, 85, 170},  {255, 0, 170}, {255, 0, 85},  {242, 165, 65}};
__constant__ signed char kCocoPairs[16][2] = {{0, 1}, {0, 2}, {1, 3}, {2, 4},   {3, 5},   {4, 6},   {5, 7},   {6, 8},
                                              {7, 9}, {8, 10}, {5, 11}, {6, 12}, {11, 13}, {12, 14}, {13, 15}, {14, 16}};

struct Prim {
  int type;            // 0 circle, 1 line, 2 box
  int ax, ay, bx, by;  // circle: centre in (ax, ay); line: end points; box: corners
  int color;
};

constexpr int kPrimsPerPerson = NKP + 16 + 1;

}  // namespace

__global__ void __launch_bounds__(256)
k_overlay(int p_max, const ses3d_person2d* __restrict__ persons, const int32_t* __restrict__ n_persons, int width,
          int height, unsigned char* __restrict__ rgb) {
  extern __shared__ __align__(16) unsigned char smem_overlay[];
  Prim* prims = reinterpret_cast<Prim*>(smem_overlay);
  __shared__ int n_prims;
  const int img = blockIdx.x;
  unsigned char* out = rgb + (size_t)img * width * height * 3;
  const int np = min(max(n_persons[img], 0), p_max);
  const int scale = max(1, width / 360);
  const int r = scale * 5, t_line = scale * 4, h_box = scale;   // box thickness 2 * scale -> half width scale
  // white background, 32-bit stores (the image start is 4-byte aligned when width * height * 3 is a multiple of 4;
  // the tail and odd sizes fall back to bytes)
  const size_t bytes = (size_t)width * height * 3;
  if ((reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
    uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
    for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) o32[i] = 0xFFFFFFFFu;
    for (size_t i = bytes / 4 * 4 + threadIdx.x; i < bytes; i += blockDim.x) out[i] = 255;
  } else {
    for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = 255;
  }
  // primitive list in drawing order: one thread per person writes that person's fixed block of slots
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    const ses3d_person2d& ps = persons[(size_t)img * p_max + p];
    Prim* q = prims + (size_t)p * kPrimsPerPerson;
    int cx[NKP], cy[NKP];
    bool drawn[NKP];
    for (int k = 0; k < NKP; ++k) {
      const ses3d_keypoint2d& kp = ps.keypoints[k];
      drawn[k] = kp.score >= 0.25f;                          // _CONF_THRESHOLD_DRAW, :19,37
      cx[k] = (int)(kp.x + 0.5f); cy[k] = (int)(kp.y + 0.5f);  // :42
      q[k] = Prim{drawn[k] ? 0 : -1, cx[k], cy[k], 0, 0, k};
    }
    for (int e = 0; e < 16; ++e) {
      const int a = kCocoPairs[e][0], b = kCocoPairs[e][1];
      q[NKP + e] = Prim{(drawn[a] && drawn[b]) ? 1 : -1, cx[a], cy[a], cx[b], cy[b], b};   // colors[pair[1]], :54
    }
    q[NKP + 16] = Prim{2, (int)(ps.bbox[0] + 0.5f) - 6, (int)(ps.bbox[1] + 0.5f) - 6, (int)(ps.bbox[2] + 0.5f) + 6,
                       (int)(ps.bbox[3] + 0.5f) + 6, 0};                                  // :57-63, id = 0 (:84)
  }
  if (threadIdx.x == 0) n_prims = np * kPrimsPerPerson;
  __syncthreads();
  for (int i = 0; i < n_prims; ++i) {
    const Prim pr = prims[i];
    if (pr.type < 0) continue;   // uniform across the CTA
    int x0, y0, x1, y1;
    if (pr.type == 0) { x0 = pr.ax - r; x1 = pr.ax + r; y0 = pr.ay - r; y1 = pr.ay + r; }
    else if (pr.type == 1) {
      const int h = (t_line + 1) / 2;
      x0 = min(pr.ax, pr.bx) - h; x1 = max(pr.ax, pr.bx) + h; y0 = min(pr.ay, pr.by) - h; y1 = max(pr.ay, pr.by) + h;
    } else { x0 = pr.ax - h_box; x1 = pr.bx + h_box; y0 = pr.ay - h_box; y1 = pr.by + h_box; }
    x0 = max(x0, 0); y0 = max(y0, 0); x1 = min(x1, width - 1); y1 = min(y1, height - 1);
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    if (bw > 0 && bh > 0) {
      const unsigned char cr = kCocoColors[pr.color][0], cg = kCocoColors[pr.color][1], cb = kCocoColors[pr.color][2];
      const long long n_px = (long long)bw * bh;
      for (long long e = threadIdx.x; e < n_px; e += blockDim.x) {
        const int x = x0 + (int)(e % bw), y = y0 + (int)(e / bw);
        bool hit;
        if (pr.type == 0) {
          const long long dx = x - pr.ax, dy = y - pr.ay;
          hit = dx * dx + dy * dy <= (long long)r * r;
        } else if (pr.type == 1) {
          const long long dx = pr.bx - pr.ax, dy = pr.by - pr.ay, px = x - pr.ax, py = y - pr.ay;
          const long long L2 = dx * dx + dy * dy, u = px * dx + py * dy;
          // 4 * dist^2 <= t^2, all in integers
          if (L2 == 0 || u <= 0) hit = 4 * (px * px + py * py) <= (long long)t_line * t_line;
          else if (u >= L2) {
            const long long qx = x - pr.bx, qy = y - pr.by;
            hit = 4 * (qx * qx + qy * qy) <= (long long)t_line * t_line;
          } else {
            hit = 4 * ((px * px + py * py) * L2 - u * u) <= (long long)t_line * t_line * L2;
          }
        } else {
          const bool inner = x >= pr.ax + h_box && x <= pr.bx - h_box && y >= pr.ay + h_box && y <= pr.by - h_box;
          hit = !inner;   // the clipped bounding box is the outer rectangle
        }
        if (hit) {
          unsigned char* px3 = out + ((size_t)y * width + x) * 3;
          px3[0] = cr; px3[1] = cg; px3[2] = cb;
        }
      }
    }
    __syncthreads();   // painter's order: the next primitive may overwrite these pixels
  }
}

cudaError_t launch_overlay(int n_images, int p_max, const ses3d_person2d* persons, const int32_t* n_persons, int width,
                           int height, unsigned char* rgb, cudaStream_t st) {
  if (n_images == 0) return cudaSuccess;
  const size_t smem = (size_t)p_max * kPrimsPerPerson * sizeof(Prim);
  if (smem > 200 * 1024) return cudaErrorInvalidConfiguration;
  static bool attr_set = false;   // raising the limit is idempotent; racing threads set the same value
  if (!attr_set && smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_overlay, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  k_overlay<<<n_images, 256, smem, st>>>(p_max, persons, n_persons, width, height, rgb);
  return cudaGetLastError();
}

}  // namespace ses3d
