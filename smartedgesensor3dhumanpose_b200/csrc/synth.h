// synth.h — counter-based synthetic frame generator, host + device.
//
// Not part of the reference: it replaces the demo bag (README.md:41-47, an external download)
// as the input source for tests and benchmarks. Only IEEE add/mul/div/sqrt and integer
// arithmetic are used (no transcendental functions), and the file must be compiled without
// FMA contraction (-ffp-contract=off on the host, -fmad=false on the device), so the host and
// device generators produce bit-identical frames.
#pragma once
#include <math.h>
#include <stdint.h>

#include "ses3d.h"

#if defined(__CUDACC__)
#define SES_HD __host__ __device__ __forceinline__
#else
#define SES_HD inline
#endif

namespace ses3d_synth {

struct U4 { uint32_t v[4]; };

SES_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }

// Philox4x32-10 (Salmon et al., SC'11)
SES_HD U4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  U4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

SES_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }  // [0,1), exact

// Irwin-Hall(12) on 16-bit halves: exact integer sum, sigma = 1 (approximately Gaussian, |g| <= 6)
SES_HD float gauss12(const uint32_t w[6]) {
  int32_t s = 0;
  for (int i = 0; i < 6; ++i) s += (int32_t)(w[i] & 0xFFFFu) + (int32_t)(w[i] >> 16);
  return (float)(s - 393210) * (1.0f / 65536.0f);
}

// Upright standing COCO-17 figure, local frame x forward / y left / z up, metres. Bone lengths
// follow EdgeTPU_BodyParts_Simple::limbLength (S3D:101): shoulder-elbow 0.28, elbow-wrist 0.25,
// shoulder-hip 0.50, hip-knee 0.45, knee-ankle 0.446, nose-eye 0.05, eye-ear 0.10, shoulders 0.35 apart.
SES_HD void template_joint(int k, double out[3]) {
  const double T[17][3] = {
      {0.100, 0.000, 1.6400},                                  // 0 nose
      {0.075, 0.030, 1.671225},  {0.075, -0.030, 1.671225},    // 1 LEye 2 REye
      {-0.005, 0.085, 1.651225}, {-0.005, -0.085, 1.651225},   // 3 LEar 4 REar
      {0.000, 0.175, 1.4744},    {0.000, -0.175, 1.4744},      // 5 LShoulder 6 RShoulder
      {0.000, 0.225, 1.1989},    {0.000, -0.225, 1.1989},      // 7 LElbow 8 RElbow
      {0.060, 0.225, 0.9562},    {0.060, -0.225, 0.9562},      // 9 LWrist 10 RWrist
      {0.000, 0.135, 0.9760},    {0.000, -0.135, 0.9760},      // 11 LHip 12 RHip
      {0.000, 0.135, 0.5260},    {0.000, -0.135, 0.5260},      // 13 LKnee 14 RKnee
      {0.000, 0.135, 0.0800},    {0.000, -0.135, 0.0800}};     // 15 LAnkle 16 RAnkle
  out[0] = T[k][0]; out[1] = T[k][1]; out[2] = T[k][2];
}

struct Scene {          // per-frame people: root (x, y) and yaw (cos, sin)
  float x[64], y[64], c[64], s[64];
};
#define SES3D_SYNTH_MAX_PEOPLE 64

// Place n_people roots with >= min_separation by rejection sampling (at most 64 attempts each).
SES_HD void make_scene(const ses3d_synth_config& cfg, uint64_t frame_in, Scene& sc) {
  const uint32_t k0 = (uint32_t)cfg.seed, k1 = (uint32_t)(cfg.seed >> 32);
  // sequence mode: the scene is drawn once per sequence and then moves (people walk along their heading)
  const uint64_t T = cfg.frames_per_sequence > 0 ? (uint64_t)cfg.frames_per_sequence : 0;
  const uint64_t frame = T ? frame_in / T : frame_in;
  const float walked = T ? cfg.step_m * (float)(frame_in % T) : 0.0f;
  const uint32_t f0 = (uint32_t)frame, f1 = (uint32_t)(frame >> 32);
  const float sep2 = cfg.min_separation * cfg.min_separation;
  for (int p = 0; p < cfg.n_people; ++p) {
    float px = 0.f, py = 0.f, pc = 1.f, ps = 0.f;
    for (int attempt = 0; attempt < 64; ++attempt) {
      const U4 r = philox(f0, f1, (uint32_t)p, 0x100u + (uint32_t)attempt, k0, k1);
      px = cfg.area[0] + u01(r.v[0]) * (cfg.area[2] - cfg.area[0]);
      py = cfg.area[1] + u01(r.v[1]) * (cfg.area[3] - cfg.area[1]);
      const float a = 2.0f * u01(r.v[2]) - 1.0f, b = 2.0f * u01(r.v[3]) - 1.0f;
      const float r2 = a * a + b * b;
      bool ok = (r2 >= 0.01f) && (r2 <= 1.0f);
      if (ok) { const float rr = sqrtf(r2); pc = a / rr; ps = b / rr; }
      for (int q = 0; q < p && ok; ++q) {
        const float dx = px - sc.x[q], dy = py - sc.y[q];
        if (dx * dx + dy * dy < sep2) ok = false;
      }
      if (ok) break;
    }
    sc.x[p] = px; sc.y[p] = py; sc.c[p] = pc; sc.s[p] = ps;
  }
  if (T)
    for (int p = 0; p < cfg.n_people; ++p) { sc.x[p] += sc.c[p] * walked; sc.y[p] += sc.s[p] * walked; }
}

SES_HD void world_joint(const Scene& sc, int p, int k, double X[3]) {
  double l[3];
  template_joint(k, l);
  const double c = (double)sc.c[p], s = (double)sc.s[p];
  X[0] = (double)sc.x[p] + (c * l[0] - s * l[1]);
  X[1] = (double)sc.y[p] + (s * l[0] + c * l[1]);
  X[2] = l[2];
}

// Project into camera; returns true when in front (Z > 0.5 m) and inside the image.
SES_HD bool project(const ses3d_camera& cam, const double X[3], double& u, double& v) {
  const double* T = cam.T_cam_base;
  const double xc = T[0] * X[0] + T[1] * X[1] + T[2] * X[2] + T[3];
  const double yc = T[4] * X[0] + T[5] * X[1] + T[6] * X[2] + T[7];
  const double zc = T[8] * X[0] + T[9] * X[1] + T[10] * X[2] + T[11];
  if (!(zc > 0.5)) { u = 0; v = 0; return false; }
  u = (cam.fx * xc + cam.Tx) / zc + cam.cx;
  v = (cam.fy * yc + cam.Ty) / zc + cam.cy;
  return u >= 0.0 && u < (double)cam.width && v >= 0.0 && v < (double)cam.height;
}

// All detections of one (frame, camera): persons [p_max], gt [p_max] (nullable). Returns the count.
SES_HD int make_camera_view(const ses3d_synth_config& cfg, const ses3d_camera& cam, int cam_idx, uint64_t frame,
                            const Scene& sc, ses3d_person2d* persons, int32_t* gt) {
  const uint32_t k0 = (uint32_t)cfg.seed, k1 = (uint32_t)(cfg.seed >> 32);
  const uint32_t f0 = (uint32_t)frame, f1 = (uint32_t)(frame >> 32);
  int order[SES3D_SYNTH_MAX_PEOPLE];
  int n = 0;
  for (int p = 0; p < cfg.n_people; ++p) {  // who is seen by this camera
    int n_vis = 0;
    for (int k = 0; k < 17; ++k) {
      double X[3], u, v;
      world_joint(sc, p, k, X);
      if (project(cam, X, u, v)) ++n_vis;
    }
    if (n_vis >= cfg.min_visible && n < cfg.p_max) order[n++] = p;
  }
  for (int i = 0; i + 1 < n; ++i) {  // Fisher-Yates shuffle of the per-camera person order
    const U4 r = philox(f0, f1, (uint32_t)cam_idx, 0x200u + (uint32_t)i, k0, k1);
    const int j = i + (int)(r.v[0] % (uint32_t)(n - i));
    const int t = order[i]; order[i] = order[j]; order[j] = t;
  }
  const float var = cfg.noise_px * cfg.noise_px;
  for (int slot = 0; slot < n; ++slot) {
    const int p = order[slot];
    ses3d_person2d& out = persons[slot];
    float score_sum = 0.f;
    float bx0 = 0.f, by0 = 0.f, bx1 = 0.f, by1 = 0.f;
    bool any = false;
    const uint32_t ent = (uint32_t)cam_idx * 4096u + (uint32_t)p;
    for (int k = 0; k < 17; ++k) {
      double X[3], u, v;
      world_joint(sc, p, k, X);
      const bool vis = project(cam, X, u, v);
      const U4 ra = philox(f0, f1, ent, 0x1000u + (uint32_t)k * 4u + 0u, k0, k1);
      const U4 rb = philox(f0, f1, ent, 0x1000u + (uint32_t)k * 4u + 1u, k0, k1);
      const U4 rc = philox(f0, f1, ent, 0x1000u + (uint32_t)k * 4u + 2u, k0, k1);
      const uint32_t wx[6] = {rb.v[0], rb.v[1], rb.v[2], rb.v[3], rc.v[0], rc.v[1]};
      const U4 rd = philox(f0, f1, ent, 0x1000u + (uint32_t)k * 4u + 3u, k0, k1);
      const uint32_t wy[6] = {rc.v[2], rc.v[3], rd.v[0], rd.v[1], rd.v[2], rd.v[3]};
      ses3d_keypoint2d& kp = out.keypoints[k];
      if (vis && !(u01(ra.v[0]) < cfg.dropout)) {
        kp.x = (float)u + cfg.noise_px * gauss12(wx);
        kp.y = (float)v + cfg.noise_px * gauss12(wy);
        kp.score = 0.5f + 0.5f * u01(ra.v[1]);
        if (!any) { bx0 = bx1 = kp.x; by0 = by1 = kp.y; any = true; }
        else {
          bx0 = kp.x < bx0 ? kp.x : bx0; bx1 = kp.x > bx1 ? kp.x : bx1;
          by0 = kp.y < by0 ? kp.y : by0; by1 = kp.y > by1 ? kp.y : by1;
        }
      } else {  // dropout / not visible: low score, random pixel
        kp.x = u01(ra.v[2]) * (float)cam.width;
        kp.y = u01(ra.v[3]) * (float)cam.height;
        kp.score = 0.29f * u01(ra.v[1]);
      }
      kp.cov[0] = var; kp.cov[1] = 0.1f * var; kp.cov[2] = var;
      score_sum += kp.score;
    }
    out.score = score_sum / 17.0f;
    out.bbox[0] = bx0; out.bbox[1] = by0; out.bbox[2] = bx1; out.bbox[3] = by1;
    if (gt) gt[slot] = p;
  }
  return n;
}

}  // namespace ses3d_synth
