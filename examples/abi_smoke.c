/* abi_smoke.c — the C ABI used from plain C (C99): what a non-Python, non-ROS client links against.
 *   gcc -std=c99 -Iinclude examples/abi_smoke.c -Lsmartedgesensor3dhumanpose_b200 -lses3d -o abi_smoke
 * Exercises the host-only entry points (no GPU needed) and, when a device is present, one single-frame
 * triangulate + reproject call on a 4-camera ring (the n_frames = 1 ROS-shim case). Exit code 0 = ok. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ses3d.h"

static void look_at(double ex, double ey, double ez, double T[12]) {
  /* camera at (ex,ey,ez) looking at (0,0,1), z forward, x right, y down */
  double z[3] = {-ex, -ey, 1.0 - ez}, n = sqrt(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]);
  double x[3], y[3];
  int i;
  for (i = 0; i < 3; ++i) z[i] /= n;
  x[0] = z[1] * 1.0 - z[2] * 0.0; x[1] = z[2] * 0.0 - z[0] * 1.0; x[2] = 0.0;   /* z cross up(0,0,1) */
  n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (i = 0; i < 3; ++i) x[i] /= n;
  y[0] = z[1] * x[2] - z[2] * x[1]; y[1] = z[2] * x[0] - z[0] * x[2]; y[2] = z[0] * x[1] - z[1] * x[0];
  for (i = 0; i < 3; ++i) { T[i] = x[i]; T[4 + i] = y[i]; T[8 + i] = z[i]; }
  T[3] = -(x[0] * ex + x[1] * ey + x[2] * ez);
  T[7] = -(y[0] * ex + y[1] * ey + y[2] * ez);
  T[11] = -(z[0] * ex + z[1] * ey + z[2] * ez);
}

int main(void) {
  ses3d_params prm;
  ses3d_camera cams[4];
  ses3d_synth_config cfg;
  ses3d_person2d persons[4 * 2];
  int32_t n_persons[4];
  ses3d_assembler_config acfg;
  ses3d_assembler asmb = NULL;
  ses3d_handle h = NULL;
  uint8_t wire[4096];
  size_t n;
  int c, rc;

  ses3d_default_params(&prm);
  if (prm.min_num_valid_keypoints != 9 || prm.max_epipolar_error != 0.050) return 1;
  printf("%s\n", ses3d_version());

  for (c = 0; c < 4; ++c) {
    const double a = 6.283185307179586 * c / 4;
    memset(&cams[c], 0, sizeof(cams[c]));
    look_at(5.0 * cos(a), 5.0 * sin(a), 2.5, cams[c].T_cam_base);
    cams[c].fx = cams[c].fy = 1000.0; cams[c].cx = 640.0; cams[c].cy = 360.0;
    cams[c].width = 1280; cams[c].height = 720;
  }
  memset(&cfg, 0, sizeof(cfg));
  cfg.seed = 1; cfg.n_people = 1; cfg.p_max = 2; cfg.noise_px = 2.0f; cfg.min_separation = 0.6f; cfg.min_visible = 5;
  cfg.area[0] = -1.0f; cfg.area[1] = -1.0f; cfg.area[2] = 1.0f; cfg.area[3] = 1.0f;
  if (ses3d_synth_frames(4, cams, &cfg, 0, 1, persons, n_persons, NULL, NULL) != SES3D_OK) return 2;
  for (c = 0; c < 4; ++c) if (n_persons[c] != 1) return 3;

  /* wire format round trip */
  n = ses3d_wire_encode_person2dlist(7, 1234567890123LL, "cam_1_color_optical_frame", 0.1f, persons, n_persons[0], wire, sizeof(wire));
  if (n != 16 + 25 + 8 + 432) return 4;
  {
    ses3d_person2d back[2];
    int64_t stamp = 0;
    if (ses3d_wire_decode_person2dlist(wire, n, NULL, &stamp, NULL, 0, NULL, back, 2) != 1) return 5;
    if (stamp != 1234567890123LL || memcmp(&back[0], &persons[0], sizeof(back[0])) != 0) return 6;
  }

  /* frame assembler: four synchronous cameras -> one frame per tick */
  if (ses3d_assembler_default_config(4, &acfg) != SES3D_OK || ses3d_assembler_create(&acfg, &asmb) != SES3D_OK) return 7;
  {
    int t, frames = 0;
    for (t = 0; t < 10; ++t)
      for (c = 0; c < 4; ++c) {
        rc = ses3d_assembler_add(asmb, c, 1000000000LL + t * 40000000LL, t * 4 + c);
        if (rc < 0) return 8;
        frames += rc;
      }
    if (frames < 8) return 9;
  }
  ses3d_assembler_destroy(asmb);

  /* latest-wins mailbox replay: a 100 ms worker on a 25 Hz stream processes every third frame and always the newest */
  {
    int64_t t_ready[10], busy[10];
    uint8_t taken[10];
    int i, n_taken;
    for (i = 0; i < 10; ++i) { t_ready[i] = (int64_t)i * 40000000LL; busy[i] = 100000000LL; }
    n_taken = ses3d_mailbox_replay(10, t_ready, busy, taken, NULL);
    if (n_taken < 3 || n_taken > 5 || !taken[0] || !taken[9] || taken[1]) return 18;
  }

  /* the compute path needs a GPU: without one the library must say so instead of computing on the CPU */
  rc = ses3d_create(4, cams, &prm, 0, &h);
  if (rc != SES3D_OK) {
    printf("no device: %s\n", ses3d_last_error_string());
    return rc == SES3D_E_CUDA ? 0 : 10;
  }
  {
    ses3d_person_cov out3d[8];
    ses3d_person2d out2d[4 * 8];
    int32_t n3 = 0, n2[4];
    rc = ses3d_process_batch(h, 1, 2, persons, n_persons, 8, out3d, &n3, out2d, n2, NULL, SES3D_HOST_BUFFERS, NULL);
    if (rc != SES3D_OK) { printf("%s\n", ses3d_last_error_string()); return 11; }
    printf("persons3d = %d, nose = (%.3f, %.3f, %.3f), reprojected persons per camera = %d %d %d %d\n", n3,
           out3d[0].keypoints[SES3D_FBP_NOSE].x, out3d[0].keypoints[SES3D_FBP_NOSE].y, out3d[0].keypoints[SES3D_FBP_NOSE].z,
           n2[0], n2[1], n2[2], n2[3]);
    if (n3 != 1 || n2[0] != 1) return 12;

    /* the per-message call of a node: the second call of a shape captured a CUDA graph, later calls replay it -
     * same record every time */
    {
      int rep;
      for (rep = 0; rep < 5; ++rep) {
        ses3d_person_cov r3d[8];
        ses3d_person2d r2d[4 * 8];
        int32_t rn3 = 0, rn2[4];
        rc = ses3d_process_batch(h, 1, 2, persons, n_persons, 8, r3d, &rn3, r2d, rn2, NULL, SES3D_HOST_BUFFERS, NULL);
        if (rc != SES3D_OK || rn3 != n3 || memcmp(r3d, out3d, sizeof r3d) != 0 || memcmp(r2d, out2d, sizeof r2d) != 0 ||
            memcmp(rn2, n2, sizeof rn2) != 0)
          return 23;
      }
      printf("single-frame replays: identical records\n");
    }

    /* the same call through the single-process multi-GPU entry (device list; here: every visible device) */
    {
      ses3d_multi m = NULL;
      ses3d_person_cov m3d[8];
      ses3d_person2d m2d[4 * 8];
      int32_t mn3 = 0, mn2[4];
      if (ses3d_create_multi(4, cams, &prm, 0, NULL, &m) != SES3D_OK) { printf("%s\n", ses3d_last_error_string()); return 19; }
      const int32_t n_dev = ses3d_multi_device_count(m);
      if (n_dev < 1) return 20;
      rc = ses3d_multi_process_batch(m, 1, 2, persons, n_persons, 8, m3d, &mn3, m2d, mn2, NULL);
      if (rc != SES3D_OK || mn3 != n3 || mn2[0] != n2[0] || memcmp(&m3d[0], &out3d[0], sizeof(m3d[0])) != 0) return 21;
      ses3d_multi_destroy(m);
      printf("multi: %d device(s), identical record\n", n_dev);
    }
    /* 2-D overlay of the reprojected skeletons of camera 0 (pose2D_plot_node.py) on a canvas of the camera's size */
    {
      static uint8_t rgb[720 * 1280 * 3];
      int32_t n_img = n2[0];
      int i, coloured = 0;
      rc = ses3d_overlay_batch(h, 1, 8, out2d, &n_img, 1280, 720, rgb, SES3D_HOST_BUFFERS, NULL);
      if (rc != SES3D_OK) { printf("%s\n", ses3d_last_error_string()); return 22; }
      for (i = 0; i < 720 * 1280; ++i) coloured += (rgb[3 * i] != 255 || rgb[3 * i + 1] != 255 || rgb[3 * i + 2] != 255);
      printf("overlay: %d coloured pixels\n", coloured);
    }

    /* pose_prior: feed the same skeleton as 12 consecutive 30 Hz messages; it is published from the 11th on */
    {
      ses3d_prior_params pp;
      ses3d_prior pr = NULL;
      ses3d_person_cov fused[8], pred[8];
      int32_t n_in = n3, n_pub = 0, ids[8], nobs[8];
      float delay = 0.f;
      int t;
      ses3d_prior_default_params(&pp);
      if (pp.min_num_obs_track != 10 || pp.dist_threshold != 5.0) return 13;
      if (ses3d_prior_create(&pp, 1, 8, 0, &pr) != SES3D_OK) { printf("%s\n", ses3d_last_error_string()); return 14; }
      for (t = 0; t < 12; ++t) {
        const int64_t stamp = 1000000000000LL + (int64_t)t * 33333333LL;
        rc = ses3d_prior_run(pr, 1, 1, 8, out3d, &n_in, &stamp, 0, NULL, fused, pred, &n_pub, &delay, NULL,
                             SES3D_HOST_BUFFERS, NULL);
        if (rc != SES3D_OK) { printf("%s\n", ses3d_last_error_string()); return 15; }
        if ((t < 10) != (n_pub == 0)) return 16;
      }
      if (ses3d_prior_get_tracks(pr, 0, ids, nobs) != 1 || ids[0] != 0 || nobs[0] != 12) return 17;
      printf("pose_prior: track %d, %d observations, fused nose = (%.3f, %.3f, %.3f), predicted delay %.3f s\n", ids[0],
             nobs[0], fused[0].keypoints[SES3D_FBP_NOSE].x, fused[0].keypoints[SES3D_FBP_NOSE].y,
             fused[0].keypoints[SES3D_FBP_NOSE].z, delay);
      ses3d_prior_destroy(pr);
    }
  }
  ses3d_destroy(h);
  return 0;
}
