/*
 * ses3d.h — C ABI of the B200-native multi-view geometry hot path
 * (cross-view association -> per-joint DLT triangulation + UT covariance ->
 *  skeleton plausibility/merge -> reprojection into every camera).
 *
 * This header is the drop-in boundary for the two reference entry points
 *   triangulate_persons(...)      skeleton_3d/src/skeleton_3d_triang_mult_node.cpp:525-526 (called :1069)
 *   fusedSkeletonCallback(...)    pose_reprojection/src/skeleton_reproj_mult_node.cpp:139 (bound :293)
 * and for the one-time table set-up done in their main()s
 *   skeleton_3d_triang_mult_node.cpp:1184-1214, skeleton_reproj_mult_node.cpp:272-279.
 * The reference has no library target (both functions live in node executables),
 * so the ABI is derived from those signatures: ROS messages become POD mirror
 * structs of person_msgs/msg/(.msg files), tf/CameraInfo become ses3d_camera.
 *
 * Plain C, no CUDA/torch types. All functions return an int status (0 = ok,
 * <0 = error, see SES3D_E_*), never throw, and never allocate on the steady
 * state per-batch path (device scratch is grown once and kept in the handle).
 */
#ifndef SES3D_H_
#define SES3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ layouts */

#define SES3D_NUM_KEYPOINTS 17        /* detector joints, NUM_KEYPOINTS S3D:56,1098 */
#define SES3D_NUM_FUSION_KEYPOINTS 21 /* FUSION_BODY_PARTS::NUM_KEYPOINTS, fusion_body_parts.h:25 */

/* person_msgs/msg/Keypoint2D.msg:1-4 — 24 B */
typedef struct ses3d_keypoint2d {
  float x, y, score;
  float cov[3]; /* xx, xy, yy */
} ses3d_keypoint2d;

/* person_msgs/msg/Person2D.msg:1-5 — 428 B (fixed 17 keypoints, no length prefix) */
typedef struct ses3d_person2d {
  float score;
  ses3d_keypoint2d keypoints[SES3D_NUM_KEYPOINTS];
  float bbox[4]; /* x0, y0, x1, y1 */
} ses3d_person2d;

/* person_msgs/msg/KeypointWithCovariance.msg:1-3 — 80 B with natural alignment */
typedef struct ses3d_keypoint_cov {
  double x, y, z; /* geometry_msgs/Point joint */
  float score;
  float pad_;
  double cov[6]; /* xx, xy, xz, yy, yz, zz */
} ses3d_keypoint_cov;

/* person_msgs/msg/PersonCov.msg:1-8 — 1768 B; keypoints indexed by FUSION_BODY_PARTS slot */
typedef struct ses3d_person_cov {
  uint32_t id;
  float score;
  ses3d_keypoint_cov keypoints[SES3D_NUM_FUSION_KEYPOINTS];
  double bbox_center[7]; /* geometry_msgs/Pose: position xyz, orientation xyzw */
  double bbox_size[3];   /* geometry_msgs/Vector3 */
} ses3d_person_cov;

/* FUSION_BODY_PARTS slots, skeleton_3d/include/skeleton_3d/fusion_body_parts.h:4-25 */
enum {
  SES3D_FBP_NOSE = 0, SES3D_FBP_NECK = 1, SES3D_FBP_RSHOULDER = 2, SES3D_FBP_RELBOW = 3,
  SES3D_FBP_RWRIST = 4, SES3D_FBP_LSHOULDER = 5, SES3D_FBP_LELBOW = 6, SES3D_FBP_LWRIST = 7,
  SES3D_FBP_MIDHIP = 8, SES3D_FBP_RHIP = 9, SES3D_FBP_RKNEE = 10, SES3D_FBP_RANKLE = 11,
  SES3D_FBP_LHIP = 12, SES3D_FBP_LKNEE = 13, SES3D_FBP_LANKLE = 14, SES3D_FBP_REYE = 15,
  SES3D_FBP_LEYE = 16, SES3D_FBP_REAR = 17, SES3D_FBP_LEAR = 18, SES3D_FBP_HEAD = 19,
  SES3D_FBP_BELLY = 20
};

/* One camera: what getTransforms()/getIntrinsics() deliver in the reference
 * (S3D:161-228, REP:77-137). T_cam_base is the row-major 3x4 [R|t] that maps a
 * base-frame point into the camera optical frame (lookupTransform(target=cam,
 * source=base), S3D:166-167, 1192). fx..Ty are the CameraInfo P-matrix entries
 * image_geometry::PinholeCameraModel exposes (P[0],P[5],P[2],P[6],P[3],P[7]). */
typedef struct ses3d_camera {
  double T_cam_base[12];
  double fx, fy, cx, cy, Tx, Ty;
  uint32_t width, height;
} ses3d_camera;

enum { SES3D_POSE_SIMPLE = 0, SES3D_POSE_H36M = 1 };   /* param pose_method, S3D:1095,1101-1112 */
enum { SES3D_PRECISION_FP32 = 0, SES3D_PRECISION_FP64 = 1 };

/* The reference's run-time parameters and compile-time constants (S3D:43-64,149). */
typedef struct ses3d_params {
  int32_t pose_method;                 /* SES3D_POSE_SIMPLE */
  int32_t precision;                   /* SES3D_PRECISION_FP32 = the reference's float DLT */
  int32_t lm_refine;                   /* 0 = off (reference behaviour); 1 = LM refinement (not in the reference) */
  int32_t lm_max_iters;                /* 10 */
  int32_t min_num_valid_keypoints;     /* 9     S3D:57 */
  float triangulation_threshold;       /* 0.30f S3D:58,1099 */
  double max_epipolar_error;           /* 0.050 S3D:60,1097 (demo launch file: 0.045) */
  double reproj_error_max_acceptable;  /* 0.050 S3D:59 */
  double max_joint_dist_to_root;       /* 2.0   S3D:61 */
  double merge_dist_thresh;            /* 0.20  S3D:62 */
  double limb_cov_offset_sigma;        /* 0.075 S3D:149 */
} ses3d_params;

/* Optional association dump used for the bit-exact index check; every pointer may be NULL. */
typedef struct ses3d_assoc_dump {
  int32_t* hyp_of;      /* [n_frames][n_cams][p_max]: hypothesis index of each detection, -1 = none */
  int32_t* n_hyp;       /* [n_frames]: number of hypotheses H.size() after the last camera (S3D:674) */
  int32_t* n_hungarian; /* [n_frames]: how many cameras needed the Hungarian solve (S3D:628-634) */
} ses3d_assoc_dump;

typedef struct ses3d_handle_s* ses3d_handle;

/* ------------------------------------------------------------------- status */
enum {
  SES3D_OK = 0,
  SES3D_E_INVALID = -1,   /* bad argument */
  SES3D_E_CUDA = -2,      /* CUDA runtime error (see ses3d_last_error_string) */
  SES3D_E_CAPACITY = -3,  /* a frame produced more hypotheses / persons than h_max */
  SES3D_E_NOMEM = -4
};

/* flags for the *_batch calls */
enum {
  SES3D_HOST_BUFFERS = 0,   /* all data pointers are host memory (pinned recommended) */
  SES3D_DEVICE_BUFFERS = 1  /* all data pointers are device memory on the handle's GPU */
};

/* -------------------------------------------------------------- entry points */

void ses3d_default_params(ses3d_params* p);

/* Replaces the set-up in main(): S3D:1184-1211 (P, camera centres, fundamental
 * matrices F_ij for i<j in FP64 then cast to float) and REP:272-279.
 * device = CUDA device ordinal. */
int ses3d_create(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* params,
                 int32_t device, ses3d_handle* out);
int ses3d_destroy(ses3d_handle h);

/* Number of cameras the handle was created with. */
int32_t ses3d_n_cams(ses3d_handle h);

/* Read back the constant tables (host memory): P [n_cams][12] float row-major,
 * F [n_cams*(n_cams-1)/2][9] float row-major in get_fundamental_idx order (S3D:242-253). */
int ses3d_get_tables(ses3d_handle h, float* P, float* F);

/* Replaces triangulate_persons() (S3D:525-997) for a batch of frames.
 *   persons   [n_frames][n_cams][p_max]   n_persons [n_frames][n_cams]
 *   out       [n_frames][h_max]           n_out     [n_frames]
 * Output persons are in hypothesis-index order (the reference built without
 * OpenMP), after plausibility checks and merge. Fewer than two cameras with
 * detections is not an error (n_out = 0, S3D:557-560).
 * stream: a cudaStream_t cast to void*, used by device-buffer calls only. NULL means the legacy default stream
 * (stream 0), exactly as a NULL cudaStream_t does in the CUDA runtime.
 * Host-buffer calls are synchronous. A host-buffer call with n_frames = 1 and no association dump (the per-message
 * call of the ROS nodes) replays a CUDA graph the handle captured on the previous call of the same shape: one launch
 * for upload, kernels and download; results are identical to the batched call (SES3D_FRAME_GRAPH=0 disables it).
 * Device-buffer calls are STREAM-ORDERED: the kernels are enqueued on `stream`
 * and the call returns without waiting, so the caller can enqueue the next batch, or its own consumers, behind
 * it. A capacity overflow (more hypotheses than h_max in some frame) of such a call is reported by ses3d_check()
 * or by the next batch call on the handle, whichever comes first. One handle owns one set of device scratch: calls
 * issued on different streams are serialised on the device in call order. */
int ses3d_triangulate_batch(ses3d_handle h, int32_t n_frames, int32_t p_max,
                            const ses3d_person2d* persons, const int32_t* n_persons,
                            int32_t h_max, ses3d_person_cov* out, int32_t* n_out,
                            const ses3d_assoc_dump* dump, uint32_t flags, void* stream);

/* Replaces fusedSkeletonCallback() (REP:139-235) for a batch of frames.
 *   persons3d [n_frames][h_max]            n_persons3d [n_frames]
 *   out       [n_frames][n_cams][h_max]    n_out       [n_frames][n_cams] */
int ses3d_reproject_batch(ses3d_handle h, int32_t n_frames, int32_t h_max,
                          const ses3d_person_cov* persons3d, const int32_t* n_persons3d,
                          ses3d_person2d* out, int32_t* n_out, uint32_t flags, void* stream);

/* Both stages chained on the device (skeleton_3d -> pose_reprojection, the
 * PersonCovList never leaves HBM); any of out3d/n_out3d may be NULL with host
 * buffers to skip that copy. */
int ses3d_process_batch(ses3d_handle h, int32_t n_frames, int32_t p_max,
                        const ses3d_person2d* persons, const int32_t* n_persons,
                        int32_t h_max, ses3d_person_cov* out3d, int32_t* n_out3d,
                        ses3d_person2d* out2d, int32_t* n_out2d,
                        const ses3d_assoc_dump* dump, uint32_t flags, void* stream);

/* Ragged form of ses3d_process_batch. The reference's messages are variable-length lists
 * (Person2D[] persons, PersonCov[] persons), so only occupied records travel:
 *   persons_dense  all detections back to back, frame-major then camera-major; n_persons [n_frames][n_cams]
 *                  gives the run lengths (each <= p_max)
 *   out3d          dense PersonCov records, frame-major, n_out3d [n_frames] run lengths, capacity cap3d records
 *   out2d          dense reprojected Person2D records, frame- then camera-major, n_out2d [n_frames][n_cams],
 *                  capacity cap2d records
 * Offsets are the exclusive prefix sums of the count arrays; totals are returned. Synchronous.
 * Device outputs are written straight to their final position by the pack kernels (running offsets stay on the
 * device, no host round trip between the internal chunks). Host outputs are packed on the device and leave through
 * the copy engine, chunk by chunk, overlapped with the upload and the kernels of the neighbouring chunks
 * (SES3D_RAGGED_DIRECT=1 makes the pack kernels write into pinned host memory themselves - measured slower). */
int ses3d_process_batch_ragged(ses3d_handle h, int32_t n_frames, int32_t p_max,
                               const ses3d_person2d* persons_dense, const int32_t* n_persons,
                               int32_t h_max, ses3d_person_cov* out3d, int64_t cap3d, int32_t* n_out3d,
                               ses3d_person2d* out2d, int64_t cap2d, int32_t* n_out2d,
                               int64_t* total3d, int64_t* total2d, uint32_t flags);

/* Wait for the handle's outstanding stream-ordered (device-buffer) work and report its status: SES3D_OK, or
 * SES3D_E_CAPACITY if a frame of those calls exceeded h_max (results of that call are then incomplete). */
int ses3d_check(ses3d_handle h);

/* Pin the calling host thread to the CPUs of the NUMA node `device` hangs off (read from sysfs; a no-op where that
 * information is not available). Call it before allocating the pinned buffers of a rank: on multi-socket boxes the
 * host<->device copies of the batch calls then stay on the GPU's own PCIe root. numa_node (nullable) receives the
 * node number or -1. */
int ses3d_bind_thread_to_device_numa(int32_t device, int32_t* numa_node);

/* ---------------------------------------------------------- single-process multi-GPU
 * One handle per device of the list (n_devices = 0 / devices = NULL: every visible device). A batch is cut into
 * contiguous frame ranges, range g = [g N / G, (g+1) N / G) runs on device g from its own host thread (bound to the
 * GPU's NUMA node); there is no data-path collective. Host buffers only (pinned recommended). */
typedef struct ses3d_multi_s* ses3d_multi;
int ses3d_create_multi(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* params, int32_t n_devices,
                       const int32_t* devices, ses3d_multi* out);
int ses3d_multi_destroy(ses3d_multi m);
int32_t ses3d_multi_device_count(ses3d_multi m);
ses3d_handle ses3d_multi_handle(ses3d_multi m, int32_t i);   /* the i-th device's handle (e.g. for ses3d_reserve) */
/* ses3d_process_batch over all devices; same argument meaning, host buffers. */
int ses3d_multi_process_batch(ses3d_multi m, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                              const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int32_t* n_out3d,
                              ses3d_person2d* out2d, int32_t* n_out2d, const ses3d_assoc_dump* dump);
/* ses3d_process_batch_ragged over all devices. The dense outputs come back as one segment per device: device g
 * writes its records from index seg[2g] on (= cap * first_frame_g / n_frames: its proportional share of the
 * buffer) and seg[2g+1] receives how many it wrote; frames stay in order inside a segment, segments are in device
 * order. seg3d / seg2d: [n_devices][2]. */
int ses3d_multi_process_batch_ragged(ses3d_multi m, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons_dense,
                                     const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int64_t cap3d,
                                     int32_t* n_out3d, ses3d_person2d* out2d, int64_t cap2d, int32_t* n_out2d,
                                     int64_t* seg3d, int64_t* seg2d);

/* Pre-size the handle's device scratch (otherwise grown on first use). */
int ses3d_reserve(ses3d_handle h, int32_t n_frames, int32_t p_max, int32_t h_max);

/* Diagnostics: n independent rows x cols assignment problems (cost [n][rows*cols], column-major like
 * HungarianAlgorithm::assignmentoptimal, Hungarian.h:24) through the device Munkres; assignment [n][rows]. */
int ses3d_munkres_batch(ses3d_handle h, int32_t n, int32_t rows, int32_t cols, const double* cost,
                        int32_t* assignment);

/* How many kernels the handle has launched so far (bench.py's gpu_launches). */
int64_t ses3d_launch_count(ses3d_handle h);

/* Device-time (ms, CUDA events on the launching stream) of each kernel of the
 * most recent device-buffer call when profiling was enabled with
 * ses3d_set_profiling(h, 1): order = {associate, triangulate, finalize, reproject}. */
int ses3d_set_profiling(ses3d_handle h, int32_t on);
int ses3d_last_kernel_ms(ses3d_handle h, float ms[4]);

const char* ses3d_last_error_string(void);
const char* ses3d_version(void);

/* Diagnostics: measured peak of the CUDA-core FMA pipe of `device` in TFLOP/s (fp64 = 0: FP32, 1: FP64), a
 * micro-benchmark of independent FMA chains. The benchmarks use it as the measured roofline denominator of the
 * FP32 / FP64 kernels (MEASURED_PEAKS.json holds HBM and tensor-core figures only). */
int ses3d_measure_fma_peak(int32_t device, int32_t fp64, double* tflops);

/* ------------------------------------------------- synthetic frame generator
 * Test/bench input source; not part of the reference. Counter
 * based (Philox4x32-10), IEEE-only arithmetic, so host and device variants are
 * bit-identical. gt_id (nullable) [n_frames][n_cams][p_max] = generator person
 * id of every emitted detection (-1 = empty slot). */
typedef struct ses3d_synth_config {
  uint64_t seed;
  int32_t n_people;       /* people in the scene */
  int32_t p_max;          /* slots per camera (>= n_people) */
  float dropout;          /* per-keypoint dropout probability */
  float noise_px;         /* 2-D noise sigma in pixels */
  float area[4];          /* x0, y0, x1, y1 of the floor area people stand in (base frame) */
  float min_separation;   /* metres between roots */
  int32_t min_visible;    /* emit a detection only if >= this many keypoints are inside the image */
  int32_t frames_per_sequence; /* 0: every frame is an independent scene. T > 0: frames [qT, (q+1)T) form sequence q —
                                  one scene whose people walk straight ahead (temporally coherent input for the
                                  pose_prior stage); the 2-D noise stays independent per frame */
  float step_m;           /* walking distance per frame in sequence mode (e.g. 1 m/s at 30 Hz = 0.0333) */
} ses3d_synth_config;

int ses3d_synth_frames(int32_t n_cams, const ses3d_camera* cams, const ses3d_synth_config* cfg,
                       int64_t first_frame, int32_t n_frames,
                       ses3d_person2d* persons, int32_t* n_persons, int32_t* gt_id,
                       float* gt_joints /* nullable [n_frames][n_people][17][3] */);

/* Same frames, generated on the GPU into device buffers (cams is a host array). Synchronous. */
int ses3d_synth_frames_device(int32_t n_cams, const ses3d_camera* cams, const ses3d_synth_config* cfg,
                              int64_t first_frame, int32_t n_frames,
                              ses3d_person2d* persons, int32_t* n_persons, int32_t* gt_id, void* stream);

/* ------------------------------------------------------------- frame assembler
 * Host-side restatement of the approximate-time synchroniser that defines what a "frame" is in live
 * operation (my_message_filters/sync_policies/approximate_time_vec.h:170-217, 488-626;
 * synchronizer_vec.h:147-188) plus the worker loop's gating (S3D:1029-1057). Messages are (stamp, id) pairs;
 * the caller keeps the payloads and packs emitted frames into the batch arrays (blanked camera = 0 persons). */
typedef struct ses3d_assembler_config {
  int32_t n_cams;
  uint32_t queue_size;                   /* std::max(3u, 1 + n_cams / 4)   S3D:1219 */
  int64_t inter_message_lower_bound_ns;  /* 20 ms                          S3D:1220 */
  double age_penalty;                    /* 2.0                            S3D:1221 */
  int64_t max_interval_ns;               /* < 0: unlimited (ros::DURATION_MAX) */
  double max_sync_diff_s;                /* 0.067                          S3D:64, 1051; < 0: synchroniser only -
                                            no worker gating, every synchronised tuple is emitted unchanged */
} ses3d_assembler_config;
typedef struct ses3d_assembler_s* ses3d_assembler;

int ses3d_assembler_default_config(int32_t n_cams, ses3d_assembler_config* cfg);
int ses3d_assembler_create(const ses3d_assembler_config* cfg, ses3d_assembler* out);
int ses3d_assembler_destroy(ses3d_assembler a);
/* One message of camera `cam`; returns how many frames became ready (>= 0) or an error (< 0). */
int ses3d_assembler_add(ses3d_assembler a, int32_t cam, int64_t stamp_ns, int64_t msg_id);
/* Next ready frame: ids / stamps_ns / blank [n_cams] (blank[i] = 1: camera i lags the pivot by more than
 * max_sync_diff_s and is replaced by an empty list, S3D:1049-1057), pivot = camera with the newest stamp.
 * Returns 1 when a frame was written, 0 when none is ready. */
int ses3d_assembler_pop(ses3d_assembler a, int64_t* ids, int64_t* stamps_ns, uint8_t* blank, int32_t* pivot);
/* stats: {frames emitted, frames skipped by the backwards-time rule (S3D:1043-1046), blanked cameras,
 * messages dropped by queue overflow, tuples signalled by the synchroniser} */
int ses3d_assembler_stats(ses3d_assembler a, int64_t stats[5]);

/* The 1-slot latest-wins mailbox between the synchroniser callback and the worker thread (S3D:999-1025), replayed
 * deterministically: frame i reaches the slot at t_ready_ns[i] (non-decreasing) and overwrites an unread frame; the idle
 * worker takes the slot's content and is busy for busy_ns[i]. taken[i] = 1: processed, 0: overwritten unseen;
 * t_start_ns (nullable): when processing of frame i began (-1 if dropped). Returns the number of processed frames. Use
 * it to thin a replayed frame list to what the live node would have processed at a given per-frame cost. */
int ses3d_mailbox_replay(int32_t n, const int64_t* t_ready_ns, const int64_t* busy_ns, uint8_t* taken,
                         int64_t* t_start_ns);

/* ------------------------------------------------------------- wire format
 * ROS-free (de)serialisation of the ROS 1 wire encoding of person_msgs/Person2DList and PersonCovList
 * (person_msgs/msg, std_msgs/Header, geometry_msgs Point/Pose/Vector3): replay of recorded message bodies
 * through the batch ABI. decode_* return the number of persons in the message (which may exceed `cap`; only
 * the first `cap` are written) or SES3D_E_INVALID for a truncated / malformed buffer. encode_* return the
 * encoded size; nothing is written beyond `cap` (call with buf = NULL to size the buffer). */
int ses3d_wire_decode_person2dlist(const uint8_t* buf, size_t len, uint32_t* seq, int64_t* stamp_ns, char* frame_id,
                                   size_t frame_id_cap, float* fb_delay, ses3d_person2d* persons, int32_t cap);
size_t ses3d_wire_encode_person2dlist(uint32_t seq, int64_t stamp_ns, const char* frame_id, float fb_delay,
                                      const ses3d_person2d* persons, int32_t n, uint8_t* buf, size_t cap);
int ses3d_wire_decode_personcovlist(const uint8_t* buf, size_t len, uint32_t* seq, int64_t* stamp_ns, char* frame_id,
                                    size_t frame_id_cap, int64_t* ts_per_cam_ns, float* fb_delay_per_cam,
                                    int32_t cam_cap, int32_t* n_cams, ses3d_person_cov* persons, int32_t cap);
size_t ses3d_wire_encode_personcovlist(uint32_t seq, int64_t stamp_ns, const char* frame_id, int32_t n_cams,
                                       const int64_t* ts_per_cam_ns, const float* fb_delay_per_cam,
                                       const ses3d_person_cov* persons, int32_t n, uint8_t* buf, size_t cap);

/* ------------------------------------------------------------- pose_prior
 * Replaces skeletonCallback() of pose_prior/src/pose_prior_mult_node.cpp:505-921 (bound :947) with its
 * file-scope state (g_tracks :123, g_t_prev :58, g_next_id :59, g_frame_nr :60, g_fb_delay_buffer :54):
 * per-person tracking (Hungarian on the velocity-normalised joint distance, :84-101, :548-580), the skeleton
 * model fit (unary Gaussian factors :126-145, :690, :714, :732 and bone-length range factors :384-481 solved
 * with gtsam's LevenbergMarquardtOptimizer :746-749), marginal covariances (:760-816), the constant-velocity
 * prediction (:818-831) and the track life cycle (:191-211, :839-848, :866-903).
 * The stage is stateful across the frames of one message stream, so the parallel unit is the *sequence*:
 * one handle tracks n_sequences independent streams (rigs / replays); the frames of a sequence are processed in
 * order, inside one launch when several are passed at once. Marker output is dropped (visualisation). */
typedef struct ses3d_prior_params {
  int32_t pose_method;          /* SES3D_POSE_SIMPLE  param pose_method   PRI:39,930 */
  int32_t normalize_by_height;  /* 0                  param norm_height   PRI:40,931 (selects the bone table and
                                   limb_sigma_factor 2.0 instead of 1.0, PRI:934-937) */
  int32_t min_num_obs_track;    /* 10   PRI:66 */
  int32_t lm_max_iterations;    /* 100  gtsam LevenbergMarquardtParams defaults (gtsam 4.0.3, README.md:22) */
  float min_score;              /* 0.10f PRI:50 */
  float pad_;
  double pred_noise_sigma;      /* 0.12 PRI:47 */
  double default_res_sigma;     /* 0.10 PRI:48 */
  double avg_delay;             /* 0.10 PRI:51 */
  double root_sigma_factor;     /* 100  PRI:52 */
  double t_max_unobserved;      /* 1.0  PRI:62 */
  double dist_threshold;        /* 5.0  PRI:63 */
  double merge_dist_thresh;     /* 0.20 PRI:64 */
  double lm_lambda_initial;     /* 1e-5 */
  double lm_lambda_factor;      /* 10   */
  double lm_lambda_upper_bound; /* 1e5  */
  double lm_relative_error_tol; /* 1e-5 */
  double lm_absolute_error_tol; /* 1e-5 */
  double lm_min_model_fidelity; /* 1e-3 */
} ses3d_prior_params;

typedef struct ses3d_prior_s* ses3d_prior;

void ses3d_prior_default_params(ses3d_prior_params* p);
/* max_tracks = capacity of g_tracks per sequence (a frame that would exceed it fails with SES3D_E_CAPACITY). */
int ses3d_prior_create(const ses3d_prior_params* params, int32_t n_sequences, int32_t max_tracks, int32_t device,
                       ses3d_prior* out);
int ses3d_prior_destroy(ses3d_prior p);
int ses3d_prior_reset(ses3d_prior p); /* reset() PRI:182-189, all sequences */

/* n_frames consecutive PersonCovList messages of each of n_sequences streams (sequence-major arrays):
 *   persons  [n_sequences][n_frames][h_max]   n_persons [n_sequences][n_frames]
 *   stamp_ns [n_sequences][n_frames]          header.stamp (t = stamp.toSec(), PRI:506)
 *   fb_delay [n_sequences][n_frames][n_cams]  fb_delay_per_cam (PRI:513-526); NULL = no measurement (-1)
 *   fused, pred [n_sequences][n_frames][h_max] persons3d_fused / persons3d_fused_pred (PRI:906-907), in
 *            detection order (the reference built without OpenMP); n_out [n_sequences][n_frames]
 *   pred_delay [n_sequences][n_frames]        the fb_delay_per_cam value stored in both outputs (PRI:531); nullable
 *   track_of [n_sequences][n_frames][h_max]   diagnostics, nullable: track id each detection was fused into
 * Tracker state persists in the handle between calls (streaming: n_frames = 1 per call). */
int ses3d_prior_run(ses3d_prior p, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                    const ses3d_person_cov* persons, const int32_t* n_persons, const int64_t* stamp_ns,
                    int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused, ses3d_person_cov* pred,
                    int32_t* n_out, float* pred_delay, int32_t* track_of, uint32_t flags, void* stream);

/* Ragged form of ses3d_prior_run (host buffers): the messages are variable-length lists (PersonCov[] persons,
 * PersonCovList.msg:4), so only occupied records cross PCIe.
 *   persons_dense  all input records back to back, stream-major then message-major; n_persons [n_sequences][n_frames]
 *                  gives the run lengths (each <= h_max)
 *   fused_dense / pred_dense  the published records back to back in the same order, capacity `cap` records each;
 *                  n_out [n_sequences][n_frames] run lengths, *total = records written to each of the two arrays. */
int ses3d_prior_run_ragged(ses3d_prior p, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                           const ses3d_person_cov* persons_dense, const int32_t* n_persons, const int64_t* stamp_ns,
                           int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused_dense,
                           ses3d_person_cov* pred_dense, int64_t cap, int32_t* n_out, float* pred_delay, int64_t* total);

/* Diagnostics: live tracks of one sequence (ids / num_obs [max_tracks], nullable); returns the track count or <0. */
int ses3d_prior_get_tracks(ses3d_prior p, int32_t sequence, int32_t* ids, int32_t* num_obs);
int64_t ses3d_prior_launch_count(ses3d_prior p);
/* Device time (ms, CUDA events) of the most recent device-buffer ses3d_prior_run. */
int ses3d_prior_last_kernel_ms(ses3d_prior p, float* ms);

/* ------------------------------------------------------------- visualisation
 * The numeric content of the reference's rviz markers; message assembly (headers, colours, lifetimes) stays in the node.
 *   ellipsoids [n_frames][h_max][21]  covariance ellipsoid of every joint with score > 0 (setMarkerPose S3D:279-310 ==
 *                                     PRI:237-254): orientation quaternion + axis lengths 2 x 2.7955 x sqrt(eigenvalue);
 *                                     all-zero for absent joints
 *   segments   [n_frames][h_max][22][2][3]  LINE_LIST end points of the skeleton marker, n_segments [n_frames][h_max];
 *              style 0 = skeleton_3d (S3D:898-916), 1 = pose_prior (addJointToSkeleton PRI:273-382);
 *              segment_slot [n_frames][h_max][22] (nullable) = fusion slot whose colour the segment carries
 * ellipsoids or segments may be NULL. */
typedef struct ses3d_ellipsoid {
  double qw, qx, qy, qz; /* marker.pose.orientation */
  double sx, sy, sz;     /* marker.scale */
} ses3d_ellipsoid;
enum { SES3D_MARKERS_SKELETON3D = 0, SES3D_MARKERS_POSE_PRIOR = 1 };
#define SES3D_MARKER_MAX_SEGMENTS 22
int ses3d_markers_batch(ses3d_handle h, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons3d,
                        const int32_t* n_persons3d, int32_t style, ses3d_ellipsoid* ellipsoids, double* segments,
                        int32_t* n_segments, int8_t* segment_slot, uint32_t flags, void* stream);

/* 2-D overlay renderer: the image person_msgs/scripts/pose2D_plot_node.py publishes for one Person2DList
 * (draw_humans :18-66 on a white rgb8 canvas, callback_pose :82-91; the node uses 640 x 480). Per person, in this
 * order: a filled circle (radius max(1, width/360) * 5) at every keypoint with score >= 0.25 in the COCO colour of the
 * joint, a line (thickness max(1, width/360) * 4) for every CocoPairs limb whose two joints were drawn in the colour
 * of its second joint, the bounding box grown by 6 px (thickness max(1, width/360) * 2) in colour 0.
 *   persons [n_images][p_max], n_persons [n_images], rgb [n_images][height][width][3]
 * OpenCV's exact pixel coverage is not reproduced (no OpenCV here): coverage rules are documented in kernels_overlay.cu. */
int ses3d_overlay_batch(ses3d_handle h, int32_t n_images, int32_t p_max, const ses3d_person2d* persons,
                        const int32_t* n_persons, int32_t width, int32_t height, uint8_t* rgb, uint32_t flags,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SES3D_H_ */
