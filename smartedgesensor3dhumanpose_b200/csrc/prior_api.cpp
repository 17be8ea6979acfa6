// prior_api.cpp — C ABI of the pose_prior stage (include/ses3d.h, "pose_prior" section) on top of K7.
//
// ses3d_prior_create  <- main() of pose_prior_mult_node.cpp (PRI:923-947: parameters, limb sigma factor)
// ses3d_prior_run     <- skeletonCallback (PRI:505-921) for n_sequences streams x n_frames messages
// ses3d_prior_reset   <- reset() (PRI:182-189)
// No CPU fallback: the handle needs a CUDA device. Tracker state (tracks, counters) lives in device memory owned
// by the handle; host-buffer calls stage the records through device buffers that grow once and are reused.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "launch.h"
#include "ses3d.h"

namespace ses3d {
int set_error(int code, const std::string& msg);
const char* last_error();
}

namespace {

int cuda_fail(cudaError_t e, const char* what) {
  return ses3d::set_error(SES3D_E_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

}  // namespace

struct PriorSlot {   // one in-flight chunk of streams of a ragged call
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  long long* total = nullptr;   // pinned
  unsigned char* h_stage = nullptr;   // pinned staging of the chunk's n_out / pred_delay (the caller's arrays may be pageable,
  size_t h_stage_cap = 0;             //   and an async copy into pageable memory blocks the host until the kernel is done)
  cudaError_t ensure_stage(size_t bytes) {
    if (bytes <= h_stage_cap) return cudaSuccess;
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr; h_stage_cap = 0;
    cudaError_t e = cudaMallocHost(reinterpret_cast<void**>(&h_stage), bytes + bytes / 8 + 256);
    if (e == cudaSuccess) h_stage_cap = bytes + bytes / 8 + 256;
    return e;
  }
  Buf persons, n, stamp, delay, fused, pred, n_out, delay_out, dense_in, dense_fused, dense_pred, off_in, off_out;
  void release() {
    for (Buf* b : {&persons, &n, &stamp, &delay, &fused, &pred, &n_out, &delay_out, &dense_in, &dense_fused, &dense_pred,
                   &off_in, &off_out})
      b->release();
    if (total) cudaFreeHost(total);
    if (h_stage) cudaFreeHost(h_stage);
    h_stage = nullptr; h_stage_cap = 0;
    if (done) cudaEventDestroy(done);
    if (stream) cudaStreamDestroy(stream);
    total = nullptr; done = nullptr; stream = nullptr;
  }
};

struct ses3d_prior_s {
  int device = 0;
  int n_sequences = 0, max_tracks = 0;
  ses3d::PriorTables pt;
  Buf states, tracks, order;
  Buf in_persons, in_n, in_stamp, in_delay, out_fused, out_pred, out_n, out_delay, out_track;
  static constexpr int kSlots = 3;   // ragged call: H2D of chunk i+1, kernel of chunk i and D2H of chunk i-1 overlap
  PriorSlot rs[kSlots];
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int ragged_chunk_env = 0;       // SES3D_PRIOR_RAGGED_CHUNK, read at create
  // streaming calls (a few messages, the ROS node): one contiguous record [inputs | outputs | stream states] in pinned
  // host memory and its mirror on the device - one upload, one clear, one download instead of a dozen small copies
  Buf small_dev;
  unsigned char* small_pin = nullptr;
  size_t small_pin_bytes = 0;
  int small_msgs = 4;             // SES3D_PRIOR_SMALL_MSGS: calls of at most this many messages take that path (0 = off)
  bool dev_run_pending = false;   // a device-buffer run was enqueued on a caller's stream; ev1 marks its end
  float last_ms = 0.f;
  int64_t launches = 0;
  std::mutex mu;
};

// Device-buffer runs are enqueued on the caller's stream and return at once. Every other entry point that reads or
// rewrites the tracker state (host-buffer runs, reset, get_tracks, a device run on another stream) first orders itself
// behind the last such run.
static cudaError_t wait_for_device_runs(ses3d_prior_s* h, cudaStream_t st) {
  if (!h->dev_run_pending) return cudaSuccess;
  return cudaStreamWaitEvent(st, h->ev1, 0);
}

static int prior_run_ragged_impl(ses3d_prior h, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                                 const ses3d_person_cov* persons_dense, const int32_t* n_persons, const int64_t* stamp_ns,
                                 int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused_dense,
                                 ses3d_person_cov* pred_dense, int64_t cap, int32_t* n_out, float* pred_delay, int64_t* total);

extern "C" {

void ses3d_prior_default_params(ses3d_prior_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->pose_method = SES3D_POSE_SIMPLE;
  p->normalize_by_height = 0;
  p->min_num_obs_track = 10;
  p->lm_max_iterations = 100;
  p->min_score = 0.10f;
  p->pred_noise_sigma = 0.12;
  p->default_res_sigma = 0.10;
  p->avg_delay = 0.10;
  p->root_sigma_factor = 100.0;
  p->t_max_unobserved = 1.0;
  p->dist_threshold = 5.0;
  p->merge_dist_thresh = 0.20;
  p->lm_lambda_initial = 1e-5;
  p->lm_lambda_factor = 10.0;
  p->lm_lambda_upper_bound = 1e5;
  p->lm_relative_error_tol = 1e-5;
  p->lm_absolute_error_tol = 1e-5;
  p->lm_min_model_fidelity = 1e-3;
}

int ses3d_prior_create(const ses3d_prior_params* params, int32_t n_sequences, int32_t max_tracks, int32_t device,
                       ses3d_prior* out) {
  if (!out) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_create: out is NULL");
  *out = nullptr;
  if (n_sequences < 1) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_create: n_sequences < 1");
  if (max_tracks < 1 || max_tracks > ses3d::PRIOR_MAX_TRACKS)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_create: max_tracks must be in [1, 64]");
  ses3d_prior_params prm;
  if (params) prm = *params; else ses3d_prior_default_params(&prm);
  if (prm.pose_method != SES3D_POSE_SIMPLE && prm.pose_method != SES3D_POSE_H36M)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_create: unknown pose_method");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return ses3d::set_error(SES3D_E_CUDA, "ses3d_prior_create: no CUDA device (there is no CPU path)");
  if (device < 0 || device >= n_dev) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_create: bad device ordinal");
  CU(cudaSetDevice(device));
  CU(ses3d::init_prior_kernels(device));
  ses3d_prior_s* h = new ses3d_prior_s;
  h->device = device;
  if (const char* env = getenv("SES3D_PRIOR_RAGGED_CHUNK")) h->ragged_chunk_env = atoi(env);
  if (const char* env = getenv("SES3D_PRIOR_SMALL_MSGS")) h->small_msgs = std::max(0, atoi(env));
  h->n_sequences = n_sequences;
  h->max_tracks = max_tracks;
  h->pt.prm = prm;
  h->pt.limb_sigma_factor = prm.normalize_by_height ? 2.0 : 1.0;   // PRI:934-937
  h->pt.st = nullptr;   // set by the kernel (shared-memory copy of the skeleton tables)
  auto bail = [&](cudaError_t err, const char* what) { ses3d_prior_destroy(h); return cuda_fail(err, what); };
  if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&h->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
  for (PriorSlot& sl : h->rs) {
    if ((e = cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if ((e = cudaMallocHost(reinterpret_cast<void**>(&sl.total), sizeof(long long))) != cudaSuccess) return bail(e, "cudaMallocHost");
  }
  if ((e = h->states.ensure(sizeof(ses3d::PriorSeqState) * n_sequences)) != cudaSuccess) return bail(e, "cudaMalloc states");
  if ((e = h->tracks.ensure(sizeof(ses3d::PriorTrack) * (size_t)n_sequences * max_tracks)) != cudaSuccess) return bail(e, "cudaMalloc tracks");
  if ((e = h->order.ensure((size_t)n_sequences * max_tracks)) != cudaSuccess) return bail(e, "cudaMalloc order");
  if ((e = cudaMemsetAsync(h->tracks.p, 0, sizeof(ses3d::PriorTrack) * (size_t)n_sequences * max_tracks, h->stream)) != cudaSuccess) return bail(e, "cudaMemset");
  if ((e = cudaMemsetAsync(h->order.p, 0, (size_t)n_sequences * max_tracks, h->stream)) != cudaSuccess) return bail(e, "cudaMemset");
  if ((e = ses3d::launch_prior_reset(prm, n_sequences, h->states.as<ses3d::PriorSeqState>(), false, h->stream)) != cudaSuccess) return bail(e, "k_prior_reset");
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail(e, "cudaStreamSynchronize");
  ++h->launches;
  *out = h;
  return SES3D_OK;
}

int ses3d_prior_destroy(ses3d_prior h) {
  if (!h) return SES3D_OK;
  cudaSetDevice(h->device);
  for (Buf* b : {&h->states, &h->tracks, &h->order, &h->in_persons, &h->in_n, &h->in_stamp, &h->in_delay, &h->out_fused,
                 &h->out_pred, &h->out_n, &h->out_delay, &h->out_track})
    b->release();
  for (PriorSlot& sl : h->rs) sl.release();
  h->small_dev.release();
  if (h->small_pin) cudaFreeHost(h->small_pin);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return SES3D_OK;
}

int ses3d_prior_reset(ses3d_prior h) {
  if (!h) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_reset: NULL handle");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  CU(wait_for_device_runs(h, h->stream));
  CU(ses3d::launch_prior_reset(h->pt.prm, h->n_sequences, h->states.as<ses3d::PriorSeqState>(), true, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  ++h->launches;
  return SES3D_OK;
}

int ses3d_prior_run(ses3d_prior h, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                    const ses3d_person_cov* persons, const int32_t* n_persons, const int64_t* stamp_ns,
                    int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused, ses3d_person_cov* pred,
                    int32_t* n_out, float* pred_delay, int32_t* track_of, uint32_t flags, void* stream) {
  if (!h) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run: NULL handle");
  if (n_sequences < 0 || n_sequences > h->n_sequences)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run: n_sequences exceeds the handle's");
  if (n_frames < 0 || h_max < 1 || h_max > 64) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run: bad n_frames / h_max");
  if (n_cams < 0) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run: n_cams < 0");
  if (n_sequences == 0 || n_frames == 0) return SES3D_OK;
  if (!persons || !n_persons || !stamp_ns || !fused || !pred || !n_out)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run: NULL buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  const size_t n_msg = (size_t)n_sequences * n_frames;
  const size_t rec = sizeof(ses3d_person_cov) * n_msg * h_max;
  const bool on_device = (flags & SES3D_DEVICE_BUFFERS) != 0;
  // device buffers: the kernel is ordered on the caller's stream; NULL means the legacy default stream (0), exactly
  // as a NULL cudaStream_t does in the CUDA runtime, so producers / consumers on that stream are ordered with it
  cudaStream_t st = on_device ? static_cast<cudaStream_t>(stream) : h->stream;
  if (!on_device) CU(wait_for_device_runs(h, h->stream));
  auto* states = h->states.as<ses3d::PriorSeqState>();
  auto* tracks = h->tracks.as<ses3d::PriorTrack>();
  auto* order = h->order.as<uint8_t>();
  if (fb_delay == nullptr) n_cams = 0;

  if (on_device) {
    CU(cudaStreamSynchronize(h->stream));   // host-path / reset work of earlier calls (normally idle)
    CU(wait_for_device_runs(h, st));        // a previous device run on a different stream
    CU(cudaEventRecord(h->ev0, st));
    CU(ses3d::launch_prior(h->pt, n_sequences, n_frames, h_max, h->max_tracks, states, tracks, order, persons, n_persons,
                           stamp_ns, n_cams, fb_delay, fused, pred, n_out, pred_delay, track_of, st));
    CU(cudaEventRecord(h->ev1, st));
    h->dev_run_pending = true;   // later entry points wait on ev1 before they touch the tracker state
    ++h->launches;
    // the sticky overflow flag is checked on the host-buffer path and by ses3d_prior_get_tracks
    return SES3D_OK;
  }

  if (h->small_msgs > 0 && n_msg <= (size_t)h->small_msgs) {
    size_t at = 0;
    auto take = [&at](size_t bytes) { const size_t r = at; at += (bytes + 255) & ~(size_t)255; return r; };
    const size_t i_persons = take(rec), i_n = take(4 * n_msg), i_stamp = take(8 * n_msg),
                 i_delay = take(4 * n_msg * (size_t)std::max(n_cams, 1));
    const size_t o_fused = take(rec), o_pred = take(rec), o_n = take(4 * n_msg), o_delay = take(4 * n_msg),
                 o_track = take(4 * n_msg * h_max), o_end = at;
    const size_t st_bytes = sizeof(ses3d::PriorSeqState) * (size_t)n_sequences;
    const size_t p_states = take(st_bytes);   // host side only: the stream states live in their own device buffer
    if (h->small_pin_bytes < at) {
      if (h->small_pin) cudaFreeHost(h->small_pin);
      h->small_pin = nullptr;
      h->small_pin_bytes = 0;
      CU(cudaMallocHost(reinterpret_cast<void**>(&h->small_pin), at));
      h->small_pin_bytes = at;
    }
    CU(h->small_dev.ensure(o_end));
    unsigned char* pin = h->small_pin;
    unsigned char* dv = h->small_dev.as<unsigned char>();
    std::memcpy(pin + i_persons, persons, rec);
    std::memcpy(pin + i_n, n_persons, 4 * n_msg);
    std::memcpy(pin + i_stamp, stamp_ns, 8 * n_msg);
    if (n_cams > 0) std::memcpy(pin + i_delay, fb_delay, 4 * n_msg * n_cams);
    CU(cudaMemcpyAsync(dv, pin, o_fused, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dv + o_fused, 0, o_end - o_fused, st));   // unpublished slots of the outputs read as zero records
    CU(cudaEventRecord(h->ev0, st));
    CU(ses3d::launch_prior(h->pt, n_sequences, n_frames, h_max, h->max_tracks, states, tracks, order,
                           reinterpret_cast<const ses3d_person_cov*>(dv + i_persons),
                           reinterpret_cast<const int32_t*>(dv + i_n), reinterpret_cast<const int64_t*>(dv + i_stamp), n_cams,
                           n_cams > 0 ? reinterpret_cast<const float*>(dv + i_delay) : nullptr,
                           reinterpret_cast<ses3d_person_cov*>(dv + o_fused), reinterpret_cast<ses3d_person_cov*>(dv + o_pred),
                           reinterpret_cast<int32_t*>(dv + o_n), reinterpret_cast<float*>(dv + o_delay),
                           track_of ? reinterpret_cast<int32_t*>(dv + o_track) : nullptr, st));
    CU(cudaEventRecord(h->ev1, st));
    ++h->launches;
    CU(cudaMemcpyAsync(pin + o_fused, dv + o_fused, o_end - o_fused, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(pin + p_states, states, st_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    std::memcpy(fused, pin + o_fused, rec);
    std::memcpy(pred, pin + o_pred, rec);
    std::memcpy(n_out, pin + o_n, 4 * n_msg);
    if (pred_delay) std::memcpy(pred_delay, pin + o_delay, 4 * n_msg);
    if (track_of) std::memcpy(track_of, pin + o_track, 4 * n_msg * h_max);
    const auto* hs = reinterpret_cast<const ses3d::PriorSeqState*>(pin + p_states);
    for (int i = 0; i < n_sequences; ++i)
      if (hs[i].overflow) return ses3d::set_error(SES3D_E_CAPACITY, "ses3d_prior_run: a stream needed more than max_tracks tracks");
    return SES3D_OK;
  }
  CU(h->in_persons.ensure(rec));
  CU(h->in_n.ensure(4 * n_msg));
  CU(h->in_stamp.ensure(8 * n_msg));
  CU(h->out_fused.ensure(rec));
  CU(h->out_pred.ensure(rec));
  CU(h->out_n.ensure(4 * n_msg));
  CU(h->out_delay.ensure(4 * n_msg));
  if (track_of) CU(h->out_track.ensure(4 * n_msg * h_max));
  if (n_cams > 0) CU(h->in_delay.ensure(4 * n_msg * n_cams));
  CU(cudaMemcpyAsync(h->in_persons.p, persons, rec, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->in_n.p, n_persons, 4 * n_msg, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(h->in_stamp.p, stamp_ns, 8 * n_msg, cudaMemcpyHostToDevice, st));
  if (n_cams > 0) CU(cudaMemcpyAsync(h->in_delay.p, fb_delay, 4 * n_msg * n_cams, cudaMemcpyHostToDevice, st));
  // unpublished slots of the outputs read as zero records
  CU(cudaMemsetAsync(h->out_fused.p, 0, rec, st));
  CU(cudaMemsetAsync(h->out_pred.p, 0, rec, st));
  CU(cudaEventRecord(h->ev0, st));
  CU(ses3d::launch_prior(h->pt, n_sequences, n_frames, h_max, h->max_tracks, states, tracks, order,
                         h->in_persons.as<ses3d_person_cov>(), h->in_n.as<int32_t>(), h->in_stamp.as<int64_t>(), n_cams,
                         n_cams > 0 ? h->in_delay.as<float>() : nullptr, h->out_fused.as<ses3d_person_cov>(),
                         h->out_pred.as<ses3d_person_cov>(), h->out_n.as<int32_t>(), h->out_delay.as<float>(),
                         track_of ? h->out_track.as<int32_t>() : nullptr, st));
  CU(cudaEventRecord(h->ev1, st));
  ++h->launches;
  CU(cudaMemcpyAsync(fused, h->out_fused.p, rec, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(pred, h->out_pred.p, rec, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(n_out, h->out_n.p, 4 * n_msg, cudaMemcpyDeviceToHost, st));
  if (pred_delay) CU(cudaMemcpyAsync(pred_delay, h->out_delay.p, 4 * n_msg, cudaMemcpyDeviceToHost, st));
  if (track_of) CU(cudaMemcpyAsync(track_of, h->out_track.p, 4 * n_msg * h_max, cudaMemcpyDeviceToHost, st));
  std::vector<ses3d::PriorSeqState> hs(n_sequences);
  CU(cudaMemcpyAsync(hs.data(), states, sizeof(ses3d::PriorSeqState) * n_sequences, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (const auto& s : hs)
    if (s.overflow) return ses3d::set_error(SES3D_E_CAPACITY, "ses3d_prior_run: a stream needed more than max_tracks tracks");
  return SES3D_OK;
}

int ses3d_prior_run_ragged(ses3d_prior h, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                           const ses3d_person_cov* persons_dense, const int32_t* n_persons, const int64_t* stamp_ns,
                           int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused_dense,
                           ses3d_person_cov* pred_dense, int64_t cap, int32_t* n_out, float* pred_delay, int64_t* total) {
  if (!h) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run_ragged: NULL handle");
  if (total) *total = 0;
  if (n_sequences < 0 || n_sequences > h->n_sequences)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run_ragged: n_sequences exceeds the handle's");
  if (n_frames < 0 || h_max < 1 || h_max > 64 || n_cams < 0 || cap < 0)
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run_ragged: bad n_frames / h_max / n_cams / cap");
  if (n_sequences == 0 || n_frames == 0) return SES3D_OK;
  if (!n_persons || !stamp_ns || !n_out || !total || (cap > 0 && (!fused_dense || !pred_dense)))
    return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run_ragged: NULL buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  const int rc = prior_run_ragged_impl(h, n_sequences, n_frames, h_max, persons_dense, n_persons, stamp_ns, n_cams, fb_delay,
                                       fused_dense, pred_dense, cap, n_out, pred_delay, total);
  if (rc != SES3D_OK) {
    // a failed call may have copies into the caller's buffers in flight: nothing may be written after we return
    const std::string msg = ses3d::last_error();
    for (PriorSlot& sl : h->rs) cudaStreamSynchronize(sl.stream);
    cudaStreamSynchronize(h->stream);
    ses3d::set_error(rc, msg);
  }
  return rc;
}

}  // extern "C"

static int prior_run_ragged_impl(ses3d_prior h, int32_t n_sequences, int32_t n_frames, int32_t h_max,
                                 const ses3d_person_cov* persons_dense, const int32_t* n_persons, const int64_t* stamp_ns,
                                 int32_t n_cams, const float* fb_delay, ses3d_person_cov* fused_dense,
                                 ses3d_person_cov* pred_dense, int64_t cap, int32_t* n_out, float* pred_delay, int64_t* total) {
  CU(wait_for_device_runs(h, h->stream));
  CU(cudaStreamSynchronize(h->stream));   // earlier padded / reset / device-buffer work
  const size_t rec = sizeof(ses3d_person_cov);
  if (fb_delay == nullptr) n_cams = 0;
  {  // validate everything before the first launch: a failed call must not advance the tracker state
    long long n_in_total = 0;
    for (size_t i = 0; i < (size_t)n_sequences * n_frames; ++i) n_in_total += std::min(std::max(n_persons[i], 0), h_max);
    if (n_in_total > 0 && !persons_dense) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_run_ragged: NULL persons_dense");
    // every published record stems from one input record, so cap >= the input count can never overflow; a smaller
    // capacity is accepted and checked against the actual totals chunk by chunk (the state then has advanced: documented)
  }
  // chunks of streams (streams are independent): three slots keep H2D, the kernel and D2H of neighbouring chunks busy.
  // The kernel walks the messages of a stream one after the other, so a launch takes about n_frames x the per-message
  // latency however few streams it holds: chunks stay large (B200, 2048 streams x 32 messages, ms per call:
  // 1 chunk 43.5, 8 chunks 58.8)
  int chunk = std::max(1, std::min(n_sequences, std::max(592, (n_sequences + 2) / 3)));
  if (h->ragged_chunk_env > 0) chunk = std::max(1, std::min(n_sequences, h->ragged_chunk_env));
  long long in_done = 0, out_done = 0;
  struct Pending { int slot; size_t m0, n_msg; bool active; } prev{0, 0, 0, false};
  int status = SES3D_OK;
  auto finish = [&](const Pending& pd) -> int {   // total known -> copy the dense results out
    PriorSlot& sl = h->rs[pd.slot];
    CU(cudaEventSynchronize(sl.done));
    const long long t = *sl.total;
    std::memcpy(n_out + pd.m0, sl.h_stage, 4 * pd.n_msg);
    if (pred_delay) std::memcpy(pred_delay + pd.m0, sl.h_stage + 4 * pd.n_msg, 4 * pd.n_msg);
    if (out_done + t > cap) return ses3d::set_error(SES3D_E_CAPACITY, "ses3d_prior_run_ragged: output capacity too small");
    if (t) {
      CU(cudaMemcpyAsync(fused_dense + out_done, sl.dense_fused.p, rec * (size_t)t, cudaMemcpyDeviceToHost, sl.stream));
      CU(cudaMemcpyAsync(pred_dense + out_done, sl.dense_pred.p, rec * (size_t)t, cudaMemcpyDeviceToHost, sl.stream));
    }
    out_done += t;
    return SES3D_OK;
  };
  int ci = 0;
  for (int s0 = 0; s0 < n_sequences && status == SES3D_OK; s0 += chunk, ++ci) {
    const int ns = std::min(chunk, n_sequences - s0);
    PriorSlot& sl = h->rs[ci % ses3d_prior_s::kSlots];
    cudaStream_t st = sl.stream;
    const size_t m0 = (size_t)s0 * n_frames, n_msg = (size_t)ns * n_frames;
    long long n_in = 0;
    for (size_t i = m0; i < m0 + n_msg; ++i) n_in += std::min(std::max(n_persons[i], 0), h_max);
    CU(cudaEventSynchronize(sl.done));   // the slot's previous chunk (its copies included) has left the buffers
    CU(sl.persons.ensure(rec * n_msg * h_max));
    CU(sl.n.ensure(4 * n_msg));
    CU(sl.stamp.ensure(8 * n_msg));
    CU(sl.fused.ensure(rec * n_msg * h_max));
    CU(sl.pred.ensure(rec * n_msg * h_max));
    CU(sl.n_out.ensure(4 * n_msg));
    CU(sl.delay_out.ensure(4 * n_msg));
    CU(sl.dense_in.ensure(rec * (size_t)std::max<long long>(n_in, 1)));
    CU(sl.dense_fused.ensure(rec * (size_t)std::max<long long>(n_in, 1)));   // published <= fitted
    CU(sl.dense_pred.ensure(rec * (size_t)std::max<long long>(n_in, 1)));
    CU(sl.off_in.ensure(8 * (n_msg + 1)));
    CU(sl.off_out.ensure(8 * (n_msg + 1)));
    if (n_cams > 0) CU(sl.delay.ensure(4 * n_msg * n_cams));
    CU(sl.ensure_stage(8 * n_msg));
    CU(cudaMemcpyAsync(sl.n.p, n_persons + m0, 4 * n_msg, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(sl.stamp.p, stamp_ns + m0, 8 * n_msg, cudaMemcpyHostToDevice, st));
    if (n_in) CU(cudaMemcpyAsync(sl.dense_in.p, persons_dense + in_done, rec * (size_t)n_in, cudaMemcpyHostToDevice, st));
    in_done += n_in;
    if (n_cams > 0) CU(cudaMemcpyAsync(sl.delay.p, fb_delay + m0 * n_cams, 4 * n_msg * n_cams, cudaMemcpyHostToDevice, st));
    CU(ses3d::launch_scan_counts(sl.n.as<int32_t>(), (int)n_msg, h_max, sl.off_in.as<long long>(), nullptr, st));
    CU(ses3d::launch_move_records(1, (int)n_msg, h_max, (int)rec, sl.n.as<int32_t>(), sl.off_in.as<long long>(),
                                  sl.persons.p, sl.dense_in.p, -1, st));
    CU(ses3d::launch_prior(h->pt, ns, n_frames, h_max, h->max_tracks, h->states.as<ses3d::PriorSeqState>() + s0,
                           h->tracks.as<ses3d::PriorTrack>() + (size_t)s0 * h->max_tracks,
                           h->order.as<uint8_t>() + (size_t)s0 * h->max_tracks, sl.persons.as<ses3d_person_cov>(),
                           sl.n.as<int32_t>(), sl.stamp.as<int64_t>(), n_cams, n_cams > 0 ? sl.delay.as<float>() : nullptr,
                           sl.fused.as<ses3d_person_cov>(), sl.pred.as<ses3d_person_cov>(), sl.n_out.as<int32_t>(),
                           sl.delay_out.as<float>(), nullptr, st));
    CU(ses3d::launch_scan_counts(sl.n_out.as<int32_t>(), (int)n_msg, h_max, sl.off_out.as<long long>(), nullptr, st));
    CU(ses3d::launch_move_records(0, (int)n_msg, h_max, (int)rec, sl.n_out.as<int32_t>(), sl.off_out.as<long long>(),
                                  sl.fused.p, sl.dense_fused.p, -1, st));
    CU(ses3d::launch_move_records(0, (int)n_msg, h_max, (int)rec, sl.n_out.as<int32_t>(), sl.off_out.as<long long>(),
                                  sl.pred.p, sl.dense_pred.p, -1, st));
    h->launches += 6;
    CU(cudaMemcpyAsync(sl.total, sl.off_out.as<long long>() + n_msg, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sl.h_stage, sl.n_out.p, 4 * n_msg, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sl.h_stage + 4 * n_msg, sl.delay_out.p, 4 * n_msg, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(sl.done, st));
    if (prev.active) status = finish(prev);   // overlaps with the chunk just enqueued
    prev = Pending{ci % ses3d_prior_s::kSlots, m0, n_msg, true};
  }
  if (status == SES3D_OK && prev.active) status = finish(prev);
  for (PriorSlot& sl : h->rs) {
    CU(cudaEventRecord(sl.done, sl.stream));   // covers the result copies: the slot is free once this has passed
    CU(cudaStreamSynchronize(sl.stream));
  }
  if (status != SES3D_OK) return status;
  std::vector<ses3d::PriorSeqState> hs(n_sequences);
  CU(cudaMemcpy(hs.data(), h->states.p, sizeof(ses3d::PriorSeqState) * n_sequences, cudaMemcpyDeviceToHost));
  for (const auto& st : hs)
    if (st.overflow) return ses3d::set_error(SES3D_E_CAPACITY, "ses3d_prior_run_ragged: a stream needed more than max_tracks tracks");
  *total = out_done;
  return SES3D_OK;
}

extern "C" {

int ses3d_prior_get_tracks(ses3d_prior h, int32_t sequence, int32_t* ids, int32_t* num_obs) {
  if (!h || sequence < 0 || sequence >= h->n_sequences) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_get_tracks: bad argument");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  CU(wait_for_device_runs(h, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  ses3d::PriorSeqState s;
  CU(cudaMemcpy(&s, h->states.as<ses3d::PriorSeqState>() + sequence, sizeof s, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> ord(h->max_tracks);
  CU(cudaMemcpy(ord.data(), h->order.as<uint8_t>() + (size_t)sequence * h->max_tracks, h->max_tracks, cudaMemcpyDeviceToHost));
  for (int i = 0; i < s.n_tracks; ++i) {
    ses3d::PriorTrack t;
    CU(cudaMemcpy(&t, h->tracks.as<ses3d::PriorTrack>() + (size_t)sequence * h->max_tracks + ord[i], sizeof t, cudaMemcpyDeviceToHost));
    if (ids) ids[i] = t.id;
    if (num_obs) num_obs[i] = t.num_obs;
  }
  if (s.overflow) return ses3d::set_error(SES3D_E_CAPACITY, "ses3d_prior_get_tracks: the stream overflowed max_tracks");
  return s.n_tracks;
}

int64_t ses3d_prior_launch_count(ses3d_prior h) { return h ? h->launches : 0; }

int ses3d_prior_last_kernel_ms(ses3d_prior h, float* ms) {
  if (!h || !ms) return ses3d::set_error(SES3D_E_INVALID, "ses3d_prior_last_kernel_ms: bad argument");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  CU(cudaEventSynchronize(h->ev1));
  CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return SES3D_OK;
}

}  // extern "C"
