// api.cpp — the C ABI of include/ses3d.h on top of the sm_100a kernels.
//
// ses3d_create            <- main() set-up of skeleton_3d / pose_reprojection (S3D:1184-1214, REP:272-279)
// ses3d_triangulate_batch <- triangulate_persons (S3D:525-997), batched over frames
// ses3d_reproject_batch   <- fusedSkeletonCallback (REP:139-235), batched over frames
// ses3d_process_batch     <- both, chained on the device
//
// There is no CPU fallback: every entry point that computes needs a CUDA device and fails
// with SES3D_E_CUDA otherwise. Host-buffer calls stream the batch through the handle's device slots
// (H2D, kernels and D2H of neighbouring chunks overlap; single-frame calls replay a captured CUDA graph);
// device-buffer calls run in place and are stream-ordered.
// Device scratch belongs to the handle, grows monotonically and is reused across calls.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "host_setup.h"
#include "launch.h"
#include "ses3d.h"

namespace {

thread_local std::string g_last_error = "";

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what) {
  return fail(SES3D_E_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define CU(call)                                         \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

}  // namespace
namespace ses3d {   // shared with prior_api.cpp
int set_error(int code, const std::string& msg) { return fail(code, msg); }
const char* last_error() { return g_last_error.c_str(); }
}  // namespace ses3d
namespace {

// Bumped whenever a device buffer moves: captured single-frame graphs hold raw device addresses and are rebuilt
// when the generation they were captured under is no longer current.
std::atomic<uint64_t> g_alloc_gen{1};

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_alloc_gen.fetch_add(1, std::memory_order_relaxed);
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) { cudaFree(p); g_alloc_gen.fetch_add(1, std::memory_order_relaxed); }
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

struct Scratch {  // per-slot intermediates of the triangulation path
  DevBuf hyp_det, n_hyp, n_hung, keep, tmp, nk, work, work_count, pairs, far, meta;
  void release() {
    hyp_det.release(); n_hyp.release(); n_hung.release(); keep.release(); tmp.release(); nk.release(); work.release();
    work_count.release(); pairs.release(); far.release(); meta.release();
  }
};

struct Slot {  // one in-flight chunk of a host-buffer call
  cudaStream_t stream = nullptr;
  Scratch sc;
  DevBuf persons, n_persons, out3d, n_out3d, out2d, n_out2d, hyp_of, dump_nhyp, dump_nhung;
  DevBuf in_dense, in_off, off3, off2, c3d, c2d;  // ragged calls: dense staging + offsets
  long long* totals = nullptr;                    // pinned host: {total3d, total2d} of the chunk in flight
  cudaEvent_t done = nullptr;
  // ragged host calls: the upload of a chunk runs on its own stream so that it only waits for the kernels that read
  // the slot's input buffers last (kern_done), not for the download of that older chunk's results
  cudaStream_t up = nullptr;
  cudaEvent_t up_done = nullptr, kern_done = nullptr, down_done = nullptr;
  void release() {
    if (up_done) cudaEventDestroy(up_done);
    if (kern_done) cudaEventDestroy(kern_done);
    if (down_done) cudaEventDestroy(down_done);
    up_done = kern_done = down_done = nullptr;
    if (up) cudaStreamDestroy(up);
    up = nullptr;
    sc.release(); persons.release(); n_persons.release(); out3d.release(); n_out3d.release(); out2d.release();
    n_out2d.release(); hyp_of.release(); dump_nhyp.release(); dump_nhung.release(); in_dense.release(); in_off.release(); off3.release(); off2.release();
    c3d.release(); c2d.release();
    if (totals) cudaFreeHost(totals);
    totals = nullptr;
    if (done) cudaEventDestroy(done);
    done = nullptr;
    if (stream) cudaStreamDestroy(stream);
    stream = nullptr;
  }
};

}  // namespace

// Single-frame host calls (the ROS nodes: one message set per call) replay a captured CUDA graph: input copy,
// the kernel chain, output copies and the overflow word in ONE launch, staged through pinned memory the handle owns.
struct FrameGraph {
  int p_max = 0, h_max = 0;
  uint64_t alloc_gen = 0;        // g_alloc_gen at capture
  cudaGraphExec_t exec = nullptr;
  int n_kernels = 0;             // kernel launches one replay stands for
  bool broken = false;           // capture failed once: stay on the eager path
  // one contiguous record [persons | n_persons | 3-D | n_3d | overflow word | 2-D | n_2d] in pinned host memory and,
  // mirrored, in device memory: one upload, one clear and one download per replay whatever the stage mask
  unsigned char* pin = nullptr;
  size_t pin_bytes = 0;
  DevBuf dev;
  void release() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    dev.release();
    if (pin) cudaFreeHost(pin);
    pin = nullptr;
    pin_bytes = 0;
  }
};

struct ses3d_handle_s {
  int device = 0;
  ses3d_params prm;
  ses3d::HostTables host;
  DevBuf d_camf, d_camd, d_F, d_frow, d_overflow;
  ses3d::Tables tb;
  static constexpr int kSlots = 5;   // chunks in flight: the upload runs ahead of the kernels, the download trails them
  Slot slot[kSlots];
  cudaEvent_t fork_ev = nullptr;     // device-buffer calls: the caller's stream forks into the slot streams
  int device_split = 2;              // sub-batches of a device-buffer call that run on concurrent streams
  ses3d::LaunchCfg cfg;              // SM count + tuning overrides, fixed at create
  int ragged_chunk_env = 0;          // SES3D_RAGGED_CHUNK
  // SES3D_RAGGED_DIRECT: -1 (default) = pack kernels write straight into the caller's buffers only when those are in
  // device memory; 1 = also into pinned host memory (posted PCIe writes from the SMs: measured 17.0 ms against 10.6 ms
  // for the copy-engine path per 16 384 frames of hall16 x 6, profiles/r02_e2e_timeline.json); 0 = always stage
  int ragged_direct = -1;
  // Ragged host calls: all uploads share one stream and all result downloads another (one copy per direction in
  // flight; the slot streams carry the kernels). B200, 16 384 frames of hall16 x 6, two boxes: 10.11 / 10.39 ms with
  // per-slot copy streams and 1024-frame chunks, 9.86 / ~10.1 ms with shared copy streams and 1536-frame chunks
  int ragged_one_down = 1;           // SES3D_RAGGED_ONE_DOWN=0: downloads on the slot streams
  cudaStream_t down = nullptr;
  int ragged_one_up = 1;             // SES3D_RAGGED_ONE_UP=0: uploads on the slots' own upload streams
  int ragged_slots = kSlots;         // SES3D_RAGGED_SLOTS: chunks in flight of a ragged host call (2..kSlots)
  // ragged calls: running output totals {3-D records, 2-D records} live on the device and are carried from chunk to
  // chunk by the scan kernels; scan_ev orders the scans of consecutive chunks across the slot streams
  DevBuf d_run;
  cudaEvent_t scan_ev[kSlots] = {};
  // pinned host words: [0..1] totals of the last ragged call, [2] overflow flag snapshot of the last device-buffer call
  long long* h_words = nullptr;
  // device-buffer calls return without synchronising; dev_done marks the end of the last one on its stream
  cudaEvent_t dev_done = nullptr;
  bool dev_pending = false;
  std::mutex mu;
  int64_t launches = 0;
  bool profiling = false;
  float kernel_ms[4] = {0, 0, 0, 0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending_events;
  FrameGraph fgraph[3];              // by stage mask - 1: triangulate, reproject, both
  int frame_graph = 1;               // SES3D_FRAME_GRAPH=0: single-frame host calls take the eager path
};

namespace {

using ses3d::LaunchDims;

const int kDeviceChunk = 16384;  // frames per kernel launch on the device path (bounds scratch)

// frames per launch for a given rig: the per-frame pair table (n(n-1)/2 doubles, n = C*p_max) is the largest
// scratch item; keep it under ~4 GiB per slot
int device_chunk(int n_cams, int p_max) {
  const size_t per_frame = ses3d::associate_pair_table_bytes(n_cams, p_max);
  const size_t fit = std::max<size_t>(1, ((size_t)4 << 30) / std::max<size_t>(per_frame, 1));
  return (int)std::min<size_t>(kDeviceChunk, fit);
}

struct ProfScope {  // optional CUDA-event bracket around one kernel launch
  ses3d_handle_s* h;
  int idx;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(ses3d_handle_s* h_, int idx_, cudaStream_t st_) : h(h_), idx(idx_), st(st_) {
    if (h->profiling) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~ProfScope() {
    if (h->profiling) {
      cudaEventRecord(b, st);
      h->pending_events.push_back({idx, {a, b}});
    }
  }
};

void resolve_events(ses3d_handle_s* h) {
  for (auto& pe : h->pending_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, pe.second.first, pe.second.second) == cudaSuccess) h->kernel_ms[pe.first] += ms;
    cudaEventDestroy(pe.second.first);
    cudaEventDestroy(pe.second.second);
  }
  h->pending_events.clear();
}

// K2a -> K2b -> K3 -> K4 on device pointers, stream-ordered, no synchronisation. With out2d != nullptr the last step
// is the fused K4 + K6 kernel (finalize + reproject): four launches per device chunk for the whole path.
int triangulate_on_device(ses3d_handle_s* h, Scratch& sc, cudaStream_t st, int n_frames, int p_max, int h_max,
                          const ses3d_person2d* persons, const int32_t* n_persons, ses3d_person_cov* out,
                          int32_t* n_out, int32_t* hyp_of, int32_t* n_hyp_dump, int32_t* n_hung_dump,
                          ses3d_person2d* out2d = nullptr, int32_t* n_out2d = nullptr, int32_t* overflow = nullptr) {
  const int C = h->tb.n_cams;
  if (!overflow) overflow = h->d_overflow.as<int32_t>();   // the handle's sticky flag unless the caller keeps its own
  bool need_nk = false;
  ses3d::associate_smem_bytes(C, p_max, h_max, &need_nk);
  const int dchunk = device_chunk(C, p_max);
  for (int f0 = 0; f0 < n_frames; f0 += dchunk) {
    const int nf = std::min(dchunk, n_frames - f0);
    CU(sc.pairs.ensure((size_t)nf * ses3d::associate_pair_table_bytes(C, p_max)));
    CU(sc.meta.ensure((size_t)nf * ses3d::associate_meta_bytes(C, p_max)));
    CU(sc.hyp_det.ensure((size_t)nf * h_max * C));
    CU(sc.n_hyp.ensure((size_t)nf * 4));
    CU(sc.n_hung.ensure((size_t)nf * 4));
    CU(sc.keep.ensure((size_t)nf * h_max * 4));
    CU(sc.tmp.ensure((size_t)nf * h_max * sizeof(ses3d_person_cov)));
    CU(sc.work.ensure((size_t)nf * h_max * 4 * ses3d::kTriBuckets));
    CU(sc.work_count.ensure(64));
    if (h->prm.precision == SES3D_PRECISION_FP32) CU(sc.far.ensure(ses3d::triangulate_far_scratch_bytes(h->cfg)));
    if (need_nk) CU(sc.nk.ensure((size_t)nf * C * p_max * ses3d::NKP * 3 * sizeof(float)));
    LaunchDims d{nf, p_max, h_max};
    const ses3d_person2d* pin = persons + (size_t)f0 * C * p_max;
    const int32_t* nin = n_persons + (size_t)f0 * C;
    int32_t* n_hyp = n_hyp_dump ? n_hyp_dump + f0 : sc.n_hyp.as<int32_t>();
    int32_t* n_hung = n_hung_dump ? n_hung_dump + f0 : sc.n_hung.as<int32_t>();
    {
      ProfScope ps(h, 0, st);
      CU(ses3d::launch_associate(h->cfg, h->tb, d, pin, nin, need_nk ? sc.nk.as<float>() : nullptr, sc.pairs.as<double>(),
                                 sc.meta.as<unsigned char>(), sc.hyp_det.as<int8_t>(),
                                 n_hyp, n_hung, overflow,
                                 hyp_of ? hyp_of + (size_t)f0 * C * p_max : nullptr, sc.keep.as<int32_t>(),
                                 sc.work.as<uint32_t>(), sc.work_count.as<int32_t>(), st));
    }
    {
      ProfScope ps(h, 1, st);
      CU(ses3d::launch_triangulate(h->cfg, h->tb, d, pin, sc.hyp_det.as<int8_t>(), sc.work.as<uint32_t>(),
                                   sc.work_count.as<int32_t>(), sc.tmp.as<ses3d_person_cov>(), sc.keep.as<int32_t>(),
                                   sc.far.as<float>(), sc.far.cap, st));
    }
    if (out2d) {
      ProfScope ps(h, 3, st);
      CU(ses3d::launch_finproj(h->cfg, h->tb, d, n_hyp, sc.tmp.as<ses3d_person_cov>(), sc.keep.as<int32_t>(),
                               out + (size_t)f0 * h_max, n_out + f0, out2d + (size_t)f0 * C * h_max,
                               n_out2d + (size_t)f0 * C, st));
    } else {
      ProfScope ps(h, 2, st);
      CU(ses3d::launch_finalize(h->tb, d, n_hyp, sc.tmp.as<ses3d_person_cov>(), sc.keep.as<int32_t>(),
                                out + (size_t)f0 * h_max, n_out + f0, st));
    }
    h->launches += 2 + ses3d::associate_launches(h->cfg, p_max);   // K2a (+ dense instance), K2b, K3, K4 / K4+K6
  }
  return SES3D_OK;
}

int reproject_on_device(ses3d_handle_s* h, cudaStream_t st, int n_frames, int h_max, const ses3d_person_cov* persons3d,
                        const int32_t* n_persons3d, ses3d_person2d* out, int32_t* n_out) {
  const int C = h->tb.n_cams;
  for (int f0 = 0; f0 < n_frames; f0 += kDeviceChunk) {
    const int nf = std::min(kDeviceChunk, n_frames - f0);
    ProfScope ps(h, 3, st);
    CU(ses3d::launch_reproject(h->cfg, h->tb, nf, h_max, persons3d + (size_t)f0 * h_max, n_persons3d + f0,
                               out + (size_t)f0 * C * h_max, n_out + (size_t)f0 * C, st));
    h->launches += 1;
  }
  return SES3D_OK;
}

// The kernels only ever set the flag; it is cleared at create and again after it has been reported, so the
// batch calls need no reset (and no extra synchronisation) on entry.
int overflow_error(ses3d_handle_s* h) {
  cudaMemset(h->d_overflow.p, 0, 4);
  return fail(SES3D_E_CAPACITY, "a frame produced more hypotheses than h_max");
}

// Device-buffer calls are stream-ordered: they enqueue their kernels plus a 4-byte snapshot of the sticky overflow
// flag into pinned memory and return. collect_device_status(wait = true) blocks until that work has finished
// (ses3d_check); with wait = false it only looks (start of the next call) - either way a pending capacity overflow is
// reported exactly once.
int collect_device_status(ses3d_handle_s* h, bool wait) {
  if (!h->dev_pending) return SES3D_OK;
  if (wait) {
    CU(cudaEventSynchronize(h->dev_done));
  } else {
    const cudaError_t q = cudaEventQuery(h->dev_done);
    if (q == cudaErrorNotReady) return SES3D_OK;
    if (q != cudaSuccess) return cuda_fail(q, "cudaEventQuery(dev_done)");
  }
  h->dev_pending = false;
  if (h->h_words[2]) {
    h->h_words[2] = 0;
    return overflow_error(h);
  }
  return SES3D_OK;
}

// Work that is about to be enqueued on `st` shares the handle's scratch with the last device-buffer call: order it
// behind that call on the device (no host wait).
cudaError_t order_after_device_work(ses3d_handle_s* h, cudaStream_t st) {
  if (!h->dev_pending) return cudaSuccess;
  return cudaStreamWaitEvent(st, h->dev_done, 0);
}

int check_dims(const ses3d_handle_s* h, int n_frames, int p_max, int h_max) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  if (n_frames < 0) return fail(SES3D_E_INVALID, "n_frames < 0");
  if (p_max < 1 || p_max > 127) return fail(SES3D_E_INVALID, "p_max must be in [1,127]");
  if (h_max < 1 || h_max > 1024) return fail(SES3D_E_INVALID, "h_max must be in [1,1024]");
  return SES3D_OK;
}

int host_chunk_frames(int n_frames) { return std::max(1, std::min(8192, std::max(1024, (n_frames + 5) / 6))); }

enum Stage { TRI = 1, REP = 2 };

int run_batch_locked(ses3d_handle_s* h, int stages, int n_frames, int p_max, const ses3d_person2d* persons,
                     const int32_t* n_persons, int h_max, ses3d_person_cov* io3d, int32_t* n_io3d, ses3d_person2d* out2d,
                     int32_t* n_out2d, const ses3d_assoc_dump* dump, uint32_t flags, void* stream);

// One chunk of a host-buffer call on its slot: zero the padded outputs, upload, the kernel chain, download. Everything
// is asynchronous on the slot's stream (and therefore capturable into a graph once the slot's buffers are big enough).
struct HostChunk {
  const ses3d_person2d* persons = nullptr;
  const int32_t* n_persons = nullptr;
  ses3d_person_cov* io3d = nullptr;
  int32_t* n_io3d = nullptr;
  ses3d_person2d* out2d = nullptr;
  int32_t* n_out2d = nullptr;
  int32_t *hyp_of = nullptr, *n_hyp = nullptr, *n_hung = nullptr;
};

int enqueue_host_chunk(ses3d_handle_s* h, int stages, Slot& s, int nf, int p_max, int h_max, const HostChunk& hc) {
  const int C = h->tb.n_cams;
  cudaStream_t st = s.stream;
  CU(s.out3d.ensure((size_t)nf * h_max * sizeof(ses3d_person_cov)));
  CU(s.n_out3d.ensure((size_t)nf * 4));
  // padded outputs travel whole: unused slots are zero, never stale device memory
  if ((stages & TRI) && hc.io3d) CU(cudaMemsetAsync(s.out3d.p, 0, (size_t)nf * h_max * sizeof(ses3d_person_cov), st));
  const bool fused = (stages & TRI) && (stages & REP);
  if (stages & REP) {
    CU(s.out2d.ensure((size_t)nf * C * h_max * sizeof(ses3d_person2d)));
    CU(s.n_out2d.ensure((size_t)nf * C * 4));
    CU(cudaMemsetAsync(s.out2d.p, 0, (size_t)nf * C * h_max * sizeof(ses3d_person2d), st));
  }
  if (stages & TRI) {
    CU(s.persons.ensure((size_t)nf * C * p_max * sizeof(ses3d_person2d)));
    CU(s.n_persons.ensure((size_t)nf * C * 4));
    CU(cudaMemcpyAsync(s.persons.p, hc.persons, (size_t)nf * C * p_max * sizeof(ses3d_person2d), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.n_persons.p, hc.n_persons, (size_t)nf * C * 4, cudaMemcpyHostToDevice, st));
    int32_t* d_hyp_of = nullptr;
    if (hc.hyp_of) {
      CU(s.hyp_of.ensure((size_t)nf * C * p_max * 4));
      d_hyp_of = s.hyp_of.as<int32_t>();
    }
    int32_t *d_nhyp = nullptr, *d_nhung = nullptr;   // a host chunk may span several device chunks
    if (hc.n_hyp) { CU(s.dump_nhyp.ensure((size_t)nf * 4)); d_nhyp = s.dump_nhyp.as<int32_t>(); }
    if (hc.n_hung) { CU(s.dump_nhung.ensure((size_t)nf * 4)); d_nhung = s.dump_nhung.as<int32_t>(); }
    int rc = triangulate_on_device(h, s.sc, st, nf, p_max, h_max, s.persons.as<ses3d_person2d>(),
                                   s.n_persons.as<int32_t>(), s.out3d.as<ses3d_person_cov>(),
                                   s.n_out3d.as<int32_t>(), d_hyp_of, d_nhyp, d_nhung,
                                   fused ? s.out2d.as<ses3d_person2d>() : nullptr,
                                   fused ? s.n_out2d.as<int32_t>() : nullptr);
    if (rc) return rc;
    if (hc.io3d) CU(cudaMemcpyAsync(hc.io3d, s.out3d.p, (size_t)nf * h_max * sizeof(ses3d_person_cov), cudaMemcpyDeviceToHost, st));
    if (hc.n_io3d) CU(cudaMemcpyAsync(hc.n_io3d, s.n_out3d.p, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    if (hc.hyp_of) CU(cudaMemcpyAsync(hc.hyp_of, d_hyp_of, (size_t)nf * C * p_max * 4, cudaMemcpyDeviceToHost, st));
    if (hc.n_hyp) CU(cudaMemcpyAsync(hc.n_hyp, d_nhyp, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    if (hc.n_hung) CU(cudaMemcpyAsync(hc.n_hung, d_nhung, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
  } else {
    CU(cudaMemcpyAsync(s.out3d.p, hc.io3d, (size_t)nf * h_max * sizeof(ses3d_person_cov), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.n_out3d.p, hc.n_io3d, (size_t)nf * 4, cudaMemcpyHostToDevice, st));
  }
  if (stages & REP) {
    if (!fused) {
      int rc = reproject_on_device(h, st, nf, h_max, s.out3d.as<ses3d_person_cov>(), s.n_out3d.as<int32_t>(),
                                   s.out2d.as<ses3d_person2d>(), s.n_out2d.as<int32_t>());
      if (rc) return rc;
    }
    CU(cudaMemcpyAsync(hc.out2d, s.out2d.p, (size_t)nf * C * h_max * sizeof(ses3d_person2d), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hc.n_out2d, s.n_out2d.p, (size_t)nf * C * 4, cudaMemcpyDeviceToHost, st));
  }
  return SES3D_OK;
}

// Single-frame host call through a captured graph. *done = false means "not handled, take the eager path" (first
// call of a shape: the eager path sizes the slot's scratch, the graph is captured right after it for the next call).
struct FrameStage {   // byte offsets into FrameGraph::pin / FrameGraph::dev
  size_t persons, n_persons, io3d, n_io3d, flag, out2d, n_out2d, total;
};
FrameStage frame_stage_layout(int C, int p_max, int h_max) {
  FrameStage o;
  size_t at = 0;
  auto take = [&at](size_t bytes) { const size_t r = at; at += (bytes + 255) & ~(size_t)255; return r; };
  o.persons = take((size_t)C * p_max * sizeof(ses3d_person2d));
  o.n_persons = take((size_t)C * 4);
  o.io3d = take((size_t)h_max * sizeof(ses3d_person_cov));
  o.n_io3d = take(4);
  o.flag = take(8);
  o.out2d = take((size_t)C * h_max * sizeof(ses3d_person2d));
  o.n_out2d = take((size_t)C * 4);
  o.total = at;
  return o;
}

int capture_frame_graph(ses3d_handle_s* h, int stages, int p_max, int h_max) {
  FrameGraph& g = h->fgraph[stages - 1];
  if (g.broken) return SES3D_OK;
  const int C = h->tb.n_cams;
  const FrameStage o = frame_stage_layout(C, p_max, h_max);
  if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  if (g.pin_bytes < o.total) {
    if (g.pin) cudaFreeHost(g.pin);
    g.pin = nullptr;
    g.pin_bytes = 0;
    if (cudaMallocHost(reinterpret_cast<void**>(&g.pin), o.total) != cudaSuccess) {
      cudaGetLastError();
      g.broken = true;
      return SES3D_OK;
    }
    g.pin_bytes = o.total;
  }
  if (g.dev.ensure(o.total) != cudaSuccess) {
    cudaGetLastError();
    g.broken = true;
    return SES3D_OK;
  }
  Slot& s = h->slot[0];
  cudaStream_t st = s.stream;
  unsigned char* d = g.dev.as<unsigned char>();
  auto at = [d](size_t off) { return d + off; };
  const uint64_t gen0 = g_alloc_gen.load(std::memory_order_relaxed);
  const int64_t launches0 = h->launches;
  const std::string err0 = g_last_error;
  const bool tri = (stages & TRI) != 0, rep = (stages & REP) != 0;
  // upload [up0, up1), clear [z0, z1), download [dn0, dn1)
  const size_t up0 = tri ? o.persons : o.io3d, up1 = tri ? o.io3d : o.out2d;   // reproject alone: 3-D + count + zero flag
  const size_t z0 = tri ? o.io3d : o.out2d, z1 = rep ? o.total : o.out2d;
  const size_t dn0 = z0, dn1 = z1;
  cudaGraph_t graph = nullptr;
  bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    ok = cudaMemsetAsync(at(z0), 0, z1 - z0, st) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(at(up0), g.pin + up0, up1 - up0, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok && tri)
      ok = triangulate_on_device(h, s.sc, st, 1, p_max, h_max, reinterpret_cast<const ses3d_person2d*>(at(o.persons)),
                                 reinterpret_cast<const int32_t*>(at(o.n_persons)),
                                 reinterpret_cast<ses3d_person_cov*>(at(o.io3d)), reinterpret_cast<int32_t*>(at(o.n_io3d)),
                                 nullptr, nullptr, nullptr, rep ? reinterpret_cast<ses3d_person2d*>(at(o.out2d)) : nullptr,
                                 rep ? reinterpret_cast<int32_t*>(at(o.n_out2d)) : nullptr,
                                 reinterpret_cast<int32_t*>(at(o.flag))) == SES3D_OK;
    if (ok && rep && !tri)
      ok = reproject_on_device(h, st, 1, h_max, reinterpret_cast<const ses3d_person_cov*>(at(o.io3d)),
                               reinterpret_cast<const int32_t*>(at(o.n_io3d)),
                               reinterpret_cast<ses3d_person2d*>(at(o.out2d)),
                               reinterpret_cast<int32_t*>(at(o.n_out2d))) == SES3D_OK;
    ok = ok && cudaMemcpyAsync(g.pin + dn0, at(dn0), dn1 - dn0, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    const cudaError_t ee = cudaStreamEndCapture(st, &graph);   // always ends the capture, also after a failure
    ok = ok && ee == cudaSuccess && graph != nullptr;
  }
  g.n_kernels = (int)(h->launches - launches0);
  h->launches = launches0;   // nothing ran
  // a device buffer of the library moved while capturing (the counter is process-wide: another handle growing its
  // buffers on another thread also lands here): drop this capture, the next call tries again
  const bool moved = g_alloc_gen.load(std::memory_order_relaxed) != gen0;
  if (ok && !moved) ok = cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess;
  if (graph) cudaGraphDestroy(graph);
  if (!ok || moved) {
    cudaGetLastError();
    g_last_error = err0;
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (!ok) g.broken = true;
    return SES3D_OK;
  }
  g.p_max = p_max;
  g.h_max = h_max;
  g.alloc_gen = gen0;
  return SES3D_OK;
}

int run_frame_graph(ses3d_handle_s* h, int stages, int p_max, int h_max, const ses3d_person2d* persons,
                    const int32_t* n_persons, ses3d_person_cov* io3d, int32_t* n_io3d, ses3d_person2d* out2d,
                    int32_t* n_out2d, bool* done) {
  *done = false;
  FrameGraph& g = h->fgraph[stages - 1];
  if (g.broken) return SES3D_OK;
  const int C = h->tb.n_cams;
  const bool hit = g.exec && g.p_max == p_max && g.h_max == h_max &&
                   g.alloc_gen == g_alloc_gen.load(std::memory_order_relaxed);
  if (!hit) {
    // eager call now (sizes the buffers, produces this call's results), capture for the next one
    h->frame_graph = 0;
    const int rc = run_batch_locked(h, stages, 1, p_max, persons, n_persons, h_max, io3d, n_io3d, out2d, n_out2d, nullptr, 0,
                                    nullptr);
    h->frame_graph = 1;
    *done = true;
    if (rc) return rc;
    return capture_frame_graph(h, stages, p_max, h_max);
  }
  const FrameStage o = frame_stage_layout(C, p_max, h_max);
  Slot& s = h->slot[0];
  if (stages & TRI) {
    std::memcpy(g.pin + o.persons, persons, (size_t)C * p_max * sizeof(ses3d_person2d));
    std::memcpy(g.pin + o.n_persons, n_persons, (size_t)C * 4);
  } else {
    std::memcpy(g.pin + o.io3d, io3d, (size_t)h_max * sizeof(ses3d_person_cov));
    std::memcpy(g.pin + o.n_io3d, n_io3d, 4);
  }
  *reinterpret_cast<long long*>(g.pin + o.flag) = 0;
  CU(order_after_device_work(h, s.stream));
  CU(cudaGraphLaunch(g.exec, s.stream));
  h->launches += g.n_kernels;
  CU(cudaStreamSynchronize(s.stream));
  *done = true;
  if (*reinterpret_cast<const int32_t*>(g.pin + o.flag)) return overflow_error(h);
  if (stages & TRI) {
    if (io3d) std::memcpy(io3d, g.pin + o.io3d, (size_t)h_max * sizeof(ses3d_person_cov));
    if (n_io3d) std::memcpy(n_io3d, g.pin + o.n_io3d, 4);
  }
  if (stages & REP) {
    std::memcpy(out2d, g.pin + o.out2d, (size_t)C * h_max * sizeof(ses3d_person2d));
    std::memcpy(n_out2d, g.pin + o.n_out2d, (size_t)C * 4);
  }
  return SES3D_OK;
}

// Shared implementation of the three batch entry points.
int run_batch(ses3d_handle_s* h, int stages, int n_frames, int p_max, const ses3d_person2d* persons,
              const int32_t* n_persons, int h_max, ses3d_person_cov* io3d, int32_t* n_io3d, ses3d_person2d* out2d,
              int32_t* n_out2d, const ses3d_assoc_dump* dump, uint32_t flags, void* stream) {
  std::lock_guard<std::mutex> lock(h->mu);
  return run_batch_locked(h, stages, n_frames, p_max, persons, n_persons, h_max, io3d, n_io3d, out2d, n_out2d, dump, flags,
                          stream);
}

int run_batch_locked(ses3d_handle_s* h, int stages, int n_frames, int p_max, const ses3d_person2d* persons,
                     const int32_t* n_persons, int h_max, ses3d_person_cov* io3d, int32_t* n_io3d, ses3d_person2d* out2d,
                     int32_t* n_out2d, const ses3d_assoc_dump* dump, uint32_t flags, void* stream) {
  CU(cudaSetDevice(h->device));
  const int C = h->tb.n_cams;
  if (n_frames == 0) return SES3D_OK;
  const bool dev = (flags & SES3D_DEVICE_BUFFERS) != 0;
  if (h->profiling) for (float& m : h->kernel_ms) m = 0.f;
  int32_t* hyp_of = dump ? dump->hyp_of : nullptr;
  int32_t* n_hyp_d = dump ? dump->n_hyp : nullptr;
  int32_t* n_hung_d = dump ? dump->n_hungarian : nullptr;

  {  // an overflow left behind by an earlier stream-ordered call is reported now (no waiting)
    const int rc0 = collect_device_status(h, false);
    if (rc0) return rc0;
  }
  if (dev) {
    // NULL = the legacy default stream, exactly as a NULL cudaStream_t means in the CUDA runtime: the kernels are
    // ordered with the caller's producers and consumers on that stream
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(order_after_device_work(h, st));
    // Large batches are cut into sub-batches that run the kernel chain on concurrent streams: the four kernels
    // stress different resources (FP32 pipe / FP64 pipe / shared-memory latency) and none fills an SM's issue
    // slots or its register / shared-memory budget alone, so CTAs of neighbouring stages co-reside and the
    // tail of one kernel overlaps the head of the next. Profiling runs serially so per-kernel times stay clean.
    int n_split = (h->profiling || n_frames < 4096) ? 1 : std::min(h->device_split, 3);
    if (n_split <= 1) {
      const bool fused = (stages & TRI) && (stages & REP);
      if (stages & TRI) {
        int rc = triangulate_on_device(h, h->slot[0].sc, st, n_frames, p_max, h_max, persons, n_persons, io3d, n_io3d,
                                       hyp_of, n_hyp_d, n_hung_d, fused ? out2d : nullptr, fused ? n_out2d : nullptr);
        if (rc) return rc;
      }
      if ((stages & REP) && !fused) {
        int rc = reproject_on_device(h, st, n_frames, h_max, io3d, n_io3d, out2d, n_out2d);
        if (rc) return rc;
      }
    } else {
      CU(cudaEventRecord(h->fork_ev, st));
      for (int i = 0; i < n_split; ++i) {
        Slot& s = h->slot[i];
        const int f0 = (int)((int64_t)n_frames * i / n_split), f1 = (int)((int64_t)n_frames * (i + 1) / n_split);
        const int nf = f1 - f0;
        if (s.stream != st) CU(cudaStreamWaitEvent(s.stream, h->fork_ev, 0));
        const bool fused = (stages & TRI) && (stages & REP);
        if (stages & TRI) {
          int rc = triangulate_on_device(h, s.sc, s.stream, nf, p_max, h_max, persons + (size_t)f0 * C * p_max,
                                         n_persons + (size_t)f0 * C, io3d + (size_t)f0 * h_max, n_io3d + f0,
                                         hyp_of ? hyp_of + (size_t)f0 * C * p_max : nullptr,
                                         n_hyp_d ? n_hyp_d + f0 : nullptr, n_hung_d ? n_hung_d + f0 : nullptr,
                                         fused ? out2d + (size_t)f0 * C * h_max : nullptr,
                                         fused ? n_out2d + (size_t)f0 * C : nullptr);
          if (rc) return rc;
        }
        if ((stages & REP) && !fused) {
          int rc = reproject_on_device(h, s.stream, nf, h_max, io3d + (size_t)f0 * h_max, n_io3d + f0,
                                       out2d + (size_t)f0 * C * h_max, n_out2d + (size_t)f0 * C);
          if (rc) return rc;
        }
        if (s.stream != st) CU(cudaEventRecord(s.done, s.stream));
      }
      for (int i = 0; i < n_split; ++i)
        if (h->slot[i].stream != st) CU(cudaStreamWaitEvent(st, h->slot[i].done, 0));
    }
    CU(cudaMemcpyAsync(&h->h_words[2], h->d_overflow.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(h->dev_done, st));
    h->dev_pending = true;
    if (h->profiling) {   // per-kernel timing needs the events resolved: profiling calls are synchronous
      const int rc1 = collect_device_status(h, true);
      resolve_events(h);
      return rc1;
    }
    return SES3D_OK;   // stream-ordered: no host synchronisation (ses3d_check / the next call report an overflow)
  }

  // host buffers
  if (n_frames == 1 && !dump && !h->profiling && h->frame_graph) {
    bool done = false;
    const int rc = run_frame_graph(h, stages, p_max, h_max, persons, n_persons, io3d, n_io3d, out2d, n_out2d, &done);
    if (rc || done) return rc;
  }
  // stream chunks through the slots
  for (Slot& sl : h->slot) CU(order_after_device_work(h, sl.stream));
  const int chunk = host_chunk_frames(n_frames);
  int ci = 0;
  for (int f0 = 0; f0 < n_frames; f0 += chunk, ++ci) {
    const int nf = std::min(chunk, n_frames - f0);
    HostChunk hc;
    hc.persons = persons ? persons + (size_t)f0 * C * p_max : nullptr;
    hc.n_persons = n_persons ? n_persons + (size_t)f0 * C : nullptr;
    hc.io3d = io3d ? io3d + (size_t)f0 * h_max : nullptr;
    hc.n_io3d = n_io3d ? n_io3d + f0 : nullptr;
    hc.out2d = out2d ? out2d + (size_t)f0 * C * h_max : nullptr;
    hc.n_out2d = n_out2d ? n_out2d + (size_t)f0 * C : nullptr;
    hc.hyp_of = hyp_of ? hyp_of + (size_t)f0 * C * p_max : nullptr;
    hc.n_hyp = n_hyp_d ? n_hyp_d + f0 : nullptr;
    hc.n_hung = n_hung_d ? n_hung_d + f0 : nullptr;
    const int rc = enqueue_host_chunk(h, stages, h->slot[ci % ses3d_handle_s::kSlots], nf, p_max, h_max, hc);
    if (rc) return rc;
  }
  // the sticky overflow flag rides at the end of every used stream (pinned word per slot): no extra blocking copy
  const int used = std::min(ci, (int)ses3d_handle_s::kSlots);
  for (int i = 0; i < used; ++i) {
    h->slot[i].totals[0] = 0;
    CU(cudaMemcpyAsync(&h->slot[i].totals[0], h->d_overflow.p, 4, cudaMemcpyDeviceToHost, h->slot[i].stream));
  }
  long long overflow = 0;
  for (int i = 0; i < used; ++i) {
    CU(cudaStreamSynchronize(h->slot[i].stream));
    overflow |= h->slot[i].totals[0];
  }
  resolve_events(h);
  if (overflow) return overflow_error(h);
  return SES3D_OK;
}


// Can a kernel on this device write to `p` directly? True for device memory and for pinned / registered host memory
// (unified addressing: the mapped device pointer is returned in *dev_ptr). Pageable host memory -> false.
bool device_accessible(const void* p, void** dev_ptr) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if ((at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged) && at.devicePointer) {
    *dev_ptr = at.devicePointer;
    return true;
  }
  return false;
}

// Ragged variant of process_batch (host or device buffers): dense records in, dense records out.
//
// Direct mode (outputs in device memory; opt-in for pinned host memory): the running output offsets stay on the
// device - every chunk's scan kernels start from the totals the previous chunk left in d_run (ordered by one event per
// chunk) - and the pack kernels write the occupied records straight to their final position in the caller's buffers.
// No staging copy, no host round trip between chunks; the host synchronises once at the end. Staged mode (host
// outputs, the default): the dense results are packed on the device and leave through the copy engine once the host
// knows the chunk totals - SM-issued posted writes to pinned memory reach only ~60 % of the copy engine's PCIe rate.
int run_ragged_impl(ses3d_handle_s* h, int n_frames, int p_max, const ses3d_person2d* persons_dense,
                    const int32_t* n_persons, int h_max, ses3d_person_cov* out3d, long long cap3d, int32_t* n_out3d,
                    ses3d_person2d* out2d, long long cap2d, int32_t* n_out2d, long long* total3d, long long* total2d,
                    uint32_t flags);

int run_ragged(ses3d_handle_s* h, int n_frames, int p_max, const ses3d_person2d* persons_dense,
               const int32_t* n_persons, int h_max, ses3d_person_cov* out3d, long long cap3d, int32_t* n_out3d,
               ses3d_person2d* out2d, long long cap2d, int32_t* n_out2d, long long* total3d, long long* total2d,
               uint32_t flags) {
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  const int rc = run_ragged_impl(h, n_frames, p_max, persons_dense, n_persons, h_max, out3d, cap3d, n_out3d, out2d, cap2d,
                                 n_out2d, total3d, total2d, flags);
  if (rc != SES3D_OK) {   // nothing may still be writing into the caller's buffers when a failed call returns
    const std::string msg = g_last_error;
    for (Slot& sl : h->slot) { cudaStreamSynchronize(sl.up); cudaStreamSynchronize(sl.stream); }
    g_last_error = msg;
  }
  return rc;
}

int run_ragged_impl(ses3d_handle_s* h, int n_frames, int p_max, const ses3d_person2d* persons_dense,
                    const int32_t* n_persons, int h_max, ses3d_person_cov* out3d, long long cap3d, int32_t* n_out3d,
                    ses3d_person2d* out2d, long long cap2d, int32_t* n_out2d, long long* total3d, long long* total2d,
                    uint32_t flags) {
  const int C = h->tb.n_cams;
  *total3d = 0;
  *total2d = 0;
  {
    const int rc0 = collect_device_status(h, true);   // a ragged call is synchronous anyway
    if (rc0) return rc0;
  }
  if (n_frames == 0) return SES3D_OK;
  const bool dev = (flags & SES3D_DEVICE_BUFFERS) != 0;
  const cudaMemcpyKind in_kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (h->profiling) for (float& m : h->kernel_ms) m = 0.f;
  // per-frame input record counts are needed on the host to cut the dense input into chunks
  std::vector<int32_t> counts_host;
  const int32_t* counts = n_persons;
  if (dev) {
    counts_host.resize((size_t)n_frames * C);
    CU(cudaMemcpy(counts_host.data(), n_persons, counts_host.size() * 4, cudaMemcpyDeviceToHost));
    counts = counts_host.data();
  }
  void *d3 = out3d, *d2 = out2d;
  bool direct = dev ? h->ragged_direct != 0 : h->ragged_direct > 0;
  if (direct && !dev) direct = device_accessible(out3d, &d3) && device_accessible(out2d, &d2);
  // B200, 16384 frames of hall16 x 6 (ms per call, shared copy streams): chunk 768 -> 10.1, 1536 -> 9.86, 2048 -> 10.3
  int chunk = std::max(1, std::min(4096, std::max(512, (n_frames * 3 + 31) / 32)));
  if (h->ragged_chunk_env > 0) chunk = h->ragged_chunk_env;
  long long in_done = 0, run3 = 0, run2 = 0;
  // staged mode: chunks whose kernels are enqueued but whose packed results have not been sent home yet (oldest first)
  int pending[ses3d_handle_s::kSlots];
  int n_pending = 0;
  int status = SES3D_OK;

  auto finish = [&](int slot) -> int {  // staged mode: totals known -> copy the dense results out
    Slot& s = h->slot[slot];
    CU(cudaEventSynchronize(s.done));
    const long long t3 = s.totals[0], t2 = s.totals[1];
    if (run3 + t3 > cap3d || run2 + t2 > cap2d) return fail(SES3D_E_CAPACITY, "ragged output buffer too small");
    // (the host has seen s.done: the packed records are complete, a different stream may read them)
    cudaStream_t ds = h->ragged_one_down ? h->down : s.stream;
    if (t3) CU(cudaMemcpyAsync(out3d + run3, s.c3d.p, (size_t)t3 * sizeof(ses3d_person_cov), out_kind, ds));
    if (t2) CU(cudaMemcpyAsync(out2d + run2, s.c2d.p, (size_t)t2 * sizeof(ses3d_person2d), out_kind, ds));
    if (h->ragged_one_down) {
      CU(cudaEventRecord(s.down_done, ds));
      CU(cudaStreamWaitEvent(s.stream, s.down_done, 0));
    }
    run3 += t3;
    run2 += t2;
    return SES3D_OK;
  };
  auto finish_oldest = [&]() -> int {
    const int rc = finish(pending[0]);
    for (int i = 1; i < n_pending; ++i) pending[i - 1] = pending[i];
    --n_pending;
    return rc;
  };

  if (direct) {
    CU(h->d_run.ensure(16));
    CU(cudaMemsetAsync(h->d_run.p, 0, 16, h->slot[0].stream));
  }
  int ci = 0;
  const int n_slots = dev ? (int)ses3d_handle_s::kSlots : h->ragged_slots;
  // The first chunks are short (1/4, 1/2 of the regular size): the download - the longest leg of a host call - can
  // only start once the first chunk has been uploaded and processed, so a short first chunk shortens the pipeline fill.
  int nf = 0;
  for (int f0 = 0; f0 < n_frames && status == SES3D_OK; f0 += nf, ++ci) {
    const int ramp = (!dev && n_frames >= 4 * chunk) ? (ci == 0 ? chunk / 4 : ci == 1 ? chunk / 2 : chunk) : chunk;
    nf = std::min(std::max(ramp, 1), n_frames - f0);
    const int si = ci % n_slots;
    Slot& s = h->slot[si];
    cudaStream_t st = s.stream;
    long long n_in = 0;
    for (size_t i = (size_t)f0 * C; i < (size_t)(f0 + nf) * C; ++i) n_in += std::min(std::max(counts[i], 0), p_max);
    const size_t u_in = (size_t)nf * C;
    CU(s.persons.ensure(u_in * p_max * sizeof(ses3d_person2d)));
    CU(s.n_persons.ensure(u_in * 4));
    CU(s.in_dense.ensure((size_t)std::max<long long>(n_in, 1) * sizeof(ses3d_person2d)));
    CU(s.in_off.ensure((u_in + 1) * 8));
    CU(s.out3d.ensure((size_t)nf * h_max * sizeof(ses3d_person_cov)));
    CU(s.n_out3d.ensure((size_t)nf * 4));
    CU(s.out2d.ensure(u_in * h_max * sizeof(ses3d_person2d)));
    CU(s.n_out2d.ensure(u_in * 4));
    CU(s.off3.ensure(((size_t)nf + 1) * 8));
    CU(s.off2.ensure((u_in + 1) * 8));
    if (!direct) {
      CU(s.c3d.ensure((size_t)nf * h_max * sizeof(ses3d_person_cov)));
      CU(s.c2d.ensure(u_in * h_max * sizeof(ses3d_person2d)));
    }
    // upload on the slot's own upload stream: behind the kernels of the slot's previous chunk (they read these buffers),
    // ahead of everything else - the results of that older chunk may still be on their way to the host
    cudaStream_t up = h->ragged_one_up ? h->slot[0].up : s.up;
    CU(cudaStreamWaitEvent(up, s.kern_done, 0));
    CU(cudaMemcpyAsync(s.n_persons.p, n_persons + (size_t)f0 * C, u_in * 4, in_kind, up));
    if (n_in) CU(cudaMemcpyAsync(s.in_dense.p, persons_dense + in_done, (size_t)n_in * sizeof(ses3d_person2d), in_kind, up));
    CU(cudaEventRecord(s.up_done, up));
    CU(cudaStreamWaitEvent(st, s.up_done, 0));
    in_done += n_in;
    CU(ses3d::launch_scan_counts(s.n_persons.as<int32_t>(), (int)u_in, p_max, s.in_off.as<long long>(), nullptr, st));
    CU(ses3d::launch_move_records(1, (int)u_in, p_max, (int)sizeof(ses3d_person2d), s.n_persons.as<int32_t>(),
                                  s.in_off.as<long long>(), s.persons.p, s.in_dense.p, -1, st));
    int rc = triangulate_on_device(h, s.sc, st, nf, p_max, h_max, s.persons.as<ses3d_person2d>(),
                                   s.n_persons.as<int32_t>(), s.out3d.as<ses3d_person_cov>(), s.n_out3d.as<int32_t>(),
                                   nullptr, nullptr, nullptr, s.out2d.as<ses3d_person2d>(), s.n_out2d.as<int32_t>());
    if (rc) { status = rc; break; }
    CU(cudaEventRecord(s.kern_done, st));
    if (direct) {
      // the scans continue the running totals of the previous chunk (which ran on another slot's stream)
      if (ci > 0) CU(cudaStreamWaitEvent(st, h->scan_ev[(ci - 1) % n_slots], 0));
      long long* run = h->d_run.as<long long>();
      CU(ses3d::launch_scan_counts(s.n_out3d.as<int32_t>(), nf, h_max, s.off3.as<long long>(), run, st));
      CU(ses3d::launch_scan_counts(s.n_out2d.as<int32_t>(), (int)u_in, h_max, s.off2.as<long long>(), run + 1, st));
      CU(cudaEventRecord(h->scan_ev[si], st));
      CU(ses3d::launch_move_records(0, nf, h_max, (int)sizeof(ses3d_person_cov), s.n_out3d.as<int32_t>(),
                                    s.off3.as<long long>(), s.out3d.p, d3, cap3d, st));
      CU(ses3d::launch_move_records(0, (int)u_in, h_max, (int)sizeof(ses3d_person2d), s.n_out2d.as<int32_t>(),
                                    s.off2.as<long long>(), s.out2d.p, d2, cap2d, st));
      h->launches += 6;
      CU(cudaMemcpyAsync(n_out3d + f0, s.n_out3d.p, (size_t)nf * 4, out_kind, st));
      CU(cudaMemcpyAsync(n_out2d + (size_t)f0 * C, s.n_out2d.p, u_in * 4, out_kind, st));
      continue;
    }
    CU(ses3d::launch_scan_counts(s.n_out3d.as<int32_t>(), nf, h_max, s.off3.as<long long>(), nullptr, st));
    CU(ses3d::launch_move_records(0, nf, h_max, (int)sizeof(ses3d_person_cov), s.n_out3d.as<int32_t>(),
                                  s.off3.as<long long>(), s.out3d.p, s.c3d.p, -1, st));
    CU(ses3d::launch_scan_counts(s.n_out2d.as<int32_t>(), (int)u_in, h_max, s.off2.as<long long>(), nullptr, st));
    CU(ses3d::launch_move_records(0, (int)u_in, h_max, (int)sizeof(ses3d_person2d), s.n_out2d.as<int32_t>(),
                                  s.off2.as<long long>(), s.out2d.p, s.c2d.p, -1, st));
    h->launches += 6;
    CU(cudaMemcpyAsync(&s.totals[0], s.off3.as<long long>() + nf, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&s.totals[1], s.off2.as<long long>() + u_in, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(n_out3d + f0, s.n_out3d.p, (size_t)nf * 4, out_kind, st));
    CU(cudaMemcpyAsync(n_out2d + (size_t)f0 * C, s.n_out2d.p, u_in * 4, out_kind, st));
    CU(cudaEventRecord(s.done, st));
    pending[n_pending++] = si;
    // Results leave in chunk order. The next chunk reuses the oldest pending slot once kSlots - 1 are in flight, so
    // that one is sent home first (blocking on its totals); anything else whose totals have already arrived follows
    // without waiting - the uploads of up to kSlots - 1 chunks run ahead of the kernels, which lets the upload
    // finish early and the tail of the download use the link alone.
    while (status == SES3D_OK && n_pending > 0 &&
           (n_pending >= n_slots - 1 || cudaEventQuery(h->slot[pending[0]].done) == cudaSuccess))
      status = finish_oldest();
  }
  while (!direct && status == SES3D_OK && n_pending > 0) status = finish_oldest();
  if (direct && status == SES3D_OK && ci > 0) {
    // totals + overflow flag ride at the end of the last chunk's stream, which is ordered behind every scan
    cudaStream_t st = h->slot[(ci - 1) % n_slots].stream;
    CU(cudaMemcpyAsync(&h->h_words[0], h->d_run.p, 16, cudaMemcpyDeviceToHost, st));
  }
  for (Slot& sl : h->slot) CU(cudaStreamSynchronize(sl.stream));
  resolve_events(h);
  if (status != SES3D_OK) return status;
  int32_t overflow = 0;
  CU(cudaMemcpy(&overflow, h->d_overflow.p, 4, cudaMemcpyDeviceToHost));
  if (overflow) return overflow_error(h);
  if (direct) {
    run3 = h->h_words[0];
    run2 = h->h_words[1];
    if (run3 > cap3d || run2 > cap2d) return fail(SES3D_E_CAPACITY, "ragged output buffer too small");
  }
  *total3d = run3;
  *total2d = run2;
  return SES3D_OK;
}

}  // namespace

extern "C" {

void ses3d_default_params(ses3d_params* p) {
  if (!p) return;
  p->pose_method = SES3D_POSE_SIMPLE;
  p->precision = SES3D_PRECISION_FP32;
  p->lm_refine = 0;
  p->lm_max_iters = 10;
  p->min_num_valid_keypoints = 9;
  p->triangulation_threshold = 0.30f;
  p->max_epipolar_error = 0.050;
  p->reproj_error_max_acceptable = 0.050;
  p->max_joint_dist_to_root = 2.0;
  p->merge_dist_thresh = 0.20;
  p->limb_cov_offset_sigma = 0.075;
}

int ses3d_create(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* params, int32_t device,
                 ses3d_handle* out) {
  if (!out) return fail(SES3D_E_INVALID, "out is null");
  *out = nullptr;
  if (!cams || n_cams < 2 || n_cams > 255) return fail(SES3D_E_INVALID, "need 2..255 cameras");  // S3D:1133-1136
  ses3d_params prm;
  if (params) prm = *params;
  else ses3d_default_params(&prm);
  if (prm.pose_method != SES3D_POSE_SIMPLE && prm.pose_method != SES3D_POSE_H36M)
    return fail(SES3D_E_INVALID, "pose_method must be SES3D_POSE_SIMPLE or SES3D_POSE_H36M");
  if (prm.precision != SES3D_PRECISION_FP32 && prm.precision != SES3D_PRECISION_FP64)
    return fail(SES3D_E_INVALID, "precision must be SES3D_PRECISION_FP32 or SES3D_PRECISION_FP64");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(SES3D_E_CUDA, std::string("no CUDA device: this library has no CPU path (") +
                                  (e == cudaSuccess ? "device count 0" : cudaGetErrorString(e)) + ")");
  if (device < 0 || device >= n_dev) return fail(SES3D_E_INVALID, "device ordinal out of range");
  CU(cudaSetDevice(device));
  ses3d_handle_s* h = new (std::nothrow) ses3d_handle_s;
  if (!h) return fail(SES3D_E_NOMEM, "out of host memory");
  h->device = device;
  h->prm = prm;
  if (!ses3d::build_host_tables(n_cams, cams, prm, &h->host)) {
    delete h;
    return fail(SES3D_E_INVALID, "invalid camera table (singular extrinsics or zero focal length)");
  }
  auto upload = [&](DevBuf& b, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t ee = b.ensure(bytes);
    if (ee != cudaSuccess) return ee;
    return cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t ue = upload(h->d_camf, h->host.camf.data(), h->host.camf.size() * sizeof(ses3d::CamF));
  if (ue == cudaSuccess) ue = upload(h->d_camd, h->host.camd.data(), h->host.camd.size() * sizeof(ses3d::CamD));
  if (ue == cudaSuccess) ue = upload(h->d_F, h->host.F.data(), h->host.F.size() * sizeof(float));
  if (ue == cudaSuccess) ue = upload(h->d_frow, h->host.f_row.data(), h->host.f_row.size() * sizeof(int));
  if (ue == cudaSuccess) ue = h->d_overflow.ensure(4);
  if (ue == cudaSuccess) ue = cudaMemset(h->d_overflow.p, 0, 4);
  for (int i = 0; i < ses3d_handle_s::kSlots && ue == cudaSuccess; ++i) {
    ue = cudaStreamCreateWithFlags(&h->slot[i].stream, cudaStreamNonBlocking);
    if (ue == cudaSuccess) ue = cudaMallocHost(reinterpret_cast<void**>(&h->slot[i].totals), 2 * sizeof(long long));
    if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->slot[i].done, cudaEventDisableTiming);
    if (ue == cudaSuccess) ue = cudaStreamCreateWithFlags(&h->slot[i].up, cudaStreamNonBlocking);
    if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->slot[i].up_done, cudaEventDisableTiming);
    if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->slot[i].kern_done, cudaEventDisableTiming);
    if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->slot[i].down_done, cudaEventDisableTiming);
  }
  if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming);
  if (ue == cudaSuccess) ue = cudaEventCreateWithFlags(&h->dev_done, cudaEventDisableTiming);
  for (int i = 0; i < ses3d_handle_s::kSlots && ue == cudaSuccess; ++i)
    ue = cudaEventCreateWithFlags(&h->scan_ev[i], cudaEventDisableTiming);
  if (ue == cudaSuccess) ue = cudaMallocHost(reinterpret_cast<void**>(&h->h_words), 4 * sizeof(long long));
  if (ue == cudaSuccess) { h->h_words[0] = h->h_words[1] = h->h_words[2] = h->h_words[3] = 0; }
  if (ue == cudaSuccess) ue = h->d_run.ensure(16);
  if (ue == cudaSuccess) ue = ses3d::init_kernels(&h->cfg, device);
  // the table uploads above are synchronous copies from pageable memory: they may return before the DMA has landed
  // and are not ordered against the handle's non-blocking streams - wait for them once, here
  if (ue == cudaSuccess) ue = cudaDeviceSynchronize();
  // tuning overrides are read here, once; nothing on the per-batch path touches the environment
  if (const char* env = getenv("SES3D_DEVICE_SPLIT")) h->device_split = std::max(1, atoi(env));
  if (const char* env = getenv("SES3D_RAGGED_CHUNK")) h->ragged_chunk_env = std::max(1, atoi(env));
  if (const char* env = getenv("SES3D_RAGGED_DIRECT")) h->ragged_direct = atoi(env);
  if (const char* env = getenv("SES3D_FRAME_GRAPH")) h->frame_graph = atoi(env);
  if (const char* env = getenv("SES3D_RAGGED_ONE_UP")) h->ragged_one_up = atoi(env);
  if (const char* env = getenv("SES3D_RAGGED_ONE_DOWN")) h->ragged_one_down = atoi(env);
  if (ue == cudaSuccess) ue = cudaStreamCreateWithFlags(&h->down, cudaStreamNonBlocking);
  if (const char* env = getenv("SES3D_RAGGED_SLOTS")) h->ragged_slots = std::max(2, std::min((int)ses3d_handle_s::kSlots, atoi(env)));
  if (ue != cudaSuccess) {
    ses3d_destroy(h);
    return cuda_fail(ue, "ses3d_create upload");
  }
  h->tb.n_cams = n_cams;
  h->tb.camf = h->d_camf.as<ses3d::CamF>();
  h->tb.camd = h->d_camd.as<ses3d::CamD>();
  h->tb.F = h->d_F.as<float>();
  h->tb.f_row = h->d_frow.as<int>();
  h->tb.model = h->host.model;
  h->tb.prm = prm;
  h->tb.exact_mode = 3;
  if (const char* env = getenv("SES3D_TRI_EXACT")) h->tb.exact_mode = atoi(env);
  *out = h;
  return SES3D_OK;
}

int ses3d_destroy(ses3d_handle h) {
  if (!h) return SES3D_OK;
  cudaSetDevice(h->device);
  for (FrameGraph& g : h->fgraph) g.release();
  for (Slot& s : h->slot) s.release();
  if (h->down) cudaStreamDestroy(h->down);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  if (h->dev_done) { cudaEventSynchronize(h->dev_done); cudaEventDestroy(h->dev_done); }
  for (cudaEvent_t& e : h->scan_ev) if (e) cudaEventDestroy(e);
  if (h->h_words) cudaFreeHost(h->h_words);
  h->d_run.release();
  h->d_camf.release(); h->d_camd.release(); h->d_F.release(); h->d_frow.release(); h->d_overflow.release();
  delete h;
  return SES3D_OK;
}

int32_t ses3d_n_cams(ses3d_handle h) { return h ? h->tb.n_cams : 0; }

int ses3d_get_tables(ses3d_handle h, float* P, float* F) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  if (P)
    for (int i = 0; i < h->host.n_cams; ++i) std::memcpy(P + (size_t)i * 12, h->host.camf[i].P, 12 * sizeof(float));
  if (F) std::memcpy(F, h->host.F.data(), h->host.F.size() * sizeof(float));
  return SES3D_OK;
}

int ses3d_triangulate_batch(ses3d_handle h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                            const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out, int32_t* n_out,
                            const ses3d_assoc_dump* dump, uint32_t flags, void* stream) {
  int rc = check_dims(h, n_frames, p_max, h_max);
  if (rc) return rc;
  if (n_frames > 0 && (!persons || !n_persons || !out || !n_out)) return fail(SES3D_E_INVALID, "null buffer");
  return run_batch(h, TRI, n_frames, p_max, persons, n_persons, h_max, out, n_out, nullptr, nullptr, dump, flags, stream);
}

int ses3d_reproject_batch(ses3d_handle h, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons3d,
                          const int32_t* n_persons3d, ses3d_person2d* out, int32_t* n_out, uint32_t flags,
                          void* stream) {
  int rc = check_dims(h, n_frames, 1, h_max);
  if (rc) return rc;
  if (n_frames > 0 && (!persons3d || !n_persons3d || !out || !n_out)) return fail(SES3D_E_INVALID, "null buffer");
  return run_batch(h, REP, n_frames, 1, nullptr, nullptr, h_max, const_cast<ses3d_person_cov*>(persons3d),
                   const_cast<int32_t*>(n_persons3d), out, n_out, nullptr, flags, stream);
}

int ses3d_process_batch(ses3d_handle h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                        const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int32_t* n_out3d,
                        ses3d_person2d* out2d, int32_t* n_out2d, const ses3d_assoc_dump* dump, uint32_t flags,
                        void* stream) {
  int rc = check_dims(h, n_frames, p_max, h_max);
  if (rc) return rc;
  if (n_frames > 0 && (!persons || !n_persons || !out2d || !n_out2d)) return fail(SES3D_E_INVALID, "null buffer");
  if ((flags & SES3D_DEVICE_BUFFERS) && n_frames > 0 && (!out3d || !n_out3d))
    return fail(SES3D_E_INVALID, "device-buffer calls need out3d/n_out3d (the PersonCov list lives there)");
  return run_batch(h, TRI | REP, n_frames, p_max, persons, n_persons, h_max, out3d, n_out3d, out2d, n_out2d, dump,
                   flags, stream);
}

int ses3d_process_batch_ragged(ses3d_handle h, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons_dense,
                               const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int64_t cap3d,
                               int32_t* n_out3d, ses3d_person2d* out2d, int64_t cap2d, int32_t* n_out2d,
                               int64_t* total3d, int64_t* total2d, uint32_t flags) {
  int rc = check_dims(h, n_frames, p_max, h_max);
  if (rc) return rc;
  if (!total3d || !total2d) return fail(SES3D_E_INVALID, "null totals");
  if (n_frames > 0 && (!n_persons || !out3d || !n_out3d || !out2d || !n_out2d)) return fail(SES3D_E_INVALID, "null buffer");
  long long t3 = 0, t2 = 0;
  rc = run_ragged(h, n_frames, p_max, persons_dense, n_persons, h_max, out3d, cap3d, n_out3d, out2d, cap2d, n_out2d,
                  &t3, &t2, flags);
  *total3d = t3;
  *total2d = t2;
  return rc;
}

int ses3d_markers_batch(ses3d_handle h, int32_t n_frames, int32_t h_max, const ses3d_person_cov* persons3d,
                        const int32_t* n_persons3d, int32_t style, ses3d_ellipsoid* ellipsoids, double* segments,
                        int32_t* n_segments, int8_t* segment_slot, uint32_t flags, void* stream) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  if (n_frames < 0 || h_max < 1 || h_max > 1024) return fail(SES3D_E_INVALID, "bad n_frames / h_max");
  if (style != SES3D_MARKERS_SKELETON3D && style != SES3D_MARKERS_POSE_PRIOR) return fail(SES3D_E_INVALID, "unknown marker style");
  if (n_frames == 0) return SES3D_OK;
  if (!persons3d || !n_persons3d || (segments && !n_segments)) return fail(SES3D_E_INVALID, "NULL buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  const size_t units = (size_t)n_frames * h_max;
  if (flags & SES3D_DEVICE_BUFFERS) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);   // NULL = legacy default stream
    CU(ses3d::launch_markers(h->tb.model, n_frames, h_max, style, persons3d, n_persons3d, ellipsoids, segments,
                             n_segments, segment_slot, st));
    ++h->launches;
    return SES3D_OK;
  }
  Slot& s = h->slot[0];
  cudaStream_t st = s.stream;
  // staging: the person records in out3d / n_out3d, the results in out2d (one allocation, carved below)
  const size_t b_ell = ellipsoids ? units * ses3d::NFUS * sizeof(ses3d_ellipsoid) : 0;
  const size_t b_seg = segments ? units * SES3D_MARKER_MAX_SEGMENTS * 6 * sizeof(double) : 0;
  const size_t b_n = segments ? units * 4 : 0;
  const size_t b_slot = (segments && segment_slot) ? (units * SES3D_MARKER_MAX_SEGMENTS + 7) / 8 * 8 : 0;
  CU(s.out3d.ensure(units * sizeof(ses3d_person_cov)));
  CU(s.n_out3d.ensure((size_t)n_frames * 4));
  CU(s.out2d.ensure(b_ell + b_seg + b_n + b_slot + 64));
  unsigned char* base = s.out2d.as<unsigned char>();
  ses3d_ellipsoid* d_ell = ellipsoids ? reinterpret_cast<ses3d_ellipsoid*>(base) : nullptr;
  double* d_seg = segments ? reinterpret_cast<double*>(base + b_ell) : nullptr;
  int8_t* d_slot = b_slot ? reinterpret_cast<int8_t*>(base + b_ell + b_seg) : nullptr;
  int32_t* d_n = segments ? reinterpret_cast<int32_t*>(base + b_ell + b_seg + b_slot) : nullptr;
  CU(cudaMemcpyAsync(s.out3d.p, persons3d, units * sizeof(ses3d_person_cov), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(s.n_out3d.p, n_persons3d, (size_t)n_frames * 4, cudaMemcpyHostToDevice, st));
  if (b_seg) CU(cudaMemsetAsync(base + b_ell, 0, b_seg + b_slot, st));   // unused segment slots arrive as zeros
  CU(ses3d::launch_markers(h->tb.model, n_frames, h_max, style, s.out3d.as<ses3d_person_cov>(), s.n_out3d.as<int32_t>(),
                           d_ell, d_seg, d_n, d_slot, st));
  ++h->launches;
  if (ellipsoids) CU(cudaMemcpyAsync(ellipsoids, d_ell, b_ell, cudaMemcpyDeviceToHost, st));
  if (segments) {
    CU(cudaMemcpyAsync(segments, d_seg, b_seg, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(n_segments, d_n, b_n, cudaMemcpyDeviceToHost, st));
    if (segment_slot) CU(cudaMemcpyAsync(segment_slot, d_slot, units * SES3D_MARKER_MAX_SEGMENTS, cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return SES3D_OK;
}

int ses3d_overlay_batch(ses3d_handle h, int32_t n_images, int32_t p_max, const ses3d_person2d* persons,
                        const int32_t* n_persons, int32_t width, int32_t height, uint8_t* rgb, uint32_t flags, void* stream) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  if (n_images < 0 || p_max < 1 || p_max > 1024 || width < 1 || height < 1 || width > 16384 || height > 16384)
    return fail(SES3D_E_INVALID, "bad n_images / p_max / image size");
  if (n_images == 0) return SES3D_OK;
  if (!persons || !n_persons || !rgb) return fail(SES3D_E_INVALID, "NULL buffer");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  if (flags & SES3D_DEVICE_BUFFERS) {
    CU(ses3d::launch_overlay(n_images, p_max, persons, n_persons, width, height, rgb, static_cast<cudaStream_t>(stream)));
    ++h->launches;
    return SES3D_OK;
  }
  Slot& s = h->slot[0];
  cudaStream_t st = s.stream;
  const size_t img_bytes = (size_t)width * height * 3;
  // images are rendered in groups that keep the staging below ~256 MiB
  const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_images, ((size_t)256 << 20) / img_bytes));
  CU(s.persons.ensure((size_t)group * p_max * sizeof(ses3d_person2d)));
  CU(s.n_persons.ensure((size_t)group * 4));
  CU(s.out2d.ensure((size_t)group * img_bytes + 16));
  for (int i0 = 0; i0 < n_images; i0 += group) {
    const int n = std::min(group, n_images - i0);
    CU(cudaMemcpyAsync(s.persons.p, persons + (size_t)i0 * p_max, (size_t)n * p_max * sizeof(ses3d_person2d), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.n_persons.p, n_persons + i0, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CU(ses3d::launch_overlay(n, p_max, s.persons.as<ses3d_person2d>(), s.n_persons.as<int32_t>(), width, height,
                             s.out2d.as<unsigned char>(), st));
    ++h->launches;
    CU(cudaMemcpyAsync(rgb + (size_t)i0 * img_bytes, s.out2d.p, (size_t)n * img_bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
  }
  return SES3D_OK;
}

int ses3d_check(ses3d_handle h) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  return collect_device_status(h, true);
}

int ses3d_reserve(ses3d_handle h, int32_t n_frames, int32_t p_max, int32_t h_max) {
  int rc = check_dims(h, n_frames, p_max, h_max);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  const int C = h->tb.n_cams;
  const int nf = std::min(n_frames, device_chunk(C, p_max));
  bool need_nk = false;
  ses3d::associate_smem_bytes(C, p_max, h_max, &need_nk);
  for (Slot& s : h->slot) {
    CU(s.sc.pairs.ensure((size_t)nf * ses3d::associate_pair_table_bytes(C, p_max)));
    CU(s.sc.meta.ensure((size_t)nf * ses3d::associate_meta_bytes(C, p_max)));
    CU(s.sc.hyp_det.ensure((size_t)nf * h_max * C));
    CU(s.sc.n_hyp.ensure((size_t)nf * 4));
    CU(s.sc.n_hung.ensure((size_t)nf * 4));
    CU(s.sc.keep.ensure((size_t)nf * h_max * 4));
    CU(s.sc.tmp.ensure((size_t)nf * h_max * sizeof(ses3d_person_cov)));
    CU(s.sc.work.ensure((size_t)nf * h_max * 4 * ses3d::kTriBuckets));
    CU(s.sc.work_count.ensure(64));
    if (h->prm.precision == SES3D_PRECISION_FP32) CU(s.sc.far.ensure(ses3d::triangulate_far_scratch_bytes(h->cfg)));
    if (need_nk) CU(s.sc.nk.ensure((size_t)nf * C * p_max * ses3d::NKP * 3 * sizeof(float)));
  }
  return SES3D_OK;
}

int ses3d_munkres_batch(ses3d_handle h, int32_t n, int32_t rows, int32_t cols, const double* cost, int32_t* assignment) {
  if (!h || n < 0 || !cost || !assignment) return fail(SES3D_E_INVALID, "bad argument");
  if (n == 0) return SES3D_OK;
  std::lock_guard<std::mutex> lock(h->mu);
  CU(cudaSetDevice(h->device));
  DevBuf dc, da;
  const size_t cb = (size_t)n * rows * cols * sizeof(double), ab = (size_t)n * rows * sizeof(int32_t);
  CU(dc.ensure(cb));
  cudaError_t e = da.ensure(ab);
  // everything on ONE stream: a synchronous cudaMemcpy from pageable memory may return before its DMA has landed and
  // is not ordered against the handle's non-blocking streams (seen once as a wrong assignment on a 44 x 20 problem)
  cudaStream_t st = h->slot[0].stream;
  if (e == cudaSuccess) e = cudaMemcpyAsync(dc.p, cost, cb, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = ses3d::launch_munkres_batch(n, rows, cols, dc.as<double>(), da.as<int32_t>(), st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(assignment, da.p, ab, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dc.release();
  da.release();
  h->launches += 1;
  if (e != cudaSuccess) return cuda_fail(e, "ses3d_munkres_batch");
  return SES3D_OK;
}

int64_t ses3d_launch_count(ses3d_handle h) { return h ? h->launches : 0; }

int ses3d_set_profiling(ses3d_handle h, int32_t on) {
  if (!h) return fail(SES3D_E_INVALID, "null handle");
  h->profiling = on != 0;
  return SES3D_OK;
}

int ses3d_last_kernel_ms(ses3d_handle h, float ms[4]) {
  if (!h || !ms) return fail(SES3D_E_INVALID, "null argument");
  for (int i = 0; i < 4; ++i) ms[i] = h->kernel_ms[i];
  return SES3D_OK;
}

const char* ses3d_last_error_string(void) { return g_last_error.c_str(); }
const char* ses3d_version(void) { return "ses3d 0.2.0 (sm_100a)"; }

}  // extern "C"
