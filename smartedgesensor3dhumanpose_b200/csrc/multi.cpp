// multi.cpp — single-process multi-GPU entry points and host-thread placement.
//
// "multi-GPU variant takes a device list". Frames are independent (no cross-frame state in
// skeleton_3d / pose_reprojection), so a batch is cut into contiguous frame ranges, one per device, and every range
// runs through that device's own handle on its own host thread: no data-path collective, no peer traffic. The
// caller's buffers are host memory (pinned recommended); each worker thread is bound to the CPUs of its GPU's NUMA
// node so that staging copies and driver work stay on the socket the GPU hangs off.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "ses3d.h"

namespace ses3d {
int set_error(int code, const std::string& msg);
const char* last_error();
}  // namespace ses3d

namespace {

// "0-15,32-47" -> cpu_set_t; returns the number of CPUs set
int parse_cpulist(const std::string& list, cpu_set_t* set) {
  CPU_ZERO(set);
  int n = 0;
  std::stringstream ss(list);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    if (tok.empty()) continue;
    int a = 0, b = 0;
    if (std::sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) {
    } else if (std::sscanf(tok.c_str(), "%d", &a) == 1) {
      b = a;
    } else {
      continue;
    }
    for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++n; }
  }
  return n;
}

std::string read_line(const std::string& path) {
  std::ifstream f(path);
  std::string s;
  if (f) std::getline(f, s);
  return s;
}

}  // namespace

struct ses3d_multi_s {
  std::vector<int> devices;
  std::vector<ses3d_handle> handles;
};

extern "C" {

int ses3d_bind_thread_to_device_numa(int32_t device, int32_t* numa_node) {
  if (numa_node) *numa_node = -1;
  char bus[32] = {0};
  cudaError_t e = cudaDeviceGetPCIBusId(bus, sizeof bus, device);
  if (e != cudaSuccess) return ses3d::set_error(SES3D_E_CUDA, std::string("cudaDeviceGetPCIBusId: ") + cudaGetErrorString(e));
  for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
  const std::string node_s = read_line(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
  if (node_s.empty()) return SES3D_OK;   // no sysfs (container without it): leave the thread where it is
  const int node = std::atoi(node_s.c_str());
  if (numa_node) *numa_node = node;
  if (node < 0) return SES3D_OK;         // single-node machine / unknown: nothing to do
  const std::string cpus = read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  cpu_set_t want, allowed, both;
  if (parse_cpulist(cpus, &want) == 0) return SES3D_OK;
  // stay inside whatever the process is allowed to use (cgroup / taskset); an empty intersection changes nothing
  if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return SES3D_OK;
  CPU_AND(&both, &want, &allowed);
  if (CPU_COUNT(&both) == 0) return SES3D_OK;
  pthread_setaffinity_np(pthread_self(), sizeof both, &both);
  return SES3D_OK;
}

int ses3d_create_multi(int32_t n_cams, const ses3d_camera* cams, const ses3d_params* params, int32_t n_devices,
                       const int32_t* devices, ses3d_multi* out) {
  if (!out) return ses3d::set_error(SES3D_E_INVALID, "ses3d_create_multi: out is NULL");
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return ses3d::set_error(SES3D_E_CUDA, "ses3d_create_multi: no CUDA device (this library has no CPU path)");
  if (n_devices < 0 || n_devices > 64) return ses3d::set_error(SES3D_E_INVALID, "ses3d_create_multi: bad n_devices");
  ses3d_multi_s* m = new ses3d_multi_s;
  if (n_devices == 0 || !devices) {   // all visible devices
    for (int d = 0; d < n_dev; ++d) m->devices.push_back(d);
  } else {
    m->devices.assign(devices, devices + n_devices);
  }
  for (int d : m->devices) {
    ses3d_handle h = nullptr;
    const int rc = ses3d_create(n_cams, cams, params, d, &h);
    if (rc != SES3D_OK) {
      for (ses3d_handle hh : m->handles) ses3d_destroy(hh);
      delete m;
      return rc;
    }
    m->handles.push_back(h);
  }
  *out = m;
  return SES3D_OK;
}

int ses3d_multi_destroy(ses3d_multi m) {
  if (!m) return SES3D_OK;
  for (ses3d_handle h : m->handles) ses3d_destroy(h);
  delete m;
  return SES3D_OK;
}

int32_t ses3d_multi_device_count(ses3d_multi m) { return m ? (int32_t)m->handles.size() : 0; }

ses3d_handle ses3d_multi_handle(ses3d_multi m, int32_t i) {
  return (m && i >= 0 && i < (int32_t)m->handles.size()) ? m->handles[i] : nullptr;
}

int ses3d_multi_process_batch(ses3d_multi m, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons,
                              const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int32_t* n_out3d,
                              ses3d_person2d* out2d, int32_t* n_out2d, const ses3d_assoc_dump* dump) {
  if (!m || m->handles.empty()) return ses3d::set_error(SES3D_E_INVALID, "ses3d_multi_process_batch: null handle");
  if (n_frames < 0) return ses3d::set_error(SES3D_E_INVALID, "n_frames < 0");
  if (n_frames == 0) return SES3D_OK;
  if (!persons || !n_persons || !out2d || !n_out2d) return ses3d::set_error(SES3D_E_INVALID, "null buffer");
  const int G = (int)m->handles.size();
  const int C = ses3d_n_cams(m->handles[0]);
  std::vector<int> rc(G, SES3D_OK);
  std::vector<std::string> err(G);
  std::vector<std::thread> pool;
  for (int g = 0; g < G; ++g) {
    const int f0 = (int)((int64_t)n_frames * g / G), f1 = (int)((int64_t)n_frames * (g + 1) / G);
    if (f1 <= f0) continue;
    pool.emplace_back([=, &rc, &err] {
      ses3d_bind_thread_to_device_numa(m->devices[g], nullptr);
      ses3d_assoc_dump sub, *dp = nullptr;
      if (dump) {
        sub.hyp_of = dump->hyp_of ? dump->hyp_of + (size_t)f0 * C * p_max : nullptr;
        sub.n_hyp = dump->n_hyp ? dump->n_hyp + f0 : nullptr;
        sub.n_hungarian = dump->n_hungarian ? dump->n_hungarian + f0 : nullptr;
        dp = &sub;
      }
      rc[g] = ses3d_process_batch(m->handles[g], f1 - f0, p_max, persons + (size_t)f0 * C * p_max,
                                  n_persons + (size_t)f0 * C, h_max, out3d ? out3d + (size_t)f0 * h_max : nullptr,
                                  n_out3d ? n_out3d + f0 : nullptr, out2d + (size_t)f0 * C * h_max,
                                  n_out2d + (size_t)f0 * C, dp, SES3D_HOST_BUFFERS, nullptr);
      if (rc[g] != SES3D_OK) err[g] = ses3d::last_error();   // thread-local: carry it to the caller's thread
    });
  }
  for (std::thread& t : pool) t.join();
  for (int g = 0; g < G; ++g)
    if (rc[g] != SES3D_OK) return ses3d::set_error(rc[g], "device " + std::to_string(m->devices[g]) + ": " + err[g]);
  return SES3D_OK;
}

int ses3d_multi_process_batch_ragged(ses3d_multi m, int32_t n_frames, int32_t p_max, const ses3d_person2d* persons_dense,
                                     const int32_t* n_persons, int32_t h_max, ses3d_person_cov* out3d, int64_t cap3d,
                                     int32_t* n_out3d, ses3d_person2d* out2d, int64_t cap2d, int32_t* n_out2d,
                                     int64_t* seg3d, int64_t* seg2d) {
  if (!m || m->handles.empty()) return ses3d::set_error(SES3D_E_INVALID, "ses3d_multi_process_batch_ragged: null handle");
  if (n_frames < 0 || cap3d < 0 || cap2d < 0) return ses3d::set_error(SES3D_E_INVALID, "bad n_frames / capacity");
  const int G = (int)m->handles.size();
  if (!seg3d || !seg2d) return ses3d::set_error(SES3D_E_INVALID, "null segment tables");
  for (int g = 0; g < 2 * G; ++g) seg3d[g] = seg2d[g] = 0;
  if (n_frames == 0) return SES3D_OK;
  if (!n_persons || !out3d || !n_out3d || !out2d || !n_out2d) return ses3d::set_error(SES3D_E_INVALID, "null buffer");
  const int C = ses3d_n_cams(m->handles[0]);
  // start of every shard's run in the dense input
  std::vector<long long> in_start(G + 1, 0);
  {
    long long run = 0;
    int g = 0;
    for (int f = 0; f <= n_frames; ++f) {
      while (g <= G && f == (int)((int64_t)n_frames * g / G)) in_start[g++] = run;
      if (f == n_frames) break;
      for (int c = 0; c < C; ++c) run += std::min(std::max(n_persons[(size_t)f * C + c], 0), p_max);
    }
  }
  std::vector<int> rc(G, SES3D_OK);
  std::vector<std::string> err(G);
  std::vector<std::thread> pool;
  for (int g = 0; g < G; ++g) {
    const int f0 = (int)((int64_t)n_frames * g / G), f1 = (int)((int64_t)n_frames * (g + 1) / G);
    // every shard owns the slice of the output buffers proportional to its frame range
    const long long s3 = cap3d * f0 / n_frames, e3 = cap3d * f1 / n_frames;
    const long long s2 = cap2d * f0 / n_frames, e2 = cap2d * f1 / n_frames;
    seg3d[2 * g] = s3;
    seg2d[2 * g] = s2;
    if (f1 <= f0) continue;
    pool.emplace_back([=, &rc, &err, &in_start] {
      ses3d_bind_thread_to_device_numa(m->devices[g], nullptr);
      int64_t t3 = 0, t2 = 0;
      rc[g] = ses3d_process_batch_ragged(m->handles[g], f1 - f0, p_max, persons_dense ? persons_dense + in_start[g] : nullptr,
                                         n_persons + (size_t)f0 * C, h_max, out3d + s3, e3 - s3, n_out3d + f0, out2d + s2,
                                         e2 - s2, n_out2d + (size_t)f0 * C, &t3, &t2, SES3D_HOST_BUFFERS);
      seg3d[2 * g + 1] = t3;
      seg2d[2 * g + 1] = t2;
      if (rc[g] != SES3D_OK) err[g] = ses3d::last_error();
    });
  }
  for (std::thread& t : pool) t.join();
  for (int g = 0; g < G; ++g)
    if (rc[g] != SES3D_OK) return ses3d::set_error(rc[g], "device " + std::to_string(m->devices[g]) + ": " + err[g]);
  return SES3D_OK;
}

}  // extern "C"
