// kernels_pack.cu — ragged <-> padded record movement for the ragged batch calls.
//
// The reference's messages are variable-length lists (Person2D[] persons, PersonCov[] persons;
// person_msgs/msg/Person2DList.msg:3, PersonCovList.msg:4). The kernels work on fixed-stride
// [unit][capacity] arrays; over PCIe only the occupied records should travel. These kernels turn
// per-unit counts into offsets (one-CTA scan) and copy each unit's contiguous run of records
// between the dense (ragged) and the strided layout, one warp per unit, 16-byte words when aligned.
#include "launch.h"

namespace ses3d {

// offsets[0..n] = base + exclusive prefix sum of clamp(counts[i], 0, cap); one CTA of 1024 threads.
// running (nullable): device accumulator carried from chunk to chunk of a ragged call - base = *running on entry,
// *running = base + total on exit (the scans of consecutive chunks are ordered by events), so every chunk knows
// where its dense output starts without a round trip to the host.
__global__ void __launch_bounds__(1024) k_scan_counts(const int32_t* __restrict__ counts, int n, int cap,
                                                      long long* __restrict__ offsets, long long* running) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, t * per), hi = min(n, lo + per);
  const long long base = running ? *running : 0;
  long long s = 0;
  for (int i = lo; i < hi; ++i) s += min(max(counts[i], 0), cap);
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan
    long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = base + part[t] - s;
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += min(max(counts[i], 0), cap);
  }
  if (t == 1023) {
    offsets[n] = base + part[1023];
    if (running) *running = base + part[1023];
  }
}

// direction 0: strided -> dense (pack), 1: dense -> strided (unpack). rec_words = record size / 4.
// dense_limit: capacity of the dense array in records (pack only; < 0 = unchecked). A unit whose run would end beyond
// it is skipped - the caller compares the running total with the capacity afterwards and reports the overflow - so a
// too-small caller buffer (possibly mapped host memory) is never overrun.
__global__ void __launch_bounds__(256) k_move_records(int direction, int n_units, int cap, int rec_words,
                                                      const int32_t* __restrict__ counts,
                                                      const long long* __restrict__ offsets, uint32_t* strided,
                                                      uint32_t* dense, long long dense_limit) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n_units) return;
  const int cnt = min(max(counts[warp], 0), cap);
  if (dense_limit >= 0 && offsets[warp] + cnt > dense_limit) return;
  const size_t words = (size_t)cnt * rec_words;
  uint32_t* a = strided + (size_t)warp * cap * rec_words;
  uint32_t* b = dense + (size_t)offsets[warp] * rec_words;
  const uint32_t* src = direction == 0 ? a : b;
  uint32_t* dst = direction == 0 ? b : a;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const size_t w4 = words / 4;
    for (size_t i = lane; i < w4; i += 32) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = w4 * 4 + lane; i < words; i += 32) dst[i] = src[i];
  } else {
    for (size_t i = lane; i < words; i += 32) dst[i] = src[i];
  }
}

cudaError_t launch_scan_counts(const int32_t* counts, int n, int cap, long long* offsets, long long* running,
                               cudaStream_t st) {
  k_scan_counts<<<1, 1024, 0, st>>>(counts, n, cap, offsets, running);
  return cudaGetLastError();
}

cudaError_t launch_move_records(int direction, int n_units, int cap, int rec_bytes, const int32_t* counts,
                                const long long* offsets, void* strided, void* dense, long long dense_limit,
                                cudaStream_t st) {
  if (n_units <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)(((size_t)n_units * 32 + 255) / 256);
  k_move_records<<<blocks, 256, 0, st>>>(direction, n_units, cap, rec_bytes / 4, counts, offsets,
                                         static_cast<uint32_t*>(strided), static_cast<uint32_t*>(dense), dense_limit);
  return cudaGetLastError();
}

}  // namespace ses3d
