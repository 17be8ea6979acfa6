// team.h — the cooperative-group abstraction the per-frame algorithms are written against.
//
//   team.pfor(n, f)   run f(i) for i in [0,n) spread over the team's threads, then barrier.
//                     Items must be independent (no two items write the same location).
//   team.single(f)    the leader runs f(), then barrier.
//   team.sync()       barrier with memory ordering among the team's threads.
//   team.shared_pfor(n, tag, f)   phase() followed by a pfor over the member's own n items, f(0, tag, i), skipped when
//                     tag == 0 (a lockstep member without work still passes the barrier). The member index exists so
//                     that a team of teams may spread the items of all members over all of its threads.
//   team.phase()      marks the start of an algorithm phase; a no-op except for LockstepWarpTeam, where it is a
//                     CTA-wide barrier that keeps the CTA's warps (one team each) inside the same stretch of code so
//                     that they share instruction-cache lines. Every path through an algorithm must pass the same
//                     number of phase() calls.
//
// WarpTeam: 32 lanes of one warp (barrier = __syncwarp). BlockTeam: the whole CTA
// (barrier = __syncthreads). SerialTeam: a plain loop (CPU test build only).
#pragma once
#include "common.h"

namespace ses3d {

//   team.first(n, pred)   smallest i in [0,n) with pred(i) true, n if none (same value on every thread)
//   team.min(n, f)        minimum of f(i) over [0,n) as double, DBL_MAX if empty (same value on every thread)
//   team.warp0(f)         run f(warp_team) on the team's first warp only, then barrier (warp-cooperative
//                         sub-algorithms such as the Munkres solver inside a CTA-wide team)
//   team.sum(n, f)        sum of f(i) over [0,n) as double (same value on every thread; warp / serial teams)
//   team.per_warp(n, f)   run f(warp_team, i) for i in [0,n), items spread over the team's warps, then barrier
//   team.compact(n, pred, emit)   emit(i, pos) for every i in [0,n) with pred(i), pos = number of passing items
//                         before i (stream compaction in index order); returns the count (warp / serial teams)
//   team.count(n, pred)   number of i in [0,n) with pred(i) (same value on every thread; warp / serial teams)
//   team.scan(n, f, out)  out[i] = f(0) + ... + f(i-1) for i in [0,n], n <= 32 (warp / serial teams), then barrier
struct SerialTeam {
  template <class F> void pfor(int n, F&& f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> void single(F&& f) { f(); }
  template <class P> int first(int n, P&& pred) { for (int i = 0; i < n; ++i) if (pred(i)) return i; return n; }
  template <class F> double min(int n, F&& f) {
    double m = DBL_MAX;
    for (int i = 0; i < n; ++i) { const double v = f(i); if (v < m) m = v; }
    return m;
  }
  template <class F> void warp0(F&& f) { f(*this); }
  template <class F> double sum(int n, F&& f) { double s = 0.0; for (int i = 0; i < n; ++i) s += f(i); return s; }
  template <class F> void per_warp(int n, F&& f) { for (int i = 0; i < n; ++i) f(*this, i); }
  template <class P, class E> int compact(int n, P&& pred, E&& emit) {
    int pos = 0;
    for (int i = 0; i < n; ++i) if (pred(i)) emit(i, pos++);
    return pos;
  }
  template <class P> int count(int n, P&& pred) { int c = 0; for (int i = 0; i < n; ++i) c += pred(i) ? 1 : 0; return c; }
  template <class F> void scan(int n, F&& f, int* out) {
    int run = 0;
    for (int i = 0; i < n; ++i) { out[i] = run; run += f(i); }
    out[n] = run;
  }
  void sync() {}
  void phase() {}
  template <class F> void shared_pfor(int n, int tag, F&& f) { if (tag) for (int i = 0; i < n; ++i) f(0, tag, i); }
  int rank() const { return 0; }
  int size() const { return 1; }
  int n_warps() const { return 1; }
};

// Reserve the next slot of a list shared by the team's threads (order is unspecified on the GPU).
SES_HD int team_append(int* counter) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(counter, 1);
#else
  return (*counter)++;
#endif
}

#if defined(__CUDACC__)
struct WarpTeam {
  template <class F> __device__ __forceinline__ void pfor(int n, F&& f) {
    for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) f(i);
    __syncwarp();
  }
  template <class F> __device__ __forceinline__ void single(F&& f) {
    if ((threadIdx.x & 31u) == 0) f();
    __syncwarp();
  }
  template <class P> __device__ __forceinline__ int first(int n, P&& pred) {
    const int lane = (int)(threadIdx.x & 31u);
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const unsigned b = __ballot_sync(0xffffffffu, i < n && pred(i));
      if (b) { __syncwarp(); return base + __ffs((int)b) - 1; }
    }
    __syncwarp();   // the predicate's shared-memory reads are ordered before whatever the caller writes next
    return n;
  }
  template <class F> __device__ __forceinline__ double min(int n, F&& f) {
    double m = DBL_MAX;
    for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) { const double v = f(i); if (v < m) m = v; }
    for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, m, off); if (o < m) m = o; }
    __syncwarp();
    return m;
  }
  template <class F> __device__ __forceinline__ void warp0(F&& f) { f(*this); __syncwarp(); }
  template <class F> __device__ __forceinline__ double sum(int n, F&& f) {   // same value on every lane
    double s = 0.0;
    for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) s += f(i);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __syncwarp();
    return s;
  }
  template <class F> __device__ __forceinline__ void per_warp(int n, F&& f) { for (int i = 0; i < n; ++i) f(*this, i); }
  template <class P, class E> __device__ __forceinline__ int compact(int n, P&& pred, E&& emit) {
    const unsigned lane = threadIdx.x & 31u;
    int count = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + (int)lane;
      const bool p = i < n && pred(i);
      const unsigned b = __ballot_sync(0xffffffffu, p);
      if (p) emit(i, count + __popc(b & ((1u << lane) - 1u)));
      count += __popc(b);
    }
    __syncwarp();
    return count;
  }
  template <class P> __device__ __forceinline__ int count(int n, P&& pred) {
    const int lane = (int)(threadIdx.x & 31u);
    int c = 0;
    for (int base = 0; base < n; base += 32) c += __popc(__ballot_sync(0xffffffffu, base + lane < n && pred(base + lane)));
    return c;
  }
  template <class F> __device__ __forceinline__ void scan(int n, F&& f, int* out) {   // n <= 32
    const unsigned lane = threadIdx.x & 31u;
    const int v = (int)lane < n ? f((int)lane) : 0;
    int incl = v;
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if ((int)lane >= off) incl += o;
    }
    if ((int)lane < n) out[lane] = incl - v;
    if ((int)lane == n - 1) out[n] = incl;
    if (n == 0 && lane == 0) out[0] = 0;
    __syncwarp();
  }
  __device__ __forceinline__ void sync() { __syncwarp(); }
  __device__ __forceinline__ void phase() {}
  template <class F> __device__ __forceinline__ void shared_pfor(int n, int tag, F&& f) {
    if (tag) for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) f(0, tag, i);
    __syncwarp();
  }
  __device__ __forceinline__ int rank() const { return (int)(threadIdx.x & 31u); }
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ int n_warps() const { return 1; }
};

// A warp team whose phase() is a CTA barrier: the warps of a CTA work on different items but stay in the same phase.
struct LockstepWarpTeam : WarpTeam {
  __device__ __forceinline__ void phase() { __syncthreads(); }
  // A shared phase starts with the CTA barrier; every member then runs its own items. (Spreading the items of all
  // members over all threads of the CTA - 4 x 17 joints on 128 lanes instead of 17 of 32 lanes per warp - was built and
  // measured on B200: 1.47 ms against 1.42 ms per 16 384 frames. The second barrier and the workspace indirection cost
  // more than the idle lanes.)
  template <class F> __device__ __forceinline__ void shared_pfor(int n, int tag, F&& f) {
    __syncthreads();
    if (tag) for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) f(0, tag, i);
    __syncwarp();
  }
};

struct BlockTeam {
  template <class F> __device__ __forceinline__ void pfor(int n, F&& f) {
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) f(i);
    __syncthreads();
  }
  template <class F> __device__ __forceinline__ void single(F&& f) {
    if (threadIdx.x == 0) f();
    __syncthreads();
  }
  template <class F> __device__ __forceinline__ void warp0(F&& f) {
    if (threadIdx.x < 32) { WarpTeam w; f(w); }
    __syncthreads();
  }
  // every warp of the CTA takes items round-robin and runs f(warp_team, item); then barrier
  template <class F> __device__ __forceinline__ void per_warp(int n, F&& f) {
    const int nw = (int)(blockDim.x >> 5);
    for (int i = (int)(threadIdx.x >> 5); i < n; i += nw) { WarpTeam w; f(w, i); }
    __syncthreads();
  }
  template <class P> __device__ __forceinline__ int count(int n, P&& pred) {
    int c = 0;
    for (int base = 0; base < n; base += (int)blockDim.x) {
      const int i = base + (int)threadIdx.x;
      c += __syncthreads_count(i < n && pred(i));
    }
    return c;
  }
  // stream compaction on the CTA's first warp (the lists are short), count broadcast through shared memory
  template <class P, class E> __device__ __forceinline__ int compact(int n, P&& pred, E&& emit) {
    __shared__ int s_count;
    if (threadIdx.x < 32) {
      WarpTeam w;
      const int c = w.compact(n, pred, emit);
      if (threadIdx.x == 0) s_count = c;
    }
    __syncthreads();
    const int c = s_count;
    __syncthreads();
    return c;
  }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ void phase() {}
  template <class F> __device__ __forceinline__ void shared_pfor(int n, int tag, F&& f) {
    if (tag) for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) f(0, tag, i);
    __syncthreads();
  }
  __device__ __forceinline__ int rank() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int size() const { return (int)blockDim.x; }
  __device__ __forceinline__ int n_warps() const { return (int)(blockDim.x >> 5); }
};
#endif

}  // namespace ses3d
