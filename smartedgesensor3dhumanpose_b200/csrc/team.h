// team.h — the cooperative-group abstraction the per-frame algorithms are written against.
//
//   team.pfor(n, f)   run f(i) for i in [0,n) spread over the team's threads, then barrier.
//                     Items must be independent (no two items write the same location).
//   team.single(f)    the leader runs f(), then barrier.
//   team.sync()       barrier with memory ordering among the team's threads.
//
// WarpTeam: 32 lanes of one warp (barrier = __syncwarp). BlockTeam: the whole CTA
// (barrier = __syncthreads). SerialTeam: a plain loop (CPU test build only).
#pragma once
#include "common.h"

namespace ses3d {

struct SerialTeam {
  template <class F> void pfor(int n, F&& f) { for (int i = 0; i < n; ++i) f(i); }
  template <class F> void single(F&& f) { f(); }
  void sync() {}
  int rank() const { return 0; }
  int size() const { return 1; }
};

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
  __device__ __forceinline__ void sync() { __syncwarp(); }
  __device__ __forceinline__ int rank() const { return (int)(threadIdx.x & 31u); }
  __device__ __forceinline__ int size() const { return 32; }
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
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ int rank() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int size() const { return (int)blockDim.x; }
};
#endif

}  // namespace ses3d
