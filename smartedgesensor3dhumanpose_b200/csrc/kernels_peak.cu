// kernels_peak.cu — micro-benchmark of the CUDA-core FMA pipes (diagnostics, used by the benchmarks as the measured
// roofline denominator of the FP32 / FP64 kernels; MEASURED_PEAKS.json only holds HBM and tensor-core figures).
// Every thread runs 8 independent fused-multiply-add chains, enough resident warps per SM to cover the pipe latency.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "ses3d.h"

namespace ses3d {
int set_error(int code, const std::string& msg);
}

namespace {

template <class T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T b, T c) {
  T a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (T)(threadIdx.x + j) * (T)1e-3;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a[j] * b + c;   // contracted to FMA (this file is compiled with -fmad=true)
  }
  T s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class T>
cudaError_t run(int n_sm, int iters, double* tflops) {
  const int blocks = n_sm * 8, threads = 256;
  T* out = nullptr;
  cudaError_t e = cudaMalloc(&out, sizeof(T) * (size_t)blocks * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0.0;
  for (int rep = 0; rep < 5 && e == cudaSuccess; ++rep) {   // first repetition warms up; best of the rest
    cudaEventRecord(a);
    k_fma_peak<T><<<blocks, threads>>>(out, iters, (T)0.999999, (T)1e-6);
    cudaEventRecord(b);
    e = cudaEventSynchronize(b);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, a, b);
    if (e == cudaSuccess && rep > 0 && ms > 0.f)
      best = std::max(best, 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(out);
  *tflops = best;
  return e;
}

}  // namespace

extern "C" int ses3d_measure_fma_peak(int32_t device, int32_t fp64, double* tflops) {
  if (!tflops) return ses3d::set_error(SES3D_E_INVALID, "ses3d_measure_fma_peak: tflops is NULL");
  *tflops = 0.0;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
    return ses3d::set_error(SES3D_E_CUDA, "ses3d_measure_fma_peak: no CUDA device");
  if (device < 0 || device >= n_dev) return ses3d::set_error(SES3D_E_INVALID, "ses3d_measure_fma_peak: bad device");
  cudaError_t e = cudaSetDevice(device);
  int n_sm = 0;
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = fp64 ? run<double>(n_sm, 1 << 14, tflops) : run<float>(n_sm, 1 << 16, tflops);
  if (e != cudaSuccess) return ses3d::set_error(SES3D_E_CUDA, std::string("ses3d_measure_fma_peak: ") + cudaGetErrorString(e));
  return SES3D_OK;
}
