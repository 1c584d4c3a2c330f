// microbench.cu -- two measurements the roofline of the pair evaluation rests on (SURVEY.md §7 asks for
// a MEASURED fp64 peak; MEASURED_PEAKS.json carries none):
//   cph_bench_fp64_peak   sustained DFMA issue rate of the device (independent chains, every SM busy)
//   cph_bench_seed_error  worst relative error of the MUFU.RSQ64H / MUFU.RCP64H seeds and of the
//                         refined 1/sqrt(x), 1/x the evaluation kernel uses (fastmath.cuh), against
//                         correctly rounded division / sqrt, over the ranges the kernel feeds them
// Neither touches a handle; both are exported through include/cph_b200.h for bench.py and the tests.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "../../include/cph_b200.h"
#include "fastmath.cuh"

namespace {

// 8 independent DFMA chains per thread, 4096 DFMAs per chain: the fp64 pipe is the only thing in the way
__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;   // never true: keeps the chains alive
}

__device__ __forceinline__ unsigned long long ordered_bits(double e) { return (unsigned long long)__double_as_longlong(e); }

// err[0] rsqrt seed, err[1] rcp seed, err[2] refined rsqrt, err[3] refined rcp  (max relative error, as double bits)
__global__ void seed_error_kernel(int n, double lo_rsq, double hi_rsq, double lo_rcp, double hi_rcp,
                                  unsigned long long *err) {
  double m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    // a dense, irregular sweep: golden-ratio sequence so that every mantissa pattern of the high word is visited
    const double u = (double)k * 0.6180339887498949;
    const double fr = u - floor(u);
    const double x = lo_rsq + (hi_rsq - lo_rsq) * fr;
    const double w = lo_rcp + (hi_rcp - lo_rcp) * fr;
    const double ex = 1.0 / sqrt(x), ew = 1.0 / w;
    m0 = fmax(m0, fabs(rsqrt_seed(x) - ex) / ex);
    m1 = fmax(m1, fabs(rcp_seed(w) - ew) / ew);
    m2 = fmax(m2, fabs(fast_rsqrt(x) - ex) / ex);
    m3 = fmax(m3, fabs(fast_rcp(w) - ew) / ew);
  }
  // positive doubles order like their bit patterns
  atomicMax(err + 0, ordered_bits(m0));
  atomicMax(err + 1, ordered_bits(m1));
  atomicMax(err + 2, ordered_bits(m2));
  atomicMax(err + 3, ordered_bits(m3));
}

}  // namespace

extern "C" int cph_bench_fp64_peak(int device, double *dfma_warp_instr_per_s, double *tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return CPH_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CPH_ERR_CUDA;
  double *d = nullptr;
  if (cudaMalloc((void **)&d, 64) != cudaSuccess) return CPH_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 256;
  double best = 0;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(d, 0.999999, 1e-9, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return CPH_ERR_CUDA; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = (double)blocks * threads * (double)iters * 16 * 8;   // thread-level DFMAs
    if (rep > 0) best = std::max(best, dfma / (ms * 1e-3));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  if (dfma_warp_instr_per_s) *dfma_warp_instr_per_s = best / 32.0;
  if (tflops) *tflops = 2.0 * best * 1e-12;
  return CPH_OK;
}

extern "C" int cph_bench_seed_error(int device, double *out4) {
  if (!out4) return CPH_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return CPH_ERR_CUDA;
  unsigned long long *d = nullptr;
  if (cudaMalloc((void **)&d, 4 * sizeof(unsigned long long)) != cudaSuccess) return CPH_ERR_CUDA;
  cudaMemset(d, 0, 4 * sizeof(unsigned long long));
  // 1/sqrt: r^2 between (0.8 A)^2 and (14 A)^2; 1/x: 1 + p alpha r between 1 and 2.5
  seed_error_kernel<<<1184, 256>>>(1 << 26, 0.64, 196.0, 1.0, 2.5, d);
  unsigned long long h[4];
  if (cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d); return CPH_ERR_CUDA; }
  cudaFree(d);
  for (int k = 0; k < 4; k++) memcpy(&out4[k], &h[k], 8);
  return CPH_OK;
}

extern "C" int cph_refine_order(void) { return 10 * CPH_REFINE_RSQRT + CPH_REFINE_RCP; }
