// fastmath.cuh -- fp64 1/sqrt(x) and 1/x from the hardware seeds (MUFU.RSQ64H / MUFU.RCP64H) plus one
// Newton step.  Shared by the pair evaluation (pair.cu) and the accuracy / peak microbenchmarks (microbench.cu).
#pragma once

// Order of the Newton step behind each MUFU seed (2 or 3).  Measured on B200 (cph_bench_seed_error): both seeds
// are good to 2^-20; a second-order step leaves 1.3e-12 (1/sqrt) / 1.0e-12 (1/x), a third-order step 1e-16.
// 1/r enters the LJ energy with its 12th power, so it takes the third-order step (+2 fp64 instructions);
// 1/(1 + p alpha r) only feeds the erfc polynomial, where 1e-12 is two orders below the parity tolerance.
#ifndef CPH_REFINE_RSQRT
#define CPH_REFINE_RSQRT 3
#endif
#ifndef CPH_REFINE_RCP
#define CPH_REFINE_RCP 2
#endif

__device__ __forceinline__ double rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rcp_seed(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
// x with its high word replaced by zero unless keep: a denormal (or zero) that cannot change a sum
__device__ __forceinline__ double keep_if(bool keep, double x) {
  return __hiloint2double(keep ? __double2hiint(x) : 0, __double2loint(x));
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  const double y = rsqrt_seed(x);
  const double t = x * y;
  const double e = fma(-t, y, 1.0);                     // 1 - x y^2
#if CPH_REFINE_RSQRT == 2
  const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));   // y / 2 (exponent - 1)
  return fma(h, e, y);                                  // y (1 + e/2): error 3/8 e^2
#else
  const double p = fma(0.375, e, 0.5) * e;
  return fma(y, p, y);                                  // y (1 + e/2 + 3/8 e^2)
#endif
}
template <int ORDER = CPH_REFINE_RCP>
__device__ __forceinline__ double fast_rcp(double x) {
  const double y = rcp_seed(x);
  const double e = fma(-x, y, 1.0);
  if (ORDER == 2) return fma(y, e, y);                  // y (1 + e): error e^2
  return fma(y, fma(e, e, e), y);                       // y (1 + e + e^2): error e^3
}
