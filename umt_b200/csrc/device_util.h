// Small device helpers shared by the r-z kernels.
#pragma once
// Select chains for the few dynamic corner indices (cEZ, nextC) of a zone solve: with them every per-thread zone array is indexed
// by compile-time constants only and stays in registers instead of local memory.
template <int MC> __device__ __forceinline__ double pick(const double (&a)[MC], int i) {
  double r = a[0];
#pragma unroll
  for (int k = 1; k < MC; k++) r = i == k ? a[k] : r;
  return r;
}
template <int MC> __device__ __forceinline__ void addto(double (&a)[MC], int i, double v) {   // a[i] += v (x + 0.0 == x)
#pragma unroll
  for (int k = 0; k < MC; k++) a[k] += i == k ? v : 0.0;
}
template <int MC> __device__ __forceinline__ void put(double (&a)[MC], int i, double v) {
#pragma unroll
  for (int k = 0; k < MC; k++) a[k] = i == k ? v : a[k];
}

// Dataflow sweeps (sweeprz.cu, gta_rz.cu): a quiet NaN with a payload no arithmetic produces marks "not computed yet"; consumers poll
// the values they need until they are real, so the data are their own completion flags (no counters, fences or barriers).
constexpr unsigned long long UMT_SENTINEL = 0xFFFFDEADFFFFDEADull;
__device__ __forceinline__ unsigned long long umt_ld_relaxed_u64(const double *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void umt_st_relaxed_f64(double *p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
