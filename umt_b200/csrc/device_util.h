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

