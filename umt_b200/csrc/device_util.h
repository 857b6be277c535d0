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

// Watchdog of the dataflow polling loops: a value that never turns real (a caller's array that carries the mark's bit pattern, an
// uninitialised row) must end in an error, not in a hung GPU.  Every 256 polls a thread looks at the launch's abort flag; a thread
// that has polled `limit` times raises it.  Pollers then leave their loops with whatever they hold (the sweep's result is discarded:
// the host sees the flag and returns UMT_ERR_STATE).
__device__ __forceinline__ bool umt_spin_expired(unsigned &polls, int *abortFlag, unsigned limit) {
  if ((++polls & 255u) != 0u) return false;
  if (polls >= limit) atomicExch(abortFlag, 1);
  return *reinterpret_cast<volatile int *>(abortFlag) != 0;
}

// mbarrier / 1-D TMA (cp.async.bulk) helpers for the pipelined r-z sweep (sweeprz.cu); sweep3d.cu keeps its own copies
__device__ __forceinline__ unsigned umt_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void umt_mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(umt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void umt_mbar_arrive(unsigned long long *bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(umt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umt_mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(umt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umt_mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{ .reg .pred p;\n"
      "WAIT_%=:\n"
      "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "  @p bra DONE_%=;\n"
      "  bra WAIT_%=;\n"
      "DONE_%=: }\n" ::"r"(umt_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool umt_mbar_test(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{ .reg .pred p;\n"
      "  mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "  selp.u32 %0, 1, 0, p; }\n" : "=r"(ok) : "r"(umt_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
constexpr unsigned long long UMT_L2_EVICT_FIRST = 0x12F0000000000000ull;   // the encoding CUTLASS names TMA::CacheHintSm90::EVICT_FIRST
__device__ __forceinline__ void umt_tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(umt_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(umt_smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ int umt_ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// 1/x for normal x > 0: MUFU.RCP64H seed + two Newton steps, ~1 ulp (as in sweep3d.cu); keeps IEEE division off latency-bound paths
__device__ __forceinline__ double umt_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// 8-byte asynchronous global -> shared copy (LDGSTS): no register staging, completes in the background
__device__ __forceinline__ void umt_cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(umt_smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void umt_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void umt_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
