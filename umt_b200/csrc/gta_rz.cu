// Grey transport acceleration in r-z geometry for sm_100a.
//
// Replaces snac/SweepGreyUCBrz.F90:12-133 + SweepGreyUCBrzKernelNew :137-330 (the 1-group UCB sweep of the "new" GTA
// solver), the r-z branch of snac/GTASweep.F90:113-119,149-159 (tPsiM = tInc = 0, finishing directions skipped) and
// snac/InitSweepGreyUCBrz.F90:10-235 (within-zone transfer matrices).  The angle set is the level-symmetric S2 set of
// rt/quadrz.F90:82-160: two xi-levels of (starting direction, mu < 0, mu > 0, finishing direction).
//
// Device design: like the multigroup r-z sweep (sweeprz.cu) one persistent launch sweeps every non-finishing angle; work
// items are (angle, hyperplane, chunk of zones) pulled through an atomic ticket and gated by the angle's previous plane and
// by the previous angle of the xi-level (half-angle values tPsiM / tInc, carried per level).  With one group the
// parallelism is zones-in-plane: one thread owns one zone.  tPsi (8, nc+nb) keeps PsiB(:, angle) in its tail rows as in
// the 3-D grey sweep (gta.cu), pInc (8, nc) is summed in fixed angle order afterwards (deterministic PhiInc).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>

#include "umt_internal.h"
#include "device_util.h"

namespace {

constexpr int MAXC2 = 8;
constexpr int GRZ_BLOCK = 64;   // zones per work item
constexpr double FOURALPHA = 1.82;

#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double dot2(const double *om, const double *A) { return __dadd_rn(__dmul_rn(om[0], A[0]), __dmul_rn(om[1], A[1])); }

struct GtaRZParams {
  int nc, nb, nz, nItems;
  const int *numCorner, *cOffSet, *cFP /* 0-based row; >= nc: boundary */, *cEZ;
  const double *Volume, *Area, *Afp, *Aez, *RadiusFP, *RadiusEZ, *omega /* (nAng, 2) */;
  const double *fac, *w1, *w2;
  const unsigned char *start;
  const int *level;
  const int *nextZ;
  const unsigned char *nextC;
  const WorkItem *items;
  int *counters;
  const double *sigTotal, *sigtInv, *tsa;
  double *tpsi, *pinc, *psim, *tinc;
  int *abortFlag;      // watchdog of the dataflow polling loop (device_util.h)
  unsigned spinLimit;
  // dataflow kernel
  double *psimA, *tincA;    // (nAng, nc) tPsiM / tInc as written by angle a
  const int *prevAngle;     // (nAng) previous swept angle of the level, -1: none
  const int *nHyp;          // (nAng)
};

// SweepGreyUCBrzKernelNew for one (zone, angle), split like the multigroup r-z solve (sweeprz.cu): the static half (no dependence
// on other zones of this sweep) reduces the closure to its linear form in the upstream fluxes u = tPsi(row) and in the half-angle
// values, every division included; the half on the dependency chain is a few FMAs per face:
//   src(c)  = srcS(c) + fa(c) tPsiM(c) + sum_f k1(c,f) u(c,f),   src(cez) += k2(c,f) u(c,f)
//   pInc(c) =           fa(c) tInc(c)  + sum_f k1(c,f) u(c,f),   pInc(cez) += k2(c,f) u(c,f)
//   psi(c_i) = src(c_i) inv_i, pinc(c_i) = pInc(c_i) inv_i, pushed into dz_i[f] with rz_i[f]     (corners c_i in nextC order)
// k1 = -R_fp afp + R gtau sigA, k2 = -R gtau sigA (SweepGreyUCBrz.F90:262-297).  Loops are fully unrolled and dynamic corner
// indices go through select chains (device_util.h), so the zone stays in registers.
template <int MC>
struct GtaZoneRZ {
  double srcS[MC], fa[MC], k1[MC][2], k2[MC][2], inv[MC], rz[MC][2];
  int row[MC][2], cez[MC][2], dz[MC][2], ci[MC];
  unsigned inMask, exitMask;
  int nCorner, c0;
};

template <int MC>
__device__ __forceinline__ void gta_static_rz(const GtaRZParams &P, int a, int zone0, GtaZoneRZ<MC> &Z) {
  const int nc = P.nc;
  const double om[2] = {P.omega[2 * a], P.omega[2 * a + 1]};
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  const double fac = P.fac[a];
  Z.nCorner = nCorner; Z.c0 = c0; Z.inMask = 0u; Z.exitMask = 0u;
  double Q[MC], Sigt[MC], denom[MC], area[MC], coef[MC][2], afpv[MC][2], aezv[MC][2], Rfp[MC][2], Rez[MC][2];
#pragma unroll
  for (int c = 0; c < MC; c++) {
    Q[c] = 0.0; Z.srcS[c] = 0.0; Sigt[c] = 1.0; denom[c] = 1.0; area[c] = 0.0; Z.fa[c] = 0.0; Z.ci[c] = c;
    if (c < nCorner) {
      const int cc = c0 + c;
      const double t = P.tsa[cc], vol = P.Volume[cc];
      area[c] = P.Area[cc];
      Q[c] = P.sigtInv[cc] * t;
      Z.srcS[c] = vol * t;
      Sigt[c] = P.sigTotal[cc];
      Z.fa[c] = fac * area[c];
      denom[c] = Sigt[c] * vol + fac * area[c];
      Z.ci[c] = nextC[cc];
    }
  }
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++) {
      coef[c][f] = 0.0; aezv[c][f] = 0.0; afpv[c][f] = 0.0; Rfp[c][f] = 0.0; Rez[c][f] = 0.0;
      Z.k1[c][f] = 0.0; Z.k2[c][f] = 0.0; Z.cez[c][f] = 0; Z.row[c][f] = 0;
      if (c < nCorner) {
        const int cc = c0 + c;
        afpv[c][f] = dot2(om, P.Afp + ((size_t)cc * 2 + f) * 2);
        aezv[c][f] = dot2(om, P.Aez + ((size_t)cc * 2 + f) * 2);
        Z.row[c][f] = P.cFP[cc * 2 + f];
        Z.cez[c][f] = P.cEZ[cc * 2 + f];
        Rfp[c][f] = P.RadiusFP[cc * 2 + f];
        Rez[c][f] = P.RadiusEZ[cc * 2 + f];
      }
    }
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++)
      if (c < nCorner) {
        const double afp = afpv[c][f], aez = aezv[c][f];
        const bool inc = afp < 0.0;
        if (inc) {
          const double R_afp = Rfp[c][f] * afp;
          Z.inMask |= 1u << (2 * c + f);
          Z.k1[c][f] = -R_afp;
          denom[c] -= R_afp;
        } else if (Z.row[c][f] >= nc) Z.exitMask |= 1u << (2 * c + f);
        if (aez > 0.0) {
          const double R = Rez[c][f];
          const int cez = Z.cez[c][f];
          coef[c][f] = R * aez;
          addto<MC>(denom, cez, R * aez);
          const double qcez = pick<MC>(Q, cez);
          double B0;
          if (inc) {
            const double sigA = Sigt[c] * area[c], sigA2 = sigA * sigA;
            const double gnum = aez * aez * (FOURALPHA * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
            const double gtau = gnum / (gnum + 4.0 * sigA2 * sigA2 + aez * sigA * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
            const double B1 = R * gtau * sigA;
            B0 = R * (0.5 * aez * (1.0 - gtau) * (Q[c] - qcez) - gtau * sigA * Q[c]);
            Z.k1[c][f] += B1;
            Z.k2[c][f] = -B1;
          } else {
            B0 = 0.5 * R * aez * (Q[c] - qcez);
          }
          Z.srcS[c] += B0;
          addto<MC>(Z.srcS, cez, -B0);
        }
      }
#pragma unroll
  for (int i = 0; i < MC; i++) {
    Z.inv[i] = 1.0; Z.rz[i][0] = 0.0; Z.rz[i][1] = 0.0; Z.dz[i][0] = 0; Z.dz[i][1] = 0;
    if (i < nCorner) {
      const int c = Z.ci[i];
      Z.inv[i] = 1.0 / pick<MC>(denom, c);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        double r = coef[0][f];
        int d = Z.cez[0][f];
#pragma unroll
        for (int k = 1; k < MC; k++) { r = c == k ? coef[k][f] : r; d = c == k ? Z.cez[k][f] : d; }
        Z.rz[i][f] = r;
        Z.dz[i][f] = d;
      }
    }
  }
}

// FLOW: the dataflow scheme (see gta_sweep_rz_flow_kernel): inputs are polled until they are no longer marked, the half-angle values
// live in per-angle slabs, outputs other threads wait for are stored with st.relaxed.gpu.
template <int MC, bool FLOW = false>
__device__ __forceinline__ void gta_solve_rz(const GtaRZParams &P, int a, const GtaZoneRZ<MC> &Z) {
  const int nc = P.nc;
  double *tpsi = P.tpsi + (size_t)a * (nc + P.nb);
  double *pincA = P.pinc + (size_t)a * nc;
  double *psimL = FLOW ? P.psimA + (size_t)a * nc : P.psim + (size_t)P.level[a] * nc;
  double *tincL = FLOW ? P.tincA + (size_t)a * nc : P.tinc + (size_t)P.level[a] * nc;
  const int nCorner = Z.nCorner, c0 = Z.c0;
  double src[MC], pinc[MC], pmOld[MC], tiOld[MC], u[MC][2];
  if (FLOW) {
    const int pa = P.prevAngle[a];
    const double *psimP = P.psimA + (size_t)(pa < 0 ? 0 : pa) * nc, *tincP = P.tincA + (size_t)(pa < 0 ? 0 : pa) * nc;
    bool ok;
    unsigned polls = 0;
    do {
      ok = true;
#pragma unroll
      for (int c = 0; c < MC; c++) {
        pmOld[c] = 0.0; tiOld[c] = 0.0; u[c][0] = 0.0; u[c][1] = 0.0;
        if (c < nCorner) {
          if (pa >= 0) {
            const unsigned long long v1 = umt_ld_relaxed_u64(&psimP[c0 + c]), v2 = umt_ld_relaxed_u64(&tincP[c0 + c]);
            ok = ok && v1 != UMT_SENTINEL && v2 != UMT_SENTINEL;
            pmOld[c] = __longlong_as_double((long long)v1); tiOld[c] = __longlong_as_double((long long)v2);
          }
#pragma unroll
          for (int f = 0; f < 2; f++)
            if (Z.inMask & (1u << (2 * c + f))) {
              const unsigned long long v = umt_ld_relaxed_u64(&tpsi[Z.row[c][f]]);
              ok = ok && v != UMT_SENTINEL;
              u[c][f] = __longlong_as_double((long long)v);
            }
        }
      }
      if (!ok) { __nanosleep(40); if (umt_spin_expired(polls, P.abortFlag, P.spinLimit)) break; }
    } while (!ok);
  } else {
#pragma unroll
    for (int c = 0; c < MC; c++) {
      pmOld[c] = 0.0; tiOld[c] = 0.0; u[c][0] = 0.0; u[c][1] = 0.0;
      if (c < nCorner) {
        pmOld[c] = psimL[c0 + c]; tiOld[c] = tincL[c0 + c];
#pragma unroll
        for (int f = 0; f < 2; f++)
          if (Z.inMask & (1u << (2 * c + f))) u[c][f] = __ldcg(&tpsi[Z.row[c][f]]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MC; c++) { src[c] = fma(Z.fa[c], pmOld[c], Z.srcS[c]); pinc[c] = Z.fa[c] * tiOld[c]; }
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++) {
      const double t1 = Z.k1[c][f] * u[c][f], t2 = Z.k2[c][f] * u[c][f];
      src[c] += t1; pinc[c] += t1;
      addto<MC>(src, Z.cez[c][f], t2);
      addto<MC>(pinc, Z.cez[c][f], t2);
    }
#pragma unroll
  for (int i = 0; i < MC; i++) {
    if (i < nCorner) {
      const int c = Z.ci[i];
      const double p = pick<MC>(src, c) * Z.inv[i], pi = pick<MC>(pinc, c) * Z.inv[i];
      put<MC>(src, c, p); put<MC>(pinc, c, pi);   // src now holds the corner flux
#pragma unroll
      for (int f = 0; f < 2; f++) { addto<MC>(src, Z.dz[i][f], Z.rz[i][f] * p); addto<MC>(pinc, Z.dz[i][f], Z.rz[i][f] * pi); }
    }
  }
  const bool starting = P.start[a] != 0;
  const double w1 = P.w1[a], w2 = P.w2[a];
#pragma unroll
  for (int c = 0; c < MC; c++) {
    if (c < nCorner) {
      const int cc = c0 + c;
      const double pmn = starting ? src[c] : w1 * src[c] - w2 * pmOld[c], tin = starting ? pinc[c] : w1 * pinc[c] - w2 * tiOld[c];
      pincA[cc] = pinc[c];
      if (FLOW) { umt_st_relaxed_f64(&psimL[cc], pmn); umt_st_relaxed_f64(&tincL[cc], tin); umt_st_relaxed_f64(&tpsi[cc], src[c]); }
      else { tpsi[cc] = src[c]; psimL[cc] = pmn; tincL[cc] = tin; }
#pragma unroll
      for (int f = 0; f < 2; f++)
        if (Z.exitMask & (1u << (2 * c + f))) tpsi[Z.row[c][f]] = src[c];
    }
  }
}

template <int MC>
__global__ void __launch_bounds__(GRZ_BLOCK) gta_sweep_rz_kernel(GtaRZParams P) {
  __shared__ int s_item;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&P.counters[0], 1);   // taken only when free to work on it: a ticket held ahead of time blocks a ready item behind a waiting one
    __syncthreads();
    const int it = s_item;
    if (it >= P.nItems) break;
    const WorkItem w = P.items[it];
    const int zi = w.zbeg + threadIdx.x;
    GtaZoneRZ<MC> Z;
    if (zi < w.zend) gta_static_rz<MC>(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + zi], Z);
    if (threadIdx.x == blockDim.x - 1) {
      if (w.wait_idx >= 0)
        while (ld_acquire(&P.counters[1 + w.wait_idx]) < w.wait_count) __nanosleep(20);
      if (w.pad0 >= 0)   // the previous angle of this xi-level (tPsiM / tInc chain)
        while (ld_acquire(&P.counters[1 + w.pad0]) < w.pad1) __nanosleep(20);
    }
    __syncthreads();
    if (zi < w.zend) gta_solve_rz<MC>(P, w.angle, Z);
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + w.signal_idx]) : "memory");
    }
  }
}

// Dataflow variant.  The grey sweeps are bound by the plane-to-plane chain (one group: a plane is one or two warps of zones), and a
// hop of the item kernel costs store -> fence -> counter -> poll -> barrier -> load.  Here the corner rows of tPsi and per-angle
// slabs of tPsiM / tInc are marked "not computed yet" (UMT_SENTINEL) before the launch and every thread polls exactly the values its
// zone needs until they are real: a hop is one L2 store -> load round trip, and a zone starts as soon as ITS upstream zones are
// done instead of when the whole previous plane is.  One polling thread per zone (there is only one group), so the polling
// traffic that made this scheme a loss for the multigroup sweep is 64x smaller.  Warps take 32 zones of an item through their own
// ticket; producers always hold earlier tickets than their consumers, so every polled value is being computed by a resident
// warp.  Not used with reflecting boundaries (staged launches) or zones with an intra-zone cycle.
template <int MC>
__global__ void __launch_bounds__(GRZ_BLOCK) gta_sweep_rz_flow_kernel(GtaRZParams P, int warpsPerItem) {
  const int lane = threadIdx.x & 31;
  const int nUnits = P.nItems * warpsPerItem;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&P.counters[0], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nUnits) break;
    const int it = t / warpsPerItem, sub = t - it * warpsPerItem;
    const WorkItem w = P.items[it];
    const int zi = w.zbeg + sub * 32 + lane;
    if (zi < w.zend) {
      GtaZoneRZ<MC> Z;
      gta_static_rz<MC>(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + zi], Z);
      gta_solve_rz<MC, true>(P, w.angle, Z);
    }
    __syncwarp();
  }
}

__global__ void gta_rz_mark_kernel(double *tpsi, double *psimA, double *tincA, const int *nHyp, int nc, int rows) {
  const int a = blockIdx.y;
  if (nHyp[a] == 0) return;
  const double mark = __longlong_as_double((long long)UMT_SENTINEL);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
    tpsi[(size_t)a * rows + i] = mark; psimA[(size_t)a * nc + i] = mark; tincA[(size_t)a * nc + i] = mark;
  }
}

// Chain variant: with one group the only independent chains are the xi-levels, so one CTA per level walks its angles and
// planes by itself, a __syncthreads() between planes instead of a global signal/poll round trip (see sweeprz_chain_kernel).
template <int MC>
__global__ void __launch_bounds__(256) gta_sweep_rz_chain_kernel(GtaRZParams P, const int *levelAngles, int maxAngLevel, const int *planeOff,
                                                                  int hypStride, const int *nHyp) {
  const int lev = blockIdx.x;
  for (int k = 0; k < maxAngLevel; k++) {
    const int a = levelAngles[lev * maxAngLevel + k];
    if (a < 0) break;
    const int *nextZ = P.nextZ + (size_t)a * P.nz;
    const int *off = planeOff + (size_t)a * hypStride;
    const int nh = nHyp[a];
    for (int p = 0; p < nh; p++) {
      const int zbeg = off[p], n = off[p + 1] - zbeg;
      GtaZoneRZ<MC> Z;
      if ((int)threadIdx.x < n) gta_static_rz<MC>(P, a, nextZ[zbeg + threadIdx.x], Z);   // off the chain: before the barrier
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (i != (int)threadIdx.x) gta_static_rz<MC>(P, a, nextZ[zbeg + i], Z);
        gta_solve_rz<MC>(P, a, Z);
      }
    }
  }
}

__global__ void gta_reflect_rows_kernel(double *tpsi, const int4 *ops, int rows, int nc) {
  const int4 o = ops[blockIdx.y];   // x = Minc, y = Mref, z = first boundary element, w = count
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < o.w) tpsi[(size_t)o.x * rows + nc + o.z + i] = tpsi[(size_t)o.y * rows + nc + o.z + i];
}

struct TTRZParams {
  int nz, nc, mC, nAng;
  const int *numCorner, *cOffSet, *cEZ;
  const unsigned char *nextC, *start, *finish;
  const double *Volume, *Area, *Afp, *Aez, *RadiusFP, *RadiusEZ, *omega, *weight, *fac, *w1, *w2, *sigTotal;
  double *TT;
};

// InitGreySweepUCBrz: TT(:, corners of zone) = sum over the weighted angles of w Pvv, the starting direction's response
// carried along the xi-level through Tvv
template <int MC>
__global__ void __launch_bounds__(64) gta_init_tt_rz_kernel(TTRZParams Z) {
  const int zone = blockIdx.x * blockDim.x + threadIdx.x;
  if (zone >= Z.nz) return;
  const int nCorner = Z.numCorner[zone], c0 = Z.cOffSet[zone], mC = Z.mC;
  double T[MC][MC], Tvv[MC][MC], Pvv[MC][MC];   // [column c1][row c]
  double Sigt[MC], denom[MC], coefpsi[MC][2];
  int nxez[MC], ez_exit[MC][2];
  for (int i = 0; i < MC; i++) for (int j = 0; j < MC; j++) { T[i][j] = 0.0; Tvv[i][j] = 0.0; }
  for (int c = 0; c < nCorner; c++) Sigt[c] = Z.sigTotal[c0 + c];
  for (int a = 0; a < Z.nAng; a++) {
    if (Z.finish[a]) continue;
    const double om[2] = {Z.omega[2 * a], Z.omega[2 * a + 1]};
    const double quadwt = Z.weight[a], fac = Z.fac[a];
    for (int i = 0; i < MC; i++) { nxez[i] = 0; for (int j = 0; j < MC; j++) Pvv[i][j] = 0.0; }
    for (int c = 0; c < nCorner; c++) {
      const double vol = Z.Volume[c0 + c], area = Z.Area[c0 + c];
      Pvv[c][c] = vol;
      denom[c] = Sigt[c] * vol + fac * area;
      for (int c1 = 0; c1 < nCorner; c1++) Pvv[c1][c] = Pvv[c1][c] + fac * area * Tvv[c1][c];
    }
    for (int c = 0; c < nCorner; c++) {
      const int cc = c0 + c;
      for (int f = 0; f < 2; f++) {
        const double afp = dot2(om, Z.Afp + ((size_t)cc * 2 + f) * 2);
        const double aez = dot2(om, Z.Aez + ((size_t)cc * 2 + f) * 2);
        if (afp < 0.0) denom[c] -= Z.RadiusFP[cc * 2 + f] * afp;
        if (aez > 0.0) {
          const double R = Z.RadiusEZ[cc * 2 + f];
          const int cez = Z.cEZ[cc * 2 + f];
          ez_exit[c][nxez[c]] = cez; coefpsi[c][nxez[c]] = R * aez; nxez[c]++;
          denom[cez] += R * aez;
          double B1, B2;
          if (afp < 0.0) {
            const double sigA = Sigt[c] * Z.Area[cc], sigA2 = sigA * sigA;
            const double gnum = aez * aez * (FOURALPHA * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
            const double gtau = gnum / (gnum + 4.0 * sigA2 * sigA2 + aez * sigA * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
            const double B0 = 0.5 * aez * (1.0 - gtau) * R;
            B1 = (B0 - R * gtau * sigA) / Sigt[c];
            B2 = B0 / Sigt[cez];
          } else {
            B1 = 0.5 * R * aez / Sigt[c];
            B2 = 0.5 * R * aez / Sigt[cez];
          }
          Pvv[c][c] += B1; Pvv[cez][c] -= B2; Pvv[c][cez] -= B1; Pvv[cez][cez] += B2;
        }
      }
    }
    for (int i = 0; i < nCorner; i++) {
      const int c = Z.nextC[(size_t)a * Z.nc + c0 + i];
      const double dInv = 1.0 / denom[c];
      for (int c1 = 0; c1 < nCorner; c1++) Pvv[c1][c] = dInv * Pvv[c1][c];
      for (int k = 0; k < nxez[c]; k++) {
        const int cez = ez_exit[c][k];
        const double coef = coefpsi[c][k];
        for (int c1 = 0; c1 < nCorner; c1++) Pvv[c1][cez] += coef * Pvv[c1][c];
      }
    }
    if (Z.start[a]) {
      for (int c = 0; c < nCorner; c++) for (int c1 = 0; c1 < nCorner; c1++) Tvv[c1][c] = Pvv[c1][c];
    } else {
      const double w1 = Z.w1[a], w2 = Z.w2[a];
      for (int c = 0; c < nCorner; c++)
        for (int c1 = 0; c1 < nCorner; c1++) {
          T[c1][c] = T[c1][c] + quadwt * Pvv[c1][c];
          Tvv[c1][c] = w1 * Pvv[c1][c] - w2 * Tvv[c1][c];
        }
    }
  }
  for (int c = 0; c < nCorner; c++)
    for (int c1 = 0; c1 < mC; c1++) Z.TT[(size_t)(c0 + c) * mC + c1] = c1 < nCorner ? T[c1][c] : 0.0;
}

template <class T>
int dalloc2(umt_ctx *ctx, T **p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  UMT_CUDA(ctx, cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(n, 1)));
  UMT_CUDA(ctx, cudaMemsetAsync(*p, 0, sizeof(T) * std::max<size_t>(n, 1), ctx->stream));
  return UMT_OK;
}

}  // namespace

// angle set, sweep order and work items of the r-z grey sweeps; fills the parts of GtaState that umt_gta_setup (gta.cu)
// turns into device arrays (omega, weight, nextZ, nextC, items) and uploads the r-z specific ones
int umt_gta_setup_rz(umt_ctx *ctx) {
  GtaState &g = ctx->gta;
  if (ctx->maxCorner > MAXC2 || ctx->maxcf != 2) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_gta_setup: r-z needs maxCorner <= 8 and maxcf == 2");
  if (!ctx->d_Area || !ctx->d_RadiusFP || !ctx->d_RadiusEZ) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_setup: r-z geometry (Area, RadiusFP, RadiusEZ) not set");
  if (umt_host_gta_quadrature_rz(g.omega, g.weight, g.start, g.finish, g.angDerivFac, g.tauW1, g.tauW2))
    UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_setup: r-z GTA quadrature failed");
  g.nAng = (int)g.weight.size();
  TRY(umt_host_build_order(ctx, g.omega.data(), g.nAng, g.nHyp, g.zonesInPlane, g.nextZ, g.nextC));
  for (int a = 0; a < g.nAng; a++)
    if (g.finish[a]) { g.nHyp[a] = 0; g.zonesInPlane[a].clear(); }   // finishing directions are not swept (GTASweep.F90:149)
  return UMT_OK;
}

// items + r-z device arrays; called by umt_gta_setup after the common arrays exist
int umt_gta_finish_setup_rz(umt_ctx *ctx, std::vector<WorkItem> &items) {
  GtaState &g = ctx->gta;
  TRY(umt_build_items_rz_set(ctx->nz, g.nAng, g.nHyp, g.zonesInPlane, g.nextZ, g.start, GRZ_BLOCK, items, g.level, g.nLevels, g.maxHyp));
  TRY(dalloc2(ctx, &g.d_start, g.nAng)); TRY(dalloc2(ctx, &g.d_finish, g.nAng)); TRY(dalloc2(ctx, &g.d_level, g.nAng));
  TRY(dalloc2(ctx, &g.d_fac, g.nAng)); TRY(dalloc2(ctx, &g.d_w1, g.nAng)); TRY(dalloc2(ctx, &g.d_w2, g.nAng));
  TRY(dalloc2(ctx, &g.d_psim, (size_t)g.nLevels * ctx->nc)); TRY(dalloc2(ctx, &g.d_tinc, (size_t)g.nLevels * ctx->nc));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_start, g.start.data(), g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_finish, g.finish.data(), g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_level, g.level.data(), sizeof(int) * g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_fac, g.angDerivFac.data(), sizeof(double) * g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_w1, g.tauW1.data(), sizeof(double) * g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_w2, g.tauW2.data(), sizeof(double) * g.nAng, cudaMemcpyHostToDevice));
  // tables of the chain kernel
  int maxAng = 1, maxPlane = 1;
  std::vector<int> cnt(g.nLevels, 0);
  for (int a = 0; a < g.nAng; a++) if (g.nHyp[a] > 0) maxAng = std::max(maxAng, ++cnt[g.level[a]]);
  std::vector<int> la((size_t)g.nLevels * maxAng, -1), po((size_t)g.nAng * (g.maxHyp + 1), 0), nh(g.nAng, 0);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int a = 0; a < g.nAng; a++) {
    nh[a] = g.nHyp[a];
    if (nh[a] == 0) continue;
    la[(size_t)g.level[a] * maxAng + cnt[g.level[a]]++] = a;
    int o = 0;
    for (int p = 0; p < nh[a]; p++) { po[(size_t)a * (g.maxHyp + 1) + p] = o; o += g.zonesInPlane[a][p]; maxPlane = std::max(maxPlane, g.zonesInPlane[a][p]); }
    po[(size_t)a * (g.maxHyp + 1) + nh[a]] = o;
  }
  g.rz_maxAngLevel = maxAng;
  g.rz_threads = std::max(32, std::min(256, (maxPlane + 31) / 32 * 32));
  g.rz_chain = false;   // measured at 38 k zones: item kernel 6.0 ms per grey sweep (static half overlaps the dependency wait), chain kernel 12.4 ms
  if (const char *e = getenv("UMT_GTA_RZ_KERNEL")) g.rz_chain = std::string(e) == "chain";
  {   // dataflow kernel: previous swept angle of each level; not with direct-solve zones
    std::vector<int> prevA(g.nAng, -1), lastOf(g.nLevels, -1);
    bool plain = true;
    for (int a = 0; a < g.nAng; a++) {
      if (nh[a] == 0) continue;
      prevA[a] = lastOf[g.level[a]]; lastOf[g.level[a]] = a;
      for (int z : g.nextZ[a]) if (z < 0) { plain = false; break; }
    }
    g.rz_flow = plain;   // measured at 38 k zones: 4.48 ms per grey sweep against 6.00 ms with the item kernel (UMT_GTA_RZ_KERNEL=item)
    if (const char *e = getenv("UMT_GTA_RZ_KERNEL")) g.rz_flow = plain && std::string(e) == "flow";
    TRY(dalloc2(ctx, &g.d_prevAngle, prevA.size()));
    UMT_CUDA(ctx, umt_memcpy(ctx, g.d_prevAngle, prevA.data(), sizeof(int) * prevA.size(), cudaMemcpyHostToDevice));
    TRY(dalloc2(ctx, &g.d_psimA, (size_t)g.nAng * ctx->nc)); TRY(dalloc2(ctx, &g.d_tincA, (size_t)g.nAng * ctx->nc));
  }
  TRY(dalloc2(ctx, &g.d_levelAngles, la.size())); TRY(dalloc2(ctx, &g.d_planeOff, po.size())); TRY(dalloc2(ctx, &g.d_nHyp, nh.size()));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_levelAngles, la.data(), sizeof(int) * la.size(), cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_planeOff, po.data(), sizeof(int) * po.size(), cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_nHyp, nh.data(), sizeof(int) * nh.size(), cudaMemcpyHostToDevice));
  return UMT_OK;
}

int umt_gta_launch_sweep_rz(umt_ctx *ctx) {
  GtaState &g = ctx->gta;
  const int nc = ctx->nc;
  // GTASweep.F90:113-119: tPsiM = tInc = 0; pInc of the finishing directions stays 0
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_psim, 0, sizeof(double) * (size_t)g.nLevels * nc, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_tinc, 0, sizeof(double) * (size_t)g.nLevels * nc, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_pinc, 0, sizeof(double) * (size_t)g.nAng * nc, ctx->stream));
  GtaRZParams P;
  P.nc = nc; P.nb = ctx->nb; P.nz = ctx->nz; P.nItems = g.nItems;
  P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
  P.Volume = ctx->d_Volume; P.Area = ctx->d_Area; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez; P.RadiusFP = ctx->d_RadiusFP; P.RadiusEZ = ctx->d_RadiusEZ;
  P.omega = g.d_omega; P.fac = g.d_fac; P.w1 = g.d_w1; P.w2 = g.d_w2; P.start = g.d_start; P.level = g.d_level;
  P.nextZ = g.d_nextZ; P.nextC = g.d_nextC; P.items = g.d_items; P.counters = g.d_counters;
  P.sigTotal = g.d_sigTotal; P.sigtInv = g.d_sigtInv; P.tsa = g.d_tsaSource; P.tpsi = g.d_tpsi; P.pinc = g.d_pinc; P.psim = g.d_psim; P.tinc = g.d_tinc;
  if (g.rz_chain) {
    auto ck = ctx->maxCorner <= 4 ? gta_sweep_rz_chain_kernel<4> : gta_sweep_rz_chain_kernel<MAXC2>;
    ck<<<g.nLevels, g.rz_threads, 0, ctx->stream>>>(P, g.d_levelAngles, g.rz_maxAngLevel, g.d_planeOff, g.maxHyp + 1, g.d_nHyp);
    UMT_CUDA(ctx, cudaGetLastError());
    return UMT_OK;
  }
  if (g.rz_flow && g.nStagesR <= 1) {
    P.abortFlag = ctx->d_abort; P.spinLimit = ctx->spinLimit;
    P.psimA = g.d_psimA; P.tincA = g.d_tincA; P.prevAngle = g.d_prevAngle; P.nHyp = g.d_nHyp;
    gta_rz_mark_kernel<<<dim3(std::max(1, std::min(ctx->sm_count, (nc + 255) / 256)), g.nAng), 256, 0, ctx->stream>>>(g.d_tpsi, g.d_psimA, g.d_tincA, g.d_nHyp, nc, nc + ctx->nb);
    UMT_CUDA(ctx, cudaGetLastError());
    void (*fk)(GtaRZParams, int) = ctx->maxCorner <= 4 ? gta_sweep_rz_flow_kernel<4> : gta_sweep_rz_flow_kernel<MAXC2>;
    int occF = 0;
    UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occF, fk, GRZ_BLOCK, 0));
    const int wpi = GRZ_BLOCK / 32;
    P.items = g.d_items; P.nItems = g.nItems;
    const int grid = std::max(1, std::min(ctx->sm_count * std::max(occF, 1), (g.nItems * wpi + wpi - 1) / wpi));
    fk<<<grid, GRZ_BLOCK, 0, ctx->stream>>>(P, wpi);
    UMT_CUDA(ctx, cudaGetLastError());
    UMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));   // checked after the next sync
    return UMT_OK;
  }
  void (*kern)(GtaRZParams) = ctx->maxCorner <= 4 ? gta_sweep_rz_kernel<4> : gta_sweep_rz_kernel<MAXC2>;
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, GRZ_BLOCK, 0));
  for (int sR = 0; sR < g.nStagesR; sR++) {   // one stage unless the domain has reflecting boundaries
    const int ib = g.stageItemBegin[sR], ie = g.stageItemBegin[sR + 1];
    const int ob = g.reflOpBegin[sR], oe = g.reflOpBegin[sR + 1];
    if (oe > ob) {   // snreflect: PsiB(b, Minc) <- PsiB(b, Mref), the tail rows of tPsi
      int maxN = 1;
      for (const auto &R : ctx->refl) maxN = std::max(maxN, R.n);
      gta_reflect_rows_kernel<<<dim3((maxN + 255) / 256, oe - ob), 256, 0, ctx->stream>>>(g.d_tpsi, g.d_reflOps + ob, nc + ctx->nb, nc);
    }
    if (ie == ib) continue;
    if (sR > 0) UMT_CUDA(ctx, cudaMemsetAsync(g.d_counters, 0, sizeof(int), ctx->stream));   // the ticket; plane counters persist
    P.items = g.d_items + ib; P.nItems = ie - ib;
    const int grid = std::max(1, std::min(ctx->sm_count * std::max(occ, 1), P.nItems));
    kern<<<grid, GRZ_BLOCK, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
  }
  return UMT_OK;
}

int umt_gta_launch_init_tt_rz(umt_ctx *ctx) {
  GtaState &g = ctx->gta;
  TTRZParams Z;
  Z.nz = ctx->nz; Z.nc = ctx->nc; Z.mC = ctx->maxCorner; Z.nAng = g.nAng;
  Z.numCorner = ctx->d_numCorner; Z.cOffSet = ctx->d_cOffSet; Z.cEZ = ctx->d_cEZ;
  Z.nextC = g.d_nextC; Z.start = g.d_start; Z.finish = g.d_finish;
  Z.Volume = ctx->d_Volume; Z.Area = ctx->d_Area; Z.Afp = ctx->d_Afp; Z.Aez = ctx->d_Aez; Z.RadiusFP = ctx->d_RadiusFP; Z.RadiusEZ = ctx->d_RadiusEZ;
  Z.omega = g.d_omega; Z.weight = g.d_weight; Z.fac = g.d_fac; Z.w1 = g.d_w1; Z.w2 = g.d_w2; Z.sigTotal = g.d_sigTotal; Z.TT = g.d_TT;
  void (*kern)(TTRZParams) = ctx->maxCorner <= 4 ? gta_init_tt_rz_kernel<4> : gta_init_tt_rz_kernel<MAXC2>;
  kern<<<(ctx->nz + 63) / 64, 64, 0, ctx->stream>>>(Z);
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}
