// 2-D (r,z) upstream-corner-balance sweep for sm_100a.
//
// Replaces snac/SweepUCBrz.F90:11-282 (one angle, one set) and the angle loop of
// snac/SetSweep.F90:113-170 for ndim == 2.  One persistent kernel sweeps every
// non-finishing angle.  Work items are (angle, hyperplane, chunk of zones); CTAs
// pull them through an atomic ticket.  Two dependencies gate an item:
//   * its own angle's previous hyperplane (upstream Psi1 across FP faces);
//   * the angular-derivative chain: angles of one xi-level are coupled through the
//     half-angle intensity PsiM(:,c) (SweepUCBrz.F90:212-240), so zone z of angle
//     k+1 of a level needs zone z of angle k.  The host turns that into "the latest
//     plane of the previous angle that holds a zone of this plane is complete"
//     (planes of one angle complete in order), so consecutive angles of a level
//     pipeline through the mesh instead of running back to back, and the 4P levels
//     are independent of each other.
// One thread owns one (zone, group): group index on consecutive lanes, so all loads
// and stores of Psi, STotal, Psi1, PsiM, PsiB are contiguous G*8-byte rows.
// Exiting boundary fluxes (SweepUCBrz.F90:245-266) and the finishing direction's
// Psi/PsiB <- PsiM are written by the thread that owns the corner.
#include <algorithm>
#include <cstdlib>

#include "umt_internal.h"

namespace {

constexpr int MAXC2 = 8;   // corners per zone (quads: 4; general polygons up to 8)
constexpr int RZ_BLOCK = 64;   // threads per CTA = (zone, group) pairs per work item: small items keep the per-plane latency low
                               // (measured on the 40x40-tile r-z mesh, G = 64: 64 pairs 24.5 ms, 128 pairs 36 ms)
constexpr double FOURALPHA = 1.82;   // SweepUCBrz.F90:88

struct SweepRZParams {
  int nc, nb, nz, G, NA, nItems;
  double tau;
  const int *numCorner, *cOffSet, *cFP /* 0-based row; >= nc: boundary */, *cEZ /* 0-based */;
  const double *Volume, *Area, *Afp, *Aez, *RadiusFP, *RadiusEZ, *omega;
  const double *angDerivFac, *tauW1, *tauW2;
  const unsigned char *start, *finishNext;   // finishNext[a] = FinishingDirection(a+1)
  const int *level;                          // xi-level of each angle
  const int *nextZ;
  const unsigned char *nextC;
  const WorkItem *items;
  int *counters;
  const double *psi, *stotal, *sigt;
  double *psi1, *psim;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Everything of one (zone, group) solve that does not depend on other work items: loaded and computed while the CTA's
// dependencies are still being resolved, so that after the wait only the upstream fluxes and PsiM remain to be read.
template <int MC>
struct ZoneStatic {
  double Q[MC], srcInit[MC], sumArea[MC], volSig[MC], areaFac[MC];
  double afp[MC][2], aez[MC][2], Rafp[MC][2], Raez[MC][2];
  int row[MC][2], cez[MC][2];
  int nCorner, c0;
  double sig;
};

template <int MC>
__device__ __forceinline__ void zone_static_rz(const SweepRZParams &P, int a, int zone0, int g, ZoneStatic<MC> &Z) {
  const int G = P.G, nc = P.nc;
  const double om0 = P.omega[2 * a], om1 = P.omega[2 * a + 1];
  const size_t slab = (size_t)(nc + P.nb) * G;
  const double *psiA = P.psi + (size_t)a * slab;
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  const double sig = P.sigt[(size_t)zone * G + g];
  const double fac = P.angDerivFac[a];
  Z.nCorner = nCorner; Z.c0 = c0; Z.sig = sig;
#pragma unroll
  for (int c = 0; c < MC; c++) {
    if (c < nCorner) {
      const size_t r = (size_t)(c0 + c) * G + g;
      const double source = P.stotal[r] + P.tau * psiA[r];
      const double area = P.Area[c0 + c], vol = P.Volume[c0 + c];
      Z.Q[c] = source;
      Z.srcInit[c] = vol * source;
      Z.sumArea[c] = fac * area;
      Z.volSig[c] = sig * vol;
      Z.areaFac[c] = area * fac;
    }
  }
#pragma unroll
  for (int c = 0; c < MC; c++) {
    if (c < nCorner) {
      const int cc = c0 + c;
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double *Af = P.Afp + ((size_t)cc * 2 + f) * 2;
        const double *Ae = P.Aez + ((size_t)cc * 2 + f) * 2;
        const double afp = __dadd_rn(__dmul_rn(om0, Af[0]), __dmul_rn(om1, Af[1]));
        const double aez = __dadd_rn(__dmul_rn(om0, Ae[0]), __dmul_rn(om1, Ae[1]));
        Z.afp[c][f] = afp; Z.aez[c][f] = aez;
        Z.row[c][f] = P.cFP[cc * 2 + f];
        Z.cez[c][f] = P.cEZ[cc * 2 + f];
        Z.Rafp[c][f] = 0.0; Z.Raez[c][f] = 0.0;
        if (afp < 0.0) {
          Z.Rafp[c][f] = P.RadiusFP[cc * 2 + f] * afp;
          Z.sumArea[c] -= Z.Rafp[c][f];
        }
        if (aez > 0.0) {
          Z.Raez[c][f] = P.RadiusEZ[cc * 2 + f] * aez;
          Z.sumArea[Z.cez[c][f]] += Z.Raez[c][f];
        }
      }
    }
  }
}

// SweepUCBrz.F90:103-243 for one (zone, group), second half: upstream fluxes, closure, corner solves, PsiM, exits.
// Accumulation order into src is the reference's (corner-major, face-minor).
template <int MC>
__device__ __forceinline__ void zone_solve_rz(const SweepRZParams &P, int a, int g, const ZoneStatic<MC> &Z) {
  const int G = P.G, nc = P.nc;
  const size_t slab = (size_t)(nc + P.nb) * G;
  double *psi1A = P.psi1 + (size_t)a * slab;
  double *psimL = P.psim + (size_t)P.level[a] * nc * G;
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  const int nCorner = Z.nCorner, c0 = Z.c0;
  const double sig = Z.sig;
  double src[MC], psifp[MC][2], pm[MC];
#pragma unroll
  for (int c = 0; c < MC; c++) {
    if (c < nCorner) {
      src[c] = Z.srcInit[c];
      pm[c] = psimL[(size_t)(c0 + c) * G + g];
#pragma unroll
      for (int f = 0; f < 2; f++) psifp[c][f] = Z.afp[c][f] < 0.0 ? __ldcg(&psi1A[(size_t)Z.row[c][f] * G + g]) : 0.0;
    }
  }
  for (int c = 0; c < nCorner; c++) {
    const double area = P.Area[c0 + c];
#pragma unroll
    for (int f = 0; f < 2; f++) {
      const double afp = Z.afp[c][f], aez = Z.aez[c][f];
      if (afp < 0.0) src[c] -= Z.Rafp[c][f] * psifp[c][f];
      if (aez > 0.0) {
        const int cez = Z.cez[c][f];
        double sez;
        if (afp < 0.0) {
          const double R = P.RadiusEZ[(c0 + c) * 2 + f];
          const double sigA = sig * area, sigA2 = sigA * sigA;
          const double gnum = aez * aez * (FOURALPHA * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
          const double gden = area * (4.0 * sigA * sigA2 + aez * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
          sez = R * (area * gnum * (sig * psifp[c][f] - Z.Q[c]) + 0.5 * aez * gden * (Z.Q[c] - Z.Q[cez])) / (gnum + gden * sig);
        } else {
          sez = 0.5 * Z.Raez[c][f] * (Z.Q[c] - Z.Q[cez]) / sig;
        }
        src[c] += sez;
        src[cez] -= sez;
      }
    }
  }
  for (int i = 0; i < nCorner; i++) {
    const int c = nextC[c0 + i];
    const double p = (src[c] + Z.areaFac[c] * pm[c]) / (Z.sumArea[c] + Z.volSig[c]);
    src[c] = p;   // src now holds the corner flux
#pragma unroll
    for (int f = 0; f < 2; f++)
      if (Z.aez[c][f] > 0.0) src[Z.cez[c][f]] += Z.Raez[c][f] * p;
  }
  // half-angle intensity for the next angle of the level; exiting boundary fluxes; finishing direction
  const bool starting = P.start[a] != 0, fin = P.finishNext[a] != 0;
  double *psi1N = psi1A + slab;   // slab of angle a+1 (only touched when it is a finishing direction)
  const double w1 = P.tauW1[a], w2 = P.tauW2[a];
  for (int c = 0; c < nCorner; c++) {
    const int cc = c0 + c;
    const size_t r = (size_t)cc * G + g;
    const double p = src[c];
    const double pmn = starting ? p : w1 * p - w2 * pm[c];
    psimL[r] = pmn;
    psi1A[r] = p;
    if (fin) psi1N[r] = pmn;   // Psi(:,c,Angle+1) <- PsiM (becomes Psi when the buffers trade roles on savePsi)
#pragma unroll
    for (int f = 0; f < 2; f++) {
      const int row = Z.row[c][f];
      if (row >= nc && Z.afp[c][f] > 0.0) {
        psi1A[(size_t)row * G + g] = p;             // PsiB(:,b,Angle)   <- Psi1(:,c)
        if (fin) psi1N[(size_t)row * G + g] = pmn;  // PsiB(:,b,Angle+1) <- PsiM(:,c)
      }
    }
  }
}

// One (zone, group) pair per thread (items hold at most blockDim pairs): the static half runs before the dependency wait.
template <int MC>
__global__ void __launch_bounds__(RZ_BLOCK) sweeprz_kernel(SweepRZParams P) {
  __shared__ int s_item;
  const int G = P.G;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&P.counters[0], 1);   // taken only when free to work on it: a ticket held ahead of time blocks a ready item behind a waiting one
    __syncthreads();
    const int it = s_item;
    if (it >= P.nItems) break;
    const WorkItem w = P.items[it];
    const int npairs = (w.zend - w.zbeg) * G;
    const int zi = threadIdx.x / G, g = threadIdx.x - zi * G;
    const bool fits = npairs <= (int)blockDim.x;   // false only for G > blockDim (one zone per item, several groups per thread)
    const bool active = fits && (int)threadIdx.x < npairs;
    ZoneStatic<MC> Z;
    if (active) zone_static_rz<MC>(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + w.zbeg + zi], g, Z);
    if (threadIdx.x == blockDim.x - 1) {   // the last thread polls (it is the one most likely to have no pair)
      if (w.wait_idx >= 0)
        while (ld_acquire(&P.counters[1 + w.wait_idx]) < w.wait_count) __nanosleep(20);
      if (w.pad0 >= 0)   // second dependency: the previous angle of this xi-level (PsiM chain)
        while (ld_acquire(&P.counters[1 + w.pad0]) < w.pad1) __nanosleep(20);
    }
    __syncthreads();
    if (active) zone_solve_rz<MC>(P, w.angle, g, Z);
    if (!fits)
      for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
        const int z2 = idx / G, g2 = idx - z2 * G;
        zone_static_rz<MC>(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + w.zbeg + z2], g2, Z);
        zone_solve_rz<MC>(P, w.angle, g2, Z);
      }
    __syncthreads();
    if (threadIdx.x == 0) {   // publish the CTA's rows (acq_rel is enough: the writes were observed through the barrier) and signal
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + w.signal_idx]) : "memory");
    }
  }
}

}  // namespace

// Work items of the RZ sweep in a topological order of both dependencies (host side, once per schedule).
int umt_build_items_rz(umt_ctx *ctx, std::vector<WorkItem> &items, int zpi) {
  int maxHyp = 0;
  return umt_build_items_rz_set(ctx->nz, ctx->NA, ctx->nHyp, ctx->zonesInPlane, ctx->nextZ, ctx->h_start, zpi, items, ctx->h_level, ctx->nLevels, maxHyp);
}

// the same for any r-z angle set (the Sn set above, the GTA set in gta_rz.cu); angles with nHyp == 0 (finishing directions) get no items
int umt_build_items_rz_set(int nz, int NA, const std::vector<int> &nHypV, const std::vector<std::vector<int>> &zonesInPlaneV,
                           const std::vector<std::vector<int>> &nextZV, const std::vector<unsigned char> &startV, int zpi,
                           std::vector<WorkItem> &items, std::vector<int> &levelOut, int &nLevelsOut, int &maxHypOut) {
  struct { const std::vector<int> &nHyp; const std::vector<std::vector<int>> &zonesInPlane, &nextZ; const std::vector<unsigned char> &h_start;
           std::vector<int> h_level; int nLevels; } cx{nHypV, zonesInPlaneV, nextZV, startV, {}, 0};
  auto *ctx = &cx;
  int maxHyp = 0;
  for (int a = 0; a < NA; a++) maxHyp = std::max(maxHyp, ctx->nHyp[a]);
  maxHypOut = maxHyp;
  // xi-levels: a level starts at a starting direction; finishing directions are not swept
  std::vector<int> level(NA, 0), prev(NA, -1);
  int lev = -1, last = -1;
  for (int a = 0; a < NA; a++) {
    if (ctx->h_start[a] || lev < 0) { lev++; last = -1; }
    level[a] = lev;
    if (ctx->nHyp[a] == 0) continue;
    prev[a] = last;
    last = a;
  }
  ctx->h_level = level;
  ctx->nLevels = lev + 1;
  std::vector<std::vector<int>> planeOf(NA), nItemsPlane(NA), planeStart(NA), tdone(NA), dep2(NA);
  struct Key { int t, a, p; };
  std::vector<Key> keys;
  for (int a = 0; a < NA; a++) {
    const int nh = ctx->nHyp[a];
    if (nh == 0) continue;
    planeOf[a].assign(nz, 0);
    planeStart[a].assign(nh + 1, 0);
    nItemsPlane[a].assign(nh, 0);
    tdone[a].assign(nh, 0);
    dep2[a].assign(nh, -1);
    for (int p = 0; p < nh; p++) {
      const int n = ctx->zonesInPlane[a][p];
      planeStart[a][p + 1] = planeStart[a][p] + n;
      nItemsPlane[a][p] = (n + zpi - 1) / zpi;
      for (int i = planeStart[a][p]; i < planeStart[a][p + 1]; i++) planeOf[a][std::abs(ctx->nextZ[a][i]) - 1] = p;
    }
    const int pa = prev[a];
    for (int p = 0; p < nh; p++) {
      int t = p > 0 ? tdone[a][p - 1] : 0;
      if (pa >= 0) {
        int q = 0;
        for (int i = planeStart[a][p]; i < planeStart[a][p + 1]; i++) q = std::max(q, planeOf[pa][std::abs(ctx->nextZ[a][i]) - 1]);
        dep2[a][p] = q;
        t = std::max(t, tdone[pa][q]);
      }
      tdone[a][p] = t + 1;
      keys.push_back({t + 1, a, p});
    }
  }
  std::stable_sort(keys.begin(), keys.end(), [](const Key &x, const Key &y) { return x.t < y.t; });
  items.clear();
  for (const Key &k : keys) {
    const int a = k.a, p = k.p;
    const int n = ctx->zonesInPlane[a][p], z0 = planeStart[a][p];
    for (int j = 0; j < nItemsPlane[a][p]; j++) {
      WorkItem w;
      w.angle = a;
      w.zbeg = z0 + j * zpi;
      w.zend = std::min(z0 + n, w.zbeg + zpi);
      w.wait_idx = p > 0 ? a * maxHyp + p - 1 : -1;
      w.wait_count = p > 0 ? nItemsPlane[a][p - 1] : 0;
      w.signal_idx = a * maxHyp + p;
      w.pad0 = dep2[a][p] >= 0 ? prev[a] * maxHyp + dep2[a][p] : -1;
      w.pad1 = dep2[a][p] >= 0 ? nItemsPlane[prev[a]][dep2[a][p]] : 0;
      items.push_back(w);
    }
  }
  levelOut = ctx->h_level;
  nLevelsOut = ctx->nLevels;
  return UMT_OK;
}

int umt_launch_sweeprz(umt_ctx *ctx, int /*savePsi*/) {
  if (ctx->maxCorner > MAXC2 || ctx->maxcf != 2)
    UMT_FAIL(ctx, UMT_ERR_ARG, "RZ sweep supports maxCorner <= %d and maxcf == 2 (got %d, %d)", MAXC2, ctx->maxCorner, ctx->maxcf);
  if (!ctx->d_level || !ctx->d_psim) UMT_FAIL(ctx, UMT_ERR_STATE, "RZ sweep: schedule/state not finalized");
  SweepRZParams P;
  P.nc = ctx->nc; P.nb = ctx->nb; P.nz = ctx->nz; P.G = ctx->G; P.NA = ctx->NA; P.nItems = ctx->nItems;
  P.tau = ctx->tau;
  P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
  P.Volume = ctx->d_Volume; P.Area = ctx->d_Area; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez;
  P.RadiusFP = ctx->d_RadiusFP; P.RadiusEZ = ctx->d_RadiusEZ; P.omega = ctx->d_omega;
  P.angDerivFac = ctx->d_angDerivFac; P.tauW1 = ctx->d_tauW1; P.tauW2 = ctx->d_tauW2;
  P.start = ctx->d_start; P.finishNext = ctx->d_finishNext; P.level = ctx->d_level;
  P.nextZ = ctx->d_nextZ; P.nextC = ctx->d_nextC; P.items = ctx->d_items; P.counters = ctx->d_counters;
  P.psi = ctx->d_psi; P.stotal = ctx->d_stotal; P.sigt = ctx->d_sigt; P.psi1 = ctx->d_psi1; P.psim = ctx->d_psim;
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int) * (1 + ctx->nCounters), ctx->stream));
  // Set%PsiM = 0 at the start of every flux pass (SetSweep.F90:94-96)
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_psim, 0, sizeof(double) * (size_t)ctx->nLevels * ctx->nc * ctx->G, ctx->stream));
  void (*kern)(SweepRZParams) = ctx->maxCorner <= 4 ? sweeprz_kernel<4> : sweeprz_kernel<MAXC2>;
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RZ_BLOCK, 0));
  if (occ < 1) occ = 1;
  const int grid = std::max(1, std::min(ctx->sm_count * occ, ctx->nItems));
  kern<<<grid, RZ_BLOCK, 0, ctx->stream>>>(P);
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches += 1;
  return UMT_OK;
}
