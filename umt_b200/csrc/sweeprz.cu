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
#include <string>

#include "umt_internal.h"
#include "device_util.h"

namespace {

constexpr int MAXC2 = 8;   // corners per zone (quads: 4; general polygons up to 8)
constexpr int RZ_BLOCK = 64;   // threads per CTA = (zone, group) pairs per work item: small items keep the per-plane latency low
                               // (measured on the 40x40-tile r-z mesh, G = 64: 64 pairs 24.5 ms, 128 pairs 36 ms)
constexpr double FOURALPHA = 1.82;   // SweepUCBrz.F90:88
#ifndef RZ_MIN_CTAS
#define RZ_MIN_CTAS 1
#endif

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
  // chain kernel: one CTA per (xi-level, block of gb groups)
  int gb, nGroupBlocks, maxAngLevel, hypStride;
  const int *levelAngles;   // (nLevels, maxAngLevel) swept angles of each level in order, -1 padded
  const int *planeOff;      // (NA, hypStride) first zone of each plane in nextZ(:,a); planeOff[a][nHyp[a]] = nz
  const int *nHyp;          // (NA)
  // dataflow kernel
  double *psimA;            // (NA, nc, G) PsiM as written by angle a
  const int *prevAngle;     // (NA) previous swept angle of the level, -1: none (starting direction)
  int *abortFlag;           // watchdog of the polling loops (device_util.h)
  unsigned spinLimit;
};

// Dataflow variant: a quiet NaN with a payload no arithmetic produces marks "not computed yet".  The corner rows of Psi1 and the
// per-angle PsiM slabs are filled with it before the launch; a consumer polls the values it needs until they are real, so the
// data are their own completion flags: no counters, no fences (every 8-byte value is validated by itself), no barriers.
constexpr unsigned long long RZ_SENTINEL = UMT_SENTINEL;   // device_util.h

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Dynamic corner indices (cEZ, nextC) would put the per-thread zone arrays in local memory, and with a few hundred threads per SM
// those 700-byte frames fall out of L1: every "array" access then costs an L2 round trip inside a dependent chain.  All loops
// over corners and faces are therefore fully unrolled and the few dynamic accesses go through select chains, which keeps the
// whole zone in registers.
// Everything of one (zone, group) solve that does not depend on other work items, computed while the CTA's dependencies are still
// being resolved.  The corner-balance closure is linear in the upstream fluxes u = Psi1(row) and in PsiM, so the static half goes
// all the way to that linear form -- including every division -- and the half on the dependency chain is a handful of FMAs:
//   src(c)  = srcS(c) + sum_f k1(c,f) u(c,f)            (k1 = -R_fp afp + A1: incident face, and its share of the EZ closure)
//   src(cez(c,f)) += k2(c,f) u(c,f)                     (k2 = -A1)
//   psi(c_i) = (src(c_i) + areaFac_i PsiM(c_i)) inv_i,  src(dz_i[f]) += rz_i[f] psi(c_i)    for the corners c_i in nextC order
// with sez(c,f) = A1 u + A0 (SweepUCBrz.F90:170-205; A0 is folded into srcS).  Same arithmetic as the reference up to the order in
// which the terms of src are added and the reciprocal of the denominator.
template <int MC>
struct ZoneStatic {
  double srcS[MC], k1[MC][2], k2[MC][2];
  double areaFac[MC], inv[MC], rz[MC][2];   // areaFac: by corner; inv, rz: by solve step
  int row[MC][2], cez[MC][2], dz[MC][2], ci[MC];
  unsigned inMask, exitMask;   // bit 2c+f: omega.A_fp < 0 (incident) / omega.A_fp > 0 on a boundary face (exiting)
  int nCorner, c0;
};

template <int MC>
__device__ __forceinline__ void zone_static_rz(const SweepRZParams &P, int a, int zone0, int g, ZoneStatic<MC> &Z) {
  const int G = P.G, nc = P.nc;
  const double om0 = P.omega[2 * a], om1 = P.omega[2 * a + 1];
  const size_t slab = (size_t)(nc + P.nb) * G;
  const double *psiA = P.psi + (size_t)a * slab;
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  const double sig = P.sigt[(size_t)zone * G + g];
  const double fac = P.angDerivFac[a];
  Z.nCorner = nCorner; Z.c0 = c0; Z.inMask = 0u; Z.exitMask = 0u;
  double Q[MC], sumArea[MC], volSig[MC], area[MC], aez[MC][2], Raez[MC][2];
#pragma unroll
  for (int c = 0; c < MC; c++) {
    Q[c] = 0.0; Z.srcS[c] = 0.0; sumArea[c] = 1.0; volSig[c] = 0.0; Z.areaFac[c] = 0.0; area[c] = 0.0; Z.ci[c] = c;
    if (c < nCorner) {
      const size_t r = (size_t)(c0 + c) * G + g;
      const double source = P.stotal[r] + P.tau * psiA[r];
      const double ar = P.Area[c0 + c], vol = P.Volume[c0 + c];
      Q[c] = source;
      Z.srcS[c] = vol * source;
      sumArea[c] = fac * ar;
      volSig[c] = sig * vol;
      Z.areaFac[c] = ar * fac;
      area[c] = ar;
      Z.ci[c] = nextC[c0 + c];
    }
  }
  double afpv[MC][2], rfp[MC][2], rez[MC][2];
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++) {
      aez[c][f] = 0.0; afpv[c][f] = 0.0; rfp[c][f] = 0.0; rez[c][f] = 0.0; Raez[c][f] = 0.0;
      Z.k1[c][f] = 0.0; Z.k2[c][f] = 0.0; Z.row[c][f] = 0; Z.cez[c][f] = 0;
      if (c < nCorner) {
        const int cc = c0 + c;
        const double *Af = P.Afp + ((size_t)cc * 2 + f) * 2;
        const double *Ae = P.Aez + ((size_t)cc * 2 + f) * 2;
        afpv[c][f] = __dadd_rn(__dmul_rn(om0, Af[0]), __dmul_rn(om1, Af[1]));
        aez[c][f] = __dadd_rn(__dmul_rn(om0, Ae[0]), __dmul_rn(om1, Ae[1]));
        Z.row[c][f] = P.cFP[cc * 2 + f];
        Z.cez[c][f] = P.cEZ[cc * 2 + f];
        rfp[c][f] = P.RadiusFP[cc * 2 + f];
        rez[c][f] = P.RadiusEZ[cc * 2 + f];
      }
    }
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++)
      if (c < nCorner) {
        const double afp = afpv[c][f], az = aez[c][f];
        const bool inc = afp < 0.0;
        if (inc) {
          Z.inMask |= 1u << (2 * c + f);
          const double Rafp = rfp[c][f] * afp;
          Z.k1[c][f] = -Rafp;
          sumArea[c] -= Rafp;
        } else if (afp > 0.0 && Z.row[c][f] >= nc) Z.exitMask |= 1u << (2 * c + f);
        if (az > 0.0) {
          const int cez = Z.cez[c][f];
          const double R = rez[c][f], qcez = pick<MC>(Q, cez);
          Raez[c][f] = R * az;
          addto<MC>(sumArea, cez, Raez[c][f]);
          double A0;
          if (inc) {
            const double ar = area[c];
            const double sigA = sig * ar, sigA2 = sigA * sigA;
            const double gnum = az * az * (FOURALPHA * sigA2 + az * (4.0 * sigA + 3.0 * az));
            const double gden = ar * (4.0 * sigA * sigA2 + az * (6.0 * sigA2 + 2.0 * az * (2.0 * sigA + az)));
            const double rd = R / (gnum + gden * sig);
            const double A1 = rd * (ar * gnum * sig);
            A0 = rd * (0.5 * az * gden * (Q[c] - qcez) - ar * gnum * Q[c]);
            Z.k1[c][f] += A1;
            Z.k2[c][f] = -A1;
          } else {
            A0 = 0.5 * Raez[c][f] * (Q[c] - qcez) / sig;
          }
          Z.srcS[c] += A0;
          addto<MC>(Z.srcS, cez, -A0);
        }
      }
  // the corner solves in nextC order: everything but the fluxes themselves
#pragma unroll
  for (int i = 0; i < MC; i++) {
    Z.inv[i] = 1.0; Z.rz[i][0] = 0.0; Z.rz[i][1] = 0.0; Z.dz[i][0] = 0; Z.dz[i][1] = 0;
    if (i < nCorner) {
      const int c = Z.ci[i];
      Z.inv[i] = 1.0 / (pick<MC>(sumArea, c) + pick<MC>(volSig, c));
#pragma unroll
      for (int f = 0; f < 2; f++) {
        double r = Raez[0][f];
        int d = Z.cez[0][f];
#pragma unroll
        for (int k = 1; k < MC; k++) { r = c == k ? Raez[k][f] : r; d = c == k ? Z.cez[k][f] : d; }
        Z.rz[i][f] = r;   // 0 unless the EZ face is outgoing
        Z.dz[i][f] = d;
      }
    }
  }
}

// The half of the solve on the dependency chain: upstream fluxes and PsiM in, corner fluxes, PsiM, exiting fluxes out.
template <int MC, bool FLOW = false>
__device__ __forceinline__ void zone_solve_rz(const SweepRZParams &P, int a, int g, const ZoneStatic<MC> &Z) {
  const int G = P.G, nc = P.nc;
  const size_t slab = (size_t)(nc + P.nb) * G;
  double *psi1A = P.psi1 + (size_t)a * slab;
  double *psimL = FLOW ? P.psimA + (size_t)a * nc * G : P.psim + (size_t)P.level[a] * nc * G;
  const int nCorner = Z.nCorner, c0 = Z.c0;
  double src[MC], u[MC][2], pm[MC];
  if (FLOW) {
    // poll the upstream fluxes (rows >= nc are boundary values: inputs, never marked) and the previous angle's PsiM
    const int pa = P.prevAngle[a];
    const double *psimP = P.psimA + (size_t)(pa < 0 ? 0 : pa) * nc * G;
    bool ok;
    unsigned polls = 0;
    do {
      ok = true;
#pragma unroll
      for (int c = 0; c < MC; c++) {
        pm[c] = 0.0;
        u[c][0] = 0.0; u[c][1] = 0.0;
        if (c < nCorner) {
          if (pa >= 0) {
            const unsigned long long v = umt_ld_relaxed_u64(&psimP[(size_t)(c0 + c) * G + g]);
            ok = ok && v != RZ_SENTINEL;
            pm[c] = __longlong_as_double((long long)v);
          }
#pragma unroll
          for (int f = 0; f < 2; f++)
            if (Z.inMask & (1u << (2 * c + f))) {
              const unsigned long long v = umt_ld_relaxed_u64(&psi1A[(size_t)Z.row[c][f] * G + g]);
              ok = ok && v != RZ_SENTINEL;
              u[c][f] = __longlong_as_double((long long)v);
            }
        }
      }
      if (!ok) { __nanosleep(40); if (umt_spin_expired(polls, P.abortFlag, P.spinLimit)) break; }
    } while (!ok);
  }
#pragma unroll
  for (int c = 0; c < MC; c++) {
    src[c] = Z.srcS[c];
    if (!FLOW) {
      pm[c] = 0.0;
      u[c][0] = 0.0; u[c][1] = 0.0;
      if (c < nCorner) {
        pm[c] = psimL[(size_t)(c0 + c) * G + g];
#pragma unroll
        for (int f = 0; f < 2; f++)
          if (Z.inMask & (1u << (2 * c + f))) u[c][f] = __ldcg(&psi1A[(size_t)Z.row[c][f] * G + g]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MC; c++)
#pragma unroll
    for (int f = 0; f < 2; f++) {
      src[c] = fma(Z.k1[c][f], u[c][f], src[c]);
      addto<MC>(src, Z.cez[c][f], Z.k2[c][f] * u[c][f]);
    }
#pragma unroll
  for (int i = 0; i < MC; i++) {
    if (i < nCorner) {
      const int c = Z.ci[i];
      const double p = (pick<MC>(src, c) + pick<MC>(Z.areaFac, c) * pick<MC>(pm, c)) * Z.inv[i];
      put<MC>(src, c, p);   // src now holds the corner flux
      addto<MC>(src, Z.dz[i][0], Z.rz[i][0] * p);
      addto<MC>(src, Z.dz[i][1], Z.rz[i][1] * p);
    }
  }
  // half-angle intensity for the next angle of the level; exiting boundary fluxes; finishing direction
  const bool starting = P.start[a] != 0, fin = P.finishNext[a] != 0;
  double *psi1N = psi1A + slab;   // slab of angle a+1 (only touched when it is a finishing direction)
  const double w1 = P.tauW1[a], w2 = P.tauW2[a];
#pragma unroll
  for (int c = 0; c < MC; c++) {
    if (c < nCorner) {
      const int cc = c0 + c;
      const size_t r = (size_t)cc * G + g;
      const double p = src[c];
      const double pmn = starting ? p : w1 * p - w2 * pm[c];
      if (FLOW) { umt_st_relaxed_f64(&psimL[r], pmn); umt_st_relaxed_f64(&psi1A[r], p); }
      else { psimL[r] = pmn; psi1A[r] = p; }
      if (fin) psi1N[r] = pmn;   // Psi(:,c,Angle+1) <- PsiM (becomes Psi when the buffers trade roles on savePsi)
#pragma unroll
      for (int f = 0; f < 2; f++) {
        if (Z.exitMask & (1u << (2 * c + f))) {
          const int row = Z.row[c][f];
          psi1A[(size_t)row * G + g] = p;             // PsiB(:,b,Angle)   <- Psi1(:,c)
          if (fin) psi1N[(size_t)row * G + g] = pmn;  // PsiB(:,b,Angle+1) <- PsiM(:,c)
        }
      }
    }
  }
}

// One (zone, group) pair per thread (items hold at most blockDim pairs): the static half runs before the dependency wait.
template <int MC>
__global__ void __launch_bounds__(RZ_BLOCK, RZ_MIN_CTAS) sweeprz_kernel(SweepRZParams P) {
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

// ---------------------------------------------------------------------------------------------------------------------------
// Record kernel (quads, G <= 64).  Everything of a zone solve that does not depend on the group -- omega.A products, radii, the
// group-independent part of sumArea, connectivity, the solve order -- is the same for the G threads of a zone and for every
// sweep until the mesh or the schedule changes.  rz_rec_build_kernel computes it once per (angle, zone) into a 384-byte record
// in sweep order; the sweep's CTA copies the records of its item into shared memory and its threads keep only the
// group-dependent values in registers (Q, the closure coefficients A1, the reciprocals), indexing the rest in shared memory
// with the dynamic corner numbers directly.  Against sweeprz_kernel: a third of the instructions, two dependent levels of global
// loads fewer ahead of the solve, and few enough registers for twice the resident warps (the sweep is bound by the latency of
// the pair solves times the resident warps, not by the plane-to-plane chain: 4x the zones take 3.9x the time).
struct alignas(16) RZRec {
  double vol[4], area[4], areaFac[4], sumArea[4];   // by corner; sumArea: angDerivFac*Area - sum R.afp (incident) + sum R.aez (incoming EZ)
  double k1b[4][2], az[4][2], rez[4][2];            // -R_fp afp of incident faces (else 0); omega.A_ez where > 0 (else 0); RadiusEZ
  int row[4][2];                                    // Psi1 row behind the FP face (>= nc: boundary element)
  int c0, zone, nCorner;
  unsigned inMask, exitMask;
  unsigned char cez[4][2], ci[4];                   // corner across the EZ face; corners in solve order (nextC)
};
static_assert(sizeof(RZRec) == 384, "RZRec layout");
#ifndef RZ_REC_MIN_CTAS
#define RZ_REC_MIN_CTAS 10
#endif

// Canonical quad labelling.  The corners of a quad swept in a generic direction are solved source -> its two neighbours ->
// sink, and the two neighbours never exchange flux (they are opposite corners).  Relabelling the corners by that position
// (p0 = first corner of nextC, p1 = p0+1, p2 = p0+3, p3 = p0+2 in the zone's cyclic numbering) and ordering the two face
// slots of a corner as (toward local+1, toward local+3) makes the corner across every EZ face and the solve order compile-time
// constants (RZ_NB below), so the register arrays of the solve need no select chains at all (a third of the instructions of
// the labelled-by-corner version).  Any topological order gives the reference's corner fluxes; only the order in which the two
// pushes into the sink are added differs from nextC's.  Zones that do not fit (triangles, flows like 0>1>3>2) are counted and
// the whole mesh then takes the by-corner records.
__device__ __host__ constexpr int rz_nb(int p, int f) { return f == 0 ? (p == 0 ? 1 : p == 1 ? 3 : p == 2 ? 0 : 2) : (p == 0 ? 2 : p == 1 ? 0 : p == 2 ? 3 : 1); }

__global__ void rz_rec_build_kernel(SweepRZParams P, RZRec *recs, int canonMode, int *nonCanon) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)P.NA * P.nz) return;
  const int a = (int)(idx / P.nz);
  if (P.nHyp[a] == 0) return;
  const int zone0 = P.nextZ[idx], zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone], nc = P.nc;
  const double om0 = P.omega[2 * a], om1 = P.omega[2 * a + 1], fac = P.angDerivFac[a];
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  RZRec R;
  R.c0 = c0; R.zone = zone; R.nCorner = nCorner; R.inMask = 0u; R.exitMask = 0u;
  for (int c = 0; c < 4; c++) {
    R.vol[c] = 0.0; R.area[c] = 0.0; R.areaFac[c] = 0.0; R.sumArea[c] = 1.0; R.ci[c] = (unsigned char)c;
    for (int f = 0; f < 2; f++) { R.k1b[c][f] = 0.0; R.az[c][f] = 0.0; R.rez[c][f] = 0.0; R.row[c][f] = 0; R.cez[c][f] = 0; }
    if (c < nCorner) {
      const double ar = P.Area[c0 + c];
      R.vol[c] = P.Volume[c0 + c]; R.area[c] = ar; R.areaFac[c] = ar * fac; R.sumArea[c] = fac * ar; R.ci[c] = nextC[c0 + c];
    }
  }
  for (int c = 0; c < nCorner; c++)
    for (int f = 0; f < 2; f++) {
      const int cc = c0 + c;
      const double *Af = P.Afp + ((size_t)cc * 2 + f) * 2, *Ae = P.Aez + ((size_t)cc * 2 + f) * 2;
      const double afp = __dadd_rn(__dmul_rn(om0, Af[0]), __dmul_rn(om1, Af[1]));
      const double az = __dadd_rn(__dmul_rn(om0, Ae[0]), __dmul_rn(om1, Ae[1]));
      const int row = P.cFP[cc * 2 + f], cez = P.cEZ[cc * 2 + f];
      R.row[c][f] = row; R.cez[c][f] = (unsigned char)cez; R.rez[c][f] = P.RadiusEZ[cc * 2 + f];
      if (afp < 0.0) {
        const double Rafp = __dmul_rn(P.RadiusFP[cc * 2 + f], afp);
        R.inMask |= 1u << (2 * c + f);
        R.k1b[c][f] = -Rafp;
        R.sumArea[c] = __dadd_rn(R.sumArea[c], -Rafp);
      } else if (afp > 0.0 && row >= nc) R.exitMask |= 1u << (2 * c + f);
      if (az > 0.0) {
        R.az[c][f] = az;
        R.sumArea[cez] = __dadd_rn(R.sumArea[cez], __dmul_rn(R.rez[c][f], az));
      }
    }
  if (canonMode) {
    const int L0 = R.ci[0];
    const int L[4] = {L0, (L0 + 1) & 3, (L0 + 3) & 3, (L0 + 2) & 3};
    bool ok = nCorner == 4 && R.ci[3] == L[3] && ((R.ci[1] == L[1] && R.ci[2] == L[2]) || (R.ci[1] == L[2] && R.ci[2] == L[1]));
    RZRec C;
    C.c0 = c0; C.zone = zone; C.nCorner = nCorner; C.inMask = 0u; C.exitMask = 0u;
    for (int p = 0; p < 4; p++) {
      const int c = L[p];
      C.vol[p] = R.vol[c]; C.area[p] = R.area[c]; C.areaFac[p] = R.areaFac[c]; C.sumArea[p] = R.sumArea[c]; C.ci[p] = (unsigned char)c;
      for (int f = 0; f < 2; f++) {
        const int target = L[rz_nb(p, f)];
        const int fl = R.cez[c][0] == target ? 0 : (R.cez[c][1] == target ? 1 : -1);
        if (fl < 0 || R.cez[c][0] == R.cez[c][1]) { ok = false; C.k1b[p][f] = 0.0; C.az[p][f] = 0.0; C.rez[p][f] = 0.0; C.row[p][f] = 0; C.cez[p][f] = 0; continue; }
        C.k1b[p][f] = R.k1b[c][fl]; C.az[p][f] = R.az[c][fl]; C.rez[p][f] = R.rez[c][fl]; C.row[p][f] = R.row[c][fl];
        C.cez[p][f] = (unsigned char)rz_nb(p, f);
        if (R.inMask & (1u << (2 * c + fl))) C.inMask |= 1u << (2 * p + f);
        if (R.exitMask & (1u << (2 * c + fl))) C.exitMask |= 1u << (2 * p + f);
      }
    }
    if (!ok) atomicAdd(nonCanon, 1);
    recs[idx] = C;
    return;
  }
  recs[idx] = R;
}

// FLOW: the dataflow scheme of sweeprz_flow_kernel (below) on top of the records.  CANON: records labelled by solve position
// (see rz_rec_build_kernel): array index = position, R.ci[p] = local corner of position p (for the row addresses only).
template <bool FLOW, bool CANON>
__global__ void __launch_bounds__(RZ_BLOCK, RZ_REC_MIN_CTAS) sweeprz_rec_kernel(SweepRZParams P, const RZRec *recs) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  RZRec *s_rec = reinterpret_cast<RZRec *>(s_raw);
  __shared__ int s_item;
  const int G = P.G, nc = P.nc;
  const size_t slab = (size_t)(nc + P.nb) * G;
#ifdef RZ_TICKET_AHEAD   // A/B: the next ticket is taken while the current item is worked on (hides the atomic's round trip)
  __shared__ int s_next[2];
  if (threadIdx.x == 0) s_next[0] = atomicAdd(&P.counters[0], 1);
  __syncthreads();
  int it = s_next[0], buf = 0;
  for (;;) {
    if (it >= P.nItems) break;
    int nxt = 0;
    if (threadIdx.x == 0) nxt = atomicAdd(&P.counters[0], 1);
#else
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&P.counters[0], 1);
    __syncthreads();
    const int it = s_item;
    if (it >= P.nItems) break;
#endif
    const WorkItem w = P.items[it];
    const int a = w.angle, nrec = w.zend - w.zbeg, npairs = nrec * G;
    {   // the item's records are contiguous: 24 16-byte words each
      const uint4 *src4 = reinterpret_cast<const uint4 *>(recs + (size_t)a * P.nz + w.zbeg);
      uint4 *dst4 = reinterpret_cast<uint4 *>(s_rec);
      for (int k = threadIdx.x; k < nrec * 24; k += blockDim.x) dst4[k] = __ldg(src4 + k);
    }
    __syncthreads();
    const int zi = threadIdx.x / G, g = threadIdx.x - zi * G;
    const bool active = (int)threadIdx.x < npairs;
    const RZRec &R = s_rec[active ? zi : 0];
#define LC(c) (CANON ? (int)R.ci[c] : (c))
    const int nCorner = R.nCorner, c0 = R.c0;
    const double *psiA = P.psi + (size_t)a * slab;
    double *psi1A = P.psi1 + (size_t)a * slab;
    double *psimL = FLOW ? P.psimA + (size_t)a * nc * G : P.psim + (size_t)P.level[a] * nc * G;
    double srcS[4], A1[4][2], inv[4];
    if (active) {   // group-dependent static half
      const double sig = P.sigt[(size_t)R.zone * G + g], sigInv = 1.0 / sig;
      double Q[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        Q[c] = 0.0;
        if (c < nCorner) { const size_t r = (size_t)(c0 + LC(c)) * G + g; Q[c] = P.stotal[r] + P.tau * psiA[r]; }
        srcS[c] = R.vol[c] * Q[c];
      }
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < 2; f++) {
          A1[c][f] = 0.0;
          const double az = R.az[c][f];
          if (az > 0.0) {
            const int cez = CANON ? rz_nb(c, f) : (int)R.cez[c][f];
            const double Rr = R.rez[c][f], dq = Q[c] - pick<4>(Q, cez);
            double A0;
            if (R.inMask & (1u << (2 * c + f))) {
              const double ar = R.area[c];
              const double sigA = sig * ar, sigA2 = sigA * sigA;
              const double gnum = az * az * (FOURALPHA * sigA2 + az * (4.0 * sigA + 3.0 * az));
              const double gden = ar * (4.0 * sigA * sigA2 + az * (6.0 * sigA2 + 2.0 * az * (2.0 * sigA + az)));
              const double rd = Rr / (gnum + gden * sig);
              A1[c][f] = rd * (ar * gnum * sig);
              A0 = rd * (0.5 * az * gden * dq - ar * gnum * Q[c]);
            } else {
              A0 = 0.5 * (Rr * az) * dq * sigInv;
            }
            srcS[c] += A0;
            addto<4>(srcS, cez, -A0);
          }
        }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int c = CANON ? i : (int)R.ci[i];
        inv[i] = 1.0 / (R.sumArea[c] + sig * R.vol[c]);
      }
    }
    if (!FLOW) {
      if (threadIdx.x == blockDim.x - 1) {
        if (w.wait_idx >= 0)
          while (ld_acquire(&P.counters[1 + w.wait_idx]) < w.wait_count) __nanosleep(20);
        if (w.pad0 >= 0)
          while (ld_acquire(&P.counters[1 + w.pad0]) < w.pad1) __nanosleep(20);
      }
      __syncthreads();
    }
    if (active) {   // the half on the dependency chain
      double src[4], pm[4], u[4][2];
      if (FLOW) {   // every lane polls the values its pair needs until they are no longer marked
        const int pa = P.prevAngle[a];
        const double *psimP = P.psimA + (size_t)(pa < 0 ? 0 : pa) * nc * G;
        bool ok;
        unsigned polls = 0;
        do {
          ok = true;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            pm[c] = 0.0; u[c][0] = 0.0; u[c][1] = 0.0;
            if (c < nCorner) {
              if (pa >= 0) {
                const unsigned long long v = umt_ld_relaxed_u64(&psimP[(size_t)(c0 + LC(c)) * G + g]);
                ok = ok && v != RZ_SENTINEL;
                pm[c] = __longlong_as_double((long long)v);
              }
#pragma unroll
              for (int f = 0; f < 2; f++)
                if (R.inMask & (1u << (2 * c + f))) {
                  const unsigned long long v = umt_ld_relaxed_u64(&psi1A[(size_t)R.row[c][f] * G + g]);
                  ok = ok && v != RZ_SENTINEL;
                  u[c][f] = __longlong_as_double((long long)v);
                }
            }
          }
          if (!ok) { __nanosleep(40); if (umt_spin_expired(polls, P.abortFlag, P.spinLimit)) break; }
        } while (!ok);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        src[c] = srcS[c];
        if (!FLOW) {
          pm[c] = 0.0; u[c][0] = 0.0; u[c][1] = 0.0;
          if (c < nCorner) {
            pm[c] = psimL[(size_t)(c0 + LC(c)) * G + g];
#pragma unroll
            for (int f = 0; f < 2; f++)
              if (R.inMask & (1u << (2 * c + f))) u[c][f] = __ldcg(&psi1A[(size_t)R.row[c][f] * G + g]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < 2; f++) {
          src[c] = fma(R.k1b[c][f] + A1[c][f], u[c][f], src[c]);
          addto<4>(src, CANON ? rz_nb(c, f) : (int)R.cez[c][f], -A1[c][f] * u[c][f]);
        }
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (i < nCorner) {
          const int c = CANON ? i : (int)R.ci[i];
          const double p = (pick<4>(src, c) + R.areaFac[c] * pick<4>(pm, c)) * inv[i];
          put<4>(src, c, p);
          addto<4>(src, CANON ? rz_nb(c, 0) : (int)R.cez[c][0], (R.rez[c][0] * R.az[c][0]) * p);
          addto<4>(src, CANON ? rz_nb(c, 1) : (int)R.cez[c][1], (R.rez[c][1] * R.az[c][1]) * p);
        }
      }
      const bool starting = P.start[a] != 0, fin = P.finishNext[a] != 0;
      double *psi1N = psi1A + slab;
      const double w1 = P.tauW1[a], w2 = P.tauW2[a];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (c < nCorner) {
          const size_t r = (size_t)(c0 + LC(c)) * G + g;
          const double p = src[c];
          const double pmn = starting ? p : w1 * p - w2 * pm[c];
          if (FLOW) { umt_st_relaxed_f64(&psimL[r], pmn); umt_st_relaxed_f64(&psi1A[r], p); }
          else { psimL[r] = pmn; psi1A[r] = p; }
          if (fin) psi1N[r] = pmn;
#pragma unroll
          for (int f = 0; f < 2; f++) {
            if (R.exitMask & (1u << (2 * c + f))) {
              const int row = R.row[c][f];
              psi1A[(size_t)row * G + g] = p;
              if (fin) psi1N[(size_t)row * G + g] = pmn;
            }
          }
        }
      }
    }
#ifdef RZ_TICKET_AHEAD
    if (threadIdx.x == 0) s_next[buf ^ 1] = nxt;
#endif
    if (!FLOW) {
      __syncthreads();
      if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + w.signal_idx]) : "memory");
      }
    }
#ifdef RZ_TICKET_AHEAD
    if (FLOW) __syncthreads();
    it = s_next[buf ^ 1]; buf ^= 1;
#endif
  }
#undef LC
}

// ---------------------------------------------------------------------------------------------------------------------------
// Pipelined record kernel (UMT_RZ_KERNEL=pipe; quads with canonical records, G <= 64, G even; NOT the default, see the launcher): the serial per-item chain of
// sweeprz_rec_kernel (ticket -> item -> records -> inputs from DRAM -> divisions -> poll -> barrier -> upstream loads -> solve ->
// stores -> fence -> signal, one CTA doing all of it for one item at a time) split over three roles, as in the 3-D plan kernel:
//   * a loader warp takes tickets ahead (QB at a time, descriptors and zone info fetched by its lanes in parallel), lands every
//     item's records and its Psi^n / STotal / Sigt rows in one of the CTA's shared-memory stages with cp.async.bulk (1-D TMA,
//     mbarrier complete_tx) and resolves BOTH dependencies of the item -- the previous hyperplane of its angle and the plane of
//     the previous angle of its xi-level that covers its zones (PsiM chain, SweepUCBrz.F90:212-240) -- with ld.acquire, then
//     makes the second arrival on the stage's `full` barrier;
//   * engines of two warps (one thread per (zone, group) pair of an item) wait on `full`, compute the group-dependent static
//     half from shared memory, read PsiM and the upstream Psi1 rows through L2, solve, store, and arrive on `empty`: they never
//     poll global memory and there is no __syncthreads() per item;
//   * a signaller warp turns `empty` arrivals into fence.acq_rel.gpu + red on the plane counters.
#ifndef RZP_NE
#define RZP_NE 2                 // engines per CTA
#endif
#ifndef RZP_MIN_CTAS
#define RZP_MIN_CTAS 3
#endif
constexpr int RZP_WPE = 2;                         // warps per engine: 64 (zone, group) pairs per item at most
constexpr int RZP_THREADS = RZP_NE * RZP_WPE * 32 + 64;
constexpr int RZP_MAX_STAGES = 12, RZP_RING = 32, RZP_CTL_BYTES = 768, RZP_ZMAX = 8;
struct RZPMeta { int angle, n, signal_idx, wait_idx, wait_count, wait2_idx, wait2_count, pad; };
struct RZPCtl {
  unsigned long long full[RZP_MAX_STAGES], empty[RZP_MAX_STAGES];
  RZPMeta meta[RZP_MAX_STAGES];
  int sigRing[RZP_RING];
  volatile int issuedCount, doneFlag, nSignaled;
  int pad;
};
static_assert(sizeof(RZPCtl) <= RZP_CTL_BYTES, "RZPCtl must fit the control block");
struct RZPGeom { int nStages, stageBytes, offSt, offSigt, offRecs, zpi; };

__global__ void __launch_bounds__(RZP_THREADS, RZP_MIN_CTAS) sweeprz_pipe_kernel(SweepRZParams P, const RZRec *__restrict__ recs,
                                                                                 const int2 *__restrict__ zinfo, RZPGeom Gm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  RZPCtl &S = *reinterpret_cast<RZPCtl *>(smem_raw);
  unsigned char *stages = smem_raw + RZP_CTL_BYTES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = P.G, nc = P.nc, NS = Gm.nStages;
  const size_t slab = (size_t)(nc + P.nb) * G;
  if (tid == 0) {
    for (int s = 0; s < NS; s++) { umt_mbar_init(&S.full[s], 2); umt_mbar_init(&S.empty[s], RZP_WPE); }
    S.issuedCount = 0; S.doneFlag = 0; S.nSignaled = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == RZP_NE * RZP_WPE) {
    // ---------------- loader warp (same scheme as sweep3d_plan_kernel's) ----------------
    int nIssued = 0, nReleased = 0, sentinels = -1;
    int kFill = 0, sFill = 0, sRel = 0, lastOk = -1, lastOk2 = -1;
    unsigned parPrev = 1;
    const unsigned rowBytes = (unsigned)G * 8u;
    const int zpi = Gm.zpi, QB = min(8, 32 / zpi);
    const int myItem = lane / zpi, myZone = lane - myItem * zpi;
    WorkItem qW, nW;
    int2 qZ = make_int2(0, 0), nZ = make_int2(0, 0);
    int qCount = 0, qPos = 0, nCount = 0, nPhase = 0, nT = 0;
    bool exhausted = false;
    qW.angle = qW.zbeg = qW.zend = qW.wait_idx = qW.wait_count = qW.signal_idx = qW.pad0 = qW.pad1 = 0; nW = qW;
    auto fetch_step = [&]() {
      if (nPhase == 0) {
        if (lane == 0) nT = atomicAdd(&P.counters[0], QB);
        nPhase = 1;
      } else if (nPhase == 1) {
        nT = __shfl_sync(0xffffffffu, nT, 0);
        nCount = max(0, min(QB, P.nItems - nT));
        if (myItem < nCount) nW = P.items[nT + myItem];
        nPhase = 2;
      } else if (nPhase == 2) {
        if (myItem < nCount && myZone < nW.zend - nW.zbeg) nZ = zinfo[(size_t)nW.angle * P.nz + nW.zbeg + myZone];
        nPhase = 3;
      }
    };
    for (;;) {
      bool progressed = false;
      if (qPos == qCount && !exhausted) {
        while (nPhase < 3) fetch_step();
        qW = nW; qZ = nZ; qCount = nCount; qPos = 0; nPhase = 0;
        if (qCount < QB) exhausted = true;
      }
      if (!exhausted && nPhase < 3) fetch_step();
      if (nReleased < nIssued) {
        int ok = 1;
        if (lane == 0) {
          const RZPMeta &m = S.meta[sRel];
          ok = m.wait_idx < 0 || m.wait_idx == lastOk || umt_ld_acquire(&P.counters[1 + m.wait_idx]) >= m.wait_count;
          if (ok) {
            lastOk = m.wait_idx;
            ok = m.wait2_idx < 0 || m.wait2_idx == lastOk2 || umt_ld_acquire(&P.counters[1 + m.wait2_idx]) >= m.wait2_count;
            if (ok) { lastOk2 = m.wait2_idx; umt_mbar_arrive(&S.full[sRel]); }
          }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          nReleased++; progressed = true;
          if (++sRel == NS) sRel = 0;
        }
      }
      if (sentinels < 0 && exhausted && qPos == qCount) sentinels = RZP_NE;
      if (sentinels != 0) {
        int free_ = 1;
        if (lane == 0) free_ = (kFill < NS || umt_mbar_test(&S.empty[sFill], parPrev)) && (sentinels > 0 || nIssued - S.nSignaled < RZP_RING);
        free_ = __shfl_sync(0xffffffffu, free_, 0);
        if (free_) {
          if (sentinels > 0) {
            if (lane == 0) { S.meta[sFill].n = -1; umt_mbar_arrive(&S.full[sFill]); umt_mbar_arrive(&S.full[sFill]); }
            sentinels--;
          } else {
            const int srcLane = qPos * zpi;
            WorkItem w;
            w.angle = __shfl_sync(0xffffffffu, qW.angle, srcLane); w.zbeg = __shfl_sync(0xffffffffu, qW.zbeg, srcLane);
            w.zend = __shfl_sync(0xffffffffu, qW.zend, srcLane); w.wait_idx = __shfl_sync(0xffffffffu, qW.wait_idx, srcLane);
            w.wait_count = __shfl_sync(0xffffffffu, qW.wait_count, srcLane); w.signal_idx = __shfl_sync(0xffffffffu, qW.signal_idx, srcLane);
            w.pad0 = __shfl_sync(0xffffffffu, qW.pad0, srcLane); w.pad1 = __shfl_sync(0xffffffffu, qW.pad1, srcLane);
            int2 zi;
            zi.x = __shfl_sync(0xffffffffu, qZ.x, (srcLane + lane) & 31); zi.y = __shfl_sync(0xffffffffu, qZ.y, (srcLane + lane) & 31);
            qPos++;
            const int n = w.zend - w.zbeg;
            const size_t first = (size_t)w.angle * P.nz + w.zbeg;
            const unsigned bytes = (unsigned)n * (9u * rowBytes + (unsigned)sizeof(RZRec));   // per zone: 4 Psi rows, 4 STotal rows, 1 Sigt row, 1 record
            unsigned char *st = stages + (size_t)sFill * Gm.stageBytes;
            if (lane == 0) {
              RZPMeta &m = S.meta[sFill];
              m.angle = w.angle; m.n = n; m.wait_idx = w.wait_idx; m.wait_count = w.wait_count; m.wait2_idx = w.pad0; m.wait2_count = w.pad1;
              S.sigRing[nIssued & (RZP_RING - 1)] = w.signal_idx;
              __threadfence_block();
              S.issuedCount = nIssued + 1;
              umt_mbar_arrive_expect_tx(&S.full[sFill], bytes);
              umt_tma_load_1d(st + Gm.offRecs, recs + first, (unsigned)(n * sizeof(RZRec)), &S.full[sFill], UMT_L2_EVICT_FIRST);
            }
            __syncwarp();
            if (lane < n) {   // zi.x = first corner of the zone, zi.y = zone
              umt_tma_load_1d(st + (size_t)lane * 4 * rowBytes, P.psi + (size_t)w.angle * slab + (size_t)zi.x * G, 4u * rowBytes, &S.full[sFill], UMT_L2_EVICT_FIRST);
              umt_tma_load_1d(st + Gm.offSt + (size_t)lane * 4 * rowBytes, P.stotal + (size_t)zi.x * G, 4u * rowBytes, &S.full[sFill], UMT_L2_EVICT_FIRST);
              umt_tma_load_1d(st + Gm.offSigt + (size_t)lane * rowBytes, P.sigt + (size_t)zi.y * G, rowBytes, &S.full[sFill], UMT_L2_EVICT_FIRST);
            }
            nIssued++;
          }
          progressed = true;
          kFill++;
          if (++sFill == NS) { sFill = 0; parPrev ^= 1u; }
        }
      }
      if (sentinels == 0 && nReleased == nIssued) break;
      if (!progressed) __nanosleep(32);
    }
    if (lane == 0) { __threadfence_block(); S.doneFlag = 1; }
    return;
  }

  if (warp == RZP_NE * RZP_WPE + 1) {
    // ---------------- signaller warp ----------------
    if (lane != 0) return;
    int k = 0, sg = 0;
    unsigned par = 0;
    for (;;) {
      const int done = S.doneFlag;
      __threadfence_block();
      const int issued = S.issuedCount;
      if (k >= issued) {
        if (done) break;
        __nanosleep(64);
        continue;
      }
      umt_mbar_wait(&S.empty[sg], par);
      int m = 1, s2 = sg + 1;
      unsigned p2 = par;
      if (s2 == NS) { s2 = 0; p2 ^= 1u; }
      while (k + m < issued && umt_mbar_test(&S.empty[s2], p2)) {
        m++;
        if (++s2 == NS) { s2 = 0; p2 ^= 1u; }
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      for (int j = 0; j < m; j++)
        asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + S.sigRing[(k + j) & (RZP_RING - 1)]]) : "memory");
      k += m; sg = s2; par = p2;
      S.nSignaled = k;
    }
    return;
  }

  // ---------------- engines: two warps, one thread per (zone, group) pair of the item ----------------
  const int eng = warp / RZP_WPE, elane = (warp - eng * RZP_WPE) * 32 + lane;
  const int zi = elane / G, g = elane - zi * G;
  const double tau = P.tau;
  for (int k = eng;; k += RZP_NE) {
    const int s = k % NS;
    umt_mbar_wait(&S.full[s], (k / NS) & 1);
    const RZPMeta m = S.meta[s];
    if (m.n < 0) break;
    if (zi < m.n) {
      const unsigned char *st = stages + (size_t)s * Gm.stageBytes;
      const RZRec &R = reinterpret_cast<const RZRec *>(st + Gm.offRecs)[zi];
      const double *sPsi = reinterpret_cast<const double *>(st) + (size_t)zi * 4 * G + g;
      const double *sSt = reinterpret_cast<const double *>(st + Gm.offSt) + (size_t)zi * 4 * G + g;
      const int a = m.angle, c0 = R.c0;
      double *psi1A = P.psi1 + (size_t)a * slab;
      double *psimL = P.psim + (size_t)P.level[a] * nc * G;
      // upstream fluxes and the previous angle's PsiM go in flight first (the item's dependencies are complete)
      double pm[4], u[4][2];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        pm[c] = __ldcg(&psimL[(size_t)(c0 + (int)R.ci[c]) * G + g]);
#pragma unroll
        for (int f = 0; f < 2; f++) {
          u[c][f] = 0.0;
          if (R.inMask & (1u << (2 * c + f))) u[c][f] = __ldcg(&psi1A[(size_t)R.row[c][f] * G + g]);
        }
      }
      // group-dependent static half from the landed rows (corners by solve position: R.ci[p] = local corner of position p)
      const double sig = reinterpret_cast<const double *>(st + Gm.offSigt)[zi * G + g], sigInv = 1.0 / sig;
      double Q[4], src[4], A1[4][2], inv[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int lc = (int)R.ci[c] * G;
        Q[c] = sSt[lc] + tau * sPsi[lc];
        src[c] = R.vol[c] * Q[c];
      }
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < 2; f++) {
          A1[c][f] = 0.0;
          const double az = R.az[c][f];
          if (az > 0.0) {
            const int cez = rz_nb(c, f);
            const double Rr = R.rez[c][f], dq = Q[c] - Q[cez];
            double A0;
            if (R.inMask & (1u << (2 * c + f))) {
              const double ar = R.area[c];
              const double sigA = sig * ar, sigA2 = sigA * sigA;
              const double gnum = az * az * (FOURALPHA * sigA2 + az * (4.0 * sigA + 3.0 * az));
              const double gden = ar * (4.0 * sigA * sigA2 + az * (6.0 * sigA2 + 2.0 * az * (2.0 * sigA + az)));
              const double rd = Rr / (gnum + gden * sig);
              A1[c][f] = rd * (ar * gnum * sig);
              A0 = rd * (0.5 * az * gden * dq - ar * gnum * Q[c]);
            } else {
              A0 = 0.5 * (Rr * az) * dq * sigInv;
            }
            src[c] += A0;
            src[cez] -= A0;
          }
        }
#pragma unroll
      for (int i = 0; i < 4; i++) inv[i] = 1.0 / (R.sumArea[i] + sig * R.vol[i]);
      // the half on the dependency chain
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < 2; f++) {
          src[c] = fma(R.k1b[c][f] + A1[c][f], u[c][f], src[c]);
          src[rz_nb(c, f)] += -A1[c][f] * u[c][f];
        }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double p = (src[c] + R.areaFac[c] * pm[c]) * inv[c];
        src[c] = p;
        src[rz_nb(c, 0)] += (R.rez[c][0] * R.az[c][0]) * p;
        src[rz_nb(c, 1)] += (R.rez[c][1] * R.az[c][1]) * p;
      }
      const bool starting = P.start[a] != 0, fin = P.finishNext[a] != 0;
      double *psi1N = psi1A + slab;
      const double w1 = P.tauW1[a], w2 = P.tauW2[a];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const size_t r = (size_t)(c0 + (int)R.ci[c]) * G + g;
        const double p = src[c];
        const double pmn = starting ? p : w1 * p - w2 * pm[c];
        psimL[r] = pmn; psi1A[r] = p;
        if (fin) psi1N[r] = pmn;
#pragma unroll
        for (int f = 0; f < 2; f++) {
          if (R.exitMask & (1u << (2 * c + f))) {
            const int row = R.row[c][f];
            psi1A[(size_t)row * G + g] = p;
            if (fin) psi1N[(size_t)row * G + g] = pmn;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) umt_mbar_arrive(&S.empty[s]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Level-chain kernel (quads with canonical records, one stage).  The r-z sweep is bound by its dependency chain, not by throughput:
// 4P xi-levels whose (A + 1) swept angles are chained through PsiM, each angle a chain of hyperplanes, and every hop of the item
// kernels costs store -> fence -> counter -> poll -> load.  Here ONE CTA owns one (xi-level, block of gb groups) -- the groups are
// independent -- and walks that chain by itself, one step (a hyperplane, or a chunk of a large one) per __syncthreads(): no
// tickets, counters, fences or polls anywhere.  The CTA is split into two groups of threads, one thread per (zone, group) pair
// of a step in each:
//   * the CHAIN group does only what depends on the previous steps: PsiM and upstream Psi1 rows through L2, a handful of FMAs per
//     face (the closure in its linear form), the corner fluxes, the stores;
//   * the STATIC group runs one step ahead: it copies the step's records into shared memory (two steps ahead), turns the landed
//     Psi^n / STotal / Sigt values (loaded one step ahead into registers) into the group-dependent coefficients (every division of
//     the zone solve) and leaves them in shared memory for the chain group.
// The hop is then the dependent half alone.  Rows written by a step are read by later steps of the same CTA through L2 (ld.cg);
// __syncthreads() orders them.
struct RZRecS {   // RZRec as it sits in shared memory at the padded stride: same members, 8-byte alignment only
  double vol[4], area[4], areaFac[4], sumArea[4];
  double k1b[4][2], az[4][2], rez[4][2];
  int row[4][2];
  int c0, zone, nCorner;
  unsigned inMask, exitMask;
  unsigned char cez[4][2], ci[4];
};
static_assert(sizeof(RZRecS) == sizeof(RZRec), "RZRecS mirrors RZRec");
struct RZStep { int angle, zbeg, n, pad; };          // zones [zbeg, zbeg + n) of nextZ(:, angle): one plane or a chunk of it
constexpr int RZL_REC_STRIDE = 392;                  // 384-byte record + 8: consecutive zones fall into different shared-memory banks
constexpr int RZL_NSTAT = 16;                        // per pair: src[4], A1[4][2], inv[4]
struct RZLParams {
  const RZRec *recs;
  const int2 *zinfo;                                 // (NA, nz) in sweep order: first corner row, zone
  const RZStep *steps;                               // (nLevels, maxSteps)
  const int *nSteps;                                 // (nLevels)
  int maxSteps, gb, nGB, ZCH, CT;                    // groups per CTA, group blocks, zones per step at most, threads per group
};

constexpr int RZL_MAX_CT = 288;                      // threads per group at most (576 per CTA: 113 registers each)
__global__ void __launch_bounds__(2 * RZL_MAX_CT, 1) sweeprz_lc_kernel(SweepRZParams P, RZLParams L) {
  extern __shared__ __align__(16) unsigned char lsm[];
  const int G = P.G, nc = P.nc, CT = L.CT, PCH = L.ZCH * L.gb;
  const size_t slab = (size_t)(nc + P.nb) * G;
  const int lev = blockIdx.x / L.nGB, g0 = (blockIdx.x - lev * L.nGB) * L.gb, gbA = min(L.gb, G - g0);
  const int tid = threadIdx.x;
  const bool chain = tid < CT;
  const int t = chain ? tid : tid - CT;
  unsigned char *recBuf = lsm;                                                   // 3 x ZCH x RZL_REC_STRIDE
  double *statBuf = reinterpret_cast<double *>(lsm + (size_t)3 * L.ZCH * RZL_REC_STRIDE);   // 2 x RZL_NSTAT x PCH
  const RZStep *steps = L.steps + (size_t)lev * L.maxSteps;
  const int nS = L.nSteps[lev];
  const double tau = P.tau;
  const int zi = t / gbA, g = g0 + (t - zi * gbA);     // my pair of every step (if the step has that many)

  auto copy_records = [&](int s) {                     // static group: records of step s -> recBuf[s % 3], padded stride
    if (s >= nS) return;
    const RZStep st = steps[s];
    const unsigned char *src = reinterpret_cast<const unsigned char *>(L.recs + (size_t)st.angle * P.nz + st.zbeg);
    unsigned char *dst = recBuf + (size_t)(s % 3) * L.ZCH * RZL_REC_STRIDE;
    for (int k = t; k < st.n * 48; k += CT) {   // 8-byte asynchronous copies (the padded stride is 8 mod 16): nobody waits for them here
      const int z = k / 48, w = k - z * 48;
      umt_cp_async8(dst + (size_t)z * RZL_REC_STRIDE + w * 8, src + (size_t)k * 8);
    }
    umt_cp_async_commit();
  };
  // inputs of my pair of a step, loaded one step ahead (static group): Psi^n and STotal of the 4 corners (local order), Sigt
  // (two register sets: the loads for step s + 2 are issued before the static half of step s + 1 consumes the other set, so they have
  // a whole iteration to land)
  double aPsi[4], aSt[4], aSig = 1.0, bPsi[4], bSt[4], bSig = 1.0;
  int2 zNext = make_int2(0, 0);                        // zone info of my pair two steps ahead
  auto load_zinfo = [&](int s) {
    if (s < nS) { const RZStep st = steps[s]; if (zi < st.n) zNext = L.zinfo[(size_t)st.angle * P.nz + st.zbeg + zi]; }
  };
  auto load_inputs = [&](int s, double (&dPsi)[4], double (&dSt)[4], double &dSig) {   // uses zNext (= zone info of step s)
    if (s >= nS) return;
    const RZStep st = steps[s];
    if (zi >= st.n) return;
    const double *psiA = P.psi + (size_t)st.angle * slab + (size_t)zNext.x * G + g;
    const double *stA = P.stotal + (size_t)zNext.x * G + g;
#pragma unroll
    for (int c = 0; c < 4; c++) { dPsi[c] = __ldcs(psiA + (size_t)c * G); dSt[c] = __ldcs(stA + (size_t)c * G); }
    dSig = __ldg(P.sigt + (size_t)zNext.y * G + g);
  };
  auto static_half = [&](int s, const double (&inPsi)[4], const double (&inSt)[4], const double inSig) {   // coefficients of my pair of step s -> statBuf[s % 2]
    const RZStep st = steps[s];
    if (zi >= st.n) return;
    const RZRecS &R = *reinterpret_cast<const RZRecS *>(recBuf + ((size_t)(s % 3) * L.ZCH + zi) * RZL_REC_STRIDE);
    double *out = statBuf + (size_t)(s % 2) * RZL_NSTAT * PCH + t;
    const double sig = inSig, sigInv = umt_rcp(sig);
    double Q[4], src[4], A1[4][2];
#pragma unroll
    for (int c = 0; c < 4; c++) {   // records are labelled by solve position: R.ci[p] = local corner of position p
      const int lc = R.ci[c];
      double a = inPsi[0], b = inSt[0];
#pragma unroll
      for (int k = 1; k < 4; k++) { a = lc == k ? inPsi[k] : a; b = lc == k ? inSt[k] : b; }
      Q[c] = b + tau * a;
      src[c] = R.vol[c] * Q[c];
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int f = 0; f < 2; f++) {
        A1[c][f] = 0.0;
        const double az = R.az[c][f];
        if (az > 0.0) {
          const int cez = rz_nb(c, f);
          const double Rr = R.rez[c][f], dq = Q[c] - Q[cez];
          double A0;
          if (R.inMask & (1u << (2 * c + f))) {
            const double ar = R.area[c];
            const double sigA = sig * ar, sigA2 = sigA * sigA;
            const double gnum = az * az * (FOURALPHA * sigA2 + az * (4.0 * sigA + 3.0 * az));
            const double gden = ar * (4.0 * sigA * sigA2 + az * (6.0 * sigA2 + 2.0 * az * (2.0 * sigA + az)));
            const double rd = Rr * umt_rcp(gnum + gden * sig);
            A1[c][f] = rd * (ar * gnum * sig);
            A0 = rd * (0.5 * az * gden * dq - ar * gnum * Q[c]);
          } else {
            A0 = 0.5 * (Rr * az) * dq * sigInv;
          }
          src[c] += A0;
          src[cez] -= A0;
        }
      }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      out[(size_t)c * PCH] = src[c];
      out[(size_t)(4 + 2 * c) * PCH] = A1[c][0];
      out[(size_t)(5 + 2 * c) * PCH] = A1[c][1];
      out[(size_t)(12 + c) * PCH] = umt_rcp(R.sumArea[c] + sig * R.vol[c]);
    }
  };
  auto chain_half = [&](int s) {
    const RZStep st = steps[s];
    if (zi >= st.n) return;
    const int a = st.angle;
    const RZRecS &R = *reinterpret_cast<const RZRecS *>(recBuf + ((size_t)(s % 3) * L.ZCH + zi) * RZL_REC_STRIDE);
    const double *in = statBuf + (size_t)(s % 2) * RZL_NSTAT * PCH + t;
    double *psi1A = P.psi1 + (size_t)a * slab;
    double *psimL = P.psim + (size_t)P.level[a] * nc * G;
    const int c0 = R.c0;
    double pm[4], u[4][2], src[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      pm[c] = __ldcg(&psimL[(size_t)(c0 + (int)R.ci[c]) * G + g]);
#pragma unroll
      for (int f = 0; f < 2; f++) {
        u[c][f] = 0.0;
        if (R.inMask & (1u << (2 * c + f))) u[c][f] = __ldcg(&psi1A[(size_t)R.row[c][f] * G + g]);
      }
      src[c] = in[(size_t)c * PCH];
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const double a1 = in[(size_t)(4 + 2 * c + f) * PCH];
        src[c] = fma(R.k1b[c][f] + a1, u[c][f], src[c]);
        src[rz_nb(c, f)] += -a1 * u[c][f];
      }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const double p = (src[c] + R.areaFac[c] * pm[c]) * in[(size_t)(12 + c) * PCH];
      src[c] = p;
      src[rz_nb(c, 0)] += (R.rez[c][0] * R.az[c][0]) * p;
      src[rz_nb(c, 1)] += (R.rez[c][1] * R.az[c][1]) * p;
    }
    const bool starting = P.start[a] != 0, fin = P.finishNext[a] != 0;
    double *psi1N = psi1A + slab;
    const double w1 = P.tauW1[a], w2 = P.tauW2[a];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const size_t r = (size_t)(c0 + (int)R.ci[c]) * G + g;
      const double p = src[c];
      const double pmn = starting ? p : w1 * p - w2 * pm[c];
      __stcg(&psimL[r], pmn); __stcg(&psi1A[r], p);
      if (fin) psi1N[r] = pmn;
#pragma unroll
      for (int f = 0; f < 2; f++) {
        if (R.exitMask & (1u << (2 * c + f))) {
          const int row = R.row[c][f];
          psi1A[(size_t)row * G + g] = p;
          if (fin) psi1N[(size_t)row * G + g] = pmn;
        }
      }
    }
  };

  // prologue: the static group brings the records of steps 0 and 1 and the inputs of step 0 in, prepares step 0, and leaves the
  // inputs of step 1 and the zone info of step 2 in flight
  const bool mine = zi < L.ZCH;
  if (!chain) {
    copy_records(0); copy_records(1);
    if (mine) { load_zinfo(0); load_inputs(0, aPsi, aSt, aSig); load_zinfo(1); }
    umt_cp_async_wait_all();
  }
  __syncthreads();
  if (!chain && mine) { static_half(0, aPsi, aSt, aSig); load_inputs(1, aPsi, aSt, aSig); load_zinfo(2); }
  __syncthreads();
  // one iteration: the chain group does step s; the static group completes the records of step s + 1 (issued an iteration ago),
  // issues the records and the inputs of step s + 2 (into the set that is free) and prepares step s + 1 from the other set
#define RZL_ITER(S, CUR_PSI, CUR_ST, CUR_SIG, NXT_PSI, NXT_ST, NXT_SIG)                              \
  if (chain) {                                                                                       \
    if (mine) chain_half(S);                                                                         \
  } else {                                                                                           \
    umt_cp_async_wait_all();                                                                         \
    asm volatile("bar.sync 1, %0;" ::"r"(CT) : "memory");                                            \
    copy_records((S) + 2);                                                                           \
    if (mine) {                                                                                      \
      load_inputs((S) + 2, NXT_PSI, NXT_ST, NXT_SIG);                                                \
      load_zinfo((S) + 3);                                                                           \
      if ((S) + 1 < nS) static_half((S) + 1, CUR_PSI, CUR_ST, CUR_SIG);                              \
    }                                                                                                \
  }                                                                                                  \
  __syncthreads();
  for (int s = 0; s < nS; s += 2) {
    RZL_ITER(s, aPsi, aSt, aSig, bPsi, bSt, bSig)
    if (s + 1 < nS) { RZL_ITER(s + 1, bPsi, bSt, bSig, aPsi, aSt, aSig) }
  }
#undef RZL_ITER
}

// Dataflow kernel.  The same work items in the same topological order, but nothing waits for a whole plane: every warp takes
// 32 (zone, group) pairs of an item, runs the static half, then each lane polls exactly the values its pair needs -- the Psi1 rows
// behind its incident faces and the previous angle's PsiM of its corners -- until they are no longer marked (RZ_SENTINEL), solves
// and stores.  A dependent hop costs one L2 store -> load round trip instead of store -> fence -> counter -> poll -> load, and a
// plane's slowest item no longer holds up the whole next plane.  Producers always hold earlier tickets than their consumers and
// a warp takes a ticket only when it is free, so every polled value is being computed by a resident warp: no deadlock.
// Not used with cycle lists / direct-solve zones (a consumer must see the OLD value there, which needs the plane order) nor
// with reflecting boundaries (staged launches).
template <int MC>
__global__ void __launch_bounds__(RZ_BLOCK, RZ_MIN_CTAS) sweeprz_flow_kernel(SweepRZParams P, int warpsPerItem) {
  const int G = P.G, lane = threadIdx.x & 31;
  const int nUnits = P.nItems * warpsPerItem;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&P.counters[0], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nUnits) break;
    const int it = t / warpsPerItem, sub = t - it * warpsPerItem;
    const WorkItem w = P.items[it];
    const int npairs = (w.zend - w.zbeg) * G;
    const int idx = sub * 32 + lane;
    if (idx < npairs) {
      const int zi = idx / G, g = idx - zi * G;
      ZoneStatic<MC> Z;
      zone_static_rz<MC>(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + w.zbeg + zi], g, Z);
      zone_solve_rz<MC, true>(P, w.angle, g, Z);
    }
    __syncwarp();
  }
}

// marks the corner rows of Psi1 and the PsiM slab of every swept angle as "not computed yet"
__global__ void sweeprz_mark_kernel(double *psi1, double *psimA, const int *nHyp, size_t slab, size_t ncG) {
  const int a = blockIdx.y;
  if (nHyp[a] == 0) return;
  ulonglong2 *p1 = reinterpret_cast<ulonglong2 *>(psi1 + (size_t)a * slab), *p2 = reinterpret_cast<ulonglong2 *>(psimA + (size_t)a * ncG);
  const ulonglong2 v = make_ulonglong2(RZ_SENTINEL, RZ_SENTINEL);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ncG / 2; i += (size_t)gridDim.x * blockDim.x) { p1[i] = v; p2[i] = v; }
}

// Chain kernel.  In r-z the groups are independent of each other and the xi-levels are independent of each other, while
// within (level, group) everything is one dependency chain: the planes of an angle in order, the angles of the level in order
// (PsiM).  One CTA therefore owns one (xi-level, block of gb groups) and walks that chain by itself: every thread takes
// (zone, group) pairs of the current plane, a __syncthreads() separates planes -- no tickets, no global counters, no device-
// scope fences, no polling.  A plane step costs the latency of one zone solve instead of the ~10 us of a global
// signal/poll round trip, which is what bounded sweeprz_kernel (planes of ~80 zones x 2400 dependent steps per level).
// The static half of the next pair is computed before the barrier (it does not depend on the previous plane).
template <int MC>
__global__ void __launch_bounds__(256) sweeprz_chain_kernel(SweepRZParams P) {
  const int lev = blockIdx.x / P.nGroupBlocks, g0 = (blockIdx.x - lev * P.nGroupBlocks) * P.gb;
  const int gb = min(P.gb, P.G - g0);
  const int tid = threadIdx.x, T = blockDim.x;
  for (int k = 0; k < P.maxAngLevel; k++) {
    const int a = P.levelAngles[lev * P.maxAngLevel + k];
    if (a < 0) break;
    const int *nextZ = P.nextZ + (size_t)a * P.nz;
    const int *off = P.planeOff + (size_t)a * P.hypStride;
    const int nh = P.nHyp[a];
    for (int p = 0; p < nh; p++) {
      const int zbeg = off[p], npairs = (off[p + 1] - zbeg) * gb;
      int idx = tid;
      ZoneStatic<MC> Z;
      int g = 0;
      if (idx < npairs) { const int zi = idx / gb; g = g0 + idx - zi * gb; zone_static_rz<MC>(P, a, nextZ[zbeg + zi], g, Z); }
      __syncthreads();   // the previous plane (and the previous angle of the level) is complete and visible to the CTA
      while (idx < npairs) {
        zone_solve_rz<MC>(P, a, g, Z);
        idx += T;
        if (idx < npairs) { const int zi = idx / gb; g = g0 + idx - zi * gb; zone_static_rz<MC>(P, a, nextZ[zbeg + zi], g, Z); }
      }
    }
  }
}

}  // namespace

// Work items of the RZ sweep in a topological order of both dependencies (host side, once per schedule).
int umt_build_items_rz(umt_ctx *ctx, std::vector<WorkItem> &items, int zpi) {
  int maxHyp = 0;
  int r = umt_build_items_rz_set(ctx->nz, ctx->NA, ctx->nHyp, ctx->zonesInPlane, ctx->nextZ, ctx->h_start, zpi, items, ctx->h_level, ctx->nLevels, maxHyp);
  if (r) return r;
  if (ctx->nStages > 1) {   // reflecting boundaries: one launch per stage, items of a stage contiguous (their relative order kept)
    std::stable_sort(items.begin(), items.end(), [&](const WorkItem &x, const WorkItem &y) { return ctx->stageOf[x.angle] < ctx->stageOf[y.angle]; });
    size_t i = 0;
    for (int st = 0; st < ctx->nStages; st++) {
      while (i < items.size() && ctx->stageOf[items[i].angle] <= st) i++;
      ctx->stageItemBegin[st + 1] = (int)i;
    }
  }
  if (ctx->device < 0) return r;
  // tables of the chain kernel: swept angles of every level in order, plane offsets, and its geometry: gb groups per CTA so
  // that a (corner, group block) access is at least one 32-byte sector, as many threads as the largest plane has pairs
  const int NA = ctx->NA, nL = ctx->nLevels;
  int maxAng = 1, maxPlane = 1;
  std::vector<int> cnt(nL, 0);
  for (int a = 0; a < NA; a++) if (ctx->nHyp[a] > 0) maxAng = std::max(maxAng, ++cnt[ctx->h_level[a]]);
  std::vector<int> la((size_t)nL * maxAng, -1), po((size_t)NA * (maxHyp + 1), 0), nh(NA, 0);
  std::fill(cnt.begin(), cnt.end(), 0);
  for (int a = 0; a < NA; a++) {
    nh[a] = ctx->nHyp[a];
    if (nh[a] == 0) continue;
    la[(size_t)ctx->h_level[a] * maxAng + cnt[ctx->h_level[a]]++] = a;
    int o = 0;
    for (int p = 0; p < nh[a]; p++) { po[(size_t)a * (maxHyp + 1) + p] = o; o += ctx->zonesInPlane[a][p]; maxPlane = std::max(maxPlane, ctx->zonesInPlane[a][p]); }
    po[(size_t)a * (maxHyp + 1) + nh[a]] = o;
  }
  int gb = std::min(ctx->G, 4);
  if (const char *e = getenv("UMT_RZ_GROUP_BLOCK")) gb = std::max(1, std::min(ctx->G, atoi(e)));
  ctx->rz_gb = gb;
  ctx->rz_maxAngLevel = maxAng;
  ctx->rz_threads = std::max(32, std::min(256, (maxPlane * gb + 31) / 32 * 32));
  ctx->rz_chain = false;   // measured at configs[1]: item kernel 16.7 ms, chain kernel 33.7 ms (a pair solve is ~12 us of dependent latency either way; the item kernel also pipelines the angles of a level)
  if (const char *e = getenv("UMT_RZ_KERNEL")) ctx->rz_chain = std::string(e) == "chain";
  auto up = [&](int **d, const std::vector<int> &h) -> int {
    if (*d) { cudaFree(*d); *d = nullptr; }
    UMT_CUDA(ctx, cudaMalloc((void **)d, sizeof(int) * std::max<size_t>(h.size(), 1)));
    UMT_CUDA(ctx, umt_memcpy(ctx, *d, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice));
    return UMT_OK;
  };
  // dataflow kernel: previous swept angle of each level; usable without cycle lists, direct-solve zones and reflecting boundaries,
  // with G even (16-byte marks) and items of at most RZ_BLOCK pairs
  std::vector<int> prevA(NA, -1);
  {
    std::vector<int> lastOf(nL, -1);
    for (int a = 0; a < NA; a++) { if (nh[a] == 0) continue; prevA[a] = lastOf[ctx->h_level[a]]; lastOf[ctx->h_level[a]] = a; }
  }
  bool plain = ctx->nStages <= 1 && ctx->G % 2 == 0 && ctx->G <= RZ_BLOCK && ((size_t)(ctx->nc + ctx->nb) * ctx->G) % 2 == 0;
  for (int a = 0; a < NA && plain; a++) {
    if (ctx->numCycles[a] > 0) plain = false;
    for (int z : ctx->nextZ[a]) if (z < 0) { plain = false; break; }
  }
  ctx->rz_flow = false;
  if (const char *e = getenv("UMT_RZ_KERNEL")) ctx->rz_flow = plain && (std::string(e) == "flow" || std::string(e) == "recflow");
  // record kernel: quads, every item's pairs fit one CTA
  ctx->rz_rec = ctx->maxCorner <= 4 && ctx->G <= RZ_BLOCK && ctx->zones_per_item * ctx->G <= RZ_BLOCK;
  if (const char *e = getenv("UMT_RZ_KERNEL")) {   // rec, recflow, pipe and lc all run from the records
    const std::string k(e);
    if (k != "rec" && k != "recflow" && k != "pipe" && k != "lc") ctx->rz_rec = false;
  }
  ctx->rz_recs_valid = false;
  if ((r = up(&ctx->d_rzPrev, prevA))) return r;
  if ((r = up(&ctx->d_rzLevelAngles, la))) return r;
  if ((r = up(&ctx->d_rzPlaneOff, po))) return r;
  return up(&ctx->d_rzNHyp, nh);
}

// the same for any r-z angle set (the Sn set above, the GTA set in gta_rz.cu); angles with nHyp == 0 (finishing directions) get no items
int umt_build_items_rz_set(int nz, int NA, const std::vector<int> &nHyp, const std::vector<std::vector<int>> &zonesInPlane,
                           const std::vector<std::vector<int>> &nextZ, const std::vector<unsigned char> &start, int zpi,
                           std::vector<WorkItem> &items, std::vector<int> &levelOut, int &nLevelsOut, int &maxHypOut) {
  int maxHyp = 0;
  for (int a = 0; a < NA; a++) maxHyp = std::max(maxHyp, nHyp[a]);
  maxHypOut = maxHyp;
  // xi-levels: a level starts at a starting direction; finishing directions are not swept
  std::vector<int> level(NA, 0), prev(NA, -1);
  int lev = -1, last = -1;
  for (int a = 0; a < NA; a++) {
    if (start[a] || lev < 0) { lev++; last = -1; }
    level[a] = lev;
    if (nHyp[a] == 0) continue;
    prev[a] = last;
    last = a;
  }
  std::vector<std::vector<int>> planeOf(NA), nItemsPlane(NA), planeStart(NA), tdone(NA), dep2(NA);
  struct Key { int t, a, p; };
  std::vector<Key> keys;
  for (int a = 0; a < NA; a++) {
    const int nh = nHyp[a];
    if (nh == 0) continue;
    planeOf[a].assign(nz, 0);
    planeStart[a].assign(nh + 1, 0);
    nItemsPlane[a].assign(nh, 0);
    tdone[a].assign(nh, 0);
    dep2[a].assign(nh, -1);
    for (int p = 0; p < nh; p++) {
      const int n = zonesInPlane[a][p];
      planeStart[a][p + 1] = planeStart[a][p] + n;
      nItemsPlane[a][p] = (n + zpi - 1) / zpi;
      for (int i = planeStart[a][p]; i < planeStart[a][p + 1]; i++) planeOf[a][std::abs(nextZ[a][i]) - 1] = p;
    }
    const int pa = prev[a];
    for (int p = 0; p < nh; p++) {
      int t = p > 0 ? tdone[a][p - 1] : 0;
      if (pa >= 0) {
        int q = 0;
        for (int i = planeStart[a][p]; i < planeStart[a][p + 1]; i++) q = std::max(q, planeOf[pa][std::abs(nextZ[a][i]) - 1]);
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
    const int n = zonesInPlane[a][p], z0 = planeStart[a][p];
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
  levelOut = level;
  nLevelsOut = lev + 1;
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
  P.abortFlag = ctx->d_abort; P.spinLimit = ctx->spinLimit;
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int) * (1 + ctx->nCounters), ctx->stream));
  // Set%PsiM = 0 at the start of every flux pass (SetSweep.F90:94-96)
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_psim, 0, sizeof(double) * (size_t)ctx->nLevels * ctx->nc * ctx->G, ctx->stream));
  if (ctx->rz_chain && ctx->nStages <= 1) {
    P.gb = ctx->rz_gb; P.nGroupBlocks = (ctx->G + ctx->rz_gb - 1) / ctx->rz_gb; P.maxAngLevel = ctx->rz_maxAngLevel; P.hypStride = ctx->maxHyp + 1;
    P.levelAngles = ctx->d_rzLevelAngles; P.planeOff = ctx->d_rzPlaneOff; P.nHyp = ctx->d_rzNHyp;
    void (*ck)(SweepRZParams) = ctx->maxCorner <= 4 ? sweeprz_chain_kernel<4> : sweeprz_chain_kernel<MAXC2>;
    ck<<<ctx->nLevels * P.nGroupBlocks, ctx->rz_threads, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
    ctx->last_launches += 1;
    return UMT_OK;
  }
  if (ctx->rz_flow && !ctx->rz_rec && ctx->nStages <= 1) {
    const size_t ncG = (size_t)ctx->nc * ctx->G, slab = (size_t)(ctx->nc + ctx->nb) * ctx->G;
    if (!ctx->d_rzPsimA) UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_rzPsimA, sizeof(double) * ncG * ctx->NA));
    P.psimA = ctx->d_rzPsimA; P.prevAngle = ctx->d_rzPrev;
    sweeprz_mark_kernel<<<dim3(ctx->sm_count * 2, ctx->NA), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_rzPsimA, ctx->d_rzNHyp, slab, ncG);
    UMT_CUDA(ctx, cudaGetLastError());
    void (*fk)(SweepRZParams, int) = ctx->maxCorner <= 4 ? sweeprz_flow_kernel<4> : sweeprz_flow_kernel<MAXC2>;
    int occ = 0;
    UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fk, RZ_BLOCK, 0));
    if (occ < 1) occ = 1;
    const int wpi = (std::min(RZ_BLOCK, ctx->zones_per_item * ctx->G) + 31) / 32;
    const int grid = std::max(1, std::min(ctx->sm_count * occ, (ctx->nItems * wpi + RZ_BLOCK / 32 - 1) / (RZ_BLOCK / 32)));
    fk<<<grid, RZ_BLOCK, 0, ctx->stream>>>(P, wpi);
    UMT_CUDA(ctx, cudaGetLastError());
    UMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));   // checked after the next sync
    ctx->last_launches += 2;
    return UMT_OK;
  }
  if (ctx->rz_rec) {
    P.nHyp = ctx->d_rzNHyp;
    RZRec *recs = static_cast<RZRec *>(ctx->d_rzRecs);
    if (!ctx->rz_recs_valid) {
      const size_t n = (size_t)ctx->NA * ctx->nz;
      if (ctx->d_rzRecs) { cudaFree(ctx->d_rzRecs); ctx->d_rzRecs = nullptr; }
      UMT_CUDA(ctx, cudaMalloc(&ctx->d_rzRecs, sizeof(RZRec) * n));
      recs = static_cast<RZRec *>(ctx->d_rzRecs);
      // canonical labelling first; if any zone does not fit, the by-corner records for the whole mesh
      int h_bad = 0;
      if (!ctx->d_rzBad) UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_rzBad, sizeof(int)));
      int *d_bad = ctx->d_rzBad;
      UMT_CUDA(ctx, cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
      bool canon = true;
      if (const char *e = getenv("UMT_RZ_CANON")) canon = atoi(e) != 0;
      if (canon) {
        rz_rec_build_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(P, recs, 1, d_bad);
        UMT_CUDA(ctx, cudaGetLastError());
        UMT_CUDA(ctx, cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        canon = h_bad == 0;
      }
      if (!canon) {
        rz_rec_build_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(P, recs, 0, d_bad);
        UMT_CUDA(ctx, cudaGetLastError());
        UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      }
      ctx->rz_canon = canon;
      ctx->rz_recs_valid = true;
      ctx->rz_lc_zch = 0;   // the level-chain kernel's step table belongs to the old schedule
      // what the loader warp of the pipelined kernel needs per (angle, zone in sweep order): first corner row, zone
      std::vector<int2> zinfo(n, make_int2(0, 0));
      for (int a = 0; a < ctx->NA; a++) {
        if (ctx->nHyp[a] == 0) continue;
        for (int i = 0; i < ctx->nz; i++) {
          const int z = std::abs(ctx->nextZ[a][i]) - 1;
          zinfo[(size_t)a * ctx->nz + i] = make_int2(ctx->h_cOffSet[z], z);
        }
      }
      if (ctx->d_zinfo) { cudaFree(ctx->d_zinfo); ctx->d_zinfo = nullptr; }
      UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_zinfo, sizeof(int2) * n));
      UMT_CUDA(ctx, umt_memcpy(ctx, ctx->d_zinfo, zinfo.data(), sizeof(int2) * n, cudaMemcpyHostToDevice));
    }
    const bool flow = ctx->rz_flow && ctx->nStages <= 1;
    bool lc = ctx->rz_canon && !flow && ctx->nStages <= 1;
    if (const char *e = getenv("UMT_RZ_KERNEL")) { if (std::string(e) != "lc") lc = false; } else lc = false;
    if (lc) {
      // steps of every xi-level: its swept angles in order, their planes in order, large planes cut into chunks of ZCH zones
      int gb = 2;
      if (const char *e = getenv("UMT_RZ_LC_GB")) gb = std::max(1, std::min(8, atoi(e)));
      gb = std::min(gb, ctx->G);
      int maxPlane = 1;
      for (int a = 0; a < ctx->NA; a++) for (int p = 0; p < ctx->nHyp[a]; p++) maxPlane = std::max(maxPlane, ctx->zonesInPlane[a][p]);
      int ZCH = std::min(maxPlane, std::min(RZL_MAX_CT / gb, (int)(220000 / (3 * RZL_REC_STRIDE + 2 * RZL_NSTAT * 8 * gb))));
      if (const char *e = getenv("UMT_RZ_LC_ZONES")) ZCH = std::max(1, std::min(ZCH, atoi(e)));
      if (!ctx->d_rzSteps || ctx->rz_lc_zch != ZCH) {
        const int nL = ctx->nLevels;
        std::vector<std::vector<RZStep>> per(nL);
        for (int a = 0; a < ctx->NA; a++) {
          if (ctx->nHyp[a] == 0) continue;
          int z0 = 0;
          for (int p = 0; p < ctx->nHyp[a]; p++) {
            const int n = ctx->zonesInPlane[a][p];
            for (int o = 0; o < n; o += ZCH) per[ctx->h_level[a]].push_back(RZStep{a, z0 + o, std::min(ZCH, n - o), 0});
            z0 += n;
          }
        }
        size_t maxSteps = 1;
        for (auto &v : per) maxSteps = std::max(maxSteps, v.size());
        std::vector<RZStep> flat((size_t)nL * maxSteps, RZStep{0, 0, 0, 0});
        std::vector<int> ns(nL, 0);
        for (int l = 0; l < nL; l++) { ns[l] = (int)per[l].size(); std::copy(per[l].begin(), per[l].end(), flat.begin() + (size_t)l * maxSteps); }
        if (ctx->d_rzSteps) { cudaFree(ctx->d_rzSteps); ctx->d_rzSteps = nullptr; }
        if (ctx->d_rzNSteps) { cudaFree(ctx->d_rzNSteps); ctx->d_rzNSteps = nullptr; }
        UMT_CUDA(ctx, cudaMalloc(&ctx->d_rzSteps, sizeof(RZStep) * flat.size()));
        UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_rzNSteps, sizeof(int) * nL));
        UMT_CUDA(ctx, umt_memcpy(ctx, ctx->d_rzSteps, flat.data(), sizeof(RZStep) * flat.size(), cudaMemcpyHostToDevice));
        UMT_CUDA(ctx, umt_memcpy(ctx, ctx->d_rzNSteps, ns.data(), sizeof(int) * nL, cudaMemcpyHostToDevice));
        ctx->rz_lc_zch = ZCH; ctx->rz_lc_maxSteps = (int)maxSteps;
      }
      RZLParams L;
      L.recs = recs; L.zinfo = ctx->d_zinfo; L.steps = static_cast<const RZStep *>(ctx->d_rzSteps); L.nSteps = ctx->d_rzNSteps;
      L.maxSteps = ctx->rz_lc_maxSteps; L.gb = gb; L.nGB = (ctx->G + gb - 1) / gb; L.ZCH = ZCH; L.CT = (ZCH * gb + 31) / 32 * 32;
      const size_t smemL = (size_t)3 * ZCH * RZL_REC_STRIDE + (size_t)2 * RZL_NSTAT * 8 * ZCH * gb;
      UMT_CUDA(ctx, cudaFuncSetAttribute(sweeprz_lc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemL));
      sweeprz_lc_kernel<<<ctx->nLevels * L.nGB, 2 * L.CT, smemL, ctx->stream>>>(P, L);
      UMT_CUDA(ctx, cudaGetLastError());
      ctx->last_launches += 1;
      return UMT_OK;
    }
    bool pipe = ctx->rz_canon && !flow && ctx->G % 2 == 0 && ctx->zones_per_item <= RZP_ZMAX && ctx->zones_per_item * ctx->G <= 64;
    // measured at configs[1] size (40x40 tiles, G = 64): 12-17 ms against the record kernel's 5.2 ms.  The r-z sweep has 4 xi-levels of 5
    // chained angles and ~80-zone planes: about a thousand items are ready at any time, so items a CTA holds ahead of their turn
    // (tickets in batches, in-order release within the CTA) keep engines idle behind a waiting item while ready items sit in another
    // CTA's queue; the record kernel takes a ticket only when its CTA is free.  Kept for experiments: UMT_RZ_KERNEL=pipe.
    if (const char *e = getenv("UMT_RZ_KERNEL")) { if (std::string(e) != "pipe") pipe = false; } else pipe = false;
    if (pipe) {
      RZPGeom gm;
      const int zpi = ctx->zones_per_item, G = ctx->G;
      gm.zpi = zpi;
      gm.offSt = zpi * 4 * G * 8;
      gm.offSigt = 2 * gm.offSt;
      gm.offRecs = gm.offSigt + zpi * G * 8;
      gm.stageBytes = (gm.offRecs + zpi * (int)sizeof(RZRec) + 127) / 128 * 128;
      const int budget = (227 * 1024) / RZP_MIN_CTAS - 1024 - RZP_CTL_BYTES;
      gm.nStages = std::min(RZP_MAX_STAGES, std::max(RZP_NE + 1, std::min(RZP_NE + 4, budget / gm.stageBytes)));
      if (const char *e = getenv("UMT_RZ_STAGES")) gm.nStages = std::max(RZP_NE + 1, std::min(RZP_MAX_STAGES, atoi(e)));
      const size_t smemP = RZP_CTL_BYTES + (size_t)gm.nStages * gm.stageBytes;
      UMT_CUDA(ctx, cudaFuncSetAttribute(sweeprz_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
      int occP = 0;
      UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occP, sweeprz_pipe_kernel, RZP_THREADS, smemP));
      if (occP < 1) UMT_FAIL(ctx, UMT_ERR_CUDA, "sweeprz_pipe_kernel does not fit on an SM");
      if (const char *e = getenv("UMT_RZ_CTAS_PER_SM")) occP = std::max(1, std::min(occP, atoi(e)));
      const int nSt = std::max(1, ctx->nStages);
      for (int st = 0; st < nSt; st++) {   // reflecting boundaries: snreflect, then the angles of this stage (one stage otherwise)
        int begin = 0, end = ctx->nItems;
        if (ctx->nStages > 1) {
          begin = ctx->stageItemBegin[st]; end = ctx->stageItemBegin[st + 1];
          int r = UMT_OK;
          if (ctx->have_comm_order && !ctx->shared.empty()) r = umt_exchange_stage(ctx, st);   // SendFlux / RecvFlux of this sweep step
          if (r) return r;
          r = umt_launch_reflect(ctx, st);
          if (r) return r;
          if (end == begin) continue;
          if (st > 0) UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), ctx->stream));
        }
        P.items = ctx->d_items + begin; P.nItems = end - begin;
        const int grid = std::max(1, std::min(ctx->sm_count * occP, (P.nItems + 7) / 8));
        sweeprz_pipe_kernel<<<grid, RZP_THREADS, smemP, ctx->stream>>>(P, recs, ctx->d_zinfo, gm);
        UMT_CUDA(ctx, cudaGetLastError());
        ctx->last_launches += 1;
      }
      return UMT_OK;
    }
    const size_t smem = sizeof(RZRec) * (size_t)ctx->zones_per_item;
    void (*rk)(SweepRZParams, const RZRec *) = flow ? (ctx->rz_canon ? sweeprz_rec_kernel<true, true> : sweeprz_rec_kernel<true, false>)
                                                    : (ctx->rz_canon ? sweeprz_rec_kernel<false, true> : sweeprz_rec_kernel<false, false>);
    if (flow) {
      const size_t ncG = (size_t)ctx->nc * ctx->G, slabE = (size_t)(ctx->nc + ctx->nb) * ctx->G;
      if (!ctx->d_rzPsimA) UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_rzPsimA, sizeof(double) * ncG * ctx->NA));
      P.psimA = ctx->d_rzPsimA; P.prevAngle = ctx->d_rzPrev;
      sweeprz_mark_kernel<<<dim3(ctx->sm_count * 2, ctx->NA), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_rzPsimA, ctx->d_rzNHyp, slabE, ncG);
      UMT_CUDA(ctx, cudaGetLastError());
      ctx->last_launches += 1;
    }
    int occ = 0;
    UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rk, RZ_BLOCK, smem));
    if (occ < 1) occ = 1;
    const int nSt = std::max(1, ctx->nStages);
    for (int st = 0; st < nSt; st++) {   // reflecting boundaries: snreflect, then the angles of this stage (one stage otherwise)
      int begin = 0, end = ctx->nItems;
      if (ctx->nStages > 1) {
        begin = ctx->stageItemBegin[st]; end = ctx->stageItemBegin[st + 1];
        int r = UMT_OK;
        if (ctx->have_comm_order && !ctx->shared.empty()) r = umt_exchange_stage(ctx, st);   // SendFlux / RecvFlux of this sweep step
        if (r) return r;
        r = umt_launch_reflect(ctx, st);
        if (r) return r;
        if (end == begin) continue;
        if (st > 0) UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), ctx->stream));
      }
      P.items = ctx->d_items + begin; P.nItems = end - begin;
      const int grid = std::max(1, std::min(ctx->sm_count * occ, P.nItems));
      rk<<<grid, RZ_BLOCK, smem, ctx->stream>>>(P, recs);
      UMT_CUDA(ctx, cudaGetLastError());
      ctx->last_launches += 1;
    }
    if (flow) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    return UMT_OK;
  }
  void (*kern)(SweepRZParams) = ctx->maxCorner <= 4 ? sweeprz_kernel<4> : sweeprz_kernel<MAXC2>;
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, RZ_BLOCK, 0));
  if (occ < 1) occ = 1;
  if (ctx->nStages <= 1) {
    const int grid = std::max(1, std::min(ctx->sm_count * occ, ctx->nItems));
    kern<<<grid, RZ_BLOCK, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
    ctx->last_launches += 1;
    return UMT_OK;
  }
  for (int st = 0; st < ctx->nStages; st++) {   // reflecting boundaries: snreflect, then the angles of this stage
    const int begin = ctx->stageItemBegin[st], end = ctx->stageItemBegin[st + 1];
    int r = UMT_OK;
    if (ctx->have_comm_order && !ctx->shared.empty()) r = umt_exchange_stage(ctx, st);   // SendFlux / RecvFlux of this sweep step
    if (r) return r;
    r = umt_launch_reflect(ctx, st);
    if (r) return r;
    if (end == begin) continue;
    if (st > 0) UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), ctx->stream));   // the ticket; plane counters persist
    P.items = ctx->d_items + begin; P.nItems = end - begin;
    const int grid = std::max(1, std::min(ctx->sm_count * occ, P.nItems));
    kern<<<grid, RZ_BLOCK, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
    ctx->last_launches += 1;
  }
  return UMT_OK;
}
