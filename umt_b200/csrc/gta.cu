// Grey transport acceleration (GTA) on the device: 3-D, the "new" GTA solver (the variant the
// reference's own GPU path implements, gpu/SweepGreyUCBxyz_OMPOL.F90 / rt/GTASolver_OMPOL.F90).
//
//   rt/setGTAOpacity.F90:10-113 (setGTAOpacityNEW)             -> gta_opacity_kernel
//   rt/getCollisionRate.F90:10-97                              -> collision_rate_kernel
//   snac/GTASweep.F90:9-168 (GTA%ID == 1)                      -> gta_device_sweep (TsaSource, angle loop)
//   snac/SweepGreyUCBxyz.F90:12-355 (KernelNew)                -> gta_sweep_kernel
//   snac/InitSweepGreyUCBxyz.F90:10-253                        -> gta_init_tt_kernel
//   snac/UpdateScalarIntensity.F90:11-170                      -> gta_scalar_kernel
//   rt/GreySweep.F90:12-48 (GreySweepNEW)                      -> gta_grey_sweep
//   rt/scat_prod.F90, scat_prod1.F90, rt/GTASolver.F90:42-425  -> umt_gta_solve (BiCGSTAB, host loop, device vectors)
//   rt/addGreyCorrections.F90:70-91                            -> add_corrections_kernel
//
// The grey sweep has one group, so the parallel axes are zones-in-plane x the 8 S2 ordinates: one thread
// owns one (zone, angle) and solves its corners; the same ticket / per-(angle, plane) counter scheme as the
// multigroup sweep runs all 8 angles in one persistent launch.  The incident-only part pInc is kept per
// angle and summed in fixed angle order (deterministic PhiInc, no float atomics).  All vectors of the
// Krylov iteration stay on the device; only the scalars of the inner products come back to the host.
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include <string>

#include "umt_internal.h"
#include "device_util.h"

namespace {

constexpr int MAXC = 8, MAXCF = 3;
constexpr double FOURALPHA = 1.82;
constexpr double PI = 3.14159265358979323846;

#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double dot3(const double *om, const double *A) {
  return __dadd_rn(__dadd_rn(__dmul_rn(om[0], A[0]), __dmul_rn(om[1], A[1])), __dmul_rn(om[2], A[2]));
}

struct GtaSweepParams {
  int nc, nb, nz, nItems;
  const int *numCorner, *cOffSet, *nCFaces, *cFP, *cEZ;
  const double *Volume, *Afp, *Aez, *omega;
  const int *nextZ;
  const unsigned char *nextC;
  const WorkItem *items;
  int *counters;
  const double *sigTotal, *sigtInv, *tsa;
  double *tpsi, *pinc;
  int *abortFlag;         // dataflow kernel: watchdog of the polling loop (device_util.h)
  unsigned spinLimit;
  const unsigned *pollMask;   // dataflow kernel: (nAng, nz in sweep order) bit 3c+f: the zone behind FP face f of corner c sits in an earlier plane
};

// SweepGreyUCBxyzKernelNew for one (zone, angle)
__device__ void gta_solve_zone(const GtaSweepParams &P, int a, int zone0) {
  const int nc = P.nc;
  const double om[3] = {P.omega[3 * a], P.omega[3 * a + 1], P.omega[3 * a + 2]};
  double *tpsi = P.tpsi + (size_t)a * (nc + P.nb);
  double *pincA = P.pinc + (size_t)a * nc;
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  double Q[MAXC], src[MAXC], denom[MAXC], pinc[MAXC];
  int nxez[MAXC], ez_exit[MAXC][MAXCF];
  double coefpsi[MAXC][MAXCF];
  for (int c = 0; c < nCorner; c++) {
    const double t = P.tsa[c0 + c];
    Q[c] = P.sigtInv[c0 + c] * t;
    src[c] = P.Volume[c0 + c] * t;
    pinc[c] = 0.0;
    nxez[c] = 0;
  }
  for (int c = 0; c < nCorner; c++) {
    const int cc = c0 + c;
    const double sigv = P.Volume[cc] * P.sigTotal[cc];
    double dn = sigv;
    const int nCF = P.nCFaces[cc];
    double afp[MAXCF], psifp[MAXCF];
    for (int f = 0; f < nCF; f++) {
      afp[f] = dot3(om, P.Afp + ((size_t)cc * MAXCF + f) * 3);
      psifp[f] = 0.0;
      if (afp[f] > 0.0) dn += afp[f];
      else if (afp[f] < 0.0) {
        psifp[f] = __ldcg(&tpsi[P.cFP[cc * MAXCF + f]]);
        src[c] -= afp[f] * psifp[f];
        pinc[c] -= afp[f] * psifp[f];
      }
    }
    for (int f = 0; f < nCF; f++) {
      const double aez = dot3(om, P.Aez + ((size_t)cc * MAXCF + f) * 3);
      const int cez = P.cEZ[cc * MAXCF + f];
      if (cez > c) {
        if (aez > 0.0) { ez_exit[c][nxez[c]] = cez; coefpsi[c][nxez[c]] = aez; nxez[c]++; }
        else if (aez < 0.0) { ez_exit[cez][nxez[cez]] = c; coefpsi[cez][nxez[cez]] = -aez; nxez[cez]++; }
      }
      if (aez > 0.0) {
        double psi_opp = 0.0, area_opp = 0.0;
        dn += aez;
        int ifp = (f + 1) % nCF;
        if (afp[ifp] < 0.0) { area_opp = -afp[ifp]; psi_opp = -afp[ifp] * psifp[ifp]; }
        for (int k = 2; k <= nCF - 2; k++) {
          ifp = (ifp + 1) % nCF;
          if (afp[ifp] < 0.0) { area_opp -= afp[ifp]; psi_opp -= afp[ifp] * psifp[ifp]; }
        }
        double sez;
        if (area_opp > 0.0) {
          psi_opp = psi_opp / area_opp;
          const double sigv2 = sigv * sigv;
          const double gnum = aez * aez * (FOURALPHA * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
          const double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + aez * sigv * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
          sez = gtau * sigv * (psi_opp - Q[c]) + 0.5 * aez * (1.0 - gtau) * (Q[c] - Q[cez]);
          pinc[c] += gtau * sigv * psi_opp;
          pinc[cez] -= gtau * sigv * psi_opp;
        } else {
          sez = 0.5 * aez * (Q[c] - Q[cez]);
        }
        src[c] += sez;
        src[cez] -= sez;
      }
    }
    denom[c] = dn;
  }
  for (int i = 0; i < nCorner; i++) {
    const int c = nextC[c0 + i];
    const double p = src[c] / denom[c];
    const double q = pinc[c] / denom[c];
    src[c] = p;
    pinc[c] = q;
    for (int k = 0; k < nxez[c]; k++) {
      src[ez_exit[c][k]] += coefpsi[c][k] * p;
      pinc[ez_exit[c][k]] += coefpsi[c][k] * q;
    }
  }
  for (int c = 0; c < nCorner; c++) {
    const int cc = c0 + c;
    tpsi[cc] = src[c];
    pincA[cc] = pinc[c];
    const int nCF = P.nCFaces[cc];
    for (int f = 0; f < nCF; f++) {
      const int row = P.cFP[cc * MAXCF + f];
      if (row >= nc && dot3(om, P.Afp + ((size_t)cc * MAXCF + f) * 3) > 0.0) tpsi[row] = src[c];   // PsiB(b, Angle) <- tPsi
    }
  }
}

// The same zone solve spread over one warp: lane = (corner, face slot) = (lane >> 2, lane & 3), so the 24 corner faces of a hex
// load their area vectors, connectivity and upstream fluxes in parallel (coalesced: a zone's corners are contiguous), the closure
// terms of all EZ faces are evaluated at once, each corner *pulls* the contribution of the neighbour across each of its EZ faces
// (deterministic, no atomics), and only the 8 corner solves remain sequential (warp shuffles).  Needs three faces on every corner
// and at most 8 corners; other zones take gta_solve_zone on lane 0.
struct GtaZoneStatic {   // everything that does not depend on other work items (loaded before the dependency wait)
  double afp, aez, vol, sigv, q, tsaVol;
  int cez, row, zone0, nCorner, c0, myNext;
  bool fast;
};

__device__ __forceinline__ void gta_zone_static(const GtaSweepParams &P, int a, int zone0, int lane, GtaZoneStatic &Z) {
  const double om[3] = {P.omega[3 * a], P.omega[3 * a + 1], P.omega[3 * a + 2]};
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  const int c = lane >> 2, f = lane & 3;
  Z.zone0 = zone0; Z.nCorner = nCorner; Z.c0 = c0;
  const bool threeFaces = lane >= nCorner || P.nCFaces[c0 + lane] == 3;
  Z.fast = nCorner <= MAXC && __all_sync(0xffffffffu, threeFaces);
  Z.afp = 0.0; Z.aez = 0.0; Z.vol = 0.0; Z.sigv = 0.0; Z.q = 0.0; Z.tsaVol = 0.0; Z.cez = c < MAXC ? c : 0; Z.row = 0; Z.myNext = 0;
  if (!Z.fast) return;
  if (lane < nCorner) Z.myNext = P.nextC[(size_t)a * P.nc + c0 + lane];   // the corner order does not depend on other zones: off the chain
  if (c < nCorner) {
    const int cc = c0 + c;
    const double t = P.tsa[cc];
    Z.vol = P.Volume[cc];
    Z.sigv = Z.vol * P.sigTotal[cc];
    Z.q = P.sigtInv[cc] * t;
    Z.tsaVol = Z.vol * t;
    if (f < 3) {
      Z.afp = dot3(om, P.Afp + ((size_t)cc * MAXCF + f) * 3);
      Z.aez = dot3(om, P.Aez + ((size_t)cc * MAXCF + f) * 3);
      Z.cez = P.cEZ[cc * MAXCF + f];
      Z.row = P.cFP[cc * MAXCF + f];
    }
  }
}

// FLOW (dataflow kernel): the corner rows of tPsi were marked "not computed yet" before the launch; every face lane polls the value
// behind its incident face until it is real (the data are their own completion flags), and the corner fluxes are stored with
// relaxed device-scope stores -- no counters, no fences, no barriers, and a zone starts as soon as ITS upstream zones are done.
template <bool FLOW>
__device__ __forceinline__ void gta_zone_solve_warp(const GtaSweepParams &P, int a, int lane, const GtaZoneStatic &Z, unsigned pollMask = 0u) {
  const unsigned FULL = 0xffffffffu;
  const int nc = P.nc;
  double *tpsi = P.tpsi + (size_t)a * (nc + P.nb);
  double *pincA = P.pinc + (size_t)a * nc;
  const int nCorner = Z.nCorner, c0 = Z.c0;
  const int c = lane >> 2, f = lane & 3;
  const bool valid = c < nCorner && f < 3;
  const double afp = Z.afp, aez = Z.aez, sigv = Z.sigv, q = Z.q;
  const int cez = Z.cez;
  double psifp = 0.0;
  if (FLOW) {
    // Only faces whose upstream zone the schedule puts in an earlier plane are waited for.  An incident face the schedule does not
    // order (omega.A = 0 up to rounding: the sign is noise and so is the weight of its flux) takes what is there, 0 if nothing yet --
    // what the counter kernel and the reference read from a zeroed tPsi.
    const bool inc = valid && afp < 0.0;
    const bool need = inc && Z.row < nc && ((pollMask >> (3 * c + f)) & 1u);
    if (inc && !need) {
      const unsigned long long v = umt_ld_relaxed_u64(&tpsi[Z.row]);
      psifp = v == UMT_SENTINEL ? 0.0 : __longlong_as_double((long long)v);
    }
    unsigned polls = 0;
    for (;;) {
      unsigned long long v = 0ull;
      if (need) v = umt_ld_relaxed_u64(&tpsi[Z.row]);
      const bool ok = !need || v != UMT_SENTINEL;
      if (need) psifp = __longlong_as_double((long long)v);
      if (__all_sync(FULL, ok)) break;
      __nanosleep(40);
      if (__any_sync(FULL, umt_spin_expired(polls, P.abortFlag, P.spinLimit))) break;
    }
  } else {
    psifp = (valid && afp < 0.0) ? __ldcg(&tpsi[Z.row]) : 0.0;
  }
  const int myNext = Z.myNext;
  // the FP face "opposite" EZ face f is face (f+1) mod 3 of the same corner (SweepGreyUCBxyz.F90:263-270)
  const int lop = (lane & ~3) | (f == 2 ? 0 : f + 1);
  const double afpo = __shfl_sync(FULL, afp, lop), psio = __shfl_sync(FULL, psifp, lop);
  const double qcez = __shfl_sync(FULL, q, cez << 2);
  double dsrc = 0.0, dpinc = 0.0, dden = 0.0, sez = 0.0, ginc = 0.0;
  if (valid) {
    if (afp > 0.0) dden += afp;
    else if (afp < 0.0) { dsrc -= afp * psifp; dpinc -= afp * psifp; }
    if (aez > 0.0) {
      dden += aez;
      if (afpo < 0.0) {
        const double area_opp = -afpo;
        const double psi_opp = (-afpo * psio) / area_opp;
        const double sigv2 = sigv * sigv;
        const double gnum = aez * aez * (FOURALPHA * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
        const double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + aez * sigv * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
        sez = gtau * sigv * (psi_opp - q) + 0.5 * aez * (1.0 - gtau) * (q - qcez);
        ginc = gtau * sigv * psi_opp;
      } else {
        sez = 0.5 * aez * (q - qcez);
      }
      dsrc += sez;
      dpinc += ginc;
    }
  }
  // pull what the neighbour across my EZ face pushes into my corner: its face whose cEZ points back at me
  int fb = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int ck = __shfl_sync(FULL, cez, (cez << 2) | k);
    if (ck == c) fb = k;
  }
  const int ln = (cez << 2) | fb;
  const double sez_n = __shfl_sync(FULL, sez, ln), ginc_n = __shfl_sync(FULL, ginc, ln), aez_n = __shfl_sync(FULL, aez, ln);
  if (valid && aez_n > 0.0) { dsrc -= sez_n; dpinc -= ginc_n; }
  // per-corner sums over its face lanes (slot 3 carries zeros)
  dsrc += __shfl_xor_sync(FULL, dsrc, 1); dsrc += __shfl_xor_sync(FULL, dsrc, 2);
  dpinc += __shfl_xor_sync(FULL, dpinc, 1); dpinc += __shfl_xor_sync(FULL, dpinc, 2);
  dden += __shfl_xor_sync(FULL, dden, 1); dden += __shfl_xor_sync(FULL, dden, 2);
  double src = Z.tsaVol + dsrc, pinc = dpinc;
  const double denom = sigv + dden;
  // corner solves in nextC order, each pushed into its downstream corners (SweepGreyUCBxyz.F90:320-335)
  for (int i = 0; i < nCorner; i++) {
    const int cs = __shfl_sync(FULL, myNext, i);
    const double d_c = __shfl_sync(FULL, denom, cs << 2);
    const double psi = __shfl_sync(FULL, src, cs << 2) / d_c;
    const double pin = __shfl_sync(FULL, pinc, cs << 2) / d_c;
    if (c == cs) { src = psi; pinc = pin; }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double ak = __shfl_sync(FULL, aez, (cs << 2) | k);
      const int tk = __shfl_sync(FULL, cez, (cs << 2) | k);
      if (ak > 0.0 && c == tk && c != cs) { src += ak * psi; pinc += ak * pin; }
    }
  }
  if (c < nCorner && f == 0) {
    pincA[c0 + c] = pinc;
    if (FLOW) umt_st_relaxed_f64(&tpsi[c0 + c], src); else tpsi[c0 + c] = src;
  }
  if (valid && Z.row >= nc && afp > 0.0) tpsi[Z.row] = src;   // PsiB(b, Angle) <- tPsi
}

constexpr int GTA_WARPS = 8;   // warps per CTA = zones per work item

#ifndef GTA_MINB
#define GTA_MINB 4   // CTAs per SM the grey sweep is compiled for (register cap 65536 / (256 GTA_MINB))
#endif
__global__ void __launch_bounds__(GTA_WARPS * 32, GTA_MINB) gta_sweep_kernel(GtaSweepParams P) {
  __shared__ int s_item;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&P.counters[0], 1);   // taken only when free to work on it: a ticket held ahead of time blocks a ready item behind a waiting one
    __syncthreads();
    const int it = s_item;
    if (it >= P.nItems) break;
    const WorkItem w = P.items[it];
    const int zi = w.zbeg + warp;
    const bool active = zi < w.zend;
    GtaZoneStatic Z;
    if (active) gta_zone_static(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + zi], lane, Z);
    if (threadIdx.x == blockDim.x - 1 && w.wait_idx >= 0)
      while (ld_acquire(&P.counters[1 + w.wait_idx]) < w.wait_count) __nanosleep(20);
    __syncthreads();
    if (active) {
      if (Z.fast) gta_zone_solve_warp<false>(P, w.angle, lane, Z);
      else if (lane == 0) gta_solve_zone(P, w.angle, Z.zone0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + w.signal_idx]) : "memory");
    }
  }
}

// Dataflow variant (meshes whose zones all take the warp solve, no reflecting boundaries, no direct-solve zones): every warp takes
// ONE zone of the item list through its own ticket, in the list's topological order, so the zones a warp polls for are held by
// warps with earlier tickets: no deadlock.  A plane-to-plane hop is one L2 store -> load round trip instead of
// store -> fence -> counter -> poll -> barrier -> load.
__global__ void __launch_bounds__(GTA_WARPS * 32, GTA_MINB) gta_sweep_flow_kernel(GtaSweepParams P) {
  const int lane = threadIdx.x & 31;
  const int nUnits = P.nItems * GTA_WARPS;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&P.counters[0], 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nUnits) break;
    const int it = t / GTA_WARPS, sub = t - it * GTA_WARPS;
    const WorkItem w = P.items[it];
    const int zi = w.zbeg + sub;
    if (zi < w.zend) {
      GtaZoneStatic Z;
      gta_zone_static(P, w.angle, P.nextZ[(size_t)w.angle * P.nz + zi], lane, Z);
      gta_zone_solve_warp<true>(P, w.angle, lane, Z, P.pollMask[(size_t)w.angle * P.nz + zi]);   // every zone is "fast" (host check)
    }
    __syncwarp();
  }
}

// marks the corner rows of tPsi of every angle as "not computed yet"
__global__ void gta_mark_kernel(double *tpsi, int nc, int rows) {
  const double mark = __longlong_as_double((long long)UMT_SENTINEL);
  double *t = tpsi + (size_t)blockIdx.y * rows;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) t[i] = mark;
}

// snreflect for the grey sweeps: tPsi tail rows are PsiB(:, angle)
__global__ void gta_reflect_kernel(double *tpsi, const int4 *ops, int rows, int nc) {
  const int4 o = ops[blockIdx.y];   // x = Minc, y = Mref, z = first boundary element, w = count
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < o.w) tpsi[(size_t)o.x * rows + nc + o.z + i] = tpsi[(size_t)o.y * rows + nc + o.z + i];
}

// TsaSource = wtiso (GreySigScat P + GreySource)   (GTASweep.F90:84-89)
__global__ void gta_tsa_kernel(const double *P, const double *sigScat, const double *greySource, double wtiso, double *tsa, int n) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) tsa[c] = wtiso * (sigScat[c] * P[c] + greySource[c]);
}

// PhiInc = sum_a w_a pInc_a in angle order (SweepGreyUCBxyz.F90:126-128)
__global__ void gta_phiinc_kernel(const double *pinc, const double *w, int nAng, int nc, double *phiInc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  double s = 0.0;
  for (int a = 0; a < nAng; a++) s = s + w[a] * pinc[(size_t)a * nc + c];
  phiInc[c] = s;
}

struct ZoneParams {
  int nz, nc, mC, nAng;
  const int *numCorner, *cOffSet, *nCFaces, *cEZ;
  const double *Volume, *Afp, *Aez, *omega, *weight, *sigTotal, *sigScat, *greySource, *phiInc;
  double *TT, *P;
  double wtiso;
};

// InitGreySweepUCBxyz: TT(:, corners of zone) = sum over the GTA angles of w Pvv
__global__ void __launch_bounds__(64) gta_init_tt_kernel(ZoneParams Z) {
  const int zone = blockIdx.x * blockDim.x + threadIdx.x;
  if (zone >= Z.nz) return;
  const int nCorner = Z.numCorner[zone], c0 = Z.cOffSet[zone], mC = Z.mC;
  double T[MAXC][MAXC];   // T[c][c1] accumulates TT(c, c0+c1)
  for (int i = 0; i < MAXC; i++) for (int j = 0; j < MAXC; j++) T[i][j] = 0.0;
  for (int a = 0; a < Z.nAng; a++) {
    const double om[3] = {Z.omega[3 * a], Z.omega[3 * a + 1], Z.omega[3 * a + 2]};
    const double quadwt = Z.weight[a];
    int nxez[MAXC], need[MAXC], ez_exit[MAXC][MAXCF];
    double denom[MAXC], coefpsi[MAXC][MAXCF], Sigt[MAXC], Pvv[MAXC][MAXC];   // Pvv[row][col]
    for (int i = 0; i < MAXC; i++) { nxez[i] = 0; need[i] = 0; for (int j = 0; j < MAXC; j++) Pvv[i][j] = 0.0; }
    for (int c = 0; c < nCorner; c++) { Pvv[c][c] = Z.Volume[c0 + c]; Sigt[c] = Z.sigTotal[c0 + c]; }
    for (int c = 0; c < nCorner; c++) {
      const int cc = c0 + c;
      const double sigv = Z.Volume[cc] * Sigt[c];
      double dn = sigv;
      const int nCF = Z.nCFaces[cc];
      double afp[MAXCF];
      for (int f = 0; f < nCF; f++) {
        afp[f] = dot3(om, Z.Afp + ((size_t)cc * MAXCF + f) * 3);
        if (afp[f] > 0.0) dn += afp[f];
      }
      for (int f = 0; f < nCF; f++) {
        const double aez = dot3(om, Z.Aez + ((size_t)cc * MAXCF + f) * 3);
        const int cez = Z.cEZ[cc * MAXCF + f];
        if (cez > c) {
          if (aez > 0.0) { need[cez]++; ez_exit[c][nxez[c]] = cez; coefpsi[c][nxez[c]] = aez; nxez[c]++; }
          else if (aez < 0.0) { need[c]++; ez_exit[cez][nxez[cez]] = c; coefpsi[cez][nxez[cez]] = -aez; nxez[cez]++; }
        }
        if (aez > 0.0) {
          dn += aez;
          double area_opp = 0.0;
          if (nCF == 3) {
            const int ifp = (f + 1) % 3;
            if (afp[ifp] < 0.0) area_opp = -afp[ifp];
          } else {
            int ifp = f;
            for (int k = 1; k <= nCF - 2; k++) { ifp = (ifp + 1) % nCF; if (afp[ifp] < 0.0) area_opp -= afp[ifp]; }
          }
          double B1, B2;
          if (area_opp > 0.0) {
            const double sigv2 = sigv * sigv;
            const double gnum = aez * aez * (FOURALPHA * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
            const double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + aez * sigv * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
            const double B0 = 0.5 * aez * (1.0 - gtau);
            B1 = (B0 - gtau * sigv) / Sigt[c];
            B2 = B0 / Sigt[cez];
          } else {
            B1 = 0.5 * aez / Sigt[c];
            B2 = 0.5 * aez / Sigt[cez];
          }
          Pvv[c][c] += B1; Pvv[cez][c] -= B2; Pvv[c][cez] -= B1; Pvv[cez][cez] += B2;
        }
      }
      denom[c] = dn;
    }
    for (int i = 0; i < nCorner; i++) {
      int c = 0;   // minloc: first minimum
      for (int k = 1; k < nCorner; k++) if (need[k] < need[c]) c = k;
      const double dInv = 1.0 / denom[c];
      for (int c1 = 0; c1 < nCorner; c1++) Pvv[c1][c] = dInv * Pvv[c1][c];
      for (int k = 0; k < nxez[c]; k++) {
        const int cez = ez_exit[c][k];
        const double coef = coefpsi[c][k];
        need[cez]--;
        for (int c1 = 0; c1 < nCorner; c1++) Pvv[c1][cez] += coef * Pvv[c1][c];
      }
      need[c] = 99;
    }
    for (int c1 = 0; c1 < nCorner; c1++)
      for (int c = 0; c < nCorner; c++) T[c][c1] = T[c][c1] + quadwt * Pvv[c][c1];
  }
  for (int c1 = 0; c1 < nCorner; c1++)
    for (int c = 0; c < mC; c++) Z.TT[(size_t)(c0 + c1) * mC + c] = c < nCorner ? T[c][c1] : 0.0;
}

// InitGreySweepUCBxyz for meshes of hexahedra (8 corners, 3 faces each) and the 8 S2 ordinates: the thread-per-zone kernel above keeps
// 1.5 KB of dynamically indexed arrays per thread in local memory and took 5.7 ms at 192 k zones.  Here 64 threads work on one zone:
// thread (angle a, row r) owns row r of Pvv in registers (the elimination acts on the rows independently, InitSweepGreyUCBxyz.F90's
// loop over c1) and, as corner r, computes that corner's share of the set-up (denominator, closure coefficients B1/B2, exit
// coefficients); the 8 lanes of an angle trade these by shuffles.  The angle sum runs in angle order through shared memory.
constexpr int TTH_ZONES = 2;   // zones per CTA (128 threads)
__global__ void __launch_bounds__(TTH_ZONES * 64) gta_init_tt_hex_kernel(ZoneParams Z) {
  __shared__ double sm[TTH_ZONES][8][8][9];
  const int zl = threadIdx.x >> 6, t = threadIdx.x & 63, a = t >> 3, r = t & 7;
  const int zone = blockIdx.x * TTH_ZONES + zl;
  const bool live = zone < Z.nz;
  const int zq = live ? zone : Z.nz - 1;            // idle threads of the last CTA redo the last zone (shuffles stay convergent)
  const int c0 = Z.cOffSet[zq], cc = c0 + r;
  const int base = threadIdx.x & 24;                // first lane of my angle's group of 8 within the warp
  const double om[3] = {Z.omega[3 * a], Z.omega[3 * a + 1], Z.omega[3 * a + 2]};
  // ---- my corner's share of the set-up -------------------------------------------------------------------------------------
  const double vol = Z.Volume[cc], sigt = Z.sigTotal[cc], sigv = vol * sigt;
  double afp[3], aez[3], B1[3], B2[3], xcoef[3];
  int cez[3];
  double dn = sigv;
#pragma unroll
  for (int f = 0; f < 3; f++) {
    afp[f] = dot3(om, Z.Afp + ((size_t)cc * MAXCF + f) * 3);
    if (afp[f] > 0.0) dn += afp[f];
  }
  unsigned mine = 0;                                // bits 4f..4f+2: cEZ of face f, bit 4f+3: Pvv closure terms present (aez > 0)
#pragma unroll
  for (int f = 0; f < 3; f++) {
    aez[f] = dot3(om, Z.Aez + ((size_t)cc * MAXCF + f) * 3);
    cez[f] = Z.cEZ[cc * MAXCF + f];
    B1[f] = B2[f] = 0.0;
    if (aez[f] > 0.0) {
      dn += aez[f];
      const int ifp = (f + 1) % 3;
      const double area_opp = afp[ifp] < 0.0 ? -afp[ifp] : 0.0;
      const double sigtN = Z.sigTotal[c0 + cez[f]], ae = aez[f];
      if (area_opp > 0.0) {
        const double sigv2 = sigv * sigv;
        const double gnum = ae * ae * (FOURALPHA * sigv2 + ae * (4.0 * sigv + 3.0 * ae));
        const double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + ae * sigv * (6.0 * sigv2 + 2.0 * ae * (2.0 * sigv + ae)));
        const double B0 = 0.5 * ae * (1.0 - gtau);
        B1[f] = (B0 - gtau * sigv) / sigt;
        B2[f] = B0 / sigtN;
      } else {
        B1[f] = 0.5 * ae / sigt;
        B2[f] = 0.5 * ae / sigtN;
      }
      mine |= 8u << (4 * f);
    }
    mine |= (unsigned)cez[f] << (4 * f);
    // exit r -> cez through this face: the pair's lower corner decides (its A_ez, as the reference's loop over cez > c does)
    if (cez[f] > r) xcoef[f] = aez[f] > 0.0 ? aez[f] : 0.0;
    else {
      const int nb = c0 + cez[f];
      double alow = 0.0;
#pragma unroll
      for (int g = 0; g < 3; g++)
        if (Z.cEZ[nb * MAXCF + g] == r) alow = dot3(om, Z.Aez + ((size_t)nb * MAXCF + g) * 3);
      xcoef[f] = alow < 0.0 ? -alow : 0.0;
    }
    if (xcoef[f] > 0.0) mine |= 0x1000u << f;       // bit 12+f: face f is an exit of this corner
  }
  // ---- row r of Pvv and the dependency counts (4 bits per corner, the same word in the 8 lanes of an angle) -----------------------
  double row[8];
#pragma unroll
  for (int c = 0; c < 8; c++) row[c] = c == r ? vol : 0.0;
  unsigned need = 0;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const unsigned w = __shfl_sync(0xffffffffu, mine, base + c);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double b1 = __shfl_sync(0xffffffffu, B1[f], base + c), b2 = __shfl_sync(0xffffffffu, B2[f], base + c);
      const int ce = (w >> (4 * f)) & 7;
      if ((w >> (4 * f)) & 8u) {                    // Pvv(c,c) += B1, Pvv(cez,c) -= B2, Pvv(c,cez) -= B1, Pvv(cez,cez) += B2
        if (r == c) { row[c] += b1; addto(row, ce, -b1); }
        if (r == ce) { row[c] -= b2; addto(row, ce, b2); }
      }
      if ((w >> (12 + f)) & 1u) need += 1u << (4 * ce);
    }
  }
  // ---- elimination in the order of the fewest unresolved upstream corners (first minimum, as minloc) -----------------------------
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    int c = 0;
    unsigned best = need & 15u;
#pragma unroll
    for (int k = 1; k < 8; k++) { const unsigned v = (need >> (4 * k)) & 15u; if (v < best) { best = v; c = k; } }
    const double dInv = 1.0 / __shfl_sync(0xffffffffu, dn, base + c);
    const double v = dInv * pick(row, c);
    put(row, c, v);
    const unsigned w = __shfl_sync(0xffffffffu, mine, base + c);
#pragma unroll
    for (int f = 0; f < 3; f++) {
      const double coef = __shfl_sync(0xffffffffu, xcoef[f], base + c);
      if ((w >> (12 + f)) & 1u) {
        const int ce = (w >> (4 * f)) & 7;
        need -= 1u << (4 * ce);
        addto(row, ce, coef * v);
      }
    }
    need |= 15u << (4 * c);                         // done
  }
  // ---- TT(c, c0 + c1) = sum over the angles, in angle order, of w_a Pvv_a(c, c1) ------------------------------------------------
#pragma unroll
  for (int c = 0; c < 8; c++) sm[zl][a][r][c] = row[c];
  __syncthreads();
  if (!live) return;
  const int c1 = t >> 3, c = t & 7;                 // consecutive threads write consecutive TT entries
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s = s + Z.weight[k] * sm[zl][k][c][c1];
  Z.TT[(size_t)(c0 + c1) * Z.mC + c] = s;
}

// ScalarIntensityDecompose + ScalarIntensitySolve for one zone
__global__ void __launch_bounds__(64) gta_scalar_kernel(ZoneParams Z, int withSource) {
  const int zone = blockIdx.x * blockDim.x + threadIdx.x;
  if (zone >= Z.nz) return;
  const int n = Z.numCorner[zone], c0 = Z.cOffSet[zone], mC = Z.mC;
  double Phi[MAXC];
  double *TT = Z.TT;
#define TTF(i, j) TT[(size_t)(c0 + (j)) * mC + (i)]   /* TT(i+1, c0+j+1) */
  if (withSource) {
    for (int c = 0; c < n; c++) {
      double ph = Z.phiInc[c0 + c];
      for (int cc = 0; cc < n; cc++) {
        ph = ph + TTF(cc, c) * Z.wtiso * Z.greySource[c0 + cc];
        TTF(cc, c) = -Z.wtiso * Z.sigScat[c0 + cc] * TTF(cc, c);
      }
      Phi[c] = ph;
      TTF(c, c) = 1.0 + TTF(c, c);
    }
    for (int i = 0; i < n; i++) {
      double t = 0.0;
      for (int k = 0; k < i; k++) t = t + TTF(k, i) * TTF(i, k);
      TTF(i, i) = TTF(i, i) - t;
      const double diagInv = 1.0 / TTF(i, i);
      for (int j = i + 1; j < n; j++) {
        t = 0.0;
        double v = 0.0;
        for (int k = 0; k < i; k++) { t = t + TTF(k, i) * TTF(j, k); v = v + TTF(k, j) * TTF(i, k); }
        TTF(j, i) = TTF(j, i) - t;
        TTF(i, j) = diagInv * (TTF(i, j) - v);
      }
    }
  } else {
    for (int c = 0; c < n; c++) Phi[c] = Z.phiInc[c0 + c];
  }
  for (int j = 1; j < n; j++) {
    double t = 0.0;
    for (int i = 0; i < j; i++) t = t - TTF(i, j) * Phi[i];
    Phi[j] = Phi[j] + t;
  }
  Phi[n - 1] = Phi[n - 1] / TTF(n - 1, n - 1);
  for (int k = n - 2; k >= 0; k--) {
    double t = 0.0;
    for (int i = k + 1; i < n; i++) t = t + Phi[i] * TTF(i, k);
    Phi[k] = (Phi[k] - t) / TTF(k, k);
  }
  for (int c = 0; c < n; c++) Z.P[c0 + c] = Phi[c];
#undef TTF
}

// ScalarIntensitySolve alone (the LU factors are in place) for meshes of hexahedra: 8 lanes per zone, lane j holds column j of the
// factors (TT(0..7, c0+j), one 64-byte row of the array) in registers.  Forward substitution by columns, back substitution with every
// partial sum accumulated in ascending index order: the same sums in the same order as the one-thread version above.
__global__ void __launch_bounds__(256) gta_scalar_solve_hex_kernel(ZoneParams Z) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  int zone = gt >> 3;
  const int j = gt & 7, base = threadIdx.x & 24;
  const bool live = zone < Z.nz;
  if (!live) zone = Z.nz - 1;
  const int c0 = Z.cOffSet[zone];
  double col[8];   // col[i] = TT(i, c0 + j)
  const double2 *src = reinterpret_cast<const double2 *>(Z.TT + (size_t)(c0 + j) * 8);
#pragma unroll
  for (int i = 0; i < 4; i++) { const double2 v = src[i]; col[2 * i] = v.x; col[2 * i + 1] = v.y; }
  double phi = Z.phiInc[c0 + j];
  // Phi(j) = Phi(j) + sum_{i<j} -TT(i,j) Phi(i)
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const double pi = __shfl_sync(0xffffffffu, j == i ? phi + t : 0.0, base + i);   // lane i's value is final once i terms are in
    if (j == i) { phi = phi + t; t = 0.0; }
    if (j > i) t = t - col[i] * pi;
  }
  if (j == 7) phi = phi + t;
  // Phi(k) = (Phi(k) - sum_{i>k} Phi(i) TT(i,k)) / TT(k,k), k = 7 .. 0: TT(i,k) is element i of column k, lane k's col[i]
  double diag = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) diag = j == i ? col[i] : diag;
  if (j == 7) phi = phi / diag;
#pragma unroll
  for (int k = 6; k >= 0; k--) {
    double s = 0.0;
#pragma unroll
    for (int i = k + 1; i < 8; i++) {
      const double pi = __shfl_sync(0xffffffffu, phi, base + i);
      s = s + pi * col[i];
    }
    if (j == k) phi = (phi - s) / diag;
  }
  if (live) Z.P[c0 + j] = phi;
}

// setGTAOpacityNEW per corner; Chi (nc, G) rescaled in place
__global__ void gta_opacity_kernel(int nc, int G, double tau, const int *c2z, const double *siga, const double *sigs, const double *eta,
                                   const double *Volume, double *chi, double *sigTotal, double *sigScat, double *sigScatVol, double *sigtInv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int zone = c2z[c];
  double SigtInvAve = 0.0, Sigt2InvAve = 0.0, SigaAve = 0.0;
  for (int g = 0; g < G; g++) {
    const double a = siga[(size_t)zone * G + g], s = sigs[(size_t)zone * G + g];
    const double SigtInv = 1.0 / (a + s + tau);
    const double x = chi[(size_t)c * G + g];
    SigtInvAve = SigtInvAve + x * SigtInv;
    Sigt2InvAve = Sigt2InvAve + x * SigtInv * SigtInv;
    SigaAve = SigaAve + x * a * SigtInv;
    chi[(size_t)c * G + g] = x * SigtInv;
  }
  double greysigt, greysiga, greysigs;
  if (SigtInvAve > 0.0) {
    for (int g = 0; g < G; g++) chi[(size_t)c * G + g] = chi[(size_t)c * G + g] / SigtInvAve;
    greysigt = SigtInvAve / Sigt2InvAve;
    greysiga = tau + (1.0 - eta[c]) * SigaAve / SigtInvAve;
    greysigs = greysigt - greysiga;
  } else {
    greysigt = tau; greysiga = tau; greysigs = 0.0;
  }
  const double scatRatio = greysigs / greysigt;
  if (scatRatio <= 1.0e-10) { sigScat[c] = 0.0; sigTotal[c] = greysiga; }
  else { sigScat[c] = greysigs; sigTotal[c] = greysigt; }
  sigScatVol[c] = sigScat[c] * Volume[c];
  sigtInv[c] = 1.0 / sigTotal[c];
}

// getCollisionRate
__global__ void collision_rate_kernel(int nc, int G, const int *c2z, const double *eta, const double *siga, const double *sigs,
                                      const double *phi, double *greySource, int residualFlag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  const int zone = c2z[c];
  double s = 0.0;
  for (int g = 0; g < G; g++) s = s + (eta[c] * siga[(size_t)zone * G + g] + sigs[(size_t)zone * G + g]) * phi[(size_t)c * G + g];
  greySource[c] = residualFlag == 0 ? s : s - greySource[c];
}

// PhiTotal(g,c) += GreyCorrection(c) Chi(g,c)
__global__ void add_corrections_kernel(size_t n, int G, const double *corr, const double *chi, double *phi) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) phi[i] = phi[i] + corr[i / G] * chi[i];
}

// radEnergy(zone) = sum_c V_c sum_g PhiTotal(g,c) / VolumeZone   (GTASolver.F90:128-142)
__global__ void rad_energy_kernel(int nz, int G, const int *numCorner, const int *cOffSet, const double *Volume, const double *phi,
                                  double *volZone, double *radEnergy) {
  const int zone = blockIdx.x * blockDim.x + threadIdx.x;
  if (zone >= nz) return;
  double e = 0.0, vz = 0.0;
  for (int c = cOffSet[zone]; c < cOffSet[zone] + numCorner[zone]; c++) {
    double sumRad = 0.0;
    for (int g = 0; g < G; g++) sumRad = sumRad + phi[(size_t)c * G + g];
    e = e + Volume[c] * sumRad;
    vz = vz + Volume[c];
  }
  volZone[zone] = vz;
  radEnergy[zone] = e / vz;
}

// deterministic reductions: per-block partials in block order, then one thread adds them up in order
__global__ void __launch_bounds__(256) dot_partial_kernel(const double *x, const double *y, const double *w, int n, double *partial) {
  __shared__ double red[256];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (y ? x[i] * y[i] : x[i]) * w[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k]; __syncthreads(); }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
// (one CTA: the partials are fetched by all threads at once, then summed by one thread in index order -- deterministic, and not a
// chain of dependent global loads)
constexpr int RED_MAX = 1024;
__global__ void __launch_bounds__(256) sum_partials_kernel(const double *partial, int n, double *out) {
  __shared__ double sh[RED_MAX];
  for (int i = threadIdx.x; i < n; i += blockDim.x) sh[i] = partial[i];
  __syncthreads();
  if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < n; i++) s += sh[i]; *out = s; }
}

// zone corrections and the error norms of GTASolver.F90:331-375; partial[0..nb)=errL2, [nb..2nb)=phiL2, [2nb..3nb)=max rel err
__global__ void __launch_bounds__(256) zone_error_kernel(int nz, const int *numCorner, const int *cOffSet, const double *Volume,
                                                         const double *volZone, const double *corr, const double *radEnergy,
                                                         double *pzOld, double *partial) {
  __shared__ double r0[256], r1[256], r2[256];
  double e = 0.0, p = 0.0, m = 0.0;
  for (int zone = blockIdx.x * blockDim.x + threadIdx.x; zone < nz; zone += gridDim.x * blockDim.x) {
    double pz = 0.0;
    for (int c = cOffSet[zone]; c < cOffSet[zone] + numCorner[zone]; c++) pz = pz + Volume[c] * corr[c];
    pz = pz / volZone[zone];
    const double errZone = pz - pzOld[zone];
    e += volZone[zone] * (errZone * errZone);
    const double phiNew = radEnergy[zone] + pz;
    p += volZone[zone] * (phiNew * phiNew);
    if (phiNew != 0.0) m = fmax(m, fabs(errZone / phiNew));
    if (phiNew != phiNew || errZone != errZone) m = CUDART_INF;   // NaN: the reference aborts here
    pzOld[zone] = pz;
  }
  r0[threadIdx.x] = e; r1[threadIdx.x] = p; r2[threadIdx.x] = m;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) { r0[threadIdx.x] += r0[threadIdx.x + k]; r1[threadIdx.x] += r1[threadIdx.x + k]; r2[threadIdx.x] = fmax(r2[threadIdx.x], r2[threadIdx.x + k]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x] = r0[0]; partial[gridDim.x + blockIdx.x] = r1[0]; partial[2 * gridDim.x + blockIdx.x] = r2[0]; }
}
__global__ void __launch_bounds__(256) zone_error_finish_kernel(const double *partial, int nb, double *out3) {
  __shared__ double sh[3 * RED_MAX];
  for (int i = threadIdx.x; i < 3 * nb; i += blockDim.x) sh[i] = partial[i];
  __syncthreads();
  if (threadIdx.x < 3) {   // one thread per quantity, each in index order
    const double *q = sh + threadIdx.x * nb;
    double v = 0.0;
    if (threadIdx.x < 2) for (int i = 0; i < nb; i++) v += q[i];
    else for (int i = 0; i < nb; i++) v = fmax(v, q[i]);
    out3[threadIdx.x] = v;
  }
}

// Krylov vector updates (GTASolver.F90:259-322); evaluation order as written there
__global__ void k_sub(double *out, const double *a, const double *b, size_t n) {             // out = a - b
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}
__global__ void k_axmy(double *y, double alpha, const double *x, size_t n) {                  // y = y - alpha x
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = y[i] - alpha * x[i];
}
__global__ void k_corr(double *corr, double alpha, const double *D, double omega, const double *R, size_t n, int withR) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) corr[i] = withR ? corr[i] + alpha * D[i] + omega * R[i] : corr[i] + alpha * D[i];
}
__global__ void k_dir(double *D, const double *R, double beta, double omega, const double *A, size_t n) {   // D = R + beta (D - omega A)
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) D[i] = R[i] + beta * (D[i] - omega * A[i]);
}

inline unsigned nblk(size_t n, int b = 256) { return (unsigned)((n + b - 1) / b); }

template <class T>
int dalloc(umt_ctx *ctx, T **p, size_t n, bool zero = true) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  UMT_CUDA(ctx, cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(n, 1)));
  if (zero) UMT_CUDA(ctx, cudaMemsetAsync(*p, 0, sizeof(T) * std::max<size_t>(n, 1), ctx->stream));
  return UMT_OK;
}

constexpr int RED_BLOCKS = 296;   // 2 per SM
static_assert(RED_BLOCKS <= RED_MAX, "the finishing kernels stage the partials in shared memory");

int need_gta(umt_ctx *ctx) {
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  if (!ctx->gta.have_opacity) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA opacities not set (umt_gta_set_opacity / umt_gta_compute_opacity)");
  return UMT_OK;
}

double gta_wtiso(const umt_ctx *ctx) { return ctx->ndim == 3 ? 1.0 / (4.0 * PI) : 1.0 / (2.0 * PI); }   // Size_mod.F90:278-281

void zone_params(umt_ctx *ctx, ZoneParams &Z, double *P) {
  GtaState &g = ctx->gta;
  Z.nz = ctx->nz; Z.nc = ctx->nc; Z.mC = ctx->maxCorner; Z.nAng = g.nAng;
  Z.numCorner = ctx->d_numCorner; Z.cOffSet = ctx->d_cOffSet; Z.nCFaces = ctx->d_nCFaces; Z.cEZ = ctx->d_cEZ;
  Z.Volume = ctx->d_Volume; Z.Afp = ctx->d_Afp; Z.Aez = ctx->d_Aez; Z.omega = g.d_omega; Z.weight = g.d_weight;
  Z.sigTotal = g.d_sigTotal; Z.sigScat = g.d_sigScat; Z.greySource = g.d_greySource; Z.phiInc = g.d_phiInc;
  Z.TT = g.d_TT; Z.P = P;
  Z.wtiso = gta_wtiso(ctx);
}

// GTASweep (GTA%ID = 1): d_P (nc) in; d_PsiB (nAng, nb) in/out; leaves PhiInc in g.d_phiInc
int gta_device_sweep(umt_ctx *ctx, const double *d_P, double *d_PsiB) {
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nb = ctx->nb, rows = nc + nb;
  TRY(umt_gta_exchange(ctx, d_PsiB));   // SendFlux / RecvFlux of every angle (GTASweep.F90:139-146): lagged one grey sweep
  gta_tsa_kernel<<<nblk(nc), 256, 0, ctx->stream>>>(d_P, g.d_sigScat, g.d_greySource, gta_wtiso(ctx), g.d_tsaSource, nc);
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_tpsi, 0, sizeof(double) * (size_t)rows * g.nAng, ctx->stream));
  if (nb > 0)
    UMT_CUDA(ctx, cudaMemcpy2DAsync(g.d_tpsi + nc, sizeof(double) * rows, d_PsiB, sizeof(double) * nb, sizeof(double) * nb, g.nAng,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_counters, 0, sizeof(int) * (1 + g.nCounters), ctx->stream));
  if (ctx->ndim == 2) {
    TRY(umt_gta_launch_sweep_rz(ctx));   // gta_rz.cu
    gta_phiinc_kernel<<<nblk(nc), 256, 0, ctx->stream>>>(g.d_pinc, g.d_weight, g.nAng, nc, g.d_phiInc);
    if (nb > 0)
      UMT_CUDA(ctx, cudaMemcpy2DAsync(d_PsiB, sizeof(double) * nb, g.d_tpsi + nc, sizeof(double) * rows, sizeof(double) * nb, g.nAng,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return UMT_OK;
  }
  GtaSweepParams P;
  P.nc = nc; P.nb = nb; P.nz = ctx->nz; P.nItems = g.nItems;
  P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.nCFaces = ctx->d_nCFaces; P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
  P.Volume = ctx->d_Volume; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez; P.omega = g.d_omega;
  P.nextZ = g.d_nextZ; P.nextC = g.d_nextC; P.items = g.d_items; P.counters = g.d_counters;
  P.sigTotal = g.d_sigTotal; P.sigtInv = g.d_sigtInv; P.tsa = g.d_tsaSource; P.tpsi = g.d_tpsi; P.pinc = g.d_pinc;
  if (g.flow3d && g.nStagesR <= 1) {
    P.abortFlag = ctx->d_abort; P.spinLimit = ctx->spinLimit; P.pollMask = g.d_pollMask;
    gta_mark_kernel<<<dim3(std::max(1, std::min(2 * ctx->sm_count, (nc + 255) / 256)), g.nAng), 256, 0, ctx->stream>>>(g.d_tpsi, nc, rows);
    int occF = 0;
    UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occF, gta_sweep_flow_kernel, GTA_WARPS * 32, 0));
    const int gridF = std::max(1, std::min(ctx->sm_count * std::max(occF, 1), g.nItems));
    gta_sweep_flow_kernel<<<gridF, GTA_WARPS * 32, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
    UMT_CUDA(ctx, cudaMemcpyAsync(ctx->h_abort, ctx->d_abort, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));   // checked after the next sync
    gta_phiinc_kernel<<<nblk(nc), 256, 0, ctx->stream>>>(g.d_pinc, g.d_weight, g.nAng, nc, g.d_phiInc);
    if (nb > 0)
      UMT_CUDA(ctx, cudaMemcpy2DAsync(d_PsiB, sizeof(double) * nb, g.d_tpsi + nc, sizeof(double) * rows, sizeof(double) * nb, g.nAng,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    return UMT_OK;
  }
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gta_sweep_kernel, GTA_WARPS * 32, 0));
  for (int sR = 0; sR < g.nStagesR; sR++) {
    const int ib = g.stageItemBegin[sR], ie = g.stageItemBegin[sR + 1];
    const int ob = g.reflOpBegin[sR], oe = g.reflOpBegin[sR + 1];
    if (oe > ob) {   // PsiB(b, Minc) <- PsiB(b, Mref) on the reflecting boundaries (snreflect.F90:60-76)
      int maxN = 1;
      for (const auto &R : ctx->refl) maxN = std::max(maxN, R.n);
      gta_reflect_kernel<<<dim3((maxN + 255) / 256, oe - ob), 256, 0, ctx->stream>>>(g.d_tpsi, g.d_reflOps + ob, rows, nc);
    }
    if (ie == ib) continue;
    if (sR > 0) UMT_CUDA(ctx, cudaMemsetAsync(g.d_counters, 0, sizeof(int), ctx->stream));   // the ticket; plane counters persist
    P.items = g.d_items + ib; P.nItems = ie - ib;
    const int grid = std::max(1, std::min(ctx->sm_count * std::max(occ, 1), P.nItems));
    gta_sweep_kernel<<<grid, GTA_WARPS * 32, 0, ctx->stream>>>(P);
    UMT_CUDA(ctx, cudaGetLastError());
  }
  gta_phiinc_kernel<<<nblk(nc), 256, 0, ctx->stream>>>(g.d_pinc, g.d_weight, g.nAng, nc, g.d_phiInc);
  if (nb > 0)
    UMT_CUDA(ctx, cudaMemcpy2DAsync(d_PsiB, sizeof(double) * nb, g.d_tpsi + nc, sizeof(double) * rows, sizeof(double) * nb, g.nAng,
                                    cudaMemcpyDeviceToDevice, ctx->stream));
  return UMT_OK;
}

// GreySweepNEW: sweep, then the per-zone solves; d_P in/out
int gta_grey_sweep(umt_ctx *ctx, double *d_P, double *d_PsiB, int withSource) {
  TRY(gta_device_sweep(ctx, d_P, d_PsiB));
  ZoneParams Z;
  zone_params(ctx, Z, d_P);
  if (withSource) {
    if (ctx->gta.tt_decomposed) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA transfer matrices already decomposed: call umt_gta_init_tt before another withSource sweep");
    ctx->gta.tt_decomposed = true;
  }
  if (!withSource && ctx->gta.hexTT) gta_scalar_solve_hex_kernel<<<nblk((size_t)ctx->nz * 8, 256), 256, 0, ctx->stream>>>(Z);
  else gta_scalar_kernel<<<nblk(ctx->nz, 64), 64, 0, ctx->stream>>>(Z, withSource);
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}

int device_dot(umt_ctx *ctx, const double *x, const double *y, double *result) {
  GtaState &g = ctx->gta;
  dot_partial_kernel<<<RED_BLOCKS, 256, 0, ctx->stream>>>(x, y, g.d_sigScatVol, ctx->nc, g.d_red);
  sum_partials_kernel<<<1, 256, 0, ctx->stream>>>(g.d_red, RED_BLOCKS, g.d_red + 3 * RED_BLOCKS);
  TRY(umt_allreduce_f64(ctx, g.d_red + 3 * RED_BLOCKS, 1, 0));   // MPIAllReduce(sum) of scat_prod / scat_prod1
  UMT_CUDA(ctx, cudaMemcpyAsync(result, g.d_red + 3 * RED_BLOCKS, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

}  // namespace

void umt_gta_release(umt_ctx *ctx) {
  GtaState &g = ctx->gta;
  void *p[] = {g.d_omega, g.d_weight, g.d_nextZ, g.d_nextC, g.d_items, g.d_counters, g.d_sigTotal, g.d_sigtInv, g.d_sigScat, g.d_sigScatVol,
               g.d_greySource, g.d_tsaSource, g.d_phiInc, g.d_correction, g.d_chi, g.d_TT, g.d_tpsi, g.d_pinc, g.d_vec[0], g.d_vec[1], g.d_vec[2],
               g.d_vec[3], g.d_vecB[0], g.d_vecB[1], g.d_vecB[2], g.d_vecB[3], g.d_radEnergy, g.d_pzOld, g.d_volZone, g.d_red, g.d_P, g.d_PB,
               g.d_start, g.d_finish, g.d_level, g.d_fac, g.d_w1, g.d_w2, g.d_psim, g.d_tinc, g.d_levelAngles, g.d_planeOff, g.d_nHyp, g.d_reflOps,
               g.d_prevAngle, g.d_psimA, g.d_tincA, g.d_pollMask};
  for (void *q : p) if (q) cudaFree(q);
  g = GtaState();
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
// GTA angle set (level-symmetric S2, rt/quadxyz.F90 + rtquad.F90:95-105), its sweep order, work items, device arrays
extern "C" int umt_gta_setup(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->ndim != 3 && ctx->ndim != 2) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_setup: 2-D (r-z) or 3-D meshes");
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_setup: host-only context");
  if (!ctx->have_conn || !ctx->have_geom || ctx->h_zoneOpp.empty()) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_setup: needs full connectivity and geometry");
  if (ctx->ndim == 3 && (ctx->maxCorner > MAXC || ctx->maxcf != MAXCF)) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_gta_setup: maxCorner <= 8, maxcf == 3");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nd = ctx->ndim;
  std::vector<WorkItem> items;
  const int nz = ctx->nz, nc = ctx->nc, nb = ctx->nb;
  if (nd == 2) {
    TRY(umt_gta_setup_rz(ctx));
  } else {
  g.nAng = 8;
  const double mu = 0.577350269189625;   // QuadratureData_mod.F90:711-713
  const int sx[8] = {1, -1, -1, 1, 1, -1, -1, 1}, sy[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sz[8] = {1, 1, 1, 1, -1, -1, -1, -1};
  g.omega.resize(24); g.weight.resize(8);
  double sum = 0.0;
  for (int a = 0; a < 8; a++) {
    g.omega[3 * a] = sx[a] * mu; g.omega[3 * a + 1] = sy[a] * mu; g.omega[3 * a + 2] = sz[a] * mu;
    g.weight[a] = 0.5 * PI * 1.0;
    sum += g.weight[a];
  }
  const double fac = 1.0 / ((1.0 / (4.0 * PI)) * sum);
  for (double &w : g.weight) w = fac * w;
  TRY(umt_host_build_order(ctx, g.omega.data(), g.nAng, g.nHyp, g.zonesInPlane, g.nextZ, g.nextC));
  }
  g.maxHyp = *std::max_element(g.nHyp.begin(), g.nHyp.end());
  std::vector<int> h_nextZ((size_t)g.nAng * nz);
  std::vector<unsigned char> h_nextC((size_t)g.nAng * nc);
  const int zpi = GTA_WARPS;   // one zone per warp of the sweep CTA
  std::vector<std::vector<int>> start(g.nAng), nIt(g.nAng);
  for (int a = 0; a < g.nAng; a++) {
    std::copy(g.nextZ[a].begin(), g.nextZ[a].end(), h_nextZ.begin() + (size_t)a * nz);
    for (int i = 0; i < nc; i++) h_nextC[(size_t)a * nc + i] = (unsigned char)(g.nextC[a][i] - 1);
    start[a].assign(g.nHyp[a] + 1, 0); nIt[a].assign(g.nHyp[a], 0);
    for (int p = 0; p < g.nHyp[a]; p++) { start[a][p + 1] = start[a][p] + g.zonesInPlane[a][p]; nIt[a][p] = (g.zonesInPlane[a][p] + zpi - 1) / zpi; }
  }
  if (nd == 2) TRY(umt_gta_finish_setup_rz(ctx, items));   // items chained along the xi-levels, r-z coefficient arrays
  // reflecting boundaries (GTASweep.F90:151 snreflect): angles in stages, mirror images first, one launch per stage
  g.nStagesR = 1; g.stageOf.assign(g.nAng, 0);
  TRY(umt_reflect_analyze(ctx, g.omega.data(), g.nAng, g.mref, g.stageOf));
  if (nd == 2 && !ctx->refl.empty()) {
    // r-z: levels ordered by the mirror dependencies between levels, every angle of a level its own step (as reflect.cu does for
    // the multigroup sweep); finishing directions are not swept and get no copy
    const int nL = std::max(g.nLevels, 1);
    std::vector<std::vector<int>> ldeps(nL);
    for (const auto &mr : g.mref)
      for (int a = 0; a < g.nAng; a++)
        if (mr[a] >= 0 && g.level[mr[a]] != g.level[a]) ldeps[g.level[a]].push_back(g.level[mr[a]]);
    std::vector<int> lstage;
    umt_level_stages(nL, ldeps, lstage);
    int maxPos = 1;
    std::vector<int> pos(g.nAng, 0), cnt(nL, 0);
    for (int a = 0; a < g.nAng; a++) { pos[a] = cnt[g.level[a]]++; maxPos = std::max(maxPos, cnt[g.level[a]]); }
    for (int a = 0; a < g.nAng; a++) g.stageOf[a] = lstage[g.level[a]] * maxPos + pos[a];
    for (auto &mr : g.mref) for (int a = 0; a < g.nAng; a++) if (g.finish[a]) mr[a] = -1;
  }
  g.nStagesR = 1 + *std::max_element(g.stageOf.begin(), g.stageOf.end());
  if (nd == 2 && g.nStagesR > 1) {   // items of a stage contiguous, their relative order kept
    std::stable_sort(items.begin(), items.end(), [&](const WorkItem &x, const WorkItem &y) { return g.stageOf[x.angle] < g.stageOf[y.angle]; });
    g.rz_chain = false;              // the chain kernel owns whole levels; stages cut across them
  }
  g.stageItemBegin.assign(g.nStagesR + 1, 0);
  for (int sR = 0; nd == 3 && sR < g.nStagesR; sR++) {
  for (int p = 0; p < g.maxHyp; p++)
    for (int a = 0; a < g.nAng; a++) {
      if (g.stageOf[a] != sR) continue;
      if (p >= g.nHyp[a]) continue;
      for (int k = 0; k < nIt[a][p]; k++) {
        WorkItem w;
        w.angle = a; w.zbeg = start[a][p] + k * zpi; w.zend = std::min(start[a][p + 1], w.zbeg + zpi);
        w.wait_idx = p > 0 ? a * g.maxHyp + p - 1 : -1; w.wait_count = p > 0 ? nIt[a][p - 1] : 0;
        w.signal_idx = a * g.maxHyp + p; w.pad0 = -1; w.pad1 = 0;
        items.push_back(w);
      }
    }
    g.stageItemBegin[sR + 1] = (int)items.size();
  }
  if (nd == 2) {
    size_t i = 0;
    for (int sR = 0; sR < g.nStagesR; sR++) {
      while (i < items.size() && g.stageOf[items[i].angle] <= sR) i++;
      g.stageItemBegin[sR + 1] = (int)i;
    }
  }
  {
    std::vector<int4> ops;
    g.reflOpBegin.assign(g.nStagesR + 1, 0);
    for (int sR = 0; sR < g.nStagesR; sR++) {
      for (size_t k = 0; k < g.mref.size(); k++)
        for (int a = 0; a < g.nAng; a++)
          if (g.stageOf[a] == sR && g.mref[k][a] >= 0) ops.push_back(make_int4(a, g.mref[k][a], ctx->refl[k].first, ctx->refl[k].n));
      g.reflOpBegin[sR + 1] = (int)ops.size();
    }
    TRY(dalloc(ctx, &g.d_reflOps, ops.size()));
    if (!ops.empty()) UMT_CUDA(ctx, umt_memcpy(ctx, g.d_reflOps, ops.data(), sizeof(int4) * ops.size(), cudaMemcpyHostToDevice));
  }
  g.nItems = (int)items.size(); g.nCounters = g.nAng * g.maxHyp;
  // 3-D dataflow kernel (gta_sweep_flow_kernel): every zone must take the warp solve (<= 8 corners of three faces each, no
  // intra-zone cycle) and the angles must go in one launch (no reflecting boundaries).  UMT_GTA_KERNEL=item keeps the counters.
  g.flow3d = nd == 3 && g.nStagesR <= 1 && ctx->maxCorner <= MAXC;
  for (int c = 0; c < nc && g.flow3d; c++) g.flow3d = ctx->h_nCFaces[c] == 3;
  for (int a = 0; a < g.nAng && g.flow3d; a++)
    for (int z : g.nextZ[a]) if (z < 0) { g.flow3d = false; break; }
  if (const char *e = getenv("UMT_GTA_KERNEL")) if (std::string(e) == "item") g.flow3d = false;
  // InitGreySweep with 64 threads per zone: hexahedra only (8 corners of 3 faces), the 8 S2 ordinates
  g.hexTT = nd == 3 && g.nAng == 8 && ctx->maxCorner == 8;
  for (int z = 0; z < ctx->nz && g.hexTT; z++) g.hexTT = ctx->h_numCorner[z] == 8;
  for (int c = 0; c < nc && g.hexTT; c++) g.hexTT = ctx->h_nCFaces[c] == 3;
  if (const char *e = getenv("UMT_GTA_TT")) if (std::string(e) == "zone") g.hexTT = false;
  if (g.flow3d) {
    std::vector<int> zoneOf(nc), planeOf(nz);
    for (int z = 0; z < nz; z++) for (int c = 0; c < ctx->h_numCorner[z]; c++) zoneOf[ctx->h_cOffSet[z] + c] = z;
    std::vector<unsigned> mask((size_t)g.nAng * nz, 0u);
    for (int a = 0; a < g.nAng; a++) {
      for (int p = 0; p < g.nHyp[a]; p++) for (int i = start[a][p]; i < start[a][p + 1]; i++) planeOf[std::abs(g.nextZ[a][i]) - 1] = p;
      for (int i = 0; i < nz; i++) {
        const int z = std::abs(g.nextZ[a][i]) - 1, c0 = ctx->h_cOffSet[z];
        unsigned m = 0u;
        for (int c = 0; c < ctx->h_numCorner[z]; c++)
          for (int f = 0; f < 3; f++) {
            const int row = ctx->h_cFP[(size_t)(c0 + c) * 3 + f] - 1;
            if (row < nc && planeOf[zoneOf[row]] < planeOf[z]) m |= 1u << (3 * c + f);
          }
        mask[(size_t)a * nz + i] = m;
      }
    }
    TRY(dalloc(ctx, &g.d_pollMask, mask.size()));
    UMT_CUDA(ctx, umt_memcpy(ctx, g.d_pollMask, mask.data(), sizeof(unsigned) * mask.size(), cudaMemcpyHostToDevice));
  }
  TRY(dalloc(ctx, &g.d_omega, (size_t)nd * g.nAng)); TRY(dalloc(ctx, &g.d_weight, g.nAng));
  TRY(dalloc(ctx, &g.d_nextZ, h_nextZ.size())); TRY(dalloc(ctx, &g.d_nextC, h_nextC.size()));
  TRY(dalloc(ctx, &g.d_items, items.size())); TRY(dalloc(ctx, &g.d_counters, 1 + (size_t)g.nCounters));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_omega, g.omega.data(), sizeof(double) * nd * g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_weight, g.weight.data(), sizeof(double) * g.nAng, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_nextZ, h_nextZ.data(), sizeof(int) * h_nextZ.size(), cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_nextC, h_nextC.data(), h_nextC.size(), cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_items, items.data(), sizeof(WorkItem) * items.size(), cudaMemcpyHostToDevice));
  double **percorner[] = {&g.d_sigTotal, &g.d_sigtInv, &g.d_sigScat, &g.d_sigScatVol, &g.d_greySource, &g.d_tsaSource, &g.d_phiInc, &g.d_correction,
                          &g.d_vec[0], &g.d_vec[1], &g.d_vec[2], &g.d_vec[3], &g.d_P};
  for (double **p : percorner) TRY(dalloc(ctx, p, nc));
  double **perbdy[] = {&g.d_vecB[0], &g.d_vecB[1], &g.d_vecB[2], &g.d_vecB[3], &g.d_PB};
  for (double **p : perbdy) TRY(dalloc(ctx, p, (size_t)g.nAng * nb));
  TRY(dalloc(ctx, &g.d_chi, (size_t)nc * ctx->G));
  TRY(dalloc(ctx, &g.d_TT, (size_t)nc * ctx->maxCorner));
  TRY(dalloc(ctx, &g.d_tpsi, (size_t)g.nAng * (nc + nb)));
  TRY(dalloc(ctx, &g.d_pinc, (size_t)g.nAng * nc));
  TRY(dalloc(ctx, &g.d_radEnergy, nz)); TRY(dalloc(ctx, &g.d_pzOld, nz)); TRY(dalloc(ctx, &g.d_volZone, nz));
  TRY(dalloc(ctx, &g.d_red, 3 * RED_BLOCKS + 8));
  g.ready = true; g.have_opacity = false; g.tt_decomposed = false;
  TRY(umt_gta_build_exchange(ctx));   // decomposed mesh: collective over the domains
  return UMT_OK;
}

extern "C" int umt_gta_get_quadrature(umt_ctx *ctx, double *omega, double *weight) {
  if (!ctx || !ctx->gta.ready) return UMT_ERR_STATE;
  if (omega) std::copy(ctx->gta.omega.begin(), ctx->gta.omega.end(), omega);
  if (weight) std::copy(ctx->gta.weight.begin(), ctx->gta.weight.end(), weight);
  return UMT_OK;
}

// GTA%GreySigTotal, GreySigScat, GreySigScatVol handed over by the caller (GreySigtInv = 1 / GreySigTotal)
extern "C" int umt_gta_set_opacity(umt_ctx *ctx, const double *GreySigTotal, const double *GreySigScat, const double *GreySigScatVol) {
  if (!ctx || !GreySigTotal || !GreySigScat || !GreySigScatVol) return UMT_ERR_ARG;
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc;
  std::vector<double> inv(nc);
  for (int c = 0; c < nc; c++) inv[c] = 1.0 / GreySigTotal[c];
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_sigTotal, GreySigTotal, sizeof(double) * nc, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_sigScat, GreySigScat, sizeof(double) * nc, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_sigScatVol, GreySigScatVol, sizeof(double) * nc, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_sigtInv, inv.data(), sizeof(double) * nc, cudaMemcpyHostToDevice));
  g.have_opacity = true;
  return UMT_OK;
}

static int upload_c2z(umt_ctx *ctx, int **d_c2z) {
  std::vector<int> c2z(ctx->nc);
  for (int z = 0; z < ctx->nz; z++)
    for (int c = 0; c < ctx->h_numCorner[z]; c++) c2z[ctx->h_cOffSet[z] + c] = z;
  UMT_CUDA(ctx, cudaMalloc((void **)d_c2z, sizeof(int) * ctx->nc));
  UMT_CUDA(ctx, umt_memcpy(ctx, *d_c2z, c2z.data(), sizeof(int) * ctx->nc, cudaMemcpyHostToDevice));
  return UMT_OK;
}

// setGTAOpacityNEW on the device from Mat%Siga, Mat%Sigs (ngr,nz), Mat%Eta (nc), GTA%Chi (ngr,nc); Chi comes back rescaled
extern "C" int umt_gta_compute_opacity(umt_ctx *ctx, const double *Siga, const double *Sigs, const double *Eta, double *Chi) {
  if (!ctx || !Siga || !Sigs || !Eta || !Chi) return UMT_ERR_ARG;
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nz = ctx->nz, G = ctx->G;
  double *d_a = nullptr, *d_s = nullptr, *d_e = nullptr;
  int *d_c2z = nullptr;
  int rc = upload_c2z(ctx, &d_c2z);
  if (!rc) rc = dalloc(ctx, &d_a, (size_t)nz * G, false);
  if (!rc) rc = dalloc(ctx, &d_s, (size_t)nz * G, false);
  if (!rc) rc = dalloc(ctx, &d_e, nc, false);
  if (!rc) {
    umt_memcpy(ctx, d_a, Siga, sizeof(double) * nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_s, Sigs, sizeof(double) * nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_e, Eta, sizeof(double) * nc, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, g.d_chi, Chi, sizeof(double) * (size_t)nc * G, cudaMemcpyHostToDevice);
    gta_opacity_kernel<<<nblk(nc, 128), 128, 0, ctx->stream>>>(nc, G, ctx->tau, d_c2z, d_a, d_s, d_e, ctx->d_Volume, g.d_chi, g.d_sigTotal,
                                                             g.d_sigScat, g.d_sigScatVol, g.d_sigtInv);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = umt_memcpy(ctx, Chi, g.d_chi, sizeof(double) * (size_t)nc * G, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { ctx->err = std::string("umt_gta_compute_opacity: ") + cudaGetErrorString(e); rc = UMT_ERR_CUDA; }
  }
  cudaFree(d_a); cudaFree(d_s); cudaFree(d_e); cudaFree(d_c2z);
  if (!rc) g.have_opacity = true;
  return rc;
}

extern "C" int umt_gta_get_opacity(umt_ctx *ctx, double *GreySigTotal, double *GreySigScat, double *GreySigScatVol, double *GreySigtInv) {
  if (!ctx) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  GtaState &g = ctx->gta;
  const size_t n = sizeof(double) * ctx->nc;
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (GreySigTotal) UMT_CUDA(ctx, umt_memcpy(ctx, GreySigTotal, g.d_sigTotal, n, cudaMemcpyDeviceToHost));
  if (GreySigScat) UMT_CUDA(ctx, umt_memcpy(ctx, GreySigScat, g.d_sigScat, n, cudaMemcpyDeviceToHost));
  if (GreySigScatVol) UMT_CUDA(ctx, umt_memcpy(ctx, GreySigScatVol, g.d_sigScatVol, n, cudaMemcpyDeviceToHost));
  if (GreySigtInv) UMT_CUDA(ctx, umt_memcpy(ctx, GreySigtInv, g.d_sigtInv, n, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

// getCollisionRate on the device-resident PhiTotal: GreySource (nc) stays on the device and is also returned if asked
extern "C" int umt_collision_rate(umt_ctx *ctx, const double *Eta, const double *Siga, const double *Sigs, int residualFlag, double *GreySource) {
  if (!ctx || !Eta || !Siga || !Sigs) return UMT_ERR_ARG;
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  if (!ctx->d_phi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_collision_rate: no PhiTotal on the device");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nz = ctx->nz, G = ctx->G;
  double *d_a = nullptr, *d_s = nullptr, *d_e = nullptr;
  int *d_c2z = nullptr;
  int rc = upload_c2z(ctx, &d_c2z);
  if (!rc) rc = dalloc(ctx, &d_a, (size_t)nz * G, false);
  if (!rc) rc = dalloc(ctx, &d_s, (size_t)nz * G, false);
  if (!rc) rc = dalloc(ctx, &d_e, nc, false);
  if (!rc) {
    umt_memcpy(ctx, d_a, Siga, sizeof(double) * nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_s, Sigs, sizeof(double) * nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_e, Eta, sizeof(double) * nc, cudaMemcpyHostToDevice);
    collision_rate_kernel<<<nblk(nc, 128), 128, 0, ctx->stream>>>(nc, G, d_c2z, d_e, d_a, d_s, ctx->d_phi, g.d_greySource, residualFlag);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && GreySource) e = umt_memcpy(ctx, GreySource, g.d_greySource, sizeof(double) * nc, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { ctx->err = std::string("umt_collision_rate: ") + cudaGetErrorString(e); rc = UMT_ERR_CUDA; }
  }
  cudaFree(d_a); cudaFree(d_s); cudaFree(d_e); cudaFree(d_c2z);
  return rc;
}

extern "C" int umt_gta_set_source(umt_ctx *ctx, const double *GreySource) {
  if (!ctx || !GreySource) return UMT_ERR_ARG;
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  UMT_CUDA(ctx, umt_memcpy(ctx, ctx->gta.d_greySource, GreySource, sizeof(double) * ctx->nc, cudaMemcpyHostToDevice));
  return UMT_OK;
}

extern "C" int umt_gta_init_tt(umt_ctx *ctx, double *TT /* (maxCorner, nc) or NULL */) {
  if (!ctx) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->ndim == 2) {
    TRY(umt_gta_launch_init_tt_rz(ctx));
  } else {
    ZoneParams Z;
    zone_params(ctx, Z, ctx->gta.d_P);
    if (ctx->gta.hexTT) gta_init_tt_hex_kernel<<<(ctx->nz + TTH_ZONES - 1) / TTH_ZONES, TTH_ZONES * 64, 0, ctx->stream>>>(Z);
    else gta_init_tt_kernel<<<nblk(ctx->nz, 64), 64, 0, ctx->stream>>>(Z);
    UMT_CUDA(ctx, cudaGetLastError());
  }
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->gta.tt_decomposed = false;
  if (TT) UMT_CUDA(ctx, umt_memcpy(ctx, TT, ctx->gta.d_TT, sizeof(double) * (size_t)ctx->nc * ctx->maxCorner, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

// GTASweep (snac/GTASweep.F90, GTA%ID = 1) for all 8 angles: TsaSource = wtiso (GreySigScat P + GreySource); GreySource NULL keeps
// the device copy, withSource = 0 sweeps with GreySource = 0.  PsiB_gta (nbelem, 8) in/out, PhiInc (nc) out.
extern "C" int umt_gta_sweep(umt_ctx *ctx, const double *P, const double *GreySource, double *PsiB_gta, double *PhiInc, int withSource) {
  if (!ctx || !P) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nb = ctx->nb;
  if (GreySource) UMT_CUDA(ctx, umt_memcpy(ctx, g.d_greySource, GreySource, sizeof(double) * nc, cudaMemcpyHostToDevice));
  if (!withSource) UMT_CUDA(ctx, cudaMemsetAsync(g.d_greySource, 0, sizeof(double) * nc, ctx->stream));
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_P, P, sizeof(double) * nc, cudaMemcpyHostToDevice));
  if (nb > 0) {
    if (PsiB_gta) UMT_CUDA(ctx, umt_memcpy(ctx, g.d_PB, PsiB_gta, sizeof(double) * (size_t)nb * g.nAng, cudaMemcpyHostToDevice));
    else UMT_CUDA(ctx, cudaMemsetAsync(g.d_PB, 0, sizeof(double) * (size_t)nb * g.nAng, ctx->stream));
  }
  TRY(gta_device_sweep(ctx, g.d_P, g.d_PB));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  TRY(umt_check_abort(ctx, "umt_gta_sweep"));
  if (PhiInc) UMT_CUDA(ctx, umt_memcpy(ctx, PhiInc, g.d_phiInc, sizeof(double) * nc, cudaMemcpyDeviceToHost));
  if (PsiB_gta && nb > 0) UMT_CUDA(ctx, umt_memcpy(ctx, PsiB_gta, g.d_PB, sizeof(double) * (size_t)nb * g.nAng, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

// GreySweepNEW (rt/GreySweep.F90:12-48): P (nc) and PsiB_gta (nbelem, 8) in/out
extern "C" int umt_gta_grey_sweep(umt_ctx *ctx, double *P, double *PsiB_gta, int withSource) {
  if (!ctx || !P) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nb = ctx->nb;
  UMT_CUDA(ctx, umt_memcpy(ctx, g.d_P, P, sizeof(double) * nc, cudaMemcpyHostToDevice));
  if (nb > 0) {
    if (PsiB_gta) UMT_CUDA(ctx, umt_memcpy(ctx, g.d_PB, PsiB_gta, sizeof(double) * (size_t)nb * g.nAng, cudaMemcpyHostToDevice));
    else UMT_CUDA(ctx, cudaMemsetAsync(g.d_PB, 0, sizeof(double) * (size_t)nb * g.nAng, ctx->stream));
  }
  TRY(gta_grey_sweep(ctx, g.d_P, g.d_PB, withSource));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  TRY(umt_check_abort(ctx, "umt_gta_grey_sweep"));
  UMT_CUDA(ctx, umt_memcpy(ctx, P, g.d_P, sizeof(double) * nc, cudaMemcpyDeviceToHost));
  if (PsiB_gta && nb > 0) UMT_CUDA(ctx, umt_memcpy(ctx, PsiB_gta, g.d_PB, sizeof(double) * (size_t)nb * g.nAng, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

// GTASolver (rt/GTASolver.F90:42-425): BiCGSTAB on the grey corrections with the device-resident PhiTotal
// and GreySource (umt_collision_rate / umt_gta_set_source).  GreyCorrection stays on the device (umt_gta_get_correction,
// umt_add_grey_corrections).
extern "C" int umt_gta_solve(umt_ctx *ctx, double epsPoint, int maxIters, double epsGrey, int enforceHardMax, int *nGreyIterOut, double *maxRelErrOut) {
  if (!ctx) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  if (!ctx->d_phi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_solve: no PhiTotal on the device");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  GtaState &g = ctx->gta;
  const int nc = ctx->nc, nz = ctx->nz;
  const size_t nB = (size_t)ctx->nb * g.nAng;
  const double adqtSmall = 1.e-150;
  cudaStream_t st = ctx->stream;
  double *R = g.d_vec[0], *D = g.d_vec[1], *A = g.d_vec[2], *AS = g.d_vec[3];
  double *RB = g.d_vecB[0], *DB = g.d_vecB[1], *AB = g.d_vecB[2], *ASB = g.d_vecB[3];
  rad_energy_kernel<<<nblk(nz, 128), 128, 0, st>>>(nz, ctx->G, ctx->d_numCorner, ctx->d_cOffSet, ctx->d_Volume, ctx->d_phi, g.d_volZone, g.d_radEnergy);
  TRY(umt_gta_init_tt(ctx, nullptr));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_correction, 0, sizeof(double) * nc, st));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_pzOld, 0, sizeof(double) * nz, st));
  UMT_CUDA(ctx, cudaMemsetAsync(R, 0, sizeof(double) * nc, st));
  if (nB) UMT_CUDA(ctx, cudaMemsetAsync(RB, 0, sizeof(double) * nB, st));
  int nGreyIter = 1;
  TRY(gta_grey_sweep(ctx, R, RB, 1));
  UMT_CUDA(ctx, cudaMemcpyAsync(D, R, sizeof(double) * nc, cudaMemcpyDeviceToDevice, st));
  if (nB) UMT_CUDA(ctx, cudaMemcpyAsync(DB, RB, sizeof(double) * nB, cudaMemcpyDeviceToDevice, st));
  double rrOld = 0.0, maxRelErrGrey = 0.0;
  TRY(device_dot(ctx, R, nullptr, &rrOld));
  UMT_CUDA(ctx, cudaMemsetAsync(g.d_greySource, 0, sizeof(double) * nc, st));
  for (;;) {
    if (std::fabs(rrOld) < adqtSmall) {
      if (nGreyIter <= 2) UMT_CUDA(ctx, cudaMemcpyAsync(g.d_correction, R, sizeof(double) * nc, cudaMemcpyDeviceToDevice, st));
      break;
    }
    nGreyIter += 2;
    UMT_CUDA(ctx, cudaMemcpyAsync(A, D, sizeof(double) * nc, cudaMemcpyDeviceToDevice, st));
    if (nB) UMT_CUDA(ctx, cudaMemcpyAsync(AB, DB, sizeof(double) * nB, cudaMemcpyDeviceToDevice, st));
    TRY(gta_grey_sweep(ctx, A, AB, 0));
    k_sub<<<nblk(nc), 256, 0, st>>>(A, D, A, nc);
    if (nB) k_sub<<<nblk(nB), 256, 0, st>>>(AB, DB, AB, nB);
    double dAd = 0.0;
    TRY(device_dot(ctx, A, nullptr, &dAd));
    if (std::fabs(dAd) < adqtSmall) break;
    const double alpha = rrOld / dAd;
    k_axmy<<<nblk(nc), 256, 0, st>>>(R, alpha, A, nc);
    if (nB) k_axmy<<<nblk(nB), 256, 0, st>>>(RB, alpha, AB, nB);
    UMT_CUDA(ctx, cudaMemcpyAsync(AS, R, sizeof(double) * nc, cudaMemcpyDeviceToDevice, st));
    if (nB) UMT_CUDA(ctx, cudaMemcpyAsync(ASB, RB, sizeof(double) * nB, cudaMemcpyDeviceToDevice, st));
    TRY(gta_grey_sweep(ctx, AS, ASB, 0));
    k_sub<<<nblk(nc), 256, 0, st>>>(AS, R, AS, nc);
    if (nB) k_sub<<<nblk(nB), 256, 0, st>>>(ASB, RB, ASB, nB);
    double omegaNum = 0.0, omegaDen = 0.0;
    TRY(device_dot(ctx, AS, R, &omegaNum));
    TRY(device_dot(ctx, AS, AS, &omegaDen));
    if (std::fabs(omegaDen) < adqtSmall || std::fabs(omegaNum) < adqtSmall) {
      k_corr<<<nblk(nc), 256, 0, st>>>(g.d_correction, alpha, D, 0.0, R, nc, 0);
      break;
    }
    const double omegaCG = omegaNum / omegaDen;
    k_corr<<<nblk(nc), 256, 0, st>>>(g.d_correction, alpha, D, omegaCG, R, nc, 1);
    k_axmy<<<nblk(nc), 256, 0, st>>>(R, omegaCG, AS, nc);
    if (nB) k_axmy<<<nblk(nB), 256, 0, st>>>(RB, omegaCG, ASB, nB);
    double rr = 0.0;
    TRY(device_dot(ctx, R, nullptr, &rr));
    const double beta = (rr * alpha) / (rrOld * omegaCG);
    k_dir<<<nblk(nc), 256, 0, st>>>(D, R, beta, omegaCG, A, nc);
    if (nB) k_dir<<<nblk(nB), 256, 0, st>>>(DB, RB, beta, omegaCG, AB, nB);
    zone_error_kernel<<<RED_BLOCKS, 256, 0, st>>>(nz, ctx->d_numCorner, ctx->d_cOffSet, ctx->d_Volume, g.d_volZone, g.d_correction, g.d_radEnergy,
                                                  g.d_pzOld, g.d_red);
    zone_error_finish_kernel<<<1, 256, 0, st>>>(g.d_red, RED_BLOCKS, g.d_red + 3 * RED_BLOCKS);
    double e3[3];
    UMT_CUDA(ctx, cudaMemcpyAsync(e3, g.d_red + 3 * RED_BLOCKS, sizeof(double) * 3, cudaMemcpyDeviceToHost, st));
    UMT_CUDA(ctx, cudaStreamSynchronize(st));
    TRY(umt_check_abort(ctx, "umt_gta_solve"));
    if (std::isinf(e3[2])) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_solve: grey solver encountered a NaN (iteration %d)", nGreyIter);
    const double relErrL2 = e3[1] != 0.0 ? std::sqrt(std::fabs(e3[0] / e3[1])) : 0.0;
    maxRelErrGrey = std::max(e3[2], relErrL2);
    if (ctx->nRanks > 1 && ctx->transport) {   // MPIAllReduce(maxRelErrGrey, "max") (GTASolver.F90:380-381)
      UMT_CUDA(ctx, cudaMemcpyAsync(g.d_red + 3 * RED_BLOCKS + 4, &maxRelErrGrey, sizeof(double), cudaMemcpyHostToDevice, st));
      TRY(umt_allreduce_f64(ctx, g.d_red + 3 * RED_BLOCKS + 4, 1, 1));
      UMT_CUDA(ctx, cudaMemcpyAsync(&maxRelErrGrey, g.d_red + 3 * RED_BLOCKS + 4, sizeof(double), cudaMemcpyDeviceToHost, st));
      UMT_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (enforceHardMax && nGreyIter >= maxIters) break;
    else if ((maxRelErrGrey < epsPoint || nGreyIter >= maxIters) && maxRelErrGrey < epsGrey) break;
    else if (nGreyIter >= 100 * maxIters) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_gta_solve: grey solver is not converging (%d iterations)", nGreyIter);
    rrOld = rr;
  }
  UMT_CUDA(ctx, cudaStreamSynchronize(st));
  TRY(umt_check_abort(ctx, "umt_gta_solve"));
  if (nGreyIterOut) *nGreyIterOut = nGreyIter;
  if (maxRelErrOut) *maxRelErrOut = maxRelErrGrey;
  return UMT_OK;
}

extern "C" int umt_gta_get_correction(umt_ctx *ctx, double *GreyCorrection) {
  if (!ctx || !GreyCorrection) return UMT_ERR_ARG;
  if (!ctx->gta.ready) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA not set up (umt_gta_setup)");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  UMT_CUDA(ctx, umt_memcpy(ctx, GreyCorrection, ctx->gta.d_correction, sizeof(double) * ctx->nc, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

// addGreyCorrections.F90:70-91: PhiTotal += GreyCorrection * Chi on the device (Chi from umt_gta_compute_opacity)
extern "C" int umt_add_grey_corrections(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  TRY(need_gta(ctx));
  if (!ctx->d_phi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_add_grey_corrections: no PhiTotal on the device");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->nc * ctx->G;
  add_corrections_kernel<<<nblk(n), 256, 0, ctx->stream>>>(n, ctx->G, ctx->gta.d_correction, ctx->gta.d_chi, ctx->d_phi);
  UMT_CUDA(ctx, cudaGetLastError());
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}
