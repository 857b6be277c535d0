// 3-D upstream-corner-balance sweep for sm_100a.
//
// Replaces snac/SweepUCBxyz.F90:11-322 (one angle, one set) *and* the angle loop
// of snac/SetSweep.F90:113-170: one persistent kernel sweeps every angle of the
// quadrature.  Work is a list of items (angle, hyperplane, chunk of zones); CTAs
// pull items through an atomic ticket and wait on a per-(angle,plane) completion
// counter, so the hyperplane dependency of one angle hides behind other angles.
// Group index is on consecutive lanes: every load/store of Psi, STotal, Psi1, PsiB
// is a contiguous G*8-byte row.
//
// Two kernels:
//  * sweep3d_plan_kernel (the hot one): zones whose corners all have three faces
//    (hexes, prisms, tets) run from a per-(zone,angle) "plan record" built once per
//    schedule (plan_build_kernel): corners relabelled into solve order, omega.A
//    products, upstream rows, the closure polynomials of SweepUCBxyz.F90:217-252 as
//    group-independent coefficients.  A producer warp streams the records of the
//    next items into shared memory with TMA bulk copies (cp.async.bulk + mbarrier),
//    prefetches the item's Psi/STotal/Sigt rows into L2 and resolves the plane
//    dependency, so the consumer warps never spin; consumers keep the whole 8-corner
//    zone solve in registers (static indexing in solve-order space).
//  * sweep3d_generic_kernel: any zone shape (nCFaces != 3, intra-zone cycles); also
//    the per-zone slow path of the plan kernel.
//
// Upstream Psi1 rows were written a few planes earlier by other CTAs and are read
// through L2 (ld.global.cg) — L1 is not coherent across SMs.
#include <algorithm>
#include <cstddef>
#include <cstdlib>

#include "umt_internal.h"

namespace {

constexpr int MAXC = 8;    // corners per zone handled by these kernels
constexpr int MAXCF = 3;   // corner faces
constexpr double FOURALPHA = 1.82;   // SweepUCBxyz.F90:80

// address of the Psi1 row at element offset o of an angle's slab: corner rows in the Psi1 workspace (upg), boundary-element rows
// (o >= ncG) in the buffer whose tails are Set%PsiB (pbg).
// TWO is a template parameter of the zone solves: false where both are the same buffer (legacy layout, in-place savePsi sweep).
#define PLAN_ROW(o) ((TWO && (o) >= ncG ? pbg : upg) + (o))
#define PLAN_BROW(o) ((TWO ? pbg : upg) + (o))

struct Sweep3DParams {
  int nc, nb, nz, G, NA, nItems;
  int wpe, nEngines, nStages, stageBytes, offSt, offSigt, offRecs, zpi;   // PlanGeom
  int nStagesWG;           // stages per half of the warp-group build
  int qbMax;               // tickets a loader takes at a time (at most)
  double tau;
  const int *numCorner, *cOffSet, *nCFaces, *cFP /* 0-based row; >= nc: boundary */, *cEZ /* 0-based */;
  const double *Volume, *Afp, *Aez, *omega;
  const int *nextZ;
  const unsigned char *nextC;
  const WorkItem *items;
  int *counters;
  const double *psi, *stotal, *sigt;
  double *upBase;          // Psi1 workspace: slab pad1 of an item holds the corner rows its angle writes (legacy: d_psi1, slab = angle;
                           // in place: d_psi, slab = angle; ring: the ring, slab = slot)
  double *bBase;           // buffer whose slab tails are Set%PsiB (legacy: d_psi1, single-psi layout: d_psi); slab = angle
  int ncG;                 // nc * G: row offsets at or beyond it address boundary elements
  const double *weight;    // quadrature weights and PhiTotal for the phi-tally items (ring mode)
  double *phi;
  const ZoneRec *recs;
  const int2 *zinfo;       // (NA, nz) in sweep order: first corner row, zone | numCorner << 28
};

__device__ __forceinline__ double dot3_seq(const double *om, const double *A) {
  // DOT_PRODUCT order, no contraction: the signs decide incoming/outgoing exactly as on the host
  return __dadd_rn(__dadd_rn(__dmul_rn(om[0], A[0]), __dmul_rn(om[1], A[1])), __dmul_rn(om[2], A[2]));
}

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------
// generic zone solve: SweepUCBxyz.F90:103-306 for one (zone, group)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void solve_zone_generic(const Sweep3DParams &P, int a, int upIdx, int zone0, int g) {
  const int G = P.G, nc = P.nc;
  // omega . A without FMA contraction: the schedule (host snneed, plan_build_kernel) decided incident / exiting from the same
  // uncontracted sums, and a face classified differently here would read a row the schedule does not order
  const double om[3] = {P.omega[3 * a], P.omega[3 * a + 1], P.omega[3 * a + 2]};
  const size_t slab = (size_t)(nc + P.nb) * G;   // Psi, Psi1 are (G, nc+nb, NA): boundary rows follow the corner rows
  const double *psiA = P.psi + (size_t)a * slab;
  double *psi1A = P.upBase + (size_t)upIdx * slab;
  double *psibA = P.bBase + (size_t)a * slab + (size_t)nc * G;
  const unsigned char *nextC = P.nextC + (size_t)a * nc;
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int nCorner = P.numCorner[zone], c0 = P.cOffSet[zone];
  const double sig = P.sigt[(size_t)zone * G + g];

  double Q[MAXC], src[MAXC], sumArea[MAXC], vol[MAXC];
  int nxez[MAXC], ez_exit[MAXC][MAXCF];
  double coefpsi[MAXC][MAXCF];
  for (int c = 0; c < nCorner; c++) {
    const size_t r = (size_t)(c0 + c) * G + g;
    const double source = P.stotal[r] + P.tau * psiA[r];
    vol[c] = P.Volume[c0 + c];
    Q[c] = source;
    src[c] = vol[c] * source;
    nxez[c] = 0;
  }
  for (int c = 0; c < nCorner; c++) {
    const int cc = c0 + c;
    const int nCF = P.nCFaces[cc];
    double afp[MAXCF], psifp[MAXCF];
    double sa = 0.0;
    for (int f = 0; f < nCF; f++) {
      const double *A = P.Afp + ((size_t)cc * MAXCF + f) * 3;
      afp[f] = dot3_seq(om, A);
      psifp[f] = 0.0;
      if (afp[f] > 0.0) {
        sa += afp[f];
      } else if (afp[f] < 0.0) {
        const int row = P.cFP[cc * MAXCF + f];
        psifp[f] = row < nc ? __ldcg(&psi1A[(size_t)row * G + g]) : __ldcg(&psibA[(size_t)(row - nc) * G + g]);
        src[c] -= afp[f] * psifp[f];
      }
    }
    for (int f = 0; f < nCF; f++) {
      const double *A = P.Aez + ((size_t)cc * MAXCF + f) * 3;
      const double aez = dot3_seq(om, A);
      const int cez = P.cEZ[cc * MAXCF + f];
      if (cez > c) {
        if (aez > 0.0) { ez_exit[c][nxez[c]] = cez; coefpsi[c][nxez[c]] = aez; nxez[c]++; }
        else if (aez < 0.0) { ez_exit[cez][nxez[cez]] = c; coefpsi[cez][nxez[cez]] = -aez; nxez[cez]++; }
      }
      if (aez > 0.0) {
        sa += aez;
        double area_opp = 0.0, psi_opp = 0.0;
        if (nCF == 3) {
          const int ifp = (f + 1) % 3;
          if (afp[ifp] < 0.0) { psi_opp = psifp[ifp]; area_opp = -afp[ifp]; }
        } else {
          int ifp = f;
          for (int k = 0; k < nCF - 2; k++) {
            ifp = (ifp + 1) % nCF;
            if (afp[ifp] < 0.0) { area_opp -= afp[ifp]; psi_opp -= afp[ifp] * psifp[ifp]; }
          }
          if (area_opp > 0.0) psi_opp *= 1.0 / area_opp;
        }
        double sez;
        if (area_opp > 0.0) {
          const double aez2 = aez * aez, v = vol[c];
          const double sigv = sig * v, sigv2 = sigv * sigv;
          const double gnum = aez2 * (FOURALPHA * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
          const double gden = v * (4.0 * sigv * sigv2 + aez * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
          sez = (v * gnum * (sig * psi_opp - Q[c]) + 0.5 * aez * gden * (Q[c] - Q[cez])) / (gnum + gden * sig);
        } else {
          sez = 0.5 * aez * (1.0 / sig) * (Q[c] - Q[cez]);
        }
        src[c] += sez;
        src[cez] -= sez;
      }
    }
    sumArea[c] = sa;
  }
  if (zone0 > 0) {
    for (int i = 0; i < nCorner; i++) {
      const int c = nextC[c0 + i];
      const double p = src[c] / (sumArea[c] + sig * vol[c]);
      src[c] = p;   // src now holds the corner flux
      for (int k = 0; k < nxez[c]; k++) src[ez_exit[c][k]] += coefpsi[c][k] * p;
    }
  } else {
    // intra-zone cycle: Jacobi on the previous Psi1 (SweepUCBxyz.F90:283-298)
    for (int c = 0; c < nCorner; c++) {
      const double old = __ldcg(&psi1A[(size_t)(c0 + c) * G + g]);
      for (int k = 0; k < nxez[c]; k++) src[ez_exit[c][k]] += coefpsi[c][k] * old;
    }
    for (int c = 0; c < nCorner; c++) src[c] = src[c] / (sumArea[c] + sig * vol[c]);
  }
  for (int c = 0; c < nCorner; c++) {
    const int cc = c0 + c;
    psi1A[(size_t)cc * G + g] = src[c];
    const int nCF = P.nCFaces[cc];
    for (int f = 0; f < nCF; f++) {
      const int row = P.cFP[cc * MAXCF + f];
      if (row >= nc) {
        const double *A = P.Afp + ((size_t)cc * MAXCF + f) * 3;
        const double afp = dot3_seq(om, A);
        if (afp > 0.0) psibA[(size_t)(row - nc) * G + g] = src[c];
      }
    }
  }
}

__device__ __noinline__ void solve_zone_slow(const Sweep3DParams &P, int a, int upIdx, int zone0, int g) { solve_zone_generic(P, a, upIdx, zone0, g); }

__global__ void __launch_bounds__(128) sweep3d_generic_kernel(Sweep3DParams P) {
  __shared__ int s_item;
  const int G = P.G;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(&P.counters[0], 1);
    __syncthreads();
    const int it = s_item;
    if (it >= P.nItems) break;
    const WorkItem w = P.items[it];
    if (threadIdx.x == 0 && w.wait_idx >= 0) {
      while (ld_acquire(&P.counters[1 + w.wait_idx]) < w.wait_count) __nanosleep(64);
    }
    __syncthreads();
    const int *nextZ = P.nextZ + (size_t)w.angle * P.nz;
    const int npairs = (w.zend - w.zbeg) * G;
    for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
      const int zi = idx / G, g = idx - zi * G;
      solve_zone_generic(P, w.angle, w.pad1, nextZ[w.zbeg + zi], g);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.counters[1 + w.signal_idx], 1);
    }
  }
}

// ---------------------------------------------------------------------------
// plan records
// ---------------------------------------------------------------------------
struct PlanBuildParams {
  int nc, nb, nz, NA, G;
  const int *numCorner, *cOffSet, *nCFaces, *cFP, *cEZ;
  const double *Volume, *Afp, *Aez, *omega;
  const int *nextZ;
  const unsigned char *nextC;
  ZoneRec *recs;
  int *nSlow;     // [0] zones on the slow path, [1] zones on the canonical (register-resident) path
  int canon;      // try the canonical order
};

// The in-zone corner graph of a hexahedron swept in a generic direction is the cube DAG: one source corner,
// its three neighbours, their three pairwise common neighbours, one sink.  "Canonical" solve order numbers
// them 0 | 1 2 3 | 4 5 6 | 7 with 4 = common(1,2), 5 = common(1,3), 6 = common(2,3), which makes the
// downstream position of every outgoing EZ face a compile-time constant (solve_zone_canon keeps the whole
// zone in registers).  Any topological order of the DAG gives the reference's corner fluxes (only the order
// in which the pushes into a corner are added differs from nextC's).
__host__ __device__ constexpr int cn_nout(int p) { return p == 0 ? 3 : (p < 4 ? 2 : (p < 7 ? 1 : 0)); }
__host__ __device__ constexpr int cn_dst(int p, int k) { return p == 0 ? 1 + k : (p == 1 ? 4 + k : (p == 2 ? (k == 0 ? 4 : 6) : (p == 3 ? 5 + k : 7))); }
__host__ __device__ constexpr int cn_e0(int p) { return p == 0 ? 0 : (p == 1 ? 3 : (p == 2 ? 5 : (p == 3 ? 7 : 5 + p))); }

// canonical order of the zone's corners from the signs of omega.A_ez, or false if the graph is not the cube DAG
__device__ bool canonical_order(const PlanBuildParams &B, int c0, const double (&aezL)[MAXC][3], int (&order)[MAXC]) {
  int out[MAXC][3], nout[MAXC], indeg[MAXC];
  for (int c = 0; c < MAXC; c++) { nout[c] = 0; indeg[c] = 0; }
  for (int c = 0; c < MAXC; c++)
    for (int f = 0; f < 3; f++)
      if (aezL[c][f] > 0.0) {
        const int q = B.cEZ[(c0 + c) * 3 + f];
        if (q < 0 || q >= MAXC) return false;
        out[c][nout[c]++] = q; indeg[q]++;
      }
  int src = -1;
  for (int c = 0; c < MAXC; c++)
    if (indeg[c] == 0) { if (src >= 0) return false; src = c; }
  if (src < 0 || nout[src] != 3) return false;
  int n1[3] = {out[src][0], out[src][1], out[src][2]};
  for (int i = 0; i < 3; i++)      // ascending local corner id: deterministic
    for (int j = i + 1; j < 3; j++)
      if (n1[j] < n1[i]) { const int t = n1[i]; n1[i] = n1[j]; n1[j] = t; }
  for (int i = 0; i < 3; i++) if (indeg[n1[i]] != 1 || nout[n1[i]] != 2) return false;
  auto common = [&](int x, int y) {
    int r = -1, cnt = 0;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) if (out[x][i] == out[y][j]) { r = out[x][i]; cnt++; }
    return cnt == 1 ? r : -1;
  };
  const int m4 = common(n1[0], n1[1]), m5 = common(n1[0], n1[2]), m6 = common(n1[1], n1[2]);
  if (m4 < 0 || m5 < 0 || m6 < 0 || m4 == m5 || m4 == m6 || m5 == m6) return false;
  const int ms[3] = {m4, m5, m6};
  for (int i = 0; i < 3; i++) if (indeg[ms[i]] != 2 || nout[ms[i]] != 1) return false;
  const int sink = out[m4][0];
  if (out[m5][0] != sink || out[m6][0] != sink || indeg[sink] != 3 || nout[sink] != 0) return false;
  order[0] = src; order[1] = n1[0]; order[2] = n1[1]; order[3] = n1[2]; order[4] = m4; order[5] = m5; order[6] = m6; order[7] = sink;
  return true;
}

// the record of one (zone, angle) with the corners taken in the order localc[]; false: the zone needs the slow path
__device__ bool build_record(const PlanBuildParams &B, ZoneRec &R, const int NC, const int c0, const double (&om)[3], const int (&localc)[MAXC],
                             const double (&afpL)[MAXC][3], const double (&aezL)[MAXC][3]) {
  const int G = B.G;
  for (int i = 0; i < MAXC; i++) {
    R.nIn[i] = 0; R.nOut[i] = 0; R.crow[i] = c0 * G; R.coff[i] = 0; R.vol[i] = 0.0; R.sumArea[i] = 1.0;
    for (int k = 0; k < 3; k++) { R.inOff[i][k] = 0; R.inAfp[i][k] = 0.0; R.exitOff[i][k] = 0; }
  }
  for (int k = 0; k < 12; k++) { R.edge[k].ainv = 0.0; R.edge[k].cp = 0.0; R.edge[k].ha = 0.0; R.edge[k].qoff = 0; R.edge[k].hasOpp = 0; }
  unsigned exitMask = 0;
  for (int p = 0; p < NC; p++) {
    const int c = localc[p], cc = c0 + c;
    double sa = 0.0;
    for (int f = 0; f < 3; f++) {
      const int row = B.cFP[cc * 3 + f];
      R.exitOff[p][f] = row * G;
      if (afpL[c][f] > 0.0) {
        sa += afpL[c][f];
        if (row >= B.nc) exitMask |= 1u << (p * 3 + f);
      }
    }
    for (int f = 0; f < 3; f++)
      if (aezL[c][f] > 0.0) sa += aezL[c][f];
    R.sumArea[p] = sa;
    R.vol[p] = B.Volume[cc];
    R.crow[p] = cc * G;
    R.coff[p] = c * (G / 2) * 16;
  }
  // outgoing EZ faces grouped by upstream position, downstream position ascending
  int slot = 0;
  for (int p = 0; p < NC; p++) {
    const int c = localc[p];
    int nout = 0;
    bool used[3] = {false, false, false};       // FP faces already placed in an incident slot
    int slotFace[3] = {-1, -1, -1};
    for (int q = p + 1; q < NC; q++) {
      const int cq = localc[q];
      int f = -1, fq = -1;
      for (int k = 0; k < 3; k++) {
        if (B.cEZ[(c0 + c) * 3 + k] == cq) { if (f >= 0) return false; f = k; }
        if (B.cEZ[(c0 + cq) * 3 + k] == c) { if (fq >= 0) return false; fq = k; }
      }
      if (f < 0 && fq < 0) continue;
      if (f < 0 || fq < 0) return false;
      const bool sezFwd = aezL[c][f] > 0.0, sezBwd = aezL[cq][fq] > 0.0;
      // downstream push (coefpsi) is decided by the lower local corner id, SweepUCBxyz.F90:168-179
      const bool loIsP = c < cq;
      const double alo = loIsP ? aezL[c][f] : aezL[cq][fq];
      const bool pushFwd = loIsP ? alo > 0.0 : alo < 0.0;
      const bool pushBwd = loIsP ? alo < 0.0 : alo > 0.0;
      if (sezBwd || pushBwd || sezFwd != pushFwd) return false;
      if (!sezFwd) continue;
      if (slot >= 12 || nout >= 3) return false;
      const double av = aezL[c][f];
      const int ifp = (f + 1) % 3;             // the FP face "opposite" EZ face f (SweepUCBxyz.F90:187-194)
      ZoneEdge &E = R.edge[slot];
      E.ainv = 1.0 / av;
      E.cp = alo > 0.0 ? alo : -alo;
      E.ha = 0.5 * av;
      E.qoff = cq * (G / 2) * 16;
      E.hasOpp = afpL[c][ifp] < 0.0 ? 1 : 0;
      if (E.hasOpp) { slotFace[nout] = ifp; used[ifp] = true; }   // incident slot k serves edge k
      slot++; nout++;
    }
    R.nOut[p] = (unsigned char)nout;
    // the other incident faces take the free slots
    for (int f = 0; f < 3; f++) {
      if (!(afpL[c][f] < 0.0) || used[f]) continue;
      for (int k = 0; k < 3; k++)
        if (slotFace[k] < 0) { slotFace[k] = f; used[f] = true; break; }
    }
    int nin = 0;
    for (int k = 0; k < 3; k++)
      if (slotFace[k] >= 0) {
        const int f = slotFace[k];
        R.inOff[p][k] = B.cFP[(c0 + c) * 3 + f] * G;
        R.inAfp[p][k] = afpL[c][f];
        nin = k + 1;
      }
    R.nIn[p] = (unsigned char)nin;
  }
  R.exitMask = exitMask;
  if (exitMask) R.flags |= ZREC_HAS_EXIT;
  return true;
}

// one thread per (angle, position in sweep order)
__global__ void __launch_bounds__(128) plan_build_kernel(PlanBuildParams B) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)B.NA * B.nz) return;
  const int a = (int)(t / B.nz);
  ZoneRec &R = B.recs[t];
  const int zone0 = B.nextZ[t];
  const int zone = (zone0 < 0 ? -zone0 : zone0) - 1;
  const int NC = B.numCorner[zone], c0 = B.cOffSet[zone];
  const double om[3] = {B.omega[3 * a], B.omega[3 * a + 1], B.omega[3 * a + 2]};
  R.c0 = c0;
  R.zone0 = zone0;
  R.exitMask = 0u;
  R.flags = (unsigned)(NC & 15);
  bool slow = zone0 < 0 || NC > MAXC;
  for (int c = 0; c < NC && !slow; c++) slow = B.nCFaces[c0 + c] != 3;
  double afpL[MAXC][3], aezL[MAXC][3];   // by local corner
  for (int c = 0; c < MAXC; c++)
    for (int f = 0; f < 3; f++) { afpL[c][f] = 0.0; aezL[c][f] = 0.0; }
  for (int c = 0; c < NC && !slow; c++)
    for (int f = 0; f < 3; f++) {
      afpL[c][f] = dot3_seq(om, B.Afp + ((size_t)(c0 + c) * 3 + f) * 3);
      aezL[c][f] = dot3_seq(om, B.Aez + ((size_t)(c0 + c) * 3 + f) * 3);
    }
  int localc[MAXC];
  if (!slow && NC == MAXC && B.canon && canonical_order(B, c0, aezL, localc) && build_record(B, R, NC, c0, om, localc, afpL, aezL)) {
    // fast path needs exactly the static edge table
    bool ok = true;
    for (int p = 0; p < MAXC && ok; p++) {
      ok = R.nOut[p] == cn_nout(p);
      for (int k = 0; k < cn_nout(p) && ok; k++) ok = R.edge[cn_e0(p) + k].qoff == localc[cn_dst(p, k)] * (B.G / 2) * 16;
    }
    if (ok) { R.flags |= ZREC_CANON; atomicAdd(B.nSlow + 1, 1); return; }
    R.exitMask = 0u; R.flags = (unsigned)(NC & 15);
  }
  // the schedule's own corner order (nextC)
  const unsigned char *nextC = B.nextC + (size_t)a * B.nc + c0;
  int pos[MAXC];
  for (int c = 0; c < MAXC; c++) { pos[c] = -1; localc[c] = 0; }
  for (int i = 0; i < NC && !slow; i++) {
    const int c = nextC[i];
    if (c >= NC || pos[c] >= 0) slow = true; else pos[c] = i;
    localc[i] = c;
  }
  if (!slow) slow = !build_record(B, R, NC, c0, om, localc, afpL, aezL);
  if (slow) {
    R.exitMask = 0u;
    R.flags = (unsigned)(NC & 15) | ZREC_SLOW;
    atomicAdd(B.nSlow, 1);
  }
}

// ---------------------------------------------------------------------------
// mbarrier / TMA helpers (sm_90+ PTX)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{ .reg .pred p;\n"
      "WAIT_%=:\n"
      "  mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "  @p bra DONE_%=;\n"
      "  bra WAIT_%=;\n"
      "DONE_%=: }\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ double rcp_fast(double x) {
  // MUFU.RCP64H seed (2^-23) + two Newton steps: ~1 ulp for normal x > 0 (all denominators here are
  // sums of positive areas and sigma*volume)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ bool mbar_test(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{ .reg .pred p;\n"
      "  mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "  selp.u32 %0, 1, 0, p; }\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// L2 eviction-priority policies (the encodings CUTLASS names TMA::CacheHintSm90::EVICT_FIRST / EVICT_LAST)
constexpr unsigned long long L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_1d_hint(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// two consecutive groups per lane: 16-byte loads/stores, the record decode is shared by both
struct V2 { double x, y; };
__device__ __forceinline__ V2 ld_l2(const double *p) {   // upstream Psi1 rows: written by other SMs, L1 must be bypassed
  V2 r;
  asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_keep(double *p, const V2 &v) {   // Psi1 rows: wanted again from L2 by the downstream zones
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(L2_EVICT_LAST) : "memory");
}

#ifndef PLAN_MINB
#define PLAN_MINB 2            // CTAs per SM the plan kernel is compiled for (register cap 65536 / (192 * PLAN_MINB))
#endif
#ifndef PLAN_MINB2
#define PLAN_MINB2 2           // same for the variant with 4 groups per lane (NH = 2)
#endif
constexpr int PLAN_MAX_STAGES = 12;
constexpr int PLAN_ZMAX = 8;     // zones per item at most
#ifndef PLAN_NCW_DEF
#define PLAN_NCW_DEF 4
#endif
constexpr int PLAN_NCW = PLAN_NCW_DEF;      // consumer warps per CTA
constexpr int PLAN_CTL_BYTES = 896;         // barriers + per-stage metadata ahead of the stages
constexpr int PLAN_LANES = PLAN_NCW * 32;

// How the consumer warps of a CTA are grouped for a given group count (host side, once per context).
// An "engine" is the set of warps that solves one work item: the lanes of one zone (two groups per
// lane), or for small G one warp holding several zones.  Engines run independently of each other on
// their own pipeline stages; with NE engines and NE+1 stages only one stage per CTA is a landing
// buffer in flight, the others are being computed on.
struct PlanGeom {
  int wpe;          // warps per engine (1, 2 or 4)
  int nEngines;     // PLAN_NCW / wpe
  int zpi;          // zones per item
  int nStages;
  int stageBytes;   // psi | st | sigt | recs
  int offSt, offSigt, offRecs;
  size_t smemBytes;
};
static PlanGeom plan_geom(int G, int NH) {
  PlanGeom g;
  const int Gv = G / 2, LZ = std::max(Gv / NH, 1);   // 16-byte columns per zone; lanes per zone (a lane owns NH columns)
  g.wpe = LZ > 64 ? 4 : (LZ > 32 ? 2 : 1);   // an engine must hold a whole zone (LZ lanes)
  // 32 <= G <= 64: two-warp engines all the same, each warp on its own zone(s) of the item -- half as many items, tickets and
  // signals, and the warp-group build (12 consumer warps per SM) applies: -d 20 -G 64 25.6 -> 20.0 ms, -G 32 14.5 -> 13.4 ms
  if (NH == 1 && LZ >= 16 && LZ <= 32) g.wpe = 2;
  if (const char *e = getenv("UMT_PLAN_WPE_MIN")) g.wpe = std::max(LZ > 64 ? 4 : (LZ > 32 ? 2 : 1), std::min(2, atoi(e)));
  g.nEngines = PLAN_NCW / g.wpe;
  const int LE = 32 * g.wpe;
  g.zpi = std::max(1, std::min(PLAN_ZMAX, LE / LZ));
  if (const char *e = getenv("UMT_ZONES_PER_ITEM")) g.zpi = std::max(1, std::min(g.zpi, atoi(e)));
  g.offSt = NH * LE * MAXC * 16;
  g.offSigt = 2 * g.offSt;
  g.offRecs = g.offSigt + NH * LE * 16;
  g.stageBytes = (g.offRecs + g.zpi * (int)sizeof(ZoneRec) + 127) / 128 * 128;
  // as many landing stages as still let the CTAs the kernel is compiled for share an SM (228 KB, 1 KB reserved per CTA)
  const int budget = (228 * 1024) / (NH == 1 ? PLAN_MINB : PLAN_MINB2) - 1024;
  g.nStages = std::min(PLAN_MAX_STAGES, std::max(g.nEngines + 1, std::min(g.nEngines + 3, (budget - PLAN_CTL_BYTES) / g.stageBytes)));
  if (const char *e = getenv("UMT_PLAN_STAGES")) g.nStages = std::max(g.nEngines + 1, std::min(PLAN_MAX_STAGES, atoi(e)));
  g.smemBytes = PLAN_CTL_BYTES + (size_t)g.nStages * g.stageBytes;
  return g;
}

// n: zones of the item; -1: sentinel (the engine is done); -2: phi-tally item (angle = first angle of the batch, upIdx = packed
// slot / count / first flag, c0..c1 = its 16-byte columns of PhiTotal)
struct StageMeta { int angle, n, signal_idx, wait_idx, wait_count, upIdx, c0, c1; };

// Shared memory: barriers and per-stage metadata in the first PLAN_CTL_BYTES, then the stages.  One stage = one
// work item: the TMA landing area of its Psi^n / STotal / Sigt rows, [zone][corner][G] (a zone's corner
// rows are contiguous in HBM, so each is one bulk copy), and its plan records.  The consumers turn the
// Psi^n area into Q and the STotal area into the running sources in place: a lane only ever touches
// its own 16-byte column.
constexpr int PLAN_RING = 32;    // completion-signal ring between the loader and the signaller warp
struct PlanCtl {
  unsigned long long full[PLAN_MAX_STAGES], empty[PLAN_MAX_STAGES];
  StageMeta meta[PLAN_MAX_STAGES];
  int sigRing[PLAN_RING];        // signal_idx of the CTA's item k at k % PLAN_RING
  int sig2Ring[PLAN_RING];       // its second counter (-1: none)
  volatile int issuedCount;      // real items issued so far (loader -> signaller)
  volatile int doneFlag;         // loader finished: issuedCount is final
  volatile int nSignaled;        // items whose completion is published (signaller -> loader)
  int pad;
};
static_assert(sizeof(PlanCtl) <= PLAN_CTL_BYTES, "PlanCtl must fit the control block");

__device__ __forceinline__ V2 lds_v2(unsigned addr) {
  V2 r;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_v2(unsigned addr, const V2 &v) {
  asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

// One corner of the zone solve (position p of the record's solve order): incident FP fluxes (pfC, loaded
// while the previous corner was solved), the EZ closure terms of its outgoing faces, the corner flux and
// its push into the downstream corners; it also puts the next corner's incident rows in flight (pfN).
// Q and the running sources live in shared memory (16-byte columns, a lane only touches its own) because the
// downstream corner of an edge is only known from the record.  A lane owns NH columns (2 NH groups) that lie
// LZ columns apart (hs bytes in shared memory, hg doubles in global memory): every access of a warp stays a
// contiguous run of 16-byte words, and the NH independent dependency chains interleave in the FP64 pipe.
template <int NH, bool TWO>
__device__ __forceinline__ void plan_corner(const ZoneRec *__restrict__ R, const ZoneEdge *__restrict__ &E, const int p, const int NC,
                                            const V2 (&pfC)[3][NH], V2 (&pfN)[3][NH], double *__restrict__ upg, double *__restrict__ pbg, const int ncG,
                                            const unsigned qs, const unsigned ss, const V2 (&sig)[NH], const V2 (&rsig)[NH], const unsigned flags,
                                            const unsigned hs, const int hg) {
  const int nout = R->nOut[p];
  if (p + 1 < NC) {   // incident rows of the next corner
    const int nn = R->nIn[p + 1];
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (k < nn) {
        const int o = R->inOff[p + 1][k];
        const double *src = PLAN_ROW(o);
#pragma unroll
        for (int h = 0; h < NH; h++) pfN[k][h] = ld_l2(src + h * hg);
      }
  }
  const unsigned co = (unsigned)R->coff[p];
  const double vp = R->vol[p];
  V2 s[NH], qp[NH], sv[NH];
#pragma unroll
  for (int h = 0; h < NH; h++) {
    s[h] = lds_v2(ss + co + h * hs);
    qp[h] = lds_v2(qs + co + h * hs);
    sv[h].x = sig[h].x * vp; sv[h].y = sig[h].y * vp;
  }
  // incident fluxes across FP faces (SweepUCBxyz.F90:139-161); unused slots carry afp = 0 and a finite pf
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double af = R->inAfp[p][k];
#pragma unroll
    for (int h = 0; h < NH; h++) { s[h].x = fma(-af, pfC[k][h].x, s[h].x); s[h].y = fma(-af, pfC[k][h].y, s[h].y); }
  }
  // EZ faces leaving this corner (SweepUCBxyz.F90:182-252), with x = sigma V / aez:
  //   sez = V [N(x)(sigma psi_opp - Q) + D(x)(Q - Q_cez)/2] / (N(x) + x D(x)),
  //   N = 1.82 x^2 + 4 x + 3,  D = 4 x^3 + 6 x^2 + 4 x + 2   (gnum = aez^4 N, gden = V aez^3 D)
  V2 sezk[3][NH];
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (k < nout) {
      const ZoneEdge e = E[k];
      if (e.hasOpp) {
#pragma unroll
        for (int h = 0; h < NH; h++) {
          const V2 qq = lds_v2(qs + (unsigned)e.qoff + h * hs);
          const V2 po = pfC[k][h];
          {
            const double x = sv[h].x * e.ainv;
            const double N = fma(fma(FOURALPHA, x, 4.0), x, 3.0);
            const double D = fma(fma(fma(4.0, x, 6.0), x, 4.0), x, 2.0);
            const double num = fma(N, fma(sig[h].x, po.x, -qp[h].x), (0.5 * D) * (qp[h].x - qq.x));
            sezk[k][h].x = (vp * num) * rcp_fast(fma(x, D, N));
          }
          {
            const double x = sv[h].y * e.ainv;
            const double N = fma(fma(FOURALPHA, x, 4.0), x, 3.0);
            const double D = fma(fma(fma(4.0, x, 6.0), x, 4.0), x, 2.0);
            const double num = fma(N, fma(sig[h].y, po.y, -qp[h].y), (0.5 * D) * (qp[h].y - qq.y));
            sezk[k][h].y = (vp * num) * rcp_fast(fma(x, D, N));
          }
        }
      } else {
#pragma unroll
        for (int h = 0; h < NH; h++) {
          const V2 qq = lds_v2(qs + (unsigned)e.qoff + h * hs);
          sezk[k][h].x = (e.ha * (qp[h].x - qq.x)) * rsig[h].x;
          sezk[k][h].y = (e.ha * (qp[h].y - qq.y)) * rsig[h].y;
        }
      }
#pragma unroll
      for (int h = 0; h < NH; h++) { s[h].x += sezk[k][h].x; s[h].y += sezk[k][h].y; }
    }
  // corner flux, then its push into the downstream corners (SweepUCBxyz.F90:261-281)
  const double sa = R->sumArea[p];
  V2 psi[NH];
#pragma unroll
  for (int h = 0; h < NH; h++) {
    psi[h].x = s[h].x * rcp_fast(sa + sv[h].x);
    psi[h].y = s[h].y * rcp_fast(sa + sv[h].y);
    st_keep(upg + R->crow[p] + h * hg, psi[h]);
  }
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (k < nout) {
      const unsigned qa = ss + (unsigned)E[k].qoff;
      const double cp = E[k].cp;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        V2 t = lds_v2(qa + h * hs);
        t.x = fma(cp, psi[h].x, t.x - sezk[k][h].x);
        t.y = fma(cp, psi[h].y, t.y - sezk[k][h].y);
        sts_v2(qa + h * hs, t);
      }
    }
  if (flags & ZREC_HAS_EXIT) {
    const unsigned em = R->exitMask >> (p * 3);
#pragma unroll
    for (int f = 0; f < 3; f++)
      if (em & (1u << f)) {
#pragma unroll
        for (int h = 0; h < NH; h++) st_keep(PLAN_BROW(R->exitOff[p][f]) + h * hg, psi[h]);
      }
  }
  E += nout;
}

template <int NH, bool TWO>
__device__ __forceinline__ void solve_zone_plan(const double tau, const ZoneRec *__restrict__ R, double *__restrict__ upg, double *__restrict__ pbg,
                                                const int ncG, const unsigned qs, const unsigned ss, const V2 (&sig)[NH], const unsigned hs, const int hg) {
  const unsigned flags = R->flags;
  const int NC = (int)(flags & 15u);
  // Q = STotal + tau Psi^n, src = V Q (SweepUCBxyz.F90:119-126), in place over the landed rows
#pragma unroll
  for (int p = 0; p < MAXC; p++) {
    if (p < NC) {
      const unsigned co = (unsigned)R->coff[p];
      const double v = R->vol[p];
#pragma unroll
      for (int h = 0; h < NH; h++) {
        const V2 a = lds_v2(qs + co + h * hs), b = lds_v2(ss + co + h * hs);
        V2 q, s;
        q.x = fma(tau, a.x, b.x); q.y = fma(tau, a.y, b.y);
        s.x = v * q.x; s.y = v * q.y;
        sts_v2(qs + co + h * hs, q);
        sts_v2(ss + co + h * hs, s);
      }
    }
  }
  V2 rsig[NH];
#pragma unroll
  for (int h = 0; h < NH; h++) { rsig[h].x = rcp_fast(sig[h].x); rsig[h].y = rcp_fast(sig[h].y); }
  const ZoneEdge *E = R->edge;
  V2 pfA[3][NH], pfB[3][NH];
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int h = 0; h < NH; h++) { pfA[k][h].x = pfA[k][h].y = 0.0; pfB[k][h].x = pfB[k][h].y = 0.0; }
  {
    const int n0 = R->nIn[0];
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (k < n0) {
        const int o = R->inOff[0][k];
        const double *src = PLAN_ROW(o);
#pragma unroll
        for (int h = 0; h < NH; h++) pfA[k][h] = ld_l2(src + h * hg);
      }
  }
#pragma unroll 1
  for (int p = 0; p < NC; p += 2) {
    plan_corner<NH, TWO>(R, E, p, NC, pfA, pfB, upg, pbg, ncG, qs, ss, sig, rsig, flags, hs, hg);
    if (p + 1 < NC) plan_corner<NH, TWO>(R, E, p + 1, NC, pfB, pfA, upg, pbg, ncG, qs, ss, sig, rsig, flags, hs, hg);
  }
}

// Psi1 rows of other zones are never re-read by the zone that stores them, so these stores need not order the
// compiler's loads (no "memory" clobber): the incident rows of later corners may be fetched ahead of them.
__device__ __forceinline__ void st_keep_free(double *p, const V2 &v) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(L2_EVICT_LAST));
}
__device__ __forceinline__ V2 ldcg_v2(const double *p) {
  const double2 t = __ldcg(reinterpret_cast<const double2 *>(p));
  V2 r; r.x = t.x; r.y = t.y;
  return r;
}

// Zone solve for records in canonical cube order (ZREC_CANON): the downstream position of every EZ face is a
// compile-time constant, so Q, the running sources and the corner fluxes of all eight corners stay in registers,
// the code is one straight line (the three corners of a level are independent chains the scheduler interleaves)
// and shared memory is only read: the landed Psi^n / STotal columns once each, plus the record fields.
// An edge whose opposite FP face is not incident uses the same closure with N = 0, which is algebraically the
// reference's sez = aez (Q - Q_cez) / (2 sigma) (SweepUCBxyz.F90:254-256).
template <int NH, bool TWO>
__device__ __forceinline__ void solve_zone_canon(const double tau, const ZoneRec *__restrict__ R, double *__restrict__ upg, double *__restrict__ pbg,
                                                 const int ncG, const unsigned char *__restrict__ colPsi, const unsigned char *__restrict__ colSt,
                                                 const V2 (&sig)[NH], const unsigned hs, const int hg) {
  V2 Q[MAXC][NH], S[MAXC][NH];
  V2 pf[MAXC][3][NH];   // only the slots k < cn_nout(p) exist
  // incident rows of the source corner and of the first level
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int nin = R->nIn[p];
#pragma unroll
    for (int k = 0; k < cn_nout(p); k++) {
      const int o = R->inOff[p][k];
      const double *src = PLAN_ROW(o);
#pragma unroll
      for (int h = 0; h < NH; h++) {
        pf[p][k][h].x = 0.0; pf[p][k][h].y = 0.0;
        if (k < nin) pf[p][k][h] = ldcg_v2(src + h * hg);
      }
    }
  }
  // Q = STotal + tau Psi^n, src = V Q (SweepUCBxyz.F90:119-126)
#pragma unroll
  for (int p = 0; p < MAXC; p++) {
    const unsigned co = (unsigned)R->coff[p];
    const double v = R->vol[p];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      const V2 a = *reinterpret_cast<const V2 *>(colPsi + co + h * hs), b = *reinterpret_cast<const V2 *>(colSt + co + h * hs);
      Q[p][h].x = fma(tau, a.x, b.x); Q[p][h].y = fma(tau, a.y, b.y);
      S[p][h].x = v * Q[p][h].x; S[p][h].y = v * Q[p][h].y;
    }
  }
  const unsigned flags = R->flags;
#pragma unroll
  for (int p = 0; p < MAXC; p++) {
    if (p == 1) {   // second level's incident rows go in flight while the first level is solved
#pragma unroll
      for (int pp = 4; pp < 7; pp++) {
        const int nin = R->nIn[pp];
        const int o = R->inOff[pp][0];
        const double *src = PLAN_ROW(o);
#pragma unroll
        for (int h = 0; h < NH; h++) {
          pf[pp][0][h].x = 0.0; pf[pp][0][h].y = 0.0;
          if (0 < nin) pf[pp][0][h] = ldcg_v2(src + h * hg);
        }
      }
    }
    const double vp = R->vol[p];
    V2 s[NH], sv[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) { s[h] = S[p][h]; sv[h].x = sig[h].x * vp; sv[h].y = sig[h].y * vp; }
    // incident fluxes across FP faces (SweepUCBxyz.F90:139-161); unused slots carry afp = 0 and pf = 0
#pragma unroll
    for (int k = 0; k < cn_nout(p); k++) {
      const double af = R->inAfp[p][k];
#pragma unroll
      for (int h = 0; h < NH; h++) { s[h].x = fma(-af, pf[p][k][h].x, s[h].x); s[h].y = fma(-af, pf[p][k][h].y, s[h].y); }
    }
    // an incident face that is not opposite to an outgoing EZ face (distorted zones only) sits in a slot past the edges
    if (cn_nout(p) < 3 && (int)R->nIn[p] > cn_nout(p)) {
#pragma unroll
      for (int k = cn_nout(p); k < 3; k++)
        if (k < (int)R->nIn[p]) {
          const double af = R->inAfp[p][k];
          const int o = R->inOff[p][k];
          const double *src = PLAN_ROW(o);
#pragma unroll
          for (int h = 0; h < NH; h++) {
            const V2 v = ldcg_v2(src + h * hg);
            s[h].x = fma(-af, v.x, s[h].x); s[h].y = fma(-af, v.y, s[h].y);
          }
        }
    }
    // EZ faces leaving this corner (SweepUCBxyz.F90:182-252), x = sigma V / aez (see plan_corner)
    V2 sezk[3][NH];
#pragma unroll
    for (int k = 0; k < cn_nout(p); k++) {
      const ZoneEdge &e = R->edge[cn_e0(p) + k];
      const int q = cn_dst(p, k);
      const double ainv = e.ainv, nmul = e.hasOpp ? 1.0 : 0.0;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        const V2 po = pf[p][k][h];
        {
          const double x = sv[h].x * ainv;
          const double N = nmul * fma(fma(FOURALPHA, x, 4.0), x, 3.0);
          const double D = fma(fma(fma(4.0, x, 6.0), x, 4.0), x, 2.0);
          const double num = fma(N, fma(sig[h].x, po.x, -Q[p][h].x), (0.5 * D) * (Q[p][h].x - Q[q][h].x));
          sezk[k][h].x = (vp * num) * rcp_fast(fma(x, D, N));
        }
        {
          const double x = sv[h].y * ainv;
          const double N = nmul * fma(fma(FOURALPHA, x, 4.0), x, 3.0);
          const double D = fma(fma(fma(4.0, x, 6.0), x, 4.0), x, 2.0);
          const double num = fma(N, fma(sig[h].y, po.y, -Q[p][h].y), (0.5 * D) * (Q[p][h].y - Q[q][h].y));
          sezk[k][h].y = (vp * num) * rcp_fast(fma(x, D, N));
        }
        s[h].x += sezk[k][h].x; s[h].y += sezk[k][h].y;
      }
    }
    // corner flux, then its push into the downstream corners (SweepUCBxyz.F90:261-281)
    const double sa = R->sumArea[p];
    V2 psi[NH];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      psi[h].x = s[h].x * rcp_fast(sa + sv[h].x);
      psi[h].y = s[h].y * rcp_fast(sa + sv[h].y);
      st_keep_free(upg + R->crow[p] + h * hg, psi[h]);
    }
#pragma unroll
    for (int k = 0; k < cn_nout(p); k++) {
      const int q = cn_dst(p, k);
      const double cp = R->edge[cn_e0(p) + k].cp;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        S[q][h].x = fma(cp, psi[h].x, S[q][h].x - sezk[k][h].x);
        S[q][h].y = fma(cp, psi[h].y, S[q][h].y - sezk[k][h].y);
      }
    }
    if (flags & ZREC_HAS_EXIT) {
      const unsigned em = R->exitMask >> (p * 3);
#pragma unroll
      for (int f = 0; f < 3; f++)
        if (em & (1u << f)) {
#pragma unroll
          for (int h = 0; h < NH; h++) st_keep_free(PLAN_BROW(R->exitOff[p][f]) + h * hg, psi[h]);
        }
    }
  }
}

// Phi-tally item (ring mode): PhiTotal(columns c0..c1) (+)= sum over the nA angles of a retiring batch of w_a Psi1_a, the angles in
// ascending order on top of the running sum, so that PhiTotal ends up bit-identical to the fixed-order sum over all angles
// (control/getPhiTotal_OMPOL.F90:134-160 + SweepUCBxyz.F90:270).  The batch's Psi1 slabs sit in consecutive ring slots.  Everything is
// read and written through L2 (ld.cg / st.cg): other SMs wrote these rows, and ring slots are reused.
__device__ __noinline__ void phi_tally_item(const Sweep3DParams &P, const int a0, const int packed, const int c0, const int c1, const int elane,
                                            const int nLanes) {
  const int slot0 = packed & 0xffff, nA = (packed >> 16) & 0xff;
  const bool first = (packed >> 30) & 1;
  const size_t slab = (size_t)(P.nc + P.nb) * P.G;
  const double2 *src = reinterpret_cast<const double2 *>(P.upBase + (size_t)slot0 * slab);
  const size_t slab2 = slab / 2;
  double2 *phi = reinterpret_cast<double2 *>(P.phi);
  constexpr int U = 4;   // columns per lane in flight: U (nA + 1) independent 16-byte loads hide the DRAM latency of this streaming item
  for (int c = c0 + elane; c < c1; c += U * nLanes) {
    double2 s[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      s[u] = make_double2(0.0, 0.0);
      if (!first && c + u * nLanes < c1) s[u] = __ldcg(phi + c + u * nLanes);
    }
#pragma unroll 2
    for (int i = 0; i < nA; i++) {
      const double wa = P.weight[a0 + i];
      double2 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        v[u] = make_double2(0.0, 0.0);
        if (c + u * nLanes < c1) v[u] = __ldcg(src + (size_t)i * slab2 + c + u * nLanes);
      }
#pragma unroll
      for (int u = 0; u < U; u++) { s[u].x = s[u].x + wa * v[u].x; s[u].y = s[u].y + wa * v[u].y; }
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (c + u * nLanes < c1) __stcg(phi + c + u * nLanes, s[u]);
  }
}

#ifdef PLAN_MAXNREG   // explicit register cap instead of the CTAs-per-SM hint (A/B builds)
#define PLAN_BOUNDS(NH) __maxnreg__(PLAN_MAXNREG)
#else
#define PLAN_BOUNDS(NH) __launch_bounds__(WG ? PLAN_WG_THREADS : PLAN_LANES + 64, WG ? 1 : (NH == 1 ? PLAN_MINB : PLAN_MINB2))
#endif
// MODE 0: Psi1 rows and boundary-element rows in one buffer, slab = angle (legacy layout; in-place savePsi sweep of the single-psi layout)
//      1: ring slots for the Psi1 rows, boundary-element rows in the Psi buffer, phi-tally items (single-psi layout, other sweeps)
// WG: the warp-group build (one CTA of 16 warps per SM).  Warps 0-11 are consumers (three warp groups), warps 12-15 the loader and
// signaller warps of two independent halves (half h: consumer warps 6h .. 6h+5, its own control block and stages).  After the role
// split the consumer warp groups raise their register budget with setmaxnreg and the producer group lowers its own, so the SM runs
// 12 consumer warps at the register count that the 2 x 6-warp CTAs of the plain build give to 8: half as many again to hide the
// latency of the upstream rows behind.
#ifndef PLAN_WG_CONS
#define PLAN_WG_CONS 152
#endif
#ifndef PLAN_WG_PROD
#define PLAN_WG_PROD 40
#endif
constexpr int PLAN_WG_THREADS = 512, PLAN_WG_NCW = 6, PLAN_WG_CONS_REGS = PLAN_WG_CONS, PLAN_WG_PROD_REGS = PLAN_WG_PROD;
static_assert(12 * 32 * PLAN_WG_CONS + 4 * 32 * PLAN_WG_PROD <= 65536, "register file of one SM");
template <int NH, int MODE, bool WG>
__global__ void PLAN_BOUNDS(NH) sweep3d_plan_kernel(Sweep3DParams P) {
  constexpr bool TWO = MODE == 1, RING = MODE == 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int rwarp = tid >> 5;                       // warp of the CTA
  // role index within the (virtual) CTA: 0 .. NCW-1 consumers, NCW loader, NCW+1 signaller
  const int half = WG ? (rwarp < 2 * PLAN_WG_NCW ? rwarp / PLAN_WG_NCW : (rwarp - 2 * PLAN_WG_NCW) >> 1) : 0;
  const int warp = WG ? (rwarp < 2 * PLAN_WG_NCW ? rwarp - half * PLAN_WG_NCW : PLAN_WG_NCW + ((rwarp - 2 * PLAN_WG_NCW) & 1)) : rwarp;
  constexpr int NCW = WG ? PLAN_WG_NCW : PLAN_NCW;
  const int G = P.G, Gv = G >> 1;   // lanes per zone
  const int wpe = P.wpe;
  const int NS = WG ? P.nStagesWG : P.nStages, NE = WG ? PLAN_WG_NCW / wpe : P.nEngines;
  PlanCtl &S = *reinterpret_cast<PlanCtl *>(smem_raw + (WG ? half * PLAN_CTL_BYTES : 0));
  unsigned char *stages = smem_raw + (WG ? 2 * PLAN_CTL_BYTES + (size_t)half * NS * P.stageBytes : PLAN_CTL_BYTES);
  const size_t slab = (size_t)(P.nc + P.nb) * G;
  if (lane == 0 && (WG ? (rwarp == 0 || rwarp == PLAN_WG_NCW) : tid == 0)) {   // one thread per control block
    for (int s = 0; s < NS; s++) { mbar_init(&S.full[s], 2); mbar_init(&S.empty[s], wpe); }
    S.issuedCount = 0; S.doneFlag = 0; S.nSignaled = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp >= NCW) {   // producer roles (WG: the whole fourth warp group)
  if (WG) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PLAN_WG_PROD_REGS));
  if (warp == NCW) {
    // ---------------- loader warp ----------------
    // Per CTA-local sequence number k (stage k % NS, engine k % NE):
    //   issue(k):   next queued item -> TMA of its records and Psi^n/STotal/Sigt rows (needs item k-NS finished by its engine);
    //   release(k): the item's upstream plane is complete -> second arrival on full[k % NS].
    // The signaller warp (below) publishes completions, so neither this warp nor a consumer warp ever waits on a fence.
    // After the last ticket every engine gets one sentinel.
    // Everything with global-memory latency on this warp's critical path is amortised: tickets are taken QB at a
    // time, the lanes fetch the QB descriptors and their zones' info in parallel, the next batch is fetched (one
    // dependent step per loop turn) while the current one is issued, a plane already seen complete is not polled
    // again, and stage indices/parities are kept incrementally (no integer division).
    int nIssued = 0, nReleased = 0, sentinels = -1;   // sentinels < 0: items remain
    int kFill = 0, sFill = 0;                 // next sequence number to fill and its stage
    unsigned parPrev = 1;                     // parity of empty[sFill] completed by the stage's previous occupant (item kFill - NS)
    int sRel = 0;                             // stage of the next release
    int lastOk = -1;                          // wait_idx last seen complete
    const unsigned rowBytes = (unsigned)G * 8u;
    const int zpi = P.zpi, QB = min(P.qbMax, 32 / zpi);
    const int myItem = lane / zpi, myZone = lane - myItem * zpi;
    WorkItem qW, nW;                          // lane l: descriptor of item l / zpi of the current / next batch
    int2 qZ = make_int2(0, 0), nZ = make_int2(0, 0);   // info of zone l % zpi of that item
    int qCount = 0, qPos = 0, nCount = 0, nPhase = 0, nT = 0;
    bool exhausted = false;                   // no tickets left beyond the next batch
    qW.angle = qW.zbeg = qW.zend = qW.wait_idx = qW.wait_count = qW.signal_idx = qW.pad0 = qW.pad1 = 0; nW = qW;
    auto fetch_step = [&]() {                 // one dependent step of fetching the next batch
      if (nPhase == 0) {
        if (lane == 0) nT = atomicAdd(&P.counters[0], QB);
        nPhase = 1;
      } else if (nPhase == 1) {
        nT = __shfl_sync(0xffffffffu, nT, 0);
        nCount = max(0, min(QB, P.nItems - nT));
        if (myItem < nCount) nW = P.items[nT + myItem];
        nPhase = 2;
      } else if (nPhase == 2) {
        if (myItem < nCount && (!RING || nW.angle >= 0) && myZone < nW.zend - nW.zbeg) nZ = P.zinfo[(size_t)nW.angle * P.nz + nW.zbeg + myZone];
        nPhase = 3;
      }
    };
    for (;;) {
      bool progressed = false;
      if (qPos == qCount && !exhausted) {     // current batch used up: take over the next one
        while (nPhase < 3) fetch_step();
        qW = nW; qZ = nZ; qCount = nCount; qPos = 0; nPhase = 0;
        if (qCount < QB) exhausted = true;    // the ticket ran past the item list
      }
      if (!exhausted && nPhase < 3) fetch_step();
      if (nReleased < nIssued) {
        int ok = 1;
        if (lane == 0) {
          const int wi = S.meta[sRel].wait_idx;
          ok = wi < 0 || wi == lastOk || ld_acquire(&P.counters[1 + wi]) >= S.meta[sRel].wait_count;
          if (ok) { mbar_arrive(&S.full[sRel]); lastOk = wi; }
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          nReleased++; progressed = true;
          if (++sRel == NS) sRel = 0;
        }
      }
      if (sentinels < 0 && exhausted && qPos == qCount) sentinels = NE;
      if (sentinels != 0) {
        int free_ = 1;
        if (lane == 0) free_ = (kFill < NS || mbar_test(&S.empty[sFill], parPrev)) && (sentinels > 0 || nIssued - S.nSignaled < PLAN_RING);
        free_ = __shfl_sync(0xffffffffu, free_, 0);
        if (free_) {
          if (sentinels > 0) {
            if (lane == 0) { S.meta[sFill].n = -1; mbar_arrive(&S.full[sFill]); mbar_arrive(&S.full[sFill]); }
            sentinels--;
          } else {
            const int srcLane = qPos * zpi;
            WorkItem w;
            w.angle = __shfl_sync(0xffffffffu, qW.angle, srcLane); w.zbeg = __shfl_sync(0xffffffffu, qW.zbeg, srcLane);
            w.zend = __shfl_sync(0xffffffffu, qW.zend, srcLane); w.wait_idx = __shfl_sync(0xffffffffu, qW.wait_idx, srcLane);
            w.wait_count = __shfl_sync(0xffffffffu, qW.wait_count, srcLane); w.signal_idx = __shfl_sync(0xffffffffu, qW.signal_idx, srcLane);
            w.pad0 = -1; w.pad1 = w.angle;
            if (RING) { w.pad0 = __shfl_sync(0xffffffffu, qW.pad0, srcLane); w.pad1 = __shfl_sync(0xffffffffu, qW.pad1, srcLane); }
            int2 zi;
            zi.x = __shfl_sync(0xffffffffu, qZ.x, (srcLane + lane) & 31); zi.y = __shfl_sync(0xffffffffu, qZ.y, (srcLane + lane) & 31);
            qPos++;
            const bool tally = RING && w.angle < 0;   // phi-tally item: nothing to land, the engine streams straight from global memory
            const int n = tally ? 0 : w.zend - w.zbeg;
            const size_t first = tally ? 0 : (size_t)w.angle * P.nz + w.zbeg;
            unsigned bytes = 0;
            if (lane < n) bytes = rowBytes * (2u * ((unsigned)zi.y >> 28) + 1u);
            bytes = __reduce_add_sync(0xffffffffu, bytes) + (unsigned)(n * sizeof(ZoneRec));
            unsigned char *st = stages + (size_t)sFill * P.stageBytes;
            if (lane == 0) {
              S.meta[sFill].angle = tally ? -1 - w.angle : w.angle; S.meta[sFill].n = tally ? -2 : n;
              S.meta[sFill].wait_idx = w.wait_idx; S.meta[sFill].wait_count = w.wait_count;
              if (RING) { S.meta[sFill].upIdx = w.pad1; S.meta[sFill].c0 = w.zbeg; S.meta[sFill].c1 = w.zend; S.sig2Ring[nIssued & (PLAN_RING - 1)] = w.pad0; }
              S.sigRing[nIssued & (PLAN_RING - 1)] = w.signal_idx;
              __threadfence_block();
              S.issuedCount = nIssued + 1;
              if (tally) mbar_arrive(&S.full[sFill]);
              else {
                mbar_arrive_expect_tx(&S.full[sFill], bytes);
                tma_load_1d_hint(st + P.offRecs, P.recs + first, (unsigned)(n * sizeof(ZoneRec)), &S.full[sFill], L2_EVICT_FIRST);
              }
            }
            __syncwarp();
            if (lane < n) {
              const unsigned nCorner = (unsigned)zi.y >> 28;
              const int zone = zi.y & 0x0fffffff;
              tma_load_1d_hint(st + (size_t)lane * MAXC * Gv * 16, P.psi + (size_t)w.angle * slab + (size_t)zi.x * G, rowBytes * nCorner, &S.full[sFill], L2_EVICT_FIRST);
              tma_load_1d_hint(st + P.offSt + (size_t)lane * MAXC * Gv * 16, P.stotal + (size_t)zi.x * G, rowBytes * nCorner, &S.full[sFill], L2_EVICT_FIRST);
              tma_load_1d_hint(st + P.offSigt + (size_t)lane * Gv * 16, P.sigt + (size_t)zone * G, rowBytes, &S.full[sFill], L2_EVICT_FIRST);
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

  {
    // ---------------- signaller warp ----------------
    // signal(k): item k's engine has arrived on empty[k % NS] -> make its Psi1 rows visible device-wide (one fence for
    // every item found complete) and bump the completion counter of each item's plane.
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
      mbar_wait(&S.empty[sg], par);            // the oldest unpublished item
      int m = 1, s2 = sg + 1;
      unsigned p2 = par;
      if (s2 == NS) { s2 = 0; p2 ^= 1u; }
      while (k + m < issued && mbar_test(&S.empty[s2], p2)) {   // and whatever finished behind it
        m++;
        if (++s2 == NS) { s2 = 0; p2 ^= 1u; }
      }
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      for (int j = 0; j < m; j++) {
        asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + S.sigRing[(k + j) & (PLAN_RING - 1)]]) : "memory");
        const int s2i = RING ? S.sig2Ring[(k + j) & (PLAN_RING - 1)] : -1;
        if (RING && s2i >= 0) asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(&P.counters[1 + s2i]) : "memory");
      }
      k += m; sg = s2; par = p2;
      S.nSignaled = k;
    }
    return;
  }
  }
  if (WG) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PLAN_WG_CONS_REGS));

  // ---------------- consumer warps: engines of wpe warps, each on its own stages ----------------
  const int eng = warp / wpe, elane = (warp - eng * wpe) * 32 + lane;   // my engine, my lane in it
  const int LZ = Gv / NH;                                              // lanes per zone; a lane owns columns li + h LZ
  const int zi = elane / LZ, li = elane - zi * LZ;                     // my zone of the item, my first column in it
  const unsigned hs = (unsigned)LZ * 16u;
  const int hg = 2 * LZ;
  const double tau = P.tau;
  for (int k = eng;; k += NE) {
    const int s = k % NS;
    mbar_wait(&S.full[s], (k / NS) & 1);
    const StageMeta m = S.meta[s];
    if (m.n == -1) break;
    if (RING && m.n == -2) {
      phi_tally_item(P, m.angle, m.upIdx, m.c0, m.c1, elane, 32 * wpe);
    } else if (zi < m.n) {
      unsigned char *st = stages + (size_t)s * P.stageBytes;
      const ZoneRec *R = reinterpret_cast<const ZoneRec *>(st + P.offRecs) + zi;
      const int upIdx = RING ? m.upIdx : m.angle;
      double *upg = P.upBase + (size_t)upIdx * slab + 2 * li;                      // corner rows of this angle's Psi1
      double *pbg = TWO ? P.bBase + (size_t)m.angle * slab + 2 * li : upg;         // boundary-element rows (Set%PsiB(:,:,angle))
      if (R->flags & ZREC_SLOW) {
        for (int h = 0; h < NH; h++) {
          solve_zone_slow(P, m.angle, upIdx, R->zone0, 2 * (li + h * LZ));
          solve_zone_slow(P, m.angle, upIdx, R->zone0, 2 * (li + h * LZ) + 1);
        }
      } else {
        const unsigned col = (unsigned)(zi * MAXC * Gv + li) * 16u;
        V2 sig[NH];
#pragma unroll
        for (int h = 0; h < NH; h++) sig[h] = *reinterpret_cast<const V2 *>(st + P.offSigt + (size_t)(zi * Gv + li + h * LZ) * 16);
        if (R->flags & ZREC_CANON) solve_zone_canon<NH, TWO>(tau, R, upg, pbg, P.ncG, st + col, st + P.offSt + col, sig, hs, hg);
        else solve_zone_plan<NH, TWO>(tau, R, upg, pbg, P.ncG, smem_u32(st) + col, smem_u32(st + P.offSt) + col, sig, hs, hg);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.empty[s]);
  }
}

void fill_params(umt_ctx *ctx, Sweep3DParams &P) {
  P.nc = ctx->nc; P.nb = ctx->nb; P.nz = ctx->nz; P.G = ctx->G; P.NA = ctx->NA; P.nItems = ctx->nItems;
  P.tau = ctx->tau;
  const PlanGeom pg = plan_geom(ctx->G, ctx->plan_nh);
  P.wpe = pg.wpe; P.nEngines = pg.nEngines; P.nStages = pg.nStages; P.stageBytes = pg.stageBytes;
  P.offSt = pg.offSt; P.offSigt = pg.offSigt; P.offRecs = pg.offRecs; P.zpi = pg.zpi;
  P.nStagesWG = 0; P.qbMax = 4;   // tickets per batch: 8 held items too far ahead of their turn (40.5 -> 38.6 ms)
  if (const char *e = getenv("UMT_PLAN_QB")) P.qbMax = std::max(1, std::min(8, atoi(e)));
  P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.nCFaces = ctx->d_nCFaces;
  P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
  P.Volume = ctx->d_Volume; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez; P.omega = ctx->d_omega;
  P.nextZ = ctx->d_nextZ; P.nextC = ctx->d_nextC; P.items = ctx->d_items; P.counters = ctx->d_counters;
  P.psi = ctx->d_psi; P.stotal = ctx->d_stotal; P.sigt = ctx->d_sigt;
  P.upBase = ctx->d_psi1; P.bBase = ctx->psib_buf(); P.ncG = ctx->nc * ctx->G;
  P.weight = ctx->d_weight; P.phi = ctx->d_phi;
  P.recs = ctx->d_recs; P.zinfo = ctx->d_zinfo;
}

}  // namespace

int umt_sweep3d_zones_per_item(const umt_ctx *ctx) {
  if (ctx->use_plan) return plan_geom(ctx->G, ctx->plan_nh).zpi;
  int pairs_target = 512;
  if (const char *e = getenv("UMT_PAIRS_PER_ITEM")) pairs_target = std::max(1, atoi(e));
  return std::max(1, pairs_target / ctx->G);
}

// Build the per-(zone, angle) plan records on the device (once per schedule / geometry / quadrature).
int umt_build_plan3d(umt_ctx *ctx) {
  const size_t n = (size_t)ctx->NA * ctx->nz;
  if (ctx->d_recs) { cudaFree(ctx->d_recs); ctx->d_recs = nullptr; }
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_recs, n * sizeof(ZoneRec)));
  int *d_nslow = nullptr;
  UMT_CUDA(ctx, cudaMalloc((void **)&d_nslow, 2 * sizeof(int)));
  UMT_CUDA(ctx, cudaMemsetAsync(d_nslow, 0, 2 * sizeof(int), ctx->stream));
  PlanBuildParams B;
  B.nc = ctx->nc; B.nb = ctx->nb; B.nz = ctx->nz; B.NA = ctx->NA; B.G = ctx->G;
  B.numCorner = ctx->d_numCorner; B.cOffSet = ctx->d_cOffSet; B.nCFaces = ctx->d_nCFaces; B.cFP = ctx->d_cFP; B.cEZ = ctx->d_cEZ;
  B.Volume = ctx->d_Volume; B.Afp = ctx->d_Afp; B.Aez = ctx->d_Aez; B.omega = ctx->d_omega;
  B.nextZ = ctx->d_nextZ; B.nextC = ctx->d_nextC; B.recs = ctx->d_recs; B.nSlow = d_nslow;
  B.canon = 1;   // UMT_PLAN_CANON=0: every zone through the list-driven path (A/B runs, tests of that path)
  if (const char *e = getenv("UMT_PLAN_CANON")) B.canon = atoi(e) != 0;
  plan_build_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(B);
  UMT_CUDA(ctx, cudaGetLastError());
  int h_n[2] = {0, 0};
  UMT_CUDA(ctx, cudaMemcpyAsync(h_n, d_nslow, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->plan_slow_zones = h_n[0]; ctx->plan_canon_zones = h_n[1];
  cudaFree(d_nslow);
  return UMT_OK;
}

static int launch_plan(umt_ctx *ctx, const Sweep3DParams &P, int mode) {
  const int threads = PLAN_LANES + 64;   // consumers + loader warp + signaller warp
  const size_t smem = plan_geom(ctx->G, ctx->plan_nh).smemBytes;
  void (*kern)(Sweep3DParams);
  if (ctx->plan_nh == 2) kern = mode == 1 ? sweep3d_plan_kernel<2, 1, false> : sweep3d_plan_kernel<2, 0, false>;
  else kern = mode == 1 ? sweep3d_plan_kernel<1, 1, false> : sweep3d_plan_kernel<1, 0, false>;
  // warp-group build: two-warp engines only (G = 128 with two groups per lane), UMT_PLAN_WG=0 switches it off
  // (measured at -d 20 -G 128: 37.4 ms against 38.6 ms for the plain build, both with 4 tickets per batch; 40.5 ms with 8)
  bool wg = ctx->plan_nh == 1 && P.wpe == 2 && PLAN_NCW == 4;
  if (const char *e = getenv("UMT_PLAN_WG")) wg = wg && atoi(e) != 0;
  if (wg) {
    Sweep3DParams Q = P;
    const int perHalf = ((227 * 1024 - 2 * PLAN_CTL_BYTES) / 2) / P.stageBytes;
    // engines + 2 landing stages per half (more hold tickets ahead of their turn: 6 stages 46 ms, 5 stages 37.4, 4 stages 37.8)
    Q.nStagesWG = std::max(PLAN_WG_NCW / P.wpe + 1, std::min(PLAN_WG_NCW / P.wpe + 2, std::min(PLAN_MAX_STAGES, perHalf)));
    if (const char *e = getenv("UMT_PLAN_STAGES")) Q.nStagesWG = std::max(PLAN_WG_NCW / P.wpe + 1, std::min(std::min(PLAN_MAX_STAGES, perHalf), atoi(e)));
    const size_t smemWG = 2 * PLAN_CTL_BYTES + 2 * (size_t)Q.nStagesWG * P.stageBytes;
    void (*kw)(Sweep3DParams) = mode == 1 ? sweep3d_plan_kernel<1, 1, true> : sweep3d_plan_kernel<1, 0, true>;
    UMT_CUDA(ctx, cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemWG));
    const int gridWG = std::max(1, std::min(ctx->sm_count, (P.nItems + 15) / 16));
    kw<<<gridWG, PLAN_WG_THREADS, smemWG, ctx->stream>>>(Q);
    UMT_CUDA(ctx, cudaGetLastError());
    return UMT_OK;
  }
  UMT_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) UMT_FAIL(ctx, UMT_ERR_CUDA, "sweep3d_plan_kernel does not fit on an SM");
  if (const char *e = getenv("UMT_PLAN_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(e)));
  int grid = std::max(1, std::min(ctx->sm_count * occ, P.nItems));
  kern<<<grid, threads, smem, ctx->stream>>>(P);
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}

int umt_launch_sweep3d(umt_ctx *ctx, int savePsi) {
  if (ctx->maxCorner > MAXC || ctx->maxcf != MAXCF)
    UMT_FAIL(ctx, UMT_ERR_ARG, "3-D sweep supports maxCorner <= %d and maxcf == %d (got %d, %d)", MAXC, MAXCF,
             ctx->maxCorner, ctx->maxcf);
  Sweep3DParams P;
  fill_params(ctx, P);
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int) * (1 + ctx->nCounters), ctx->stream));
  if (ctx->use_plan && !ctx->d_recs) UMT_FAIL(ctx, UMT_ERR_STATE, "sweep plan not built");
  if (!ctx->d_psi1) UMT_FAIL(ctx, UMT_ERR_STATE, "Psi1 workspace not allocated");
  if (ctx->single_psi) {
    // single-psi layout (one stage, plan kernel): Set%PsiB is the tail of every Psi slab; a savePsi sweep writes Psi in place,
    // the other sweeps write the Psi1 ring and tally the batches that have to leave it
    P.bBase = ctx->d_psi;
    P.upBase = savePsi ? ctx->d_psi : ctx->d_psi1;
    if (!savePsi && ctx->nTallied > 0) {
      P.items = ctx->d_itemsRing;
      P.nItems = ctx->nItemsRing;
      int r = launch_plan(ctx, P, 1);
      if (r) return r;
      ctx->last_launches += 1;
      return UMT_OK;
    }
  }
  if (P.upBase != P.bBase) UMT_FAIL(ctx, UMT_ERR_STATE, "single-psi layout without a ring item list");
  const int mode = 0;
  // one launch per reflection stage (a single stage unless the domain has reflecting boundaries)
  for (int s = 0; s < ctx->nStages; s++) {
    const int begin = ctx->stageItemBegin[s], end = ctx->stageItemBegin[s + 1];
    if (s > 0) UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int), ctx->stream));   // the ticket; plane counters persist
    int r = UMT_OK;
    if (ctx->have_comm_order && !ctx->shared.empty()) r = umt_exchange_stage(ctx, s);   // SendFlux / RecvFlux of this sweep step
    if (r) return r;
    r = umt_launch_reflect(ctx, s);   // snreflect for the incident angles of this stage
    if (r) return r;
    if (end == begin) continue;       // nothing to sweep at this step here; the neighbours' send/recv above were still matched
    P.items = ctx->d_items + begin;
    P.nItems = end - begin;
    if (ctx->use_plan) {
      r = launch_plan(ctx, P, mode);
      if (r) return r;
    } else {
      int occ = 0;
      UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep3d_generic_kernel, 128, 0));
      if (occ < 1) occ = 1;
      int grid = std::max(1, std::min(ctx->sm_count * occ, P.nItems));
      sweep3d_generic_kernel<<<grid, 128, 0, ctx->stream>>>(P);
      UMT_CUDA(ctx, cudaGetLastError());
    }
    ctx->last_launches += 1;
  }
  return UMT_OK;
}
