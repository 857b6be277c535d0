// 3-D upstream-corner-balance sweep for sm_100a.
//
// Replaces snac/SweepUCBxyz.F90:11-322 (one angle, one set) *and* the angle loop
// of snac/SetSweep.F90:113-170: one persistent kernel sweeps every angle of the
// quadrature.  Work is a list of items (angle, hyperplane, chunk of zones) ordered
// plane-major / angle-minor; CTAs pull items through an atomic ticket and wait on a
// per-(angle,plane) completion counter, so the hyperplane dependency of one angle
// is hidden behind the other angles' planes.  Group index is on consecutive
// threads: every load/store of Psi, STotal, Psi1, PsiB is a contiguous G*8-byte row.
//
// Data flow per unknown (corner x angle x group): read STotal, read Psi(n), write
// Psi1; Sigt once per zone; geometry once per (zone, angle).  Upstream Psi1 rows
// were written a few planes earlier by other CTAs and are read through L2
// (ld.global.cg) — L1 is not coherent across SMs.
#include "umt_internal.h"

namespace {

constexpr int MAXC = 8;    // corners per zone handled by this kernel
constexpr int MAXCF = 3;   // corner faces
constexpr double FOURALPHA = 1.82;   // SweepUCBxyz.F90:80

struct Sweep3DParams {
  int nc, nb, nz, G, NA, nItems;
  double tau;
  const int *numCorner, *cOffSet, *nCFaces, *cFP /* 0-based row; >= nc: boundary */, *cEZ /* 0-based */;
  const double *Volume, *Afp, *Aez, *omega;
  const int *nextZ;
  const unsigned char *nextC;
  const WorkItem *items;
  int *counters;
  const double *psi, *stotal, *sigt;
  double *psi1, *psib;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128) sweep3d_generic_kernel(Sweep3DParams P) {
  __shared__ int s_item;
  const int G = P.G, nc = P.nc;
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

    const int a = w.angle;
    const double om0 = P.omega[3 * a], om1 = P.omega[3 * a + 1], om2 = P.omega[3 * a + 2];
    const double *psiA = P.psi + (size_t)a * nc * G;
    double *psi1A = P.psi1 + (size_t)a * nc * G;
    double *psibA = P.psib + (size_t)a * P.nb * G;
    const int *nextZ = P.nextZ + (size_t)a * P.nz;
    const unsigned char *nextC = P.nextC + (size_t)a * nc;

    const int npairs = (w.zend - w.zbeg) * G;
    for (int idx = threadIdx.x; idx < npairs; idx += blockDim.x) {
      const int zi = idx / G, g = idx - zi * G;
      const int zone0 = nextZ[w.zbeg + zi];
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
          afp[f] = om0 * A[0] + om1 * A[1] + om2 * A[2];
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
          const double aez = om0 * A[0] + om1 * A[1] + om2 * A[2];
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
            const double afp = om0 * A[0] + om1 * A[1] + om2 * A[2];
            if (afp > 0.0) psibA[(size_t)(row - nc) * G + g] = src[c];
          }
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&P.counters[1 + w.signal_idx], 1);
    }
  }
}

}  // namespace

int umt_launch_sweep3d(umt_ctx *ctx) {
  if (ctx->maxCorner > MAXC || ctx->maxcf != MAXCF)
    UMT_FAIL(ctx, UMT_ERR_ARG, "3-D sweep supports maxCorner <= %d and maxcf == %d (got %d, %d)", MAXC, MAXCF,
             ctx->maxCorner, ctx->maxcf);
  Sweep3DParams P;
  P.nc = ctx->nc; P.nb = ctx->nb; P.nz = ctx->nz; P.G = ctx->G; P.NA = ctx->NA; P.nItems = ctx->nItems;
  P.tau = ctx->tau;
  P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.nCFaces = ctx->d_nCFaces;
  P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
  P.Volume = ctx->d_Volume; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez; P.omega = ctx->d_omega;
  P.nextZ = ctx->d_nextZ; P.nextC = ctx->d_nextC; P.items = ctx->d_items; P.counters = ctx->d_counters;
  P.psi = ctx->d_psi; P.stotal = ctx->d_stotal; P.sigt = ctx->d_sigt; P.psi1 = ctx->d_psi1; P.psib = ctx->d_psib;
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof(int) * (1 + ctx->nCounters), ctx->stream));
  int occ = 0;
  UMT_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep3d_generic_kernel, 128, 0));
  if (occ < 1) occ = 1;
  int grid = ctx->sm_count * occ;
  if (grid > ctx->nItems) grid = ctx->nItems;
  if (grid < 1) grid = 1;
  sweep3d_generic_kernel<<<grid, 128, 0, ctx->stream>>>(P);
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches += 1;
  return UMT_OK;
}
