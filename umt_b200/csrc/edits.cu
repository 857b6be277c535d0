// End-of-cycle edits on the device-resident fields, so a cycle never downloads psi:
//   aux/rtedit.F90:142-232            EnergyRadiation = sum_zones sum_c V_c sum_g PhiTotal(g,c) / c,
//                                     trz(zone) = (max(ERad / (VolumeZone a c), tr4floor))^(1/4), TrMax
//   control/initializeZones.F90:25-50 Rad%radEnergy(zone), EnergyRadBOC (same sum at the start of the cycle)
//   control/BoundaryEdit.F90:55-150   RadPowerEscape(g) = sum over vacuum boundary elements and weighted angles with
//                                     omega.A_bdy > 0 of w (omega.A_bdy) Psi(g, BdyToC(b), angle)   (no mesh motion: lambdaD = 1);
//                                     r-z (:123): times geometryFactor = 2 pi and the radius of the boundary element
//   control/setEnergyDensity.F90      RadEnergyDensity(zone,g) = sum_c (V_c / VolumeZone) PhiTotal(g,c) / c
// Sums are two-stage and ordered (deterministic).
#include <algorithm>
#include <cmath>

#include "umt_internal.h"

namespace {

// one warp per zone: ERad(zone) = sum_c V_c sum_g Phi(g,c)
__global__ void __launch_bounds__(256) zone_energy_kernel(int nz, int G, const int *numCorner, const int *cOffSet, const double *Volume,
                                                          const double *phi, double *erad, double *volZone) {
  const int zone = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (zone >= nz) return;
  double e = 0.0, vz = 0.0;
  for (int c = cOffSet[zone]; c < cOffSet[zone] + numCorner[zone]; c++) {
    double s = 0.0;
    for (int g = lane; g < G; g += 32) s += phi[(size_t)c * G + g];
    e += Volume[c] * s;
    vz += Volume[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
  if (lane == 0) { erad[zone] = e; volZone[zone] = vz; }
}

__global__ void __launch_bounds__(256) sum_max_partial_kernel(const double *erad, const double *volZone, int nz, double ac, double tr4floor,
                                                              double *trz, double *partial) {
  __shared__ double r0[256], r1[256];
  double s = 0.0, m = 0.0;
  for (int z = blockIdx.x * blockDim.x + threadIdx.x; z < nz; z += gridDim.x * blockDim.x) {
    s += erad[z];
    const double t = sqrt(sqrt(fmax(erad[z] / (volZone[z] * ac), tr4floor)));
    if (trz) trz[z] = t;
    m = fmax(m, t);
  }
  r0[threadIdx.x] = s; r1[threadIdx.x] = m;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) { r0[threadIdx.x] += r0[threadIdx.x + k]; r1[threadIdx.x] = fmax(r1[threadIdx.x], r1[threadIdx.x + k]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[blockIdx.x] = r0[0]; partial[gridDim.x + blockIdx.x] = r1[0]; }
}
__global__ void sum_max_finish_kernel(const double *partial, int nb, double *out2) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0, m = 0.0;
    for (int i = 0; i < nb; i++) { s += partial[i]; m = fmax(m, partial[nb + i]); }
    out2[0] = s; out2[1] = m;
  }
}

// escape currents: one CTA per chunk of exit entries, threads over groups
__global__ void __launch_bounds__(128) escape_partial_kernel(const double *psi, const int *ec, const int *ea, const double *coef, int nExit, int perChunk,
                                                             int rows, int G, double *partial /* (nChunks, G) */) {
  const int beg = blockIdx.x * perChunk, end = min(nExit, beg + perChunk);
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double s = 0.0;
    for (int i = beg; i < end; i++) s += coef[i] * psi[((size_t)ea[i] * rows + ec[i]) * G + g];
    partial[(size_t)blockIdx.x * G + g] = s;
  }
}
__global__ void escape_finish_kernel(const double *partial, int nChunks, int G, double *out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double s = 0.0;
  for (int k = 0; k < nChunks; k++) s += partial[(size_t)k * G + g];
  out[g] = s;
}

// RadEnergyDensity(zone, g) = sum_c V_c / VolumeZone * Phi(g,c) / c   (Fortran shape (nzones, ngr))
__global__ void energy_density_kernel(int nz, int G, const int *numCorner, const int *cOffSet, const double *Volume, const double *phi,
                                      double invc, double *dens) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nz * G) return;
  const int zone = (int)(i / G), g = (int)(i % G);
  double vz = 0.0;
  for (int c = cOffSet[zone]; c < cOffSet[zone] + numCorner[zone]; c++) vz += Volume[c];
  double s = 0.0;
  for (int c = cOffSet[zone]; c < cOffSet[zone] + numCorner[zone]; c++) s = s + (invc * Volume[c] / vz) * phi[(size_t)c * G + g];
  dens[(size_t)g * nz + zone] = s;
}

}  // namespace

int umt_finalize_schedule(umt_ctx *ctx);   // umt_api.cu

// out5 = {EnergyRadiation, TrMax, PowerEscape (sum over groups), PowerIncident (0: vacuum / shared / reflecting boundaries only), sum of ERad};
// optional arrays: trz(nzones), RadPowerEscape(ngr), RadEnergyDensity(nzones, ngr)
extern "C" int umt_cycle_edits(umt_ctx *ctx, double speedLight, double radConstant, double tr4floor, double *out5, double *trz,
                               double *RadPowerEscape, double *RadEnergyDensity) {
  if (!ctx || !out5) return UMT_ERR_ARG;
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_cycle_edits: host-only context (device -1) cannot run kernels");
  if (!ctx->d_phi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_cycle_edits: no state on the device");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  int r = umt_finalize_schedule(ctx);
  if (r) return r;
  const int nz = ctx->nz, G = ctx->G, nd = ctx->ndim, NB = 296;
  const double geometryFactor = nd == 2 ? 2.0 * 3.14159265358979323846 : 1.0;   // rtedit.F90:85-93
  double *d_erad = nullptr, *d_vz = nullptr, *d_trz = nullptr, *d_part = nullptr, *d_out = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_erad, sizeof(double) * nz);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_vz, sizeof(double) * nz);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_trz, sizeof(double) * nz);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_part, sizeof(double) * 2 * NB);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, sizeof(double) * (2 + G));
  double h2[2] = {0, 0};
  if (e == cudaSuccess) {
    zone_energy_kernel<<<(unsigned)(((size_t)nz * 32 + 255) / 256), 256, 0, ctx->stream>>>(nz, G, ctx->d_numCorner, ctx->d_cOffSet, ctx->d_Volume, ctx->d_phi, d_erad, d_vz);
    sum_max_partial_kernel<<<NB, 256, 0, ctx->stream>>>(d_erad, d_vz, nz, radConstant * speedLight, tr4floor, d_trz, d_part);
    sum_max_finish_kernel<<<1, 32, 0, ctx->stream>>>(d_part, NB, d_out);
    e = cudaMemcpyAsync(h2, d_out, sizeof(double) * 2, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && trz) e = umt_memcpy(ctx, trz, d_trz, sizeof(double) * nz, cudaMemcpyDeviceToHost);
  }
  // escape through vacuum boundary elements (neither shared nor reflecting), weighted angles only
  double escape = 0.0;
  std::vector<double> hEsc(G, 0.0);
  if (e == cudaSuccess && ctx->nExit > 0) {
    std::vector<unsigned char> notVac(std::max(ctx->nb, 1), 0);
    for (const auto &s : ctx->shared) for (int b = s.first; b < s.first + s.n; b++) notVac[b] = 1;
    for (const auto &s : ctx->refl) for (int b = s.first; b < s.first + s.n; b++) notVac[b] = 1;
    if (!ctx->have_abdy) {   // boundary-element area vectors = A_fp of the corner face they sit on
      ctx->h_Abdy.assign((size_t)nd * std::max(ctx->nb, 1), 0.0);
      for (int c = 0; c < ctx->nc; c++)
        for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
          const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
          if (v > ctx->nc)
            for (int d = 0; d < nd; d++) ctx->h_Abdy[(size_t)(v - ctx->nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * ctx->maxcf + f) * nd + d];
        }
      ctx->have_abdy = true;
    }
    std::vector<int> ec, ea;
    std::vector<double> coef;
    for (int a = 0; a < ctx->NA; a++) {
      if (!(ctx->h_weight[a] > 0.0)) continue;
      const auto &bl = ctx->bdyList[a];
      for (size_t i = 0; i + 1 < bl.size(); i += 2) {
        const int b = bl[i] - 1, c = bl[i + 1] - 1;
        if (notVac[b]) continue;
        double dot = 0.0;
        for (int d = 0; d < nd; d++) dot += ctx->h_omega[(size_t)a * nd + d] * ctx->h_Abdy[(size_t)b * nd + d];
        double factor = ctx->h_weight[a] * geometryFactor;
        if (nd == 2) {
          // BoundaryEdit.F90:123: factor = weight * geometryFactor * BdyT%Radius(b), and Radius(b) is the RadiusFP of the corner face
          // the boundary element sits on (volumeUCBrz.F90:102-113 and geometryUCBrz.F90:84-85 are the same expressions)
          if (ctx->h_RadiusFP.size() != 2 * (size_t)ctx->nc) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_cycle_edits: r-z boundary edit needs RadiusFP (umt_set_geometry / umt_compute_geometry)");
          int f = -1;
          for (int k = 0; k < ctx->h_nCFaces[c]; k++) if (ctx->h_cFP[(size_t)c * ctx->maxcf + k] == ctx->nc + b + 1) f = k;
          if (f < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_cycle_edits: boundary element %d is not a face of corner %d", b + 1, c + 1);
          factor *= ctx->h_RadiusFP[(size_t)c * 2 + f];
        }
        ec.push_back(c); ea.push_back(a); coef.push_back(factor * dot);
      }
    }
    const int nE = (int)ec.size(), perChunk = 64, nChunks = (nE + perChunk - 1) / perChunk;
    if (nE > 0) {
      int *d_ec = nullptr, *d_ea = nullptr;
      double *d_coef = nullptr, *d_p = nullptr;
      e = cudaMalloc((void **)&d_ec, sizeof(int) * nE);
      if (e == cudaSuccess) e = cudaMalloc((void **)&d_ea, sizeof(int) * nE);
      if (e == cudaSuccess) e = cudaMalloc((void **)&d_coef, sizeof(double) * nE);
      if (e == cudaSuccess) e = cudaMalloc((void **)&d_p, sizeof(double) * (size_t)nChunks * G);
      if (e == cudaSuccess) {
        umt_memcpy(ctx, d_ec, ec.data(), sizeof(int) * nE, cudaMemcpyHostToDevice);
        umt_memcpy(ctx, d_ea, ea.data(), sizeof(int) * nE, cudaMemcpyHostToDevice);
        umt_memcpy(ctx, d_coef, coef.data(), sizeof(double) * nE, cudaMemcpyHostToDevice);
        escape_partial_kernel<<<nChunks, 128, 0, ctx->stream>>>(ctx->d_psi, d_ec, d_ea, d_coef, nE, perChunk, ctx->rows, G, d_p);
        escape_finish_kernel<<<(G + 127) / 128, 128, 0, ctx->stream>>>(d_p, nChunks, G, d_out + 2);
        e = cudaMemcpyAsync(hEsc.data(), d_out + 2, sizeof(double) * G, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
      }
      cudaFree(d_ec); cudaFree(d_ea); cudaFree(d_coef); cudaFree(d_p);
    }
    for (int g = 0; g < G; g++) escape += hEsc[g];   // rtedit.F90:212 sum(RadPowerEscape)
  }
  if (e == cudaSuccess && RadEnergyDensity) {
    double *d_dens = nullptr;
    e = cudaMalloc((void **)&d_dens, sizeof(double) * (size_t)nz * G);
    if (e == cudaSuccess) {
      const size_t n = (size_t)nz * G;
      energy_density_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(nz, G, ctx->d_numCorner, ctx->d_cOffSet, ctx->d_Volume, ctx->d_phi, geometryFactor / speedLight, d_dens);
      e = cudaStreamSynchronize(ctx->stream);
      if (e == cudaSuccess) e = umt_memcpy(ctx, RadEnergyDensity, d_dens, sizeof(double) * n, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_dens);
  }
  cudaFree(d_erad); cudaFree(d_vz); cudaFree(d_trz); cudaFree(d_part); cudaFree(d_out);
  if (e != cudaSuccess) UMT_FAIL(ctx, UMT_ERR_CUDA, "umt_cycle_edits: %s", cudaGetErrorString(e));
  out5[0] = geometryFactor * h2[0] / speedLight;
  out5[1] = h2[1];
  out5[2] = escape;
  out5[3] = 0.0;
  out5[4] = h2[0];
  if (RadPowerEscape) std::copy(hEsc.begin(), hEsc.end(), RadPowerEscape);
  return UMT_OK;
}
