// Scattering + emission source build (north-star item 3).  EXTENSION, PARITY UNPINNED: the mini-app reference
// allocates GSet%STotal ("fixed + scat source", mods/GroupSet_mod.F90:22), zeroes it (:74-77) and never writes it
// again — the routines that fill it (UpdateMaterialCoupling etc.) are not in the tree.
// What the tree does fix is every consumer: the sweep reads Q = STotal + tau psi^n (SweepUCBxyz.F90:121), the grey
// source is the re-emitting collision rate sum_g (Eta siga + sigs) phi (rt/getCollisionRate.F90:60-75) redistributed
// with the spectrum GTA%Chi (rt/addGreyCorrections.F90:85-86), and emission enters through Mat%EmissionRate(ngr,ncornr)
// (mods/Material_mod.F90:44).  The isotropic source consistent with those pieces is
//
//   STotal(g,c) = wtiso [ sigs(g,z) phi(g,c) + Chi(g,c) Eta(c) sum_g' siga(g',z) phi(g',c) + EmissionRate(g,c) ]
//
// built here in one pass over PhiTotal: one warp per corner, groups on lanes, the group sum by warp shuffles in a fixed
// order (deterministic).
#include "umt_internal.h"

namespace {

__global__ void __launch_bounds__(256) source_build_kernel(int nc, int G, const int *c2z, const double *siga, const double *sigs,
                                                           const double *eta, const double *chi, const double *emis, const double *phi,
                                                           double wtiso, double *stotal) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nc) return;
  const int c = warp, zone = c2z[c];
  const double *sa = siga + (size_t)zone * G, *ss = sigs + (size_t)zone * G;
  const double *ph = phi + (size_t)c * G;
  double absorbed = 0.0;
  for (int g = lane; g < G; g += 32) absorbed += sa[g] * ph[g];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) absorbed += __shfl_xor_sync(0xffffffffu, absorbed, o);
  const double reemit = eta[c] * absorbed;
  for (int g = lane; g < G; g += 32) {
    const size_t i = (size_t)c * G + g;
    stotal[i] = wtiso * (ss[g] * ph[g] + chi[i] * reemit + (emis ? emis[i] : 0.0));
  }
}

}  // namespace

extern "C" int umt_build_source(umt_ctx *ctx, const double *Siga, const double *Sigs, const double *Eta, const double *Chi,
                                const double *EmissionRate, double *STotalOut) {
  if (!ctx || !Siga || !Sigs || !Eta || !Chi) return UMT_ERR_ARG;
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_build_source: host-only context (device -1) cannot run kernels");
  if (!ctx->d_phi || !ctx->d_stotal) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_build_source: no PhiTotal / STotal on the device");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nc = ctx->nc, nz = ctx->nz, G = ctx->G;
  std::vector<int> c2z(nc);
  for (int z = 0; z < nz; z++)
    for (int c = 0; c < ctx->h_numCorner[z]; c++) c2z[ctx->h_cOffSet[z] + c] = z;
  int *d_c2z = nullptr;
  double *d_a = nullptr, *d_s = nullptr, *d_e = nullptr, *d_chi = nullptr, *d_em = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_c2z, sizeof(int) * nc);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_a, sizeof(double) * (size_t)nz * G);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_s, sizeof(double) * (size_t)nz * G);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_e, sizeof(double) * nc);
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_chi, sizeof(double) * (size_t)nc * G);
  if (e == cudaSuccess && EmissionRate) e = cudaMalloc((void **)&d_em, sizeof(double) * (size_t)nc * G);
  if (e == cudaSuccess) {
    umt_memcpy(ctx, d_c2z, c2z.data(), sizeof(int) * nc, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_a, Siga, sizeof(double) * (size_t)nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_s, Sigs, sizeof(double) * (size_t)nz * G, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_e, Eta, sizeof(double) * nc, cudaMemcpyHostToDevice);
    umt_memcpy(ctx, d_chi, Chi, sizeof(double) * (size_t)nc * G, cudaMemcpyHostToDevice);
    if (EmissionRate) umt_memcpy(ctx, d_em, EmissionRate, sizeof(double) * (size_t)nc * G, cudaMemcpyHostToDevice);
    const double wtiso = ctx->ndim == 3 ? 1.0 / (4.0 * 3.14159265358979323846) : 1.0 / (2.0 * 3.14159265358979323846);
    const unsigned blocks = (unsigned)(((size_t)nc * 32 + 255) / 256);
    source_build_kernel<<<blocks, 256, 0, ctx->stream>>>(nc, G, d_c2z, d_a, d_s, d_e, d_chi, d_em, ctx->d_phi, wtiso, ctx->d_stotal);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess && STotalOut) e = umt_memcpy(ctx, STotalOut, ctx->d_stotal, sizeof(double) * (size_t)nc * G, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_c2z); cudaFree(d_a); cudaFree(d_s); cudaFree(d_e); cudaFree(d_chi); cudaFree(d_em);
  if (e != cudaSuccess) UMT_FAIL(ctx, UMT_ERR_CUDA, "umt_build_source: %s", cudaGetErrorString(e));
  return UMT_OK;
}
