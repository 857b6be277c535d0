// Planck (black-body) group integrals and the isotropic initial radiation field.
// The reference integrates the normalised Planck function between group bounds with
// Clark's polylogarithm-based fits (B. A. Clark, J. Comput. Phys. 70 (1987) 311-329:
// eq. 32 for small epsilon, eq. 48 rational fit otherwise) in
// misc/NormalizedBlackBody.cc:28-186, and seeds Psi with wtiso*B_g(Tr) in
// aux/InitTeton.F90:82-118.  Host+device versions here share one formula.
#include <cmath>

#include "umt_internal.h"

namespace {
// fraction of the black-body energy below eps = E/kT
__host__ __device__ inline double planck_cdf(double eps) {
  if (eps < 0.01) {
    const double third_n = 0.05132991127342032;   // (1/3)/(pi^4/15)
    const double eighth_n = 0.01924871672753262;  // (1/8)/(pi^4/15)
    return eps * eps * eps * (third_n - eighth_n * eps);
  }
  const double p1 = 1.2807339766120354, p2 = 0.8578722513311724, p3 = 0.33288908614428098, p4 = 0.079984931563508915,
               p5 = 0.011878558806454416;
  const double q1 = 0.2807339758744, q2 = 0.07713864107538;
  const double num = 1.0 + eps * (p1 + eps * (p2 + eps * (p3 + eps * (p4 + eps * p5))));
  const double den = 1.0 + eps * (q1 + eps * q2);
  return 1.0 - exp(-eps) * num / den;
}

__host__ __device__ inline void planck_groups(double T, double k, double Bnorm, int ng, const double *bounds, double *B, int stride) {
  const double kT = k * T;
  if (!(kT > 0.0)) {
    for (int g = 0; g < ng; g++) B[(size_t)g * stride] = 0.0;
    return;
  }
  const double T2 = T * T, full = Bnorm * T2 * T2;
  double lo = 0.0;   // lowest bound is taken as zero, highest as infinity
  for (int g = 0; g < ng - 1; g++) {
    const double hi = planck_cdf(bounds[g + 1] / kT);
    B[(size_t)g * stride] = full * (hi - lo);
    lo = hi;
  }
  B[(size_t)(ng - 1) * stride] = full * (1.0 - lo);
}

// Psi(g,c,a) = max(wtiso * B_g(Tr(zone(c))), floor) for every angle (InitTeton.F90:95-118)
__global__ void init_psi_kernel(double *psi, const double *trz, const int *c2z, const double *bounds, int G, int nc /* rows per angle */, int NA,
                                double kb, double ac, double wtiso, double efloor) {
  extern __shared__ double sB[];   // (G) spectrum of this corner
  const int c = blockIdx.x;
  if (threadIdx.x == 0) planck_groups(trz[c2z[c]], kb, ac, G, bounds, sB, 1);
  __syncthreads();
  for (int a = 0; a < NA; a++)
    for (int g = threadIdx.x; g < G; g += blockDim.x) psi[((size_t)a * nc + c) * G + g] = fmax(wtiso * sB[g], efloor);
}
}  // namespace

extern "C" int umt_planck_groups(double T, double k, double Bnorm, int numGroups, const double *groupBounds, double *B) {
  if (numGroups < 1 || !groupBounds || !B) return UMT_ERR_ARG;
  planck_groups(T, k, Bnorm, numGroups, groupBounds, B, 1);
  return UMT_OK;
}

extern "C" int umt_init_teton(umt_ctx *ctx, const double *Trz, const double *groupBounds, double speedLight, double radConstant,
                              double wtiso, double efloor) {
  if (!ctx || !Trz || !groupBounds) return UMT_ERR_ARG;
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_init_teton: host-only context");
  if (!ctx->d_psi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_init_teton: call umt_upload_state first (allocates the device state)");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  std::vector<int> c2z(ctx->nc, 0);
  for (int z = 0; z < ctx->nz; z++)
    for (int c = 0; c < ctx->h_numCorner[z]; c++) c2z[ctx->h_cOffSet[z] + c] = z;
  int *d_c2z = nullptr;
  double *d_tr = nullptr, *d_b = nullptr;
  UMT_CUDA(ctx, cudaMalloc((void **)&d_c2z, sizeof(int) * ctx->nc));
  UMT_CUDA(ctx, cudaMalloc((void **)&d_tr, sizeof(double) * ctx->nz));
  UMT_CUDA(ctx, cudaMalloc((void **)&d_b, sizeof(double) * (ctx->G + 1)));
  UMT_CUDA(ctx, umt_memcpy(ctx, d_c2z, c2z.data(), sizeof(int) * ctx->nc, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, d_tr, Trz, sizeof(double) * ctx->nz, cudaMemcpyHostToDevice));
  UMT_CUDA(ctx, umt_memcpy(ctx, d_b, groupBounds, sizeof(double) * (ctx->G + 1), cudaMemcpyHostToDevice));
  init_psi_kernel<<<ctx->nc, 128, sizeof(double) * ctx->G, ctx->stream>>>(ctx->d_psi, d_tr, d_c2z, d_b, ctx->G, ctx->rows, ctx->NA, 1.0,
                                                                           speedLight * radConstant, wtiso, efloor);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(d_c2z); cudaFree(d_tr); cudaFree(d_b);
  if (e != cudaSuccess) UMT_FAIL(ctx, UMT_ERR_CUDA, "init_psi_kernel: %s", cudaGetErrorString(e));
  return UMT_OK;
}
