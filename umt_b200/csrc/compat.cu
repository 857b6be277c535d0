// The reference's own C seam (gpu/GPU_SweepUCBxyz.cu:520-572) forwarded to the context API: see include/teton_gpu_compat.h.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "umt_internal.h"
#include "../../include/teton_gpu_compat.h"

int umt_finalize_schedule(umt_ctx *ctx);   // umt_api.cu

namespace {
constexpr int kMaxStreams = 80;            // MAX_CUDA_STREAMS, GPU_SweepUCBxyz.cu:34
struct Slot {
  umt_ctx *ctx = nullptr;
  int nz = 0, nc = 0, nb = 0, G = 0, maxcf = 0, maxCorner = 0;
  std::vector<double> phi;
};
// The reference's seam has no context argument (its own shim keeps static device buffers per stream id, GPU_SweepUCBxyz.cu:650-758),
// so this table is the one piece of process-wide state in the library; the context API (include/umt_sweep.h) has none.
Slot g_slot[kMaxStreams];
std::mutex g_mu;                           // the reference's caller holds an omp critical around the call (SweepUCBxyzToGPU.F90:193)

[[noreturn]] void fatal(umt_ctx *ctx, const char *what, int rc) {
  std::fprintf(stderr, "gpu_sweepucbxyz (libumtsweep): %s failed with status %d: %s\n", what, rc, umt_last_error(ctx));
  std::exit(EXIT_FAILURE);                 // as CUDA_SAFE_CALL does, GPU_SweepUCBxyz.cu:18-27
}
#define MUST(ctx, call) do { int rc_ = (call); if (rc_ != UMT_OK) fatal(ctx, #call, rc_); } while (0)
#define MUST_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    std::fprintf(stderr, "gpu_sweepucbxyz (libumtsweep): %s: %s\n", #call, cudaGetErrorString(e_)); std::exit(EXIT_FAILURE); } } while (0)
}  // namespace

extern "C" void gpu_sweepucbxyz(int *Angle, int *nHyperPlanes, int *nZonesInPlane, int *nextZ, int *nextC, double *STotal,
                                double *tau, double *Psi, int *Groups, double *Volume, double *Sigt, int *nCFacesArray,
                                int *ndim, int *maxcf, int *ncorner, double *A_fp, double *omega, int *cFP, double *Psi1,
                                int *nbelem, double *A_ez, int *cEZ, int *NumAngles, double *quadwt, double *Phi, double *PsiB,
                                int *maxCorner, int *mem0solve1, int *streamIdPtr, int *totalStreams, int *savePsi,
                                int *numCycles, int *cycleOffSet, double *cyclePsi, int *cycleList, int *b0, int *nBdyElem,
                                double *PsiBMref, int *Mref, int *Geom_numCorner, int *Geom_cOffSet) {
  (void)Angle; (void)NumAngles; (void)mem0solve1; (void)totalStreams; (void)Mref;
  std::lock_guard<std::mutex> lock(g_mu);
  const int G = *Groups, nc = *ncorner, nb = *nbelem;
  int nz = 0;
  for (int i = 0; i < *nHyperPlanes; i++) nz += nZonesInPlane[i];       // totalZones, GPU_SweepUCBxyz.cu:624-628
  Slot &s = g_slot[((*streamIdPtr % kMaxStreams) + kMaxStreams) % kMaxStreams];
  if (!s.ctx || s.nz != nz || s.nc != nc || s.nb != nb || s.G != G || s.maxcf != *maxcf || s.maxCorner != *maxCorner) {
    if (s.ctx) umt_ctx_destroy(s.ctx);
    s.ctx = nullptr;
    int dev = 0;
    MUST_CUDA(nullptr, cudaGetDevice(&dev));
    MUST(nullptr, umt_ctx_create(dev, *ndim, nz, nc, nb, *maxcf, *maxCorner, G, &s.ctx));
    s.ctx->force_legacy = true;   // the caller owns Psi1 and PsiB and sees them after every call: keep them in one full workspace
    s.nz = nz; s.nc = nc; s.nb = nb; s.G = G; s.maxcf = *maxcf; s.maxCorner = *maxCorner;
    s.phi.assign((size_t)G * nc, 0.0);
  }
  umt_ctx *ctx = s.ctx;
  // like the reference, everything the caller owns is taken afresh on every call (the mesh may have moved, the
  // opacities and sources change every sweep)
  MUST(ctx, umt_set_connectivity(ctx, Geom_numCorner, Geom_cOffSet, nCFacesArray, cFP, cEZ, 0, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, nullptr));
  MUST(ctx, umt_set_geometry(ctx, Volume, A_fp, A_ez, nullptr, nullptr, nullptr, nullptr));
  MUST(ctx, umt_set_quadrature(ctx, 1, omega, quadwt, nullptr, nullptr, nullptr, nullptr, nullptr));
  MUST(ctx, umt_set_schedule(ctx, 1, *nHyperPlanes, nZonesInPlane, nextZ, nextC, *numCycles, cycleList + *cycleOffSet, 0, nullptr));
  // snreflect for the one reflecting boundary the reference's shim knows (GPU_SweepUCBxyz.cu:796-807)
  if (*nBdyElem > 0 && PsiBMref != PsiB)
    std::memcpy(PsiB + (size_t)G * *b0, PsiBMref + (size_t)G * *b0, sizeof(double) * G * (size_t)*nBdyElem);
  MUST(ctx, umt_upload_state(ctx, Psi, PsiB, Sigt, STotal, *tau));
  MUST(ctx, umt_finalize_schedule(ctx));
  // previous Psi1 (read by the "direct solve" zones, :479-496) and this angle's cyclePsi rows (initFromCycleList, :858)
  MUST_CUDA(ctx, umt_memcpy(ctx, ctx->d_psi1, Psi1, sizeof(double) * G * (size_t)nc, cudaMemcpyHostToDevice));
  if (*numCycles > 0)
    MUST_CUDA(ctx, umt_memcpy(ctx, ctx->d_cyclePsi, cyclePsi + (size_t)G * *cycleOffSet, sizeof(double) * G * (size_t)*numCycles,
                              cudaMemcpyHostToDevice));
  int iters = 0;
  MUST(ctx, umt_sweep(ctx, 0, 1, 0.0, &iters));
  // results back into the caller's arrays (:966-993)
  MUST_CUDA(ctx, umt_memcpy(ctx, Psi1, ctx->d_psi1, sizeof(double) * G * (size_t)nc, cudaMemcpyDeviceToHost));
  if (nb > 0) MUST_CUDA(ctx, umt_memcpy(ctx, PsiB, ctx->d_psi1 + (size_t)G * nc, sizeof(double) * G * (size_t)nb, cudaMemcpyDeviceToHost));
  MUST(ctx, umt_download_phi(ctx, s.phi.data()));                 // quadwt * Psi1 of this angle
  for (size_t i = 0, n = (size_t)G * nc; i < n; i++) Phi[i] += s.phi[i];   // Set%Phi += quadwt*Psi1 (:465)
  if (*numCycles > 0)
    MUST_CUDA(ctx, umt_memcpy(ctx, cyclePsi + (size_t)G * *cycleOffSet, ctx->d_cyclePsi, sizeof(double) * G * (size_t)*numCycles,
                              cudaMemcpyDeviceToHost));
  if (*savePsi == 1) std::memcpy(Psi, Psi1, sizeof(double) * G * (size_t)nc);   // (:974-978)
}

extern "C" void gpu_streamsynchronize(int *streamId) {
  std::lock_guard<std::mutex> lock(g_mu);
  Slot &s = g_slot[((*streamId % kMaxStreams) + kMaxStreams) % kMaxStreams];
  if (s.ctx) umt_synchronize(s.ctx);
}

extern "C" void gpu_devicesynchronize(void) {
  MUST_CUDA(nullptr, cudaDeviceSynchronize());
}
