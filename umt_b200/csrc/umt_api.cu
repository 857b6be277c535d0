// Host side of the C ABI (include/umt_sweep.h): context, residency of Teton's
// arrays in HBM, schedule -> work-item translation, the SetSweep/ControlSweep
// controller, and the small streaming kernels around the sweep.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "umt_internal.h"

// error text of a failed umt_ctx_create on this thread (there is no context to hold it yet)
static thread_local std::string g_create_error;

extern "C" const char *umt_version(void) { return "umt_b200 0.1.0 (sm_100a)"; }

extern "C" const char *umt_last_error(const umt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

#define NEED_DEVICE(ctx, what) do { if ((ctx)->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, what ": host-only context (device -1) cannot run kernels"); } while (0)

template <class T>
static int dev_alloc_copy(umt_ctx *ctx, T **dptr, const T *h, size_t n) {
  if (ctx->device < 0) return UMT_OK;
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  if (n == 0) n = 1;
  UMT_CUDA(ctx, cudaMalloc((void **)dptr, n * sizeof(T)));
  if (h) UMT_CUDA(ctx, umt_memcpy(ctx, *dptr, h, n * sizeof(T), cudaMemcpyHostToDevice));
  return UMT_OK;
}
#define TRY(x) do { int _r = (x); if (_r) return _r; } while (0)

extern "C" int umt_ctx_create(int device, int ndim, int nzones, int ncornr, int nbelem, int maxcf, int maxCorner,
                              int ngr, umt_ctx **out) {
  if (!out) return UMT_ERR_ARG;
  *out = nullptr;
  if (ndim < 2 || ndim > 3 || nzones < 1 || ncornr < 1 || nbelem < 0 || ngr < 1 || maxcf != ndim || maxCorner < 1) {
    g_create_error = "umt_ctx_create: bad sizes";
    return UMT_ERR_ARG;
  }
  if (device == -1) {   // host-only context: schedule / quadrature construction, no kernels
    umt_ctx *hc = new umt_ctx;
    hc->device = -1; hc->ndim = ndim; hc->nz = nzones; hc->nc = ncornr; hc->nb = nbelem;
    hc->maxcf = maxcf; hc->maxCorner = maxCorner; hc->G = ngr;
    *out = hc;
    return UMT_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    g_create_error = std::string("umt_ctx_create: no usable CUDA device (") + cudaGetErrorString(e) + ")";
    return UMT_ERR_CUDA;
  }
  umt_ctx *ctx = new umt_ctx;
  ctx->device = device; ctx->ndim = ndim; ctx->nz = nzones; ctx->nc = ncornr; ctx->nb = nbelem;
  ctx->maxcf = maxcf; ctx->maxCorner = maxCorner; ctx->G = ngr;
  e = cudaSetDevice(device);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  // the exchange stream outranks the main stream: its few CTAs (pack kernel, NCCL send/recv) must be placed as soon as a CTA of the
  // phi tally retires, not after the tally's whole grid has been dispatched (measured at 2 GPUs: the 629 MB of rows crawled at 63 GB/s
  // behind the tally and ended 2 ms after it)
  int prLo = 0, prHi = 0;
  if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
  if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->stream3, cudaStreamNonBlocking, getenv("UMT_EXCHANGE_PRIORITY") && !atoi(getenv("UMT_EXCHANGE_PRIORITY")) ? prLo : prHi);
  for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->ev[i]);
  for (int i = 0; i < 3 && e == cudaSuccess; i++) e = cudaEventCreate(&ctx->evx[i]);
  if (e != cudaSuccess) {
    g_create_error = std::string("umt_ctx_create: ") + cudaGetErrorString(e);
    delete ctx;
    return UMT_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaMalloc((void **)&ctx->d_abort, sizeof(int)) != cudaSuccess || cudaMemset(ctx->d_abort, 0, sizeof(int)) != cudaSuccess ||
      cudaHostAlloc((void **)&ctx->h_abort, sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
    g_create_error = "umt_ctx_create: watchdog flag allocation failed";
    delete ctx;
    return UMT_ERR_CUDA;
  }
  *ctx->h_abort = 0;
  if (const char *ev = getenv("UMT_SPIN_LIMIT")) ctx->spinLimit = (unsigned)std::max(256, atoi(ev));
  // L2 set-aside for evict_last lines: the sweep stores Psi1 rows with an evict_last hint so that the downstream zones find
  // them in L2; without a persisting carve-out the hint has nothing to hold on to.  UMT_L2_PERSIST_MB overrides (0 = off).
  {
    // 64 MB of the 126 MB: more slows the streaming kernels.  Set once: toggling the limit around every sweep gave the phi kernel
    // its 0.4 ms back but made bench.py hang under ncu (cudaDeviceSetLimit between profiled kernels)
    size_t want = std::min((size_t)prop.persistingL2CacheMaxSize, (size_t)64 << 20);
    if (const char *ev = getenv("UMT_L2_PERSIST_MB")) want = std::min(want, (size_t)std::max(0, atoi(ev)) << 20);
    if (ndim != 3) want = 0;   // only the 3-D plan kernel uses evict_last; the r-z kernels lose 7 % to a smaller normal L2
    if (want > 0) {
      size_t before = 0;
      if (cudaDeviceGetLimit(&before, cudaLimitPersistingL2CacheSize) != cudaSuccess) { cudaGetLastError(); before = 0; }
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();
      else { ctx->l2_persist = true; ctx->l2_persist_bytes = want; ctx->l2_persist_before = before; }
    }
    if (getenv("UMT_VERBOSE")) fprintf(stderr, "umt: persisting L2 max %d MB, set-aside %zu MB, L2 %d MB\n", prop.persistingL2CacheMaxSize >> 20, want >> 20, prop.l2CacheSize >> 20);
  }
  *out = ctx;
  return UMT_OK;
}

extern "C" int umt_ctx_destroy(umt_ctx *ctx) {
  if (!ctx) return UMT_OK;
  if (ctx->device < 0) { delete ctx; return UMT_OK; }
  cudaSetDevice(ctx->device);
  void *ptrs[] = {ctx->d_numCorner, ctx->d_cOffSet, ctx->d_nCFaces, ctx->d_cFP, ctx->d_cEZ, ctx->d_Volume, ctx->d_Afp,
                  ctx->d_Aez, ctx->d_Area, ctx->d_RadiusFP, ctx->d_RadiusEZ, ctx->d_omega, ctx->d_weight, ctx->d_nextZ,
                  ctx->d_nextC, ctx->d_items, ctx->d_counters, ctx->d_cycleList, ctx->d_cycleAngle, ctx->d_cyclePsi,
                  ctx->d_exitB, ctx->d_exitC, ctx->d_exitA, ctx->d_psi, ctx->d_psi1, ctx->d_stotal,
                  ctx->d_sigt, ctx->d_phi, ctx->d_psim, ctx->d_recs, ctx->d_zinfo, ctx->d_angDerivFac, ctx->d_tauW1, ctx->d_tauW2,
                  ctx->d_start, ctx->d_finishNext, ctx->d_level, ctx->d_reflOps, ctx->d_rzLevelAngles, ctx->d_rzPlaneOff, ctx->d_rzNHyp,
                  ctx->d_rzPrev, ctx->d_rzPsimA, ctx->d_rzRecs, ctx->d_rzBad, ctx->d_rzSteps, ctx->d_rzNSteps, ctx->d_itemsRing, ctx->d_tailSlot, ctx->d_tailW};
  for (void *p : ptrs) if (p) cudaFree(p);
  if (ctx->d_abort) cudaFree(ctx->d_abort);
  if (ctx->h_abort) cudaFreeHost(ctx->h_abort);
  for (auto &b : ctx->host_blocks) {   // umt_host_alloc blocks never freed
    if (b.second == 0) cudaFreeHost(b.first);
    else { cudaHostUnregister(b.first); munmap(b.first, b.second); }
  }
  ctx->host_blocks.clear();
  umt_exchange_release(ctx);
  umt_gta_release(ctx);
  // hand the device-wide L2 set-aside back (the last 3-D context to go restores what it found)
  if (ctx->l2_persist && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, ctx->l2_persist_before) != cudaSuccess) cudaGetLastError();
  for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 3; i++) if (ctx->evx[i]) cudaEventDestroy(ctx->evx[i]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream3) cudaStreamDestroy(ctx->stream3);
  delete ctx;
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// host staging buffers next to the GPU
// ---------------------------------------------------------------------------
// Page-locked host memory whose pages sit on the NUMA node of the context's GPU (the node of its PCIe root, from sysfs), for the
// arrays a caller hands to umt_control_sweep / umt_upload_* / umt_download_* every sweep.  With one rank per GPU on a two-socket
// box, buffers pinned wherever the rank happened to run cross the socket interconnect on every copy (measured in round 1: 8 ranks
// moved 3.3 GB per step each at 15 GB/s instead of 55).  The pages are placed by first touch under a preferred-node memory policy
// and then registered with CUDA; without NUMA information this is a plain page-locked allocation.

namespace {
int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { cudaGetLastError(); return -1; }
  for (char *c = bus; *c; c++) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
}  // namespace

extern "C" int umt_host_alloc(umt_ctx *ctx, size_t bytes, void **ptr, int *numaNode) {
  if (!ctx || !ptr || bytes == 0) return UMT_ERR_ARG;
  NEED_DEVICE(ctx, "umt_host_alloc");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  *ptr = nullptr;
  const long page = sysconf(_SC_PAGESIZE);
  const size_t len = (bytes + page - 1) / page * page;
  void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_host_alloc: mmap of %zu bytes failed", len);
  const int node = gpu_numa_node(ctx->device);
  bool bound = false;
#ifdef SYS_mbind
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
    bound = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, mask, 8 * sizeof(mask), 0) == 0;
  }
#endif
  for (size_t o = 0; o < len; o += page) static_cast<volatile char *>(p)[o] = 0;   // first touch: the pages exist where the policy says
  cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    // registering foreign pages is refused in some environments (under a profiler, in some containers): plain page-locked memory then
    cudaGetLastError();
    munmap(p, len);
    e = cudaHostAlloc(&p, len, cudaHostAllocPortable);
    if (e != cudaSuccess) UMT_FAIL(ctx, UMT_ERR_CUDA, "umt_host_alloc: %zu page-locked bytes: %s", len, cudaGetErrorString(e));
    ctx->host_blocks[p] = 0;   // length 0: a cudaHostAlloc block
    if (numaNode) *numaNode = -1;
    *ptr = p;
    return UMT_OK;
  }
  ctx->host_blocks[p] = len;
  if (numaNode) *numaNode = bound ? node : -1;
  *ptr = p;
  return UMT_OK;
}

extern "C" int umt_host_free(umt_ctx *ctx, void *ptr) {
  if (!ptr) return UMT_OK;
  if (!ctx) return UMT_ERR_ARG;
  auto it = ctx->host_blocks.find(ptr);
  if (it == ctx->host_blocks.end()) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_host_free: not a block of this context's umt_host_alloc");
  const size_t len = it->second;
  ctx->host_blocks.erase(it);
  if (len == 0) { cudaFreeHost(ptr); return UMT_OK; }
  cudaHostUnregister(ptr);
  munmap(ptr, len);
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// connectivity / geometry / quadrature
// ---------------------------------------------------------------------------
extern "C" int umt_set_connectivity(umt_ctx *ctx, const int *numCorner, const int *cOffSet, const int *nCFaces,
                                    const int *cFP, const int *cEZ, int maxFaces, const int *zoneFaces,
                                    const int *zoneOpp, const int *faceOpp, const int *CToFace,
                                    const unsigned char *BoundaryZone, const int *BdyToC) {
  if (!ctx || !numCorner || !cOffSet || !nCFaces || !cFP || !cEZ) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nz = ctx->nz, nc = ctx->nc, mcf = ctx->maxcf;
  ctx->h_numCorner.assign(numCorner, numCorner + nz);
  ctx->h_cOffSet.assign(cOffSet, cOffSet + nz);
  ctx->h_nCFaces.assign(nCFaces, nCFaces + nc);
  ctx->h_cFP.assign(cFP, cFP + (size_t)mcf * nc);
  ctx->h_cEZ.assign(cEZ, cEZ + (size_t)mcf * nc);
  for (int z = 0; z < nz; z++) {
    if (numCorner[z] < 1 || numCorner[z] > ctx->maxCorner || cOffSet[z] < 0 || cOffSet[z] + numCorner[z] > nc)
      UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_connectivity: zone %d has bad numCorner/cOffSet", z + 1);
  }
  std::vector<int> fp0((size_t)mcf * nc), ez0((size_t)mcf * nc);
  for (size_t i = 0; i < fp0.size(); i++) {
    const int v = cFP[i], f = (int)(i % mcf), c = (int)(i / mcf);
    if (f < nCFaces[c]) {
      if (v < 1 || v > nc + ctx->nb) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_connectivity: cFP(%d,%d)=%d out of range", f + 1, c + 1, v);
      if (cEZ[i] < 1 || cEZ[i] > ctx->maxCorner) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_connectivity: cEZ(%d,%d)=%d out of range", f + 1, c + 1, cEZ[i]);
    }
    fp0[i] = v - 1;
    ez0[i] = cEZ[i] - 1;
  }
  TRY(dev_alloc_copy(ctx, &ctx->d_numCorner, numCorner, nz));
  TRY(dev_alloc_copy(ctx, &ctx->d_cOffSet, cOffSet, nz));
  TRY(dev_alloc_copy(ctx, &ctx->d_nCFaces, nCFaces, nc));
  TRY(dev_alloc_copy(ctx, &ctx->d_cFP, fp0.data(), fp0.size()));
  TRY(dev_alloc_copy(ctx, &ctx->d_cEZ, ez0.data(), ez0.size()));
  ctx->maxFaces = maxFaces;
  if (zoneFaces && zoneOpp && faceOpp && CToFace) {
    ctx->h_zoneFaces.assign(zoneFaces, zoneFaces + nz);
    ctx->h_zoneOpp.assign(zoneOpp, zoneOpp + (size_t)maxFaces * nz);
    ctx->h_faceOpp.assign(faceOpp, faceOpp + (size_t)maxFaces * nz);
    ctx->h_CToFace.assign(CToFace, CToFace + (size_t)mcf * nc);
  }
  if (BoundaryZone) ctx->h_BoundaryZone.assign(BoundaryZone, BoundaryZone + nz);
  if (BdyToC) ctx->h_BdyToC.assign(BdyToC, BdyToC + ctx->nb);
  ctx->have_conn = true;
  ctx->sched_dirty = true;
  return UMT_OK;
}

extern "C" int umt_set_geometry(umt_ctx *ctx, const double *Volume, const double *A_fp, const double *A_ez,
                                const double *Area, const double *RadiusFP, const double *RadiusEZ, const double *A_bdy) {
  if (!ctx || !Volume || !A_fp || !A_ez) return UMT_ERR_ARG;
  if (ctx->ndim == 2 && (!Area || !RadiusFP || !RadiusEZ)) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_geometry: RZ needs Area, RadiusFP, RadiusEZ");
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t nc = ctx->nc, nA = (size_t)ctx->ndim * ctx->maxcf * nc;
  ctx->h_Volume.assign(Volume, Volume + nc);
  ctx->h_Afp.assign(A_fp, A_fp + nA);
  ctx->h_Aez.assign(A_ez, A_ez + nA);
  TRY(dev_alloc_copy(ctx, &ctx->d_Volume, Volume, nc));
  TRY(dev_alloc_copy(ctx, &ctx->d_Afp, A_fp, nA));
  TRY(dev_alloc_copy(ctx, &ctx->d_Aez, A_ez, nA));
  if (ctx->ndim == 2) {
    ctx->h_Area.assign(Area, Area + nc);
    ctx->h_RadiusFP.assign(RadiusFP, RadiusFP + 2 * nc);
    ctx->h_RadiusEZ.assign(RadiusEZ, RadiusEZ + 2 * nc);
    TRY(dev_alloc_copy(ctx, &ctx->d_Area, Area, nc));
    TRY(dev_alloc_copy(ctx, &ctx->d_RadiusFP, RadiusFP, 2 * nc));
    TRY(dev_alloc_copy(ctx, &ctx->d_RadiusEZ, RadiusEZ, 2 * nc));
  }
  if (A_bdy) { ctx->h_Abdy.assign(A_bdy, A_bdy + (size_t)ctx->ndim * ctx->nb); ctx->have_abdy = true; }
  ctx->have_geom = true;
  ctx->sched_dirty = true;
  return UMT_OK;
}

extern "C" int umt_set_quadrature(umt_ctx *ctx, int nAngles, const double *omega, const double *weight,
                                  const unsigned char *start, const unsigned char *finish, const double *angDerivFac,
                                  const double *w1, const double *w2) {
  if (!ctx || nAngles < 1 || !omega || !weight) return UMT_ERR_ARG;
  if (ctx->ndim == 2 && (!start || !finish || !angDerivFac || !w1 || !w2))
    UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_quadrature: RZ needs starting/finishing flags and angular-derivative coefficients");
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->NA != nAngles && ctx->d_psi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_set_quadrature: angle count changed after state upload");
  ctx->NA = nAngles;
  ctx->h_omega.assign(omega, omega + (size_t)ctx->ndim * nAngles);
  ctx->h_weight.assign(weight, weight + nAngles);
  ctx->h_start.assign(nAngles, 0);
  ctx->h_finish.assign(nAngles + 1, 0);
  if (ctx->ndim == 2) {
    ctx->h_start.assign(start, start + nAngles);
    std::copy(finish, finish + nAngles, ctx->h_finish.begin());
    ctx->h_angDerivFac.assign(angDerivFac, angDerivFac + nAngles);
    ctx->h_tauW1.assign(w1, w1 + nAngles);
    ctx->h_tauW2.assign(w2, w2 + nAngles);
  }
  TRY(dev_alloc_copy(ctx, &ctx->d_omega, omega, (size_t)ctx->ndim * nAngles));
  TRY(dev_alloc_copy(ctx, &ctx->d_weight, weight, nAngles));
  if (ctx->ndim == 2) {
    // xi-levels: a level begins at every starting direction (rt/rtquad.F90:107-127)
    ctx->h_level.assign(nAngles, 0);
    int lev = -1;
    for (int a = 0; a < nAngles; a++) { if (ctx->h_start[a] || lev < 0) lev++; ctx->h_level[a] = lev; }
    ctx->nLevels = lev + 1;
    std::vector<unsigned char> finNext(nAngles);
    for (int a = 0; a < nAngles; a++) finNext[a] = ctx->h_finish[a + 1];
    TRY(dev_alloc_copy(ctx, &ctx->d_angDerivFac, angDerivFac, nAngles));
    TRY(dev_alloc_copy(ctx, &ctx->d_tauW1, w1, nAngles));
    TRY(dev_alloc_copy(ctx, &ctx->d_tauW2, w2, nAngles));
    TRY(dev_alloc_copy(ctx, &ctx->d_start, ctx->h_start.data(), nAngles));
    TRY(dev_alloc_copy(ctx, &ctx->d_finishNext, finNext.data(), nAngles));
    TRY(dev_alloc_copy(ctx, &ctx->d_level, ctx->h_level.data(), nAngles));
  }
  ctx->nHyp.assign(nAngles, 0); ctx->numCycles.assign(nAngles, 0); ctx->cycleOffSet.assign(nAngles, 0); ctx->nBad.assign(nAngles, 0);
  ctx->zonesInPlane.assign(nAngles, {}); ctx->nextZ.assign(nAngles, {}); ctx->nextC.assign(nAngles, {});
  ctx->cycleList.assign(nAngles, {}); ctx->bdyList.assign(nAngles, {});
  ctx->have_quad = true;
  ctx->sched_dirty = true;
  return UMT_OK;
}

extern "C" int umt_build_product_quadrature(umt_ctx *ctx, int npolar, int nazimuthal, int polaraxis, int *nAngles) {
  if (!ctx) return UMT_ERR_ARG;
  std::vector<double> om, w, adf, w1, w2;
  std::vector<unsigned char> st, fi;
  int r = umt_host_product_quadrature(ctx->ndim, npolar, nazimuthal, polaraxis, om, w, st, fi, adf, w1, w2);
  if (r) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_build_product_quadrature: npolar, nazimuthal must be in 1..32, polaraxis in 1..3");
  const int NA = (int)w.size();
  if (nAngles) *nAngles = NA;
  return umt_set_quadrature(ctx, NA, om.data(), w.data(), st.data(), fi.data(), adf.data(), w1.data(), w2.data());
}

extern "C" int umt_get_quadrature(umt_ctx *ctx, double *omega, double *weight) {
  if (!ctx || !ctx->have_quad) return UMT_ERR_STATE;
  if (omega) std::copy(ctx->h_omega.begin(), ctx->h_omega.end(), omega);
  if (weight) std::copy(ctx->h_weight.begin(), ctx->h_weight.end(), weight);
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// schedule
// ---------------------------------------------------------------------------
extern "C" int umt_set_schedule(umt_ctx *ctx, int angle, int nHyperPlanes, const int *zonesInPlane, const int *nextZ,
                                const int *nextC, int numCycles, const int *cycleList, int nxBdy, const int *bdyList) {
  if (!ctx || !ctx->have_quad) return UMT_ERR_STATE;
  if (angle < 1 || angle > ctx->NA || nHyperPlanes < 0 || numCycles < 0 || nxBdy < 0) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_schedule: bad argument");
  const int a = angle - 1;
  long total = 0;
  for (int p = 0; p < nHyperPlanes; p++) total += zonesInPlane[p];
  if (nHyperPlanes > 0 && total != ctx->nz) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "umt_set_schedule: angle %d planes hold %ld zones, expected %d", angle, total, ctx->nz);
  ctx->nHyp[a] = nHyperPlanes;
  ctx->zonesInPlane[a].assign(zonesInPlane, zonesInPlane + nHyperPlanes);
  if (nHyperPlanes > 0) {
    ctx->nextZ[a].assign(nextZ, nextZ + ctx->nz);
    ctx->nextC[a].assign(nextC, nextC + ctx->nc);
    int bad = 0;
    for (int i = 0; i < ctx->nz; i++) {
      int z = nextZ[i];
      if (z == 0 || std::abs(z) > ctx->nz) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "umt_set_schedule: nextZ(%d,%d)=%d out of range", i + 1, angle, z);
      bad += z < 0;
    }
    ctx->nBad[a] = bad;
  } else {
    ctx->nextZ[a].clear(); ctx->nextC[a].clear(); ctx->nBad[a] = 0;
  }
  ctx->numCycles[a] = numCycles;
  ctx->cycleList[a].assign(cycleList, cycleList + numCycles);
  ctx->bdyList[a].assign(bdyList, bdyList + 2 * (size_t)nxBdy);
  ctx->sched_dirty = true;
  return UMT_OK;
}

extern "C" int umt_build_schedule(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  if (!ctx->have_conn || !ctx->have_geom || !ctx->have_quad || ctx->h_zoneOpp.empty())
    UMT_FAIL(ctx, UMT_ERR_STATE, "umt_build_schedule: needs full connectivity (zoneOpp, faceOpp, CToFace), geometry and quadrature");
  return umt_host_build_schedule(ctx);
}

extern "C" int umt_get_schedule_info(umt_ctx *ctx, int angle, int *nHyperPlanes, int *numCycles, int *nBadZones) {
  if (!ctx || angle < 1 || angle > ctx->NA) return UMT_ERR_ARG;
  if (nHyperPlanes) *nHyperPlanes = ctx->nHyp[angle - 1];
  if (numCycles) *numCycles = ctx->numCycles[angle - 1];
  if (nBadZones) *nBadZones = ctx->nBad[angle - 1];
  return UMT_OK;
}

extern "C" int umt_get_schedule(umt_ctx *ctx, int angle, int *zonesInPlane, int *nextZ, int *nextC, int *cycleList) {
  if (!ctx || angle < 1 || angle > ctx->NA) return UMT_ERR_ARG;
  const int a = angle - 1;
  if (zonesInPlane) std::copy(ctx->zonesInPlane[a].begin(), ctx->zonesInPlane[a].end(), zonesInPlane);
  if (nextZ) std::copy(ctx->nextZ[a].begin(), ctx->nextZ[a].end(), nextZ);
  if (nextC) std::copy(ctx->nextC[a].begin(), ctx->nextC[a].end(), nextC);
  if (cycleList) std::copy(ctx->cycleList[a].begin(), ctx->cycleList[a].end(), cycleList);
  return UMT_OK;
}

static int ensure_layout(umt_ctx *ctx);
static int seed_cycle_psi(umt_ctx *ctx);

// Translate the per-angle hyperplane lists into the device work-item list.
static int finalize_schedule(umt_ctx *ctx) {
  if (!ctx->sched_dirty) return UMT_OK;
  const int NA = ctx->NA, nz = ctx->nz, nc = ctx->nc;
  int maxHyp = 0;
  for (int a = 0; a < NA; a++) maxHyp = std::max(maxHyp, ctx->nHyp[a]);
  if (maxHyp == 0) UMT_FAIL(ctx, UMT_ERR_STATE, "no sweep schedule installed (umt_set_schedule / umt_build_schedule)");
  ctx->maxHyp = maxHyp;
  std::vector<int> h_nextZ((size_t)NA * nz, 1);
  std::vector<unsigned char> h_nextC((size_t)NA * nc, 0);
  for (int a = 0; a < NA; a++) {
    if (ctx->nHyp[a] == 0) continue;
    std::copy(ctx->nextZ[a].begin(), ctx->nextZ[a].end(), h_nextZ.begin() + (size_t)a * nz);
    for (int i = 0; i < nc; i++) {
      int v = ctx->nextC[a][i];
      if (v < 1 || v > ctx->maxCorner) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "nextC(%d,%d)=%d out of range", i + 1, a + 1, v);
      h_nextC[(size_t)a * nc + i] = (unsigned char)(v - 1);
    }
  }
  // work items: plane-major, angle-minor.  RZ angles of one xi-level are chained
  // (PsiM dependency, SweepUCBrz.F90:212-240) so RZ uses a separate launcher.
  // plan kernel: every 3-D mesh with <= 8 corners per zone (zones it cannot plan take its slow path)
  ctx->use_plan = ctx->ndim == 3 && ctx->maxcf == 3 && ctx->maxCorner <= 8 && ctx->G % 2 == 0 && ctx->G <= 256 &&
                  (double)(ctx->nc + ctx->nb) * ctx->G < 2147483647.0;   // record offsets are 32-bit elements
  if (const char *e = getenv("UMT_SWEEP3D")) if (!strcmp(e, "generic")) ctx->use_plan = false;
  // one 16-byte column (2 groups) per lane: with the register-resident canonical solve the 4-groups-per-lane variant spills
  // (measured at -d 20: G=128 39.7 vs 88 ms, G=64 24.0 vs 52.5, G=32 16.3 vs 40.5, G=16 15.5 vs 20.1); UMT_PLAN_NH=2 selects it
  ctx->plan_nh = 1;
  if (const char *e = getenv("UMT_PLAN_NH")) ctx->plan_nh = (atoi(e) == 2 && ctx->G % 4 == 0) ? 2 : 1;
  ctx->plan_ncw = 4;
  if (const char *e = getenv("UMT_PLAN_WARPS")) ctx->plan_ncw = atoi(e) == 8 ? 8 : 4;
  TRY(dev_alloc_copy(ctx, &ctx->d_nextZ, h_nextZ.data(), h_nextZ.size()));
  TRY(dev_alloc_copy(ctx, &ctx->d_nextC, h_nextC.data(), h_nextC.size()));
  if (ctx->use_plan && ctx->device >= 0) {   // the plan records first: the Psi1 ring below is sized by what is left of the HBM
    std::vector<int2> zinfo((size_t)NA * nz);
    for (int a = 0; a < NA; a++)
      for (int i = 0; i < nz; i++) {
        const int z = std::abs(h_nextZ[(size_t)a * nz + i]) - 1;
        zinfo[(size_t)a * nz + i] = make_int2(ctx->h_cOffSet[z], z | (ctx->h_numCorner[z] << 28));
      }
    TRY(dev_alloc_copy(ctx, &ctx->d_zinfo, zinfo.data(), zinfo.size()));
    TRY(umt_build_plan3d(ctx));
  }
  int pairsRZ = 64;   // one (zone, group) pair per thread of the 64-thread RZ CTA (sweeprz.cu RZ_BLOCK)
  if (const char *e = getenv("UMT_PAIRS_PER_ITEM")) pairsRZ = std::max(1, std::min(64, atoi(e)));
  const int zpi = ctx->ndim == 3 ? umt_sweep3d_zones_per_item(ctx) : std::max(1, pairsRZ / ctx->G);
  ctx->zones_per_item = zpi;
  // Angles run in batches of K with staggered starts: a batch in its growing half overlaps the
  // previous batch's shrinking half, so the work per level is steady while the Psi1 rows of
  // the last few planes of every active angle stay L2-resident for their downstream zones.
  int K = NA;
  double stagger = 0.5;
  if (ctx->use_plan) {
    // measured at -d 20 G=128 (planes of ~750 zones): K=4 41.2 ms, K=8 42.5 ms, K=2 52 ms (too few ready items per level).
    // Small meshes have small planes: keep about 2400 items (4x the engines of the GPU) per level, two batches being active
    // at any time, so that the plane-to-plane latency of one angle hides behind the others.
    double avgPlane = 0.0;
    int nSwept = 0;
    for (int a = 0; a < NA; a++) if (ctx->nHyp[a] > 0) { avgPlane += (double)nz / ctx->nHyp[a]; nSwept++; }
    avgPlane = nSwept ? avgPlane / nSwept : 1.0;
    K = 4;
    while (K < NA && 2.0 * K * avgPlane / zpi < 2400.0) K *= 2;
    K = std::min(K, NA);
  }
  if (const char *e = getenv("UMT_ANGLE_BATCH")) K = std::max(1, atoi(e));
  if (const char *e = getenv("UMT_BATCH_STAGGER")) stagger = std::max(0.0, atof(e));
  const int delta = std::max(1, (int)(stagger * maxHyp));
  std::vector<WorkItem> items;
  std::vector<std::vector<int>> nItemsPlane(NA);
  std::vector<std::vector<int>> planeStart(NA);
  for (int a = 0; a < NA; a++) {
    nItemsPlane[a].resize(ctx->nHyp[a]);
    planeStart[a].resize(ctx->nHyp[a] + 1, 0);
    for (int p = 0; p < ctx->nHyp[a]; p++) {
      planeStart[a][p + 1] = planeStart[a][p] + ctx->zonesInPlane[a][p];
      nItemsPlane[a][p] = (ctx->zonesInPlane[a][p] + zpi - 1) / zpi;
    }
  }
  // completion counters count items (generic kernel: one CTA-wide signal per item) or, for the plan
  // kernel, consumer warps (each warp holding lanes of the item signals on its own)
  const int Gv = std::max(1, ctx->G / 2);
  auto signals = [&](int nZonesInItem) { (void)nZonesInItem; (void)Gv; return 1; };   // one signal per item in both kernels
  std::vector<std::vector<int>> planeSignals(NA);
  for (int a = 0; a < NA; a++) {
    planeSignals[a].assign(ctx->nHyp[a], 0);
    for (int p = 0; p < ctx->nHyp[a]; p++) {
      const int n = ctx->zonesInPlane[a][p];
      for (int k = 0; k < nItemsPlane[a][p]; k++) planeSignals[a][p] += signals(std::min(zpi, n - k * zpi));
    }
  }
  // Reflecting boundaries couple angles (snac/snreflect.F90): an incident angle needs the exiting flux of its mirror
  // image.  Angles are therefore swept in stages (umt_reflect_stages); all angles of a stage are independent.
  TRY(umt_reflect_stages(ctx));
  ctx->stageItemBegin.assign(ctx->nStages + 1, 0);
  // Psi storage layout (umt_internal.h): one Psi + a ring for Psi1 whenever the sweep never needs an angle's previous Psi1 and all
  // angles go in one launch; otherwise the two-buffer layout.
  {
    bool single = ctx->ndim == 3 && ctx->use_plan && ctx->device >= 0 && !ctx->force_legacy && ctx->nStages == 1 && !ctx->have_comm_order;
    for (int a = 0; a < NA && single; a++) single = ctx->nHyp[a] > 0 && ctx->numCycles[a] == 0 && ctx->nBad[a] == 0;
    if (const char *e = getenv("UMT_PSI_LAYOUT")) if (!strcmp(e, "legacy")) single = false;
    ctx->single_psi_wanted = single;
    ctx->ringBatchesAuto = 0;
    if (single) {
      // as many angle batches as the free HBM holds next to what is already allocated (the current workspace counts as free)
      size_t freeB = 0, totalB = 0;
      UMT_CUDA(ctx, cudaMemGetInfo(&freeB, &totalB));
      const size_t slabB = sizeof(double) * (size_t)(nc + ctx->nb) * ctx->G;
      freeB += (size_t)ctx->psi1Slots * slabB;
      const size_t reserve = std::max((size_t)4 << 30, totalB / 10);   // left to the caller and to this library's smaller arrays
      const size_t slots = freeB > reserve ? (freeB - reserve) / slabB : 0;
      ctx->ringBatchesAuto = (int)std::min<size_t>((size_t)(NA + K - 1) / K, slots / (size_t)K);
      if (const char *e = getenv("UMT_RING_BATCHES")) ctx->ringBatchesAuto = std::max(1, atoi(e));
      // A ring that holds every batch is the two-buffer layout with one address computation more per upstream row (measured at
      // -d 20 -G 128: 41.7 against 40.3 ms per sweep): when everything fits, or is asked for, keep the legacy layout.
      const int want = ctx->ringBatchesWanted > 0 ? ctx->ringBatchesWanted : ctx->ringBatchesAuto;
      if (want >= (NA + K - 1) / K) ctx->single_psi_wanted = false;
    }
  }
  auto make_item = [&](int a, int p, int k, int upIdx) {
    const int n = ctx->zonesInPlane[a][p], z0 = planeStart[a][p];
    WorkItem w;
    w.angle = a;
    w.zbeg = z0 + k * zpi;
    w.zend = std::min(z0 + n, w.zbeg + zpi);
    w.wait_idx = p > 0 ? a * maxHyp + p - 1 : -1;
    w.wait_count = p > 0 ? planeSignals[a][p - 1] : 0;
    w.signal_idx = a * maxHyp + p;
    w.pad0 = -1;
    w.pad1 = upIdx;
    return w;
  };
  std::vector<WorkItem> ringItems;
  ctx->angleBatch = K; ctx->nBatches = 0; ctx->ringBatches = 0; ctx->nTallied = 0;
  ctx->slotOfAngle.assign(NA, 0);
  for (int a = 0; a < NA; a++) ctx->slotOfAngle[a] = a;
  int nExtraCounters = 0;
  if (ctx->ndim == 2) {
    TRY(umt_build_items_rz(ctx, items, zpi));   // PsiM chain within a xi-level: own ordering; grouped by reflection stage
  } else {
    // every angle's Psi1 in its own slab (legacy layout, and the in-place savePsi sweep of the single-psi layout)
    for (int s = 0; s < ctx->nStages; s++) {
      std::vector<int> ang;
      for (int a2 = 0; a2 < NA; a2++) if (ctx->stageOf[a2] == s) ang.push_back(a2);
      const int nA = (int)ang.size();
      const int nBatches = (nA + K - 1) / K;
      const int nLevels = maxHyp + (nBatches - 1) * delta;
      for (int lev = 0; lev < nLevels; lev++)
        for (int ia = 0; ia < nA; ia++) {
          const int a = ang[ia];
          const int p = lev - (ia / K) * delta;
          if (p < 0 || p >= ctx->nHyp[a]) continue;
          for (int k = 0; k < nItemsPlane[a][p]; k++) items.push_back(make_item(a, p, k, a));
        }
      ctx->stageItemBegin[s + 1] = (int)items.size();
    }
    // Single-psi layout, non-final sweeps: Psi1 lives in a ring of ringBatches * K slabs.  Batch b (angles bK .. bK+K-1) writes the
    // slots (b mod ringBatches) K + i; once its last plane is done its Psi1 is tallied into PhiTotal by phi-tally items spread over
    // the following levels, and batch b + ringBatches may then reuse the slots.  Counters past the plane counters: gate[b] (bumped
    // by every sweep item of batch b and by every tally item of batch b-1: a tally item of batch b waits for all of them, which
    // also keeps the tallies in batch order) and done[b] (bumped by the tally items of batch b: plane 0 of batch b + ringBatches
    // waits for it).  Every dependency of an item sits earlier in the ticket order, so the launch cannot deadlock.
    if (ctx->single_psi_wanted) {
      const int nB = (NA + K - 1) / K;
      int Rb = ctx->ringBatchesWanted > 0 ? ctx->ringBatchesWanted : ctx->ringBatchesAuto;
      Rb = std::max(1, std::min(Rb, nB));
      if (Rb < nB && Rb < 2) Rb = std::min(2, nB);       // a retiring batch and a running one at the very least
      ctx->nBatches = nB; ctx->ringBatches = Rb;
      const int tallyB = nB - Rb;                          // batches 0 .. tallyB-1 are tallied by the sweep kernel itself
      ctx->nTallied = std::min(NA, tallyB * K);
      for (int a = 0; a < NA; a++) ctx->slotOfAngle[a] = ((a / K) % Rb) * K + a % K;
      if (tallyB > 0) {
        const int columns = (int)(((size_t)nc * ctx->G) / 2);   // use_plan: G even
        int CH = 2048;
        if (const char *e = getenv("UMT_PHI_CHUNK")) CH = std::max(64, atoi(e));
        const int nPhi = (columns + CH - 1) / CH;
        // The tail of a batch (its last, small planes: one dependent hop each) finishes well after the ticket counter has passed
        // its last level, and a tally item that is taken before its batch is complete blocks the in-order queue of its CTA: the
        // tally items therefore start `lag` levels after the batch's last level and are spread over `spread` levels.
        // (measured at -d 20 -G 128, ring of 3 batches: lag 0 / spread 0.25: 57.4 ms, 0.3 / 0.2: 55.9, 0.4 / 0.1: 54.2, 0.25 / 0.5: 61.6)
        double lagF = 0.4, spreadF = 0.1;
        if (const char *e = getenv("UMT_PHI_LAG")) lagF = std::max(0.0, atof(e));
        if (const char *e = getenv("UMT_PHI_SPREAD")) spreadF = std::max(0.0, atof(e));
        const int spread = std::max(1, (int)(spreadF * maxHyp)), lag = (int)(lagF * maxHyp);
        const int gateBase = NA * maxHyp, doneBase = gateBase + nB;
        nExtraCounters = 2 * nB;
        std::vector<int> start(nB), tend(nB, 0), nItemsBatch(nB, 0);
        for (int b = 0; b < nB; b++) {
          for (int a = b * K; a < std::min(NA, (b + 1) * K); a++)
            for (int p = 0; p < ctx->nHyp[a]; p++) nItemsBatch[b] += nItemsPlane[a][p];
          start[b] = b * delta;
          if (b > 0) start[b] = std::max(start[b], start[b - 1]);
          if (b >= Rb) start[b] = std::max(start[b], tend[b - Rb] + spread);       // its slots' previous tenant is fully tallied
          tend[b] = start[b] + maxHyp + lag;                                      // first level of batch b's tally items
          if (b > 0) tend[b] = std::max(tend[b], tend[b - 1] + spread);
        }
        int nLevels = 0;
        for (int b = 0; b < nB; b++) nLevels = std::max(nLevels, std::max(start[b] + maxHyp, b < tallyB ? tend[b] + spread : 0));
        std::vector<WorkItem> lvSweep, lvPhi;
        for (int lev = 0; lev < nLevels; lev++) {
          lvSweep.clear(); lvPhi.clear();
          for (int a = 0; a < NA; a++) {
            const int b = a / K, p = lev - start[b];
            if (p < 0 || p >= ctx->nHyp[a]) continue;
            for (int k = 0; k < nItemsPlane[a][p]; k++) {
              WorkItem w = make_item(a, p, k, ctx->slotOfAngle[a]);
              if (p == 0 && b >= Rb && b - Rb < tallyB) { w.wait_idx = doneBase + b - Rb; w.wait_count = nPhi; }
              if (b < tallyB) w.pad0 = gateBase + b;
              lvSweep.push_back(w);
            }
          }
          for (int b = 0; b < tallyB; b++) {
            const int l = lev - tend[b];
            if (l < 0 || l >= spread) continue;
            const int j0 = (int)((long long)nPhi * l / spread), j1 = (int)((long long)nPhi * (l + 1) / spread);
            const int nAb = std::min(NA, (b + 1) * K) - b * K;
            for (int j = j0; j < j1; j++) {
              WorkItem w;
              w.angle = -1 - b * K;
              w.zbeg = j * CH; w.zend = std::min(columns, (j + 1) * CH);
              w.wait_idx = gateBase + b; w.wait_count = nItemsBatch[b] + (b > 0 ? nPhi : 0);
              w.signal_idx = doneBase + b;
              w.pad0 = b + 1 < tallyB ? gateBase + b + 1 : -1;
              w.pad1 = ctx->slotOfAngle[b * K] | (nAb << 16) | ((b == 0 ? 1 : 0) << 30);
              lvPhi.push_back(w);
            }
          }
          // tally items evenly among the level's sweep items (both are independent of each other within a level)
          size_t is = 0, ip = 0;
          const size_t ns = lvSweep.size(), np = lvPhi.size();
          while (is < ns || ip < np) {
            if (ip < np && (is >= ns || ip * (ns + 1) <= is * np)) ringItems.push_back(lvPhi[ip++]);
            else ringItems.push_back(lvSweep[is++]);
          }
        }
      }
    }
  }
  if (ctx->ndim == 2) ctx->stageItemBegin[ctx->nStages] = (int)items.size();
  ctx->nItems = (int)items.size();
  ctx->nItemsRing = (int)ringItems.size();
  ctx->nCounters = NA * maxHyp + nExtraCounters;
  if (ctx->single_psi_wanted && ctx->nTallied > 0) TRY(dev_alloc_copy(ctx, &ctx->d_itemsRing, ringItems.data(), ringItems.size()));
  else if (ctx->d_itemsRing) { cudaFree(ctx->d_itemsRing); ctx->d_itemsRing = nullptr; ctx->nItemsRing = 0; }
  TRY(dev_alloc_copy(ctx, &ctx->d_items, items.data(), items.size()));
  TRY(dev_alloc_copy<int>(ctx, &ctx->d_counters, nullptr, 1 + (size_t)ctx->nCounters));
  if (ctx->device >= 0 && ctx->d_psi) TRY(ensure_layout(ctx));
  // cycle lists (control/constructDynMemory.F90:56-213)
  std::vector<int> cl, ca;
  int off = 0;
  for (int a = 0; a < NA; a++) {
    ctx->cycleOffSet[a] = off;
    for (int v : ctx->cycleList[a]) {
      if (v < 1 || v > nc) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "cycleList entry %d out of range (angle %d)", v, a + 1);
      cl.push_back(v - 1); ca.push_back(a);
    }
    off += ctx->numCycles[a];
  }
  // Set%cyclePsi belongs to the list it was filled for: when the list changes (another corner set, not merely another length) the
  // old values are some other corners' fluxes.  Re-seed it from Psi the way initCyclePsi does (constructDynMemory.F90:56-109);
  // the caller's umt_init_radiation_field / umt_init_cycle_psi after the upload of Psi does the same explicitly.
  const bool cyclesChanged = !ctx->d_cyclePsi || cl != ctx->h_cycleFlat || ca != ctx->h_cycleAngleFlat;
  ctx->totalCycles = off;
  ctx->h_cycleFlat = cl; ctx->h_cycleAngleFlat = ca;
  TRY(dev_alloc_copy(ctx, &ctx->d_cycleList, cl.data(), cl.size()));
  TRY(dev_alloc_copy(ctx, &ctx->d_cycleAngle, ca.data(), ca.size()));
  if (cyclesChanged && ctx->device >= 0) {
    TRY(dev_alloc_copy<double>(ctx, &ctx->d_cyclePsi, nullptr, (size_t)std::max(off, 1) * ctx->G));
    UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_cyclePsi, 0, sizeof(double) * (size_t)std::max(off, 1) * ctx->G, ctx->stream));
    if (off > 0 && ctx->d_psi) TRY(seed_cycle_psi(ctx));
  }
  // exit lists (AngleSet BdyExit, rt/findexit.F90:296-349)
  std::vector<int> eb, ec, ea;
  ctx->exitOff.assign(NA + 1, 0);
  for (int a = 0; a < NA; a++) {
    const auto &bl = ctx->bdyList[a];
    for (size_t i = 0; i + 1 < bl.size(); i += 2) {
      if (bl[i] < 1 || bl[i] > ctx->nb || bl[i + 1] < 1 || bl[i + 1] > nc) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "bdyList entry out of range (angle %d)", a + 1);
      eb.push_back(bl[i] - 1); ec.push_back(bl[i + 1] - 1); ea.push_back(a);
    }
    ctx->exitOff[a + 1] = (int)eb.size();
  }
  ctx->nExit = (int)eb.size();
  TRY(dev_alloc_copy(ctx, &ctx->d_exitB, eb.data(), eb.size()));
  TRY(dev_alloc_copy(ctx, &ctx->d_exitC, ec.data(), ec.size()));
  TRY(dev_alloc_copy(ctx, &ctx->d_exitA, ea.data(), ea.size()));
  ctx->sched_dirty = false;
  return UMT_OK;
}

int umt_finalize_schedule(umt_ctx *ctx) { return finalize_schedule(ctx); }

// A dataflow kernel (values as their own completion flags) whose polling budget ran out raised the abort flag; its result is garbage.
int umt_check_abort(umt_ctx *ctx, const char *what) {
  if (!ctx->h_abort || *ctx->h_abort == 0) return UMT_OK;
  *ctx->h_abort = 0;
  cudaMemsetAsync(ctx->d_abort, 0, sizeof(int), ctx->stream);
  UMT_FAIL(ctx, UMT_ERR_STATE, "%s: a dataflow sweep gave up waiting for a value that never became real (an input carrying the "
           "'not computed yet' bit pattern 0xFFFFDEADFFFFDEAD, or a broken schedule); the result of this call is undefined", what);
}

// ---------------------------------------------------------------------------
// state
// ---------------------------------------------------------------------------
static int ensure_state(umt_ctx *ctx) {
  NEED_DEVICE(ctx, "device state");
  if (!ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "set the quadrature before uploading state");
  if (ctx->d_psi) return UMT_OK;
  // Psi slabs are (G, ncornr+nbelem) per angle: like Set%Psi1 in the reference (SetData_mod.F90:164) the boundary-element rows
  // follow the corner rows, so a cFP value addresses either kind of upstream row.  The Psi1 workspace (full or a ring) is
  // allocated once the schedule says which layout applies (ensure_layout); until then Set%PsiB lives in the tails of d_psi.
  const size_t G = ctx->G, nc = ctx->nc, NA = ctx->NA;
  ctx->rows = ctx->nc + ctx->nb;
  ctx->psi_elems = G * ctx->rows * NA;
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_psi, sizeof(double) * ctx->psi_elems));
  ctx->psib_in_psi = true; ctx->psi1Slots = 0;
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_stotal, sizeof(double) * G * nc));
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_sigt, sizeof(double) * G * ctx->nz));
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_phi, sizeof(double) * G * nc));
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_psi, 0, sizeof(double) * ctx->psi_elems, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_stotal, 0, sizeof(double) * G * nc, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_sigt, 0, sizeof(double) * G * ctx->nz, ctx->stream));
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_phi, 0, sizeof(double) * G * nc, ctx->stream));
  if (ctx->ndim == 2) {
    const size_t nl = (size_t)std::max(ctx->nLevels, 1);   // Set%PsiM(G,nc), one per xi-level (levels sweep concurrently)
    UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_psim, sizeof(double) * G * nc * nl));
    UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_psim, 0, sizeof(double) * G * nc * nl, ctx->stream));
  }
  return UMT_OK;
}

extern "C" int umt_upload_state(umt_ctx *ctx, const double *Psi, const double *PsiB, const double *Sigt,
                                const double *STotal, double tau) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  const size_t G = ctx->G, nc = ctx->nc, nb = ctx->nb, NA = ctx->NA, pitch = G * ctx->rows * 8;
  if (Psi) UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi, pitch, Psi, G * nc * 8, G * nc * 8, NA, cudaMemcpyHostToDevice, ctx->stream));
  if (PsiB && nb) ctx->pack_valid = ctx->recv_valid = false;
  if (PsiB && nb) UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->psib_buf() + G * nc, pitch, PsiB, G * nb * 8, G * nb * 8, NA, cudaMemcpyHostToDevice, ctx->stream));
  if (Sigt) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sigt, Sigt, sizeof(double) * G * ctx->nz, cudaMemcpyHostToDevice, ctx->stream));
  if (STotal) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_stotal, STotal, sizeof(double) * G * nc, cudaMemcpyHostToDevice, ctx->stream));
  ctx->tau = tau;
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

static int set_copy(umt_ctx *ctx, int g0, int Groups, int angle0, int NumAngles, double *Psi, double *PsiB, bool up) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  if (g0 < 0 || Groups < 1 || g0 + Groups > ctx->G || angle0 < 0 || NumAngles < 1 || angle0 + NumAngles > ctx->NA)
    UMT_FAIL(ctx, UMT_ERR_ARG, "phase-space set (g0=%d,Groups=%d,angle0=%d,NumAngles=%d) outside problem", g0, Groups, angle0, NumAngles);
  const size_t G = ctx->G, nc = ctx->nc, nb = ctx->nb;
  const cudaMemcpyKind kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  for (int a = 0; a < NumAngles; a++) {
    if (Psi) {
      double *d = ctx->d_psi + ((size_t)(angle0 + a) * ctx->rows) * G + g0, *h = Psi + (size_t)a * nc * Groups;
      if (up) UMT_CUDA(ctx, cudaMemcpy2DAsync(d, G * 8, h, (size_t)Groups * 8, (size_t)Groups * 8, nc, kind, ctx->stream));
      else UMT_CUDA(ctx, cudaMemcpy2DAsync(h, (size_t)Groups * 8, d, G * 8, (size_t)Groups * 8, nc, kind, ctx->stream));
    }
    if (PsiB && nb && up) ctx->pack_valid = ctx->recv_valid = false;
    if (PsiB && nb) {
      double *d = ctx->psib_buf() + ((size_t)(angle0 + a) * ctx->rows + nc) * G + g0, *h = PsiB + (size_t)a * nb * Groups;
      if (up) UMT_CUDA(ctx, cudaMemcpy2DAsync(d, G * 8, h, (size_t)Groups * 8, (size_t)Groups * 8, nb, kind, ctx->stream));
      else UMT_CUDA(ctx, cudaMemcpy2DAsync(h, (size_t)Groups * 8, d, G * 8, (size_t)Groups * 8, nb, kind, ctx->stream));
    }
  }
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

extern "C" int umt_upload_set(umt_ctx *ctx, int g0, int Groups, int angle0, int NumAngles, const double *Psi, const double *PsiB) {
  return set_copy(ctx, g0, Groups, angle0, NumAngles, const_cast<double *>(Psi), const_cast<double *>(PsiB), true);
}
extern "C" int umt_download_set(umt_ctx *ctx, int g0, int Groups, int angle0, int NumAngles, double *Psi, double *PsiB) {
  return set_copy(ctx, g0, Groups, angle0, NumAngles, Psi, PsiB, false);
}

// rows x NA strided download: `width` doubles per angle starting `offset` doubles into each angle slab
static int download(umt_ctx *ctx, double *h, const double *d, size_t offset, size_t width, size_t count, size_t pitch) {
  if (!ctx || !h) return UMT_ERR_ARG;
  NEED_DEVICE(ctx, "download");
  if (!d) UMT_FAIL(ctx, UMT_ERR_STATE, "no device state to download");
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (width && count) UMT_CUDA(ctx, cudaMemcpy2DAsync(h, width * 8, d + offset, pitch * 8, width * 8, count, cudaMemcpyDeviceToHost, ctx->stream));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}
extern "C" int umt_download_psi(umt_ctx *ctx, double *Psi) {
  if (!ctx) return UMT_ERR_ARG;
  return download(ctx, Psi, ctx->d_psi, 0, (size_t)ctx->G * ctx->nc, ctx->NA, (size_t)ctx->G * ctx->rows);
}
extern "C" int umt_download_psib(umt_ctx *ctx, double *PsiB) {
  if (!ctx) return UMT_ERR_ARG;
  return download(ctx, PsiB, ctx->psib_buf(), (size_t)ctx->G * ctx->nc, (size_t)ctx->G * ctx->nb, ctx->NA, (size_t)ctx->G * ctx->rows);
}
extern "C" int umt_download_phi(umt_ctx *ctx, double *Phi) {
  if (!ctx) return UMT_ERR_ARG;
  return download(ctx, Phi, ctx->d_phi, 0, (size_t)ctx->G * ctx->nc, 1, (size_t)ctx->G * ctx->nc);
}

extern "C" int umt_synchronize(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  NEED_DEVICE(ctx, "umt_synchronize");
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// streaming kernels around the sweep
// ---------------------------------------------------------------------------
// psi -> phi: PhiTotal(g,c) = sum_a w_a Psi1(g,c,a) in fixed angle order
// (control/getPhiTotal_OMPOL.F90:134-160 + SweepUCBxyz.F90:270).  One thread per
// pair of groups; every angle slab is read once, coalesced, 16 bytes per thread.
__global__ void __launch_bounds__(256) phi_reduce_kernel(const double *__restrict__ psi, const double *__restrict__ w,
                                                         const unsigned char *__restrict__ skip, double *__restrict__ phi,
                                                         size_t n, int NA, size_t stride /* doubles between angle slabs */) {
  const size_t i2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i2 >= n) return;
  if (i2 + 1 < n && (stride & 1) == 0) {
    double2 s = make_double2(0.0, 0.0);
#pragma unroll 4
    for (int a = 0; a < NA; a++) {
      if (skip && skip[a]) continue;
      const double2 v = __ldcs(reinterpret_cast<const double2 *>(psi + (size_t)a * stride + i2));
      const double wa = w[a];
      s.x = s.x + wa * v.x;
      s.y = s.y + wa * v.y;
    }
    *reinterpret_cast<double2 *>(phi + i2) = s;
  } else {
    for (size_t i = i2; i < n && i < i2 + 2; i++) {
      double s = 0.0;
      for (int a = 0; a < NA; a++) {
        if (skip && skip[a]) continue;
        s = s + w[a] * psi[(size_t)a * stride + i];
      }
      phi[i] = s;
    }
  }
}

// Psi(:,c,a) *= VolumeOld(c)/Volume(c)   (initPhiTotal_OMPOL.F90:160-165)
__global__ void scale_psi_kernel(double *psi, const double *ratio, size_t ncG, int G, int NA, size_t stride) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncG) return;
  const double r = ratio[i / G];
  for (int a = 0; a < NA; a++) psi[(size_t)a * stride + i] *= r;
}

// K6: Psi1(:,c) <- cyclePsi(:,m)  /  cyclePsi(:,m) <- Psi1(:,c)   (constructDynMemory.F90:111-213)
__global__ void cycle_copy_kernel(double *field /* (NA,rows,G) */, double *cyclePsi, const int *cycleList, const int *cycleAngle,
                                  int total, int nc /* rows per angle */, int G, int toField) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)total * G) return;
  const int m = (int)(i / G), g = (int)(i - (size_t)m * G);
  const size_t f = ((size_t)cycleAngle[m] * nc + cycleList[m]) * G + g;
  if (toField) field[f] = cyclePsi[i];
  else cyclePsi[i] = field[f];
}

// PsiB(:,b,a) <- Psi(:,c,a) on exiting boundary elements (initializeRadiationField_OMPOL.F90:116-143)
__global__ void exit_copy_kernel(const double *psi, double *psi1, const int *eb, const int *ec, const int *ea, int nExit,
                                 int nc, int rows, int G) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)nExit * G) return;
  const int m = (int)(i / G), g = (int)(i - (size_t)m * G);
  psi1[((size_t)ea[m] * rows + nc + eb[m]) * G + g] = psi[((size_t)ea[m] * rows + ec[m]) * G + g];
}

// The same tally over an explicit list of Psi1 slabs (ring slots of the angles the sweep kernel has not tallied itself), on top of
// what PhiTotal already holds when `accumulate` is set: the angles come in ascending order, so the running sum is the same
// sequence of operations as the fixed-order sum over all angles.
__global__ void __launch_bounds__(256) phi_slots_kernel(const double *__restrict__ ws, const int *__restrict__ slot, const double *__restrict__ w,
                                                        int nA, double *__restrict__ phi, size_t n, size_t stride, int accumulate) {
  const size_t i2 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i2 >= n) return;   // n and stride are even on this path (G even)
  double2 s = make_double2(0.0, 0.0);
  if (accumulate) s = *reinterpret_cast<const double2 *>(phi + i2);
#pragma unroll 4
  for (int a = 0; a < nA; a++) {
    const double2 v = __ldcs(reinterpret_cast<const double2 *>(ws + (size_t)slot[a] * stride + i2));
    const double wa = w[a];
    s.x = s.x + wa * v.x;
    s.y = s.y + wa * v.y;
  }
  *reinterpret_cast<double2 *>(phi + i2) = s;
}

// psi -> phi after the sweep(s) for the corners*groups range [off, off+n) (off even), whatever the layout:
//   legacy: all angles from the Psi1 workspace; single-psi, savePsi sweep: all angles from Psi (written in place);
//   single-psi, other sweeps: the angles still in the ring, on top of what the sweep kernel tallied.
static int launch_phi_range(umt_ctx *ctx, int savePsi, size_t off, size_t n) {
  const size_t threads = (n + 1) / 2, stride = (size_t)ctx->rows * ctx->G;
  const int grid = (int)((threads + 255) / 256);
  if (!ctx->single_psi || savePsi) {
    const double *field = ctx->single_psi ? ctx->d_psi : ctx->d_psi1;
    phi_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(field + off, ctx->d_weight, nullptr, ctx->d_phi + off, n, ctx->NA, stride);
  } else {
    const int nTail = ctx->NA - ctx->nTallied;
    phi_slots_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->d_psi1 + off, ctx->d_tailSlot, ctx->d_tailW, nTail, ctx->d_phi + off, n, stride,
                                                    ctx->nTallied > 0 ? 1 : 0);
  }
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches += 1;
  return UMT_OK;
}

static int launch_phi(umt_ctx *ctx, const double *field) {   // all angles of a full (NA, rows, G) field
  const size_t n = (size_t)ctx->nc * ctx->G;
  const size_t threads = (n + 1) / 2;
  const int grid = (int)((threads + 255) / 256);
  // starting / finishing directions carry zero weight (rt/rtquad.F90:107-127) and are not tallied
  // (SweepUCBrz.F90:233-239); weight==0 makes the product exact, no skip array needed.
  phi_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(field, ctx->d_weight, nullptr, ctx->d_phi, n, ctx->NA, (size_t)ctx->rows * ctx->G);
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches += 1;
  return UMT_OK;
}

// (Re)allocate the Psi1 workspace for the layout the schedule allows and move Set%PsiB to where that layout keeps it.
static int ensure_layout(umt_ctx *ctx) {
  const bool single = ctx->single_psi_wanted;
  const int slots = single ? std::min(ctx->NA, ctx->ringBatches * ctx->angleBatch) : ctx->NA;
  if (single) {   // what the post-sweep tally still has to add (follows the item list just built)
    std::vector<int> ts;
    std::vector<double> tw;
    for (int a = ctx->nTallied; a < ctx->NA; a++) { ts.push_back(ctx->slotOfAngle[a]); tw.push_back(ctx->h_weight[a]); }
    TRY(dev_alloc_copy(ctx, &ctx->d_tailSlot, ts.data(), ts.size()));
    TRY(dev_alloc_copy(ctx, &ctx->d_tailW, tw.data(), tw.size()));
  }
  if (ctx->d_psi1 && ctx->single_psi == single && ctx->psi1Slots == slots) return UMT_OK;
  const size_t G = ctx->G, slab = G * ctx->rows, pitch = slab * 8, off = G * ctx->nc, tailB = G * ctx->nb * 8;
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->d_psi1) {
    if (!ctx->psib_in_psi && ctx->nb > 0)     // legacy -> anything: PsiB goes home to the tails of d_psi first
      UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi + off, pitch, ctx->d_psi1 + off, pitch, tailB, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->psib_in_psi = true;
    UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    UMT_CUDA(ctx, cudaFree(ctx->d_psi1));
    ctx->d_psi1 = nullptr; ctx->psi1Slots = 0;
  }
  if (slots < 1) UMT_FAIL(ctx, UMT_ERR_CUDA, "not enough free device memory for a Psi1 ring of even one angle batch");
  cudaError_t e = cudaMalloc((void **)&ctx->d_psi1, sizeof(double) * slab * slots);
  if (e != cudaSuccess) {
    ctx->d_psi1 = nullptr;
    UMT_FAIL(ctx, UMT_ERR_CUDA, "Psi1 workspace of %d slabs (%.1f GB): %s", slots, (double)(slab * slots * 8) / 1e9, cudaGetErrorString(e));
  }
  UMT_CUDA(ctx, cudaMemsetAsync(ctx->d_psi1, 0, sizeof(double) * slab * slots, ctx->stream));
  ctx->psi1Slots = slots; ctx->single_psi = single;
  if (!single) {
    if (ctx->nb > 0) UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi1 + off, pitch, ctx->d_psi + off, pitch, tailB, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->psib_in_psi = false;
  }
  if (getenv("UMT_VERBOSE"))
    fprintf(stderr, "umt: psi layout %s, Psi1 workspace %d slabs (%.2f GB), angle batch %d, ring %d of %d batches, %d angles tallied in the sweep kernel\n",
            single ? "single" : "legacy", slots, (double)(slab * slots * 8) / 1e9, ctx->angleBatch, ctx->ringBatches, ctx->nBatches, ctx->nTallied);
  return UMT_OK;
}

// initCyclePsi (constructDynMemory.F90:56-109): cyclePsi(:,m) <- Psi(:,c,angle) for the corners on the cycle lists
static int seed_cycle_psi(umt_ctx *ctx) {
  if (ctx->totalCycles > 0) {
    const size_t n = (size_t)ctx->totalCycles * ctx->G;
    cycle_copy_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi, ctx->d_cyclePsi, ctx->d_cycleList,
                                                                       ctx->d_cycleAngle, ctx->totalCycles, ctx->rows, ctx->G, 0);
    UMT_CUDA(ctx, cudaGetLastError());
  }
  return UMT_OK;
}

extern "C" int umt_init_cycle_psi(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  NEED_DEVICE(ctx, "umt_init_cycle_psi");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  TRY(finalize_schedule(ctx));
  TRY(seed_cycle_psi(ctx));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

// Angle batches the Psi1 ring of the single-psi layout should hold (0: as many as the free device memory allows, which is also
// the fastest; >= the number of batches: every angle keeps its own slab and nothing is tallied inside the sweep kernel).
extern "C" int umt_set_psi1_ring(umt_ctx *ctx, int nBatches) {
  if (!ctx || nBatches < 0) return UMT_ERR_ARG;
  ctx->ringBatchesWanted = nBatches;
  ctx->sched_dirty = true;
  return UMT_OK;
}

// layout in force after the last schedule finalisation: info[0] 1 = single-psi, [1] Psi1 slabs allocated, [2] angle batch K,
// [3] ring size in batches, [4] batches, [5] angles tallied by the sweep kernel; bytes = device memory of Psi + Psi1 workspace
extern "C" int umt_get_psi_layout(umt_ctx *ctx, int *info6, double *bytes) {
  if (!ctx) return UMT_ERR_ARG;
  NEED_DEVICE(ctx, "umt_get_psi_layout");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->d_psi) TRY(finalize_schedule(ctx));
  if (info6) {
    info6[0] = ctx->single_psi ? 1 : 0; info6[1] = ctx->psi1Slots; info6[2] = ctx->angleBatch;
    info6[3] = ctx->ringBatches; info6[4] = ctx->nBatches; info6[5] = ctx->nTallied;
  }
  if (bytes) *bytes = 8.0 * (double)ctx->G * ctx->rows * ((double)ctx->NA + ctx->psi1Slots);
  return UMT_OK;
}

extern "C" int umt_init_phi_total(umt_ctx *ctx, const double *volRatio) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  if (volRatio) {
    double *d_r = nullptr;
    TRY(dev_alloc_copy(ctx, &d_r, volRatio, (size_t)ctx->nc));
    const size_t n = (size_t)ctx->nc * ctx->G;
    scale_psi_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi, d_r, n, ctx->G, ctx->NA, (size_t)ctx->rows * ctx->G);
    UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d_r);
  }
  TRY(launch_phi(ctx, ctx->d_psi));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

// control/setBoundarySources.F90:42 with no source profiles: Set%PsiB(:,:,:) = 0
extern "C" int umt_set_boundary_sources(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  ctx->pack_valid = ctx->recv_valid = false;
  if (ctx->nb > 0) {
    const size_t G = ctx->G;
    UMT_CUDA(ctx, cudaMemset2DAsync(ctx->psib_buf() + G * ctx->nc, G * ctx->rows * 8, 0, G * ctx->nb * 8, ctx->NA, ctx->stream));
    UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return UMT_OK;
}

extern "C" int umt_init_radiation_field(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  TRY(ensure_state(ctx));
  TRY(finalize_schedule(ctx));
  ctx->pack_valid = ctx->recv_valid = false;
  if (ctx->nExit > 0) {
    const size_t n = (size_t)ctx->nExit * ctx->G;
    exit_copy_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi, ctx->psib_buf(), ctx->d_exitB, ctx->d_exitC,
                                                                      ctx->d_exitA, ctx->nExit, ctx->nc, ctx->rows, ctx->G);
    UMT_CUDA(ctx, cudaGetLastError());
  }
  TRY(seed_cycle_psi(ctx));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// the controller: rt/ControlSweep.F90 -> snac/SetSweep.F90 -> getPhiTotal
// ---------------------------------------------------------------------------

static int sweep_impl(umt_ctx *ctx, int savePsi, int maxFluxIters, double fluxTol, int *itersDone, double *hostPhi);

extern "C" int umt_sweep(umt_ctx *ctx, int savePsi, int maxFluxIters, double fluxTol, int *itersDone) {
  return sweep_impl(ctx, savePsi, maxFluxIters, fluxTol, itersDone, nullptr);
}

// One ControlSweep with host buffers (what the Fortran caller exchanges per call): GSet%Sigt (ngr,nzones) and GSet%STotal
// (ngr,ncornr) in, Rad%PhiTotal (ngr,ncornr) out.  The phi reduction runs in chunks of corners and each chunk goes home on a
// second stream as soon as it is reduced, so the device-to-host copy overlaps the reduction.  Host buffers should be pinned.
extern "C" int umt_control_sweep(umt_ctx *ctx, const double *Sigt, const double *STotal, double tau, int savePsi, int maxFluxIters,
                                 double fluxTol, int *itersDone, double *PhiTotal) {
  if (!ctx || !PhiTotal) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  NEED_DEVICE(ctx, "umt_control_sweep");
  TRY(ensure_state(ctx));
  const size_t G = ctx->G;
  if (Sigt) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sigt, Sigt, sizeof(double) * G * ctx->nz, cudaMemcpyHostToDevice, ctx->stream));
  if (STotal) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_stotal, STotal, sizeof(double) * G * ctx->nc, cudaMemcpyHostToDevice, ctx->stream));
  ctx->tau = tau;
  return sweep_impl(ctx, savePsi, maxFluxIters, fluxTol, itersDone, PhiTotal);
}

static int sweep_impl(umt_ctx *ctx, int savePsi, int maxFluxIters, double fluxTol, int *itersDone, double *hostPhi) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  NEED_DEVICE(ctx, "umt_sweep");
  if (!ctx->have_conn || !ctx->have_geom || !ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_sweep: connectivity, geometry and quadrature must be set");
  if (!ctx->d_psi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_sweep: no state uploaded (umt_upload_state)");
  TRY(finalize_schedule(ctx));
  if (maxFluxIters < 1) maxFluxIters = 1;
  ctx->last_launches = 0;
  float ms_sweep = 0.f, ms_phi = 0.f, ms_exch = 0.f, ms_all = 0.f;
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  int iter = 0;
  const bool multi = !ctx->shared.empty();

  const bool staged = multi && ctx->have_comm_order;   // comm sets of several bins exchange step by step inside umt_launch_sweep3d
  bool overlapped = false;
  // put path: the tally kernel stores the exiting rows straight into the neighbours' receive buffers (peer memory) and its trade of the
  // exit currents tells every domain that its neighbours' rows have arrived: no send buffer, no ncclSend/ncclRecv of the rows
  const bool put = multi && !staged && ctx->put_ready;
  // restoreCommOrder + setIncidentFlux (SetSweep.F90:68-74): packs the exiting rows and tallies the exit currents -- unless the
  // tally that followed the previous sweep still describes the PsiB on the device
  if (multi && !ctx->pack_valid) {
    ctx->put_now = put;
    const int rt = umt_exchange_tally(ctx, fluxTol);
    ctx->put_now = false;
    if (rt) return rt;
    ctx->pack_valid = true; ctx->recv_valid = put;
  }
  for (;;) {
    iter++;
    UMT_TRACE(ctx, "sweep: pass %d (passCount %lld) put %d recv_valid %d pack_valid %d", iter, ctx->passCount, (int)put, (int)ctx->recv_valid, (int)ctx->pack_valid);
    UMT_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    if (multi && !staged) {   // InitExchange/SendFlux/RecvFlux: lagged psib from the previous pass
      if (!ctx->recv_valid) TRY(umt_exchange_rows(ctx));
      TRY(umt_exchange_unpack(ctx));
      ctx->recv_valid = false;
    }
    UMT_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (ctx->totalCycles > 0) {                     // initFromCycleList (meshes with cycle lists keep the legacy layout)
      const size_t n = (size_t)ctx->totalCycles * ctx->G;
      cycle_copy_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_cyclePsi, ctx->d_cycleList,
                                                                         ctx->d_cycleAngle, ctx->totalCycles, ctx->rows, ctx->G, 1);
      ctx->last_launches++;
    }
    if (ctx->ndim == 3) TRY(umt_launch_sweep3d(ctx, savePsi));
    else TRY(umt_launch_sweeprz(ctx, savePsi));
    ctx->passCount++;   // from here on "the next pass" reads the other receive buffer
    if (ctx->totalCycles > 0) {                     // updateCycleList
      const size_t n = (size_t)ctx->totalCycles * ctx->G;
      cycle_copy_kernel<<<(int)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_cyclePsi, ctx->d_cycleList,
                                                                         ctx->d_cycleAngle, ctx->totalCycles, ctx->rows, ctx->G, 0);
      ctx->last_launches++;
    }
    UMT_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    int nNotConv = 0;
    const bool last = savePsi || iter >= maxFluxIters;   // SetSweep.F90:185-187: no further pass whatever the fluxes say
    if (multi && last) {
      // setIncidentFlux + testFluxConv of the last pass, and already the transfer of the rows the NEXT pass will start from (the
      // exchange is lagged one pass), on a stream of their own: they run under the phi tally below.  nNotConv is not needed.
      UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev[3], 0));
      UMT_CUDA(ctx, cudaEventRecord(ctx->evx[0], ctx->stream3));
      ctx->xstream = ctx->stream3;
      ctx->put_now = put;
      int r = umt_exchange_tally(ctx, fluxTol);   // its trade of the exit currents is also what tells me the neighbours' puts are complete
      ctx->put_now = false;
      if (!r && !staged) {
        if (!put) r = umt_exchange_rows(ctx);
        ctx->recv_valid = !r;
      }
      ctx->xstream = nullptr;
      if (r) return r;
      ctx->pack_valid = true;
      UMT_CUDA(ctx, cudaEventRecord(ctx->evx[1], ctx->stream3));
      overlapped = true;
    } else if (multi) {                               // setIncidentFlux + testFluxConv + Allreduce(max nNotConv)
      ctx->put_now = put;
      const int rt = umt_exchange_tally(ctx, fluxTol);
      ctx->put_now = false;
      if (rt) return rt;
      ctx->pack_valid = true;
      if (put) ctx->recv_valid = true;
      TRY(umt_exchange_test_convergence(ctx, &nNotConv));
    }
    UMT_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    if (!last || !multi) UMT_CUDA(ctx, cudaEventSynchronize(ctx->ev[4]));
    if (last) break;
    float t;
    cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2]); ms_exch += t;
    cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]); ms_sweep += t;
    cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[4]); ms_exch += t;
    if (nNotConv == 0) {   // converged early: the next pass's rows can travel under the phi tally as well
      if (multi && !staged && !put) {
        UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream3, ctx->ev[4], 0));
        UMT_CUDA(ctx, cudaEventRecord(ctx->evx[0], ctx->stream3));
        ctx->xstream = ctx->stream3;
        const int r = umt_exchange_rows(ctx);
        ctx->xstream = nullptr;
        if (r) return r;
        ctx->recv_valid = true;
        UMT_CUDA(ctx, cudaEventRecord(ctx->evx[1], ctx->stream3));
        overlapped = true;
      }
      iter = -iter;   // marks "left the loop with the pass already accounted for"
      break;
    }
  }
  const bool timedInLoop = iter < 0;
  if (iter < 0) iter = -iter;

  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
  if (!hostPhi) {
    TRY(launch_phi_range(ctx, savePsi, 0, (size_t)ctx->nc * ctx->G));
  } else {
    const size_t n = (size_t)ctx->nc * ctx->G;
    const int nChunks = n >= (size_t)1 << 22 ? 8 : 1;
    const size_t per = ((n + nChunks - 1) / nChunks + 1) / 2 * 2;   // even: the kernel works on pairs
    for (int k = 0; k < nChunks; k++) {
      const size_t o = (size_t)k * per;
      if (o >= n) break;
      const size_t m = std::min(per, n - o);
      TRY(launch_phi_range(ctx, savePsi, o, m));
      UMT_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
      UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev[7], 0));
      UMT_CUDA(ctx, cudaMemcpyAsync(hostPhi + o, ctx->d_phi + o, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream2));
    }
  }
  if (overlapped) UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evx[1], 0));   // the call ends when the exchange has, too
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
  UMT_TRACE(ctx, "sweep: %d passes enqueued, waiting for the device", iter);
  UMT_CUDA(ctx, cudaEventSynchronize(ctx->ev[6]));
  UMT_TRACE(ctx, "sweep: device done");
  if (hostPhi) UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
  if (!timedInLoop) {   // the last pass (its events were not read inside the loop)
    float t;
    cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2]); ms_exch += t;
    cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]); ms_sweep += t;
  }
  if (overlapped) { float t; cudaEventElapsedTime(&t, ctx->evx[0], ctx->evx[1]); ms_exch += t; }
  cudaEventElapsedTime(&ms_phi, ctx->ev[5], ctx->ev[6]);
  cudaEventElapsedTime(&ms_all, ctx->ev[0], ctx->ev[6]);
  if (savePsi && !ctx->single_psi) {
    // legacy layout: Psi(:,:,Angle) <- Psi1 for every angle without a copy: the buffers trade roles, and the
    // boundary rows (PsiB) move over to the buffer that plays Psi1 from now on.  (Single-psi layout: the sweep wrote Psi in place.)
    std::swap(ctx->d_psi, ctx->d_psi1);
    const size_t G = ctx->G, pitch = G * ctx->rows * 8, off = G * ctx->nc;
    bool anyBad = false;
    for (int a = 0; a < ctx->NA; a++) anyBad = anyBad || ctx->nBad[a] > 0;
    if (anyBad)   // Set%Psi1 persists across the savePsi sweep in the reference: its direct-solve zones (SweepUCBxyz.F90:283-298) iterate on it
      UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi1, pitch, ctx->d_psi, pitch, G * ctx->nc * 8, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->nb > 0)
      UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi1 + off, pitch, ctx->d_psi + off, pitch, G * ctx->nb * 8, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->last_ms[0] = ms_sweep; ctx->last_ms[1] = ms_phi; ctx->last_ms[2] = ms_exch; ctx->last_ms[3] = ms_all;
  if (itersDone) *itersDone = iter;
  return umt_check_abort(ctx, "umt_sweep");
}

// ---------------------------------------------------------------------------------------------------------------------------
// Group sets of one domain, pipelined (umt_control_sweep_sets).  Single-domain contexts only: a pass is then upload -> sweep kernel
// -> phi tally in chunks, each chunk followed by its download, with no host decision in between, so a whole set can be enqueued
// without waiting for it.
// ---------------------------------------------------------------------------------------------------------------------------
static int sweep_enqueue_single(umt_ctx *ctx, int savePsi, double *hostPhi) {
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->last_launches = 0;
  const size_t nCyc = (size_t)ctx->totalCycles * ctx->G;
  if (ctx->totalCycles > 0) {                     // initFromCycleList
    cycle_copy_kernel<<<(int)((nCyc + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_cyclePsi, ctx->d_cycleList, ctx->d_cycleAngle,
                                                                          ctx->totalCycles, ctx->rows, ctx->G, 1);
    ctx->last_launches++;
  }
  if (ctx->ndim == 3) TRY(umt_launch_sweep3d(ctx, savePsi));
  else TRY(umt_launch_sweeprz(ctx, savePsi));
  ctx->passCount++;
  if (ctx->totalCycles > 0) {                     // updateCycleList
    cycle_copy_kernel<<<(int)((nCyc + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_psi1, ctx->d_cyclePsi, ctx->d_cycleList, ctx->d_cycleAngle,
                                                                          ctx->totalCycles, ctx->rows, ctx->G, 0);
    ctx->last_launches++;
  }
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  const size_t n = (size_t)ctx->nc * ctx->G;
  const int nChunks = n >= (size_t)1 << 22 ? 8 : 1;
  const size_t per = ((n + nChunks - 1) / nChunks + 1) / 2 * 2;
  for (int k = 0; k < nChunks; k++) {
    const size_t o = (size_t)k * per;
    if (o >= n) break;
    const size_t m = std::min(per, n - o);
    TRY(launch_phi_range(ctx, savePsi, o, m));
    UMT_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
    UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev[7], 0));
    UMT_CUDA(ctx, cudaMemcpyAsync(hostPhi + o, ctx->d_phi + o, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream2));
  }
  UMT_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));   // the set's kernels are done: the next set's sweep may take the SMs
  return UMT_OK;
}

static int sweep_finish_single(umt_ctx *ctx, int savePsi) {
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  UMT_CUDA(ctx, cudaEventSynchronize(ctx->ev[6]));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream2));
  float ms_sweep = 0.f, ms_phi = 0.f, ms_all = 0.f;
  cudaEventElapsedTime(&ms_sweep, ctx->ev[2], ctx->ev[3]);
  cudaEventElapsedTime(&ms_phi, ctx->ev[3], ctx->ev[6]);
  cudaEventElapsedTime(&ms_all, ctx->ev[0], ctx->ev[6]);
  if (savePsi && !ctx->single_psi) {   // legacy layout: the buffers trade roles (see sweep_impl)
    std::swap(ctx->d_psi, ctx->d_psi1);
    const size_t G = ctx->G, pitch = G * ctx->rows * 8, off = G * ctx->nc;
    bool anyBad = false;
    for (int a = 0; a < ctx->NA; a++) anyBad = anyBad || ctx->nBad[a] > 0;
    if (anyBad) UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi1, pitch, ctx->d_psi, pitch, G * ctx->nc * 8, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->nb > 0)
      UMT_CUDA(ctx, cudaMemcpy2DAsync(ctx->d_psi1 + off, pitch, ctx->d_psi + off, pitch, G * ctx->nb * 8, ctx->NA, cudaMemcpyDeviceToDevice, ctx->stream));
    UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->last_ms[0] = ms_sweep; ctx->last_ms[1] = ms_phi; ctx->last_ms[2] = 0.0; ctx->last_ms[3] = ms_all;
  return umt_check_abort(ctx, "umt_control_sweep_sets");
}

extern "C" int umt_control_sweep_sets(umt_ctx *const *ctxs, int n, const double *const *Sigt, const double *const *STotal, double tau,
                                      int savePsi, int maxFluxIters, double fluxTol, int *itersDone, double *const *PhiTotal) {
  if (!ctxs || n < 1 || !PhiTotal) return UMT_ERR_ARG;
  bool pipeline = true;
  for (int k = 0; k < n; k++) {
    if (!ctxs[k] || !PhiTotal[k]) return UMT_ERR_ARG;
    pipeline = pipeline && ctxs[k]->device >= 0 && ctxs[k]->device == ctxs[0]->device && ctxs[k]->shared.empty();
  }
  if (itersDone) *itersDone = 1;
  if (!pipeline) {   // domains with neighbours: every set runs its own flux iteration and exchange, one after the other
    int worst = 0;
    for (int k = 0; k < n; k++) {
      int it = 0;
      TRY(umt_control_sweep(ctxs[k], Sigt ? Sigt[k] : nullptr, STotal ? STotal[k] : nullptr, tau, savePsi, maxFluxIters, fluxTol, &it, PhiTotal[k]));
      worst = std::max(worst, it);
    }
    if (itersDone) *itersDone = worst;
    return UMT_OK;
  }
  umt_ctx *c0 = ctxs[0];
  UMT_CUDA(c0, cudaSetDevice(c0->device));
  for (int k = 0; k < n; k++) {   // everything that may allocate, build or wait happens before the first copy is queued
    umt_ctx *ctx = ctxs[k];
    if (!ctx->have_conn || !ctx->have_geom || !ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_control_sweep_sets: connectivity, geometry and quadrature must be set");
    TRY(ensure_state(ctx));
    TRY(finalize_schedule(ctx));
  }
  // uploads of all sets on ONE stream, in set order: set 0 has the whole link first, set k+1 arrives while set k is swept
  cudaStream_t up = c0->stream3;
  auto enqueue = [&](int k) -> int {
    umt_ctx *ctx = ctxs[k];
    const size_t G = ctx->G;
    if (Sigt && Sigt[k]) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_sigt, Sigt[k], sizeof(double) * G * ctx->nz, cudaMemcpyHostToDevice, up));
    if (STotal && STotal[k]) UMT_CUDA(ctx, cudaMemcpyAsync(ctx->d_stotal, STotal[k], sizeof(double) * G * ctx->nc, cudaMemcpyHostToDevice, up));
    UMT_CUDA(ctx, cudaEventRecord(ctx->evx[2], up));
    UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evx[2], 0));
    // a persistent sweep kernel owns every SM it gets: the tally of the previous set goes first, or it would wait a whole sweep
    if (k > 0) UMT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctxs[k - 1]->ev[6], 0));
    ctx->tau = tau;
    return sweep_enqueue_single(ctx, savePsi, PhiTotal[k]);
  };
  int rc = UMT_OK, nQueued = 0;
  for (; nQueued < n && rc == UMT_OK; nQueued++) rc = enqueue(nQueued);
  if (rc) nQueued--;   // the set that failed to enqueue is not waited for; the ones before it are brought to a consistent end
  for (int k = 0; k < nQueued; k++) { const int r = sweep_finish_single(ctxs[k], savePsi); if (r && !rc) rc = r; }
  if (rc) cudaStreamSynchronize(up);
  return rc;
}

extern "C" int umt_last_sweep_times(umt_ctx *ctx, double *ms4) {
  if (!ctx || !ms4) return UMT_ERR_ARG;
  for (int i = 0; i < 4; i++) ms4[i] = ctx->last_ms[i];
  return UMT_OK;
}
extern "C" int umt_last_sweep_launches(umt_ctx *ctx, int *n) {
  if (!ctx || !n) return UMT_ERR_ARG;
  *n = ctx->last_launches;
  return UMT_OK;
}
