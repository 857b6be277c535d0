// Boundary-flux (psib) exchange between spatial domains, one domain per context.
//
// Replaces rt/findexit.F90:102-294 (incident/exiting classification of shared boundary
// elements, ListSend / ListRecv), rt/InitExchange.F90, rt/SendFlux.F90:68-80,
// rt/TestSend.F90, rt/RecvFlux.F90:64-78 (pack, persistent MPI send/recv, unpack),
// rt/setIncidentFlux.F90:71-146 (per-bin exit currents exchanged with the neighbours) and
// rt/testFluxConv.F90:55-105 (relative change of the incident currents) plus the
// MPI_Allreduce(max nNotConv) of snac/SetSweep.F90:199.
//
// B200 design: the exiting rows of every angle toward one neighbour are packed by one kernel
// into one contiguous send buffer (the same kernel tallies the exit currents, so PsiB is read
// once), all neighbours are served by a single ncclGroup of ncclSend/ncclRecv over NVLink
// (device to device, no host staging), and one kernel scatters the received rows into the
// incident PsiB rows.  Semantics are the reference's with one angle per comm set: pass k
// sweeps with what the neighbours produced in pass k-1 (lagged one flux pass).
//
// Transports: NCCL (dlopen'ed; one rank per GPU) and an in-process group of contexts
// (several domains in one process, e.g. on one GPU: tests, or more domains than GPUs).
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <mutex>

#include "umt_internal.h"

// ---------------------------------------------------------------------------
// transports
// ---------------------------------------------------------------------------
struct UmtTransport {
  virtual ~UmtTransport() {}
  // send `sendBytes[s]` bytes from sendPtr[s] to neighbour s, receive recvBytes[s] into recvPtr[s]; device pointers
  virtual int exchange(umt_ctx *ctx, const std::vector<const void *> &sendPtr, const std::vector<size_t> &sendBytes,
                       const std::vector<void *> &recvPtr, const std::vector<size_t> &recvBytes) = 0;
  virtual int allreduce_max(umt_ctx *ctx, int *d_value) = 0;   // in place, device int
  virtual int allreduce_f64(umt_ctx *ctx, double *d_vals, int n, int op) = 0;   // in place, device doubles; op 0 sum, 1 max
};

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
      if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) { error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define SYM(field, name) do { *(void **)(&field) = dlsym(h, name); if (!field) { error = std::string("dlsym ") + name; h = nullptr; return false; } } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

#define UMT_NCCL(ctx, call)                                                                                       \
  do {                                                                                                            \
    ncclResult_t _r = (call);                                                                                     \
    if (_r != ncclSuccess) UMT_FAIL(ctx, UMT_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(_r));          \
  } while (0)

// stream the exchange work of a context currently goes to: its main stream, or the second stream while umt_sweep overlaps the
// post-sweep exchange with the phi tally
static inline cudaStream_t xstream(const umt_ctx *ctx) { return ctx->xstream ? ctx->xstream : ctx->stream; }

struct NcclTransport : UmtTransport {
  ncclComm_t comm = nullptr;
  ~NcclTransport() override { if (comm) g_nccl.CommDestroy(comm); }
  int exchange(umt_ctx *ctx, const std::vector<const void *> &sp, const std::vector<size_t> &sb, const std::vector<void *> &rp,
               const std::vector<size_t> &rb) override {
    UMT_NCCL(ctx, g_nccl.GroupStart());
    for (size_t s = 0; s < ctx->shared.size(); s++) {
      const int peer = ctx->shared[s].neighbor;
      if (sb[s]) UMT_NCCL(ctx, g_nccl.Send(sp[s], sb[s], ncclInt8, peer, comm, xstream(ctx)));
      if (rb[s]) UMT_NCCL(ctx, g_nccl.Recv(rp[s], rb[s], ncclInt8, peer, comm, xstream(ctx)));
    }
    UMT_NCCL(ctx, g_nccl.GroupEnd());
    return UMT_OK;
  }
  int allreduce_max(umt_ctx *ctx, int *d_value) override {
    UMT_NCCL(ctx, g_nccl.AllReduce(d_value, d_value, 1, ncclInt32, ncclMax, comm, xstream(ctx)));
    return UMT_OK;
  }
  int allreduce_f64(umt_ctx *ctx, double *d_vals, int n, int op) override {
    UMT_NCCL(ctx, g_nccl.AllReduce(d_vals, d_vals, (size_t)n, ncclFloat64, op ? ncclMax : ncclSum, comm, xstream(ctx)));
    return UMT_OK;
  }
};

// Contexts of one process acting as ranks 0..n-1 (each driven by its own host thread).
struct LocalGroup {
  std::mutex m;
  std::condition_variable cv;
  int n = 0, arrived = 0;
  unsigned long generation = 0;
  std::vector<umt_ctx *> members;
  int redux = 0;
  std::vector<double> fbuf;   // (n ranks, count) staging of allreduce_f64
  int refs = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    const unsigned long gen = generation;
    if (++arrived == n) { arrived = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != gen; });
  }
};

struct LocalTransport : UmtTransport {
  LocalGroup *grp = nullptr;
  std::vector<void *> recvPtrNow;   // where my neighbours must deliver in the current exchange (read by them)
  std::vector<size_t> recvBytesNow;
  ~LocalTransport() override {
    bool last;
    { std::lock_guard<std::mutex> lk(grp->m); last = --grp->refs == 0; }
    if (last) delete grp;
  }
  int exchange(umt_ctx *ctx, const std::vector<const void *> &sp, const std::vector<size_t> &sb, const std::vector<void *> &rp,
               const std::vector<size_t> &rb) override {
    recvPtrNow = rp; recvBytesNow = rb;
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));   // my send buffers are packed, my receive buffers are free
    grp->barrier();
    for (size_t s = 0; s < ctx->shared.size(); s++) {
      umt_ctx *peer = grp->members[ctx->shared[s].neighbor];
      auto *pt = static_cast<LocalTransport *>(peer->transport);
      int t = -1;
      for (size_t k = 0; k < peer->shared.size(); k++)
        if (peer->shared[k].neighbor == ctx->myRank) t = (int)k;
      if (t < 0 || pt->recvBytesNow[t] != sb[s])
        UMT_FAIL(ctx, UMT_ERR_STATE, "local exchange: rank %d and rank %d disagree on the message size of their shared boundary", ctx->myRank, ctx->shared[s].neighbor);
      if (sb[s]) UMT_CUDA(ctx, cudaMemcpyAsync(pt->recvPtrNow[t], sp[s], sb[s], cudaMemcpyDefault, xstream(ctx)));
    }
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
    grp->barrier();
    return UMT_OK;
  }
  int allreduce_max(umt_ctx *ctx, int *d_value) override {
    int v = 0;
    UMT_CUDA(ctx, cudaMemcpyAsync(&v, d_value, sizeof(int), cudaMemcpyDeviceToHost, xstream(ctx)));
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
    { std::lock_guard<std::mutex> lk(grp->m); grp->redux = std::max(grp->redux, v); }
    grp->barrier();
    { std::lock_guard<std::mutex> lk(grp->m); v = grp->redux; }
    grp->barrier();
    if (ctx->myRank == 0) { std::lock_guard<std::mutex> lk(grp->m); grp->redux = 0; }
    grp->barrier();
    UMT_CUDA(ctx, cudaMemcpyAsync(d_value, &v, sizeof(int), cudaMemcpyHostToDevice, xstream(ctx)));
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
    return UMT_OK;
  }
  int allreduce_f64(umt_ctx *ctx, double *d_vals, int n, int op) override {
    std::vector<double> v(n);
    UMT_CUDA(ctx, cudaMemcpyAsync(v.data(), d_vals, sizeof(double) * n, cudaMemcpyDeviceToHost, xstream(ctx)));
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
    {
      std::lock_guard<std::mutex> lk(grp->m);
      if (grp->fbuf.size() < (size_t)grp->n * n) grp->fbuf.resize((size_t)grp->n * n);
    }
    grp->barrier();   // everybody sees the buffer at its final size
    { std::lock_guard<std::mutex> lk(grp->m); std::copy(v.begin(), v.end(), grp->fbuf.begin() + (size_t)ctx->myRank * n); }
    grp->barrier();
    for (int i = 0; i < n; i++) {   // rank order: every rank gets the same bits
      double a = grp->fbuf[i];
      for (int r = 1; r < grp->n; r++) { const double b = grp->fbuf[(size_t)r * n + i]; a = op ? std::max(a, b) : a + b; }
      v[i] = a;
    }
    grp->barrier();   // nobody overwrites the buffer while another rank still reads it
    UMT_CUDA(ctx, cudaMemcpyAsync(d_vals, v.data(), sizeof(double) * n, cudaMemcpyHostToDevice, xstream(ctx)));
    UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
    return UMT_OK;
  }
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// K7a: sendbuf(:, i) <- PsiB(:, b_i, a_i) for the rows of one chunk, and the chunk's share of the exit current
// sum_i w_a (omega_a . A_bdy(b_i)) sum_g PsiB(g, b_i, a)   (SendFlux.F90:68-77 + setIncidentFlux.F90:84-108)
__global__ void __launch_bounds__(256) pack_tally_kernel(const double *__restrict__ psi1, const long long *__restrict__ srcRow,
                                                         const double *__restrict__ coef, const PackChunk *__restrict__ chunks,
                                                         double *__restrict__ sendbuf, double *__restrict__ partial, int G) {
  const PackChunk ch = chunks[blockIdx.x];
  double acc = 0.0;
  const long long n = (long long)(ch.rowEnd - ch.rowBeg) * G;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int r = ch.rowBeg + (int)(i / G), g = (int)(i % G);
    const double v = __ldcg(&psi1[srcRow[r] * G + g]);
    sendbuf[(size_t)r * G + g] = v;   // the local send buffer, or the neighbour's receive buffer (put path: peer memory)
    acc += coef[r] * v;
  }
  __shared__ double red[256];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[ch.slot] = red[0];
}

// exit current per angle = its chunks' partial sums in chunk order (deterministic)
__global__ void tally_finish_kernel(const double *partial, const int *nChunksOfAngle, int maxChunks, double *exitFlux, int NA) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= NA) return;
  double s = 0.0;
  for (int k = 0; k < nChunksOfAngle[a]; k++) s += partial[(size_t)a * maxChunks + k];
  exitFlux[a] = s;
}

// K7b: PsiB(:, b_i, a_i) <- recvbuf(:, i)   (RecvFlux.F90:68-78)
__global__ void __launch_bounds__(256) unpack_kernel(double *__restrict__ psi1, const long long *__restrict__ dstRow,
                                                     const double *__restrict__ recvbuf, long long n, int G) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / G;
  psi1[dstRow[r] * G + (i - r * G)] = recvbuf[i];
}

// setIncidentFlux.F90:128-146 + testFluxConv.F90:55-105.  Comm sets hold binsPerSet consecutive bins (1 by default: the bin's
// weight in its set's total incident flux is then 1); a comm set has converged when every bin that carries more than 0.1 % of the
// set's incident flux changed by no more than tol; nNotConv counts the comm sets that have not.
// incRecv: (nShared, NA) exit currents received from the neighbours.
__global__ void flux_conv_kernel(const double *incRecv, int nShared, int NA, const int *binOfAngle, int nBins, int binsPerSet, double *incFlux,
                                 double *incFluxOld, double tol, double floorFlux, int *nNotConv) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int b = 0; b < nBins; b++) { incFluxOld[b] = incFlux[b]; incFlux[b] = 0.0; }
  for (int s = 0; s < nShared; s++)
    for (int a = 0; a < NA; a++) incFlux[binOfAngle[a]] += incRecv[(size_t)s * NA + a];
  int notConv = 0;
  for (int b0 = 0; b0 < nBins; b0 += binsPerSet) {
    double total = 0.0;
    for (int b = b0; b < b0 + binsPerSet; b++) total += incFlux[b];
    bool conv = true;
    for (int b = b0; b < b0 + binsPerSet; b++) {
      double rel = 0.0;
      if (!(fabs(total) < floorFlux) && total != 0.0) {
        const double weight = incFlux[b] / total;
        if (weight > 0.001) rel = fabs(incFlux[b] - incFluxOld[b]) / incFlux[b];
      }
      if (!(rel <= tol)) conv = false;
    }
    if (!conv) notConv++;
  }
  *nNotConv = notConv;
}

template <class T>
int upload(umt_ctx *ctx, T **d, const std::vector<T> &h) {
  if (*d) { cudaFree(*d); *d = nullptr; }
  UMT_CUDA(ctx, cudaMalloc((void **)d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) UMT_CUDA(ctx, umt_memcpy(ctx, *d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return UMT_OK;
}

// angle sets as the reference forms them without reflecting boundaries (control/decomposeAngleSets.F90:
// 3-D every angle its own set :280-285, 2-D one set per xi-level :150-172): [set0, set1) of angle a
void angle_set_range(const umt_ctx *ctx, int a, int &a0, int &a1) {
  if (ctx->ndim == 3 || ctx->h_level.empty()) { a0 = a; a1 = a + 1; return; }
  a0 = a; a1 = a + 1;
  while (a0 > 0 && ctx->h_level[a0 - 1] == ctx->h_level[a]) a0--;
  while (a1 < ctx->NA && ctx->h_level[a1] == ctx->h_level[a]) a1++;
}

// does *this* rank compute the incident test of angle a on the boundary shared with `neighbor`?
// findexit.F90:137-143: the lower rank takes the first half of the angle set, the higher rank the rest.
bool i_decide(const umt_ctx *ctx, int neighbor, int a) {
  int a0, a1;
  angle_set_range(ctx, a, a0, a1);
  const int half = (a1 - a0) / 2;
  const bool firstHalf = a - a0 < half;
  return ctx->myRank < neighbor ? firstHalf : !firstHalf;
}

int need_abdy(umt_ctx *ctx) {
  if (ctx->have_abdy) return UMT_OK;
  if (!ctx->have_geom || !ctx->have_conn) UMT_FAIL(ctx, UMT_ERR_STATE, "exchange setup needs connectivity and geometry");
  const int nd = ctx->ndim;
  ctx->h_Abdy.assign((size_t)nd * std::max(ctx->nb, 1), 0.0);
  for (int c = 0; c < ctx->nc; c++)
    for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
      const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
      if (v > ctx->nc)
        for (int d = 0; d < nd; d++) ctx->h_Abdy[(size_t)(v - ctx->nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * ctx->maxcf + f) * nd + d];
    }
  ctx->have_abdy = true;
  return UMT_OK;
}

}  // namespace

void umt_exchange_release(umt_ctx *ctx) {
  for (auto &s : ctx->shared) {
    if (s.peer_is_ipc) for (double *q : s.peer_recv) if (q) cudaIpcCloseMemHandle(q);
    s.peer_recv[0] = s.peer_recv[1] = nullptr; s.peer_is_ipc = false;
    if (s.d_recvbuf2) { cudaFree(s.d_recvbuf2); s.d_recvbuf2 = nullptr; }
    void *p[] = {s.d_send_row, s.d_recv_row, s.d_send_coef, s.d_chunks, s.d_partial, s.d_nChunksOfAngle, s.d_sendbuf, s.d_recvbuf,
                 s.d_gsend, s.d_grecv, s.d_gsendbuf, s.d_grecvbuf, s.d_stage_send, s.d_stage_recv};
    for (void *q : p) if (q) cudaFree(q);
    s.d_gsend = s.d_grecv = nullptr; s.d_gsendbuf = s.d_grecvbuf = nullptr; s.d_stage_send = s.d_stage_recv = nullptr;
    s.d_send_row = s.d_recv_row = nullptr; s.d_send_coef = nullptr; s.d_chunks = nullptr; s.d_partial = nullptr;
    s.d_nChunksOfAngle = nullptr; s.d_sendbuf = s.d_recvbuf = nullptr;
  }
  ctx->put_ready = false;
  void *p[] = {ctx->d_exitFlux, ctx->d_incRecv, ctx->d_incFlux, ctx->d_incFluxOld, ctx->d_binOfAngle, ctx->d_nNotConv};
  for (void *q : p) if (q) cudaFree(q);
  ctx->d_exitFlux = ctx->d_incRecv = ctx->d_incFlux = ctx->d_incFluxOld = nullptr; ctx->d_binOfAngle = nullptr; ctx->d_nNotConv = nullptr;
  delete ctx->transport;
  ctx->transport = nullptr;
}

// ---------------------------------------------------------------------------
// setup API
// ---------------------------------------------------------------------------
extern "C" int umt_add_shared_boundary(umt_ctx *ctx, int neighborRank, int firstBdyElem, int nBdyElem) {
  if (!ctx) return UMT_ERR_ARG;
  if (neighborRank < 0 || firstBdyElem < 1 || nBdyElem < 1 || firstBdyElem - 1 + nBdyElem > ctx->nb)
    UMT_FAIL(ctx, UMT_ERR_ARG, "umt_add_shared_boundary: elements %d..%d outside 1..%d", firstBdyElem, firstBdyElem + nBdyElem - 1, ctx->nb);
  for (const auto &s : ctx->shared)
    if (s.neighbor == neighborRank) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_add_shared_boundary: neighbour %d already has a shared boundary", neighborRank);
  SharedBdy s;
  s.neighbor = neighborRank; s.first = firstBdyElem - 1; s.n = nBdyElem;
  ctx->shared.push_back(s);
  ctx->exch_dirty = true;
  ctx->sched_dirty = true;   // exit lists put shared elements last
  return UMT_OK;
}

extern "C" int umt_set_rank(umt_ctx *ctx, int myRank, int nRanks) {
  if (!ctx || myRank < 0 || nRanks < 1 || myRank >= nRanks) return UMT_ERR_ARG;
  ctx->myRank = myRank; ctx->nRanks = nRanks;
  ctx->exch_dirty = true;
  return UMT_OK;
}

// findexit.F90:128-160: sign of omega . A_bdy for the angles this rank decides, (nBdyElem, NA) bytes, 0 elsewhere
extern "C" int umt_get_incident_test(umt_ctx *ctx, int sharedIndex, signed char *incTest) {
  if (!ctx || !incTest || sharedIndex < 0 || sharedIndex >= (int)ctx->shared.size()) return UMT_ERR_ARG;
  if (!ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_get_incident_test: quadrature not set");
  int r = need_abdy(ctx);
  if (r) return r;
  const SharedBdy &s = ctx->shared[sharedIndex];
  const int nd = ctx->ndim;
  std::memset(incTest, 0, (size_t)s.n * ctx->NA);
  for (int a = 0; a < ctx->NA; a++) {
    if (!i_decide(ctx, s.neighbor, a)) continue;
    for (int b = 0; b < s.n; b++) {
      double dot = 0.0;
      for (int d = 0; d < nd; d++) dot += ctx->h_omega[(size_t)a * nd + d] * ctx->h_Abdy[(size_t)(s.first + b) * nd + d];
      incTest[(size_t)a * s.n + b] = dot < 0.0 ? -1 : (dot > 0.0 ? 1 : 0);
    }
  }
  return UMT_OK;
}

// the neighbour's umt_get_incident_test output for the same boundary (its signs are the opposite of mine: findexit.F90:205)
extern "C" int umt_set_incident_test(umt_ctx *ctx, int sharedIndex, const signed char *incTestNeighbor) {
  if (!ctx || !incTestNeighbor || sharedIndex < 0 || sharedIndex >= (int)ctx->shared.size()) return UMT_ERR_ARG;
  SharedBdy &s = ctx->shared[sharedIndex];
  s.incTestR.assign(incTestNeighbor, incTestNeighbor + (size_t)s.n * ctx->NA);
  ctx->exch_dirty = true;
  return UMT_OK;
}

extern "C" int umt_nccl_unique_id(unsigned char *id128) {
  if (!id128) return UMT_ERR_ARG;
  std::lock_guard<std::mutex> lk(g_nccl_mutex);
  if (!g_nccl.load()) return UMT_ERR_NCCL;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return UMT_ERR_NCCL;
  std::memcpy(id128, &id, 128);
  return UMT_OK;
}

extern "C" int umt_set_comm(umt_ctx *ctx, int myRank, int nRanks, const unsigned char *id128) {
  if (!ctx || !id128 || myRank < 0 || myRank >= nRanks) return UMT_ERR_ARG;
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_set_comm: host-only context");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (!g_nccl.load()) UMT_FAIL(ctx, UMT_ERR_NCCL, "NCCL not available: %s", g_nccl.error.c_str());
  }
  delete ctx->transport;
  ctx->transport = nullptr;
  auto *t = new NcclTransport;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&t->comm, nRanks, id, myRank);
  if (r != ncclSuccess) { delete t; UMT_FAIL(ctx, UMT_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); }
  ctx->transport = t;
  ctx->myRank = myRank; ctx->nRanks = nRanks;
  ctx->nccl_comm = t->comm;
  ctx->exch_dirty = true;
  return UMT_OK;
}

// contexts of this process become ranks 0..n-1 of an in-process group (each must then be driven by its own thread)
extern "C" int umt_connect_local(umt_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return UMT_ERR_ARG;
  LocalGroup *g = new LocalGroup;
  g->n = n; g->members.assign(ctxs, ctxs + n); g->refs = n;
  for (int r = 0; r < n; r++) {
    umt_ctx *c = ctxs[r];
    if (!c || c->device < 0) { delete g; return UMT_ERR_ARG; }
    delete c->transport;
    auto *t = new LocalTransport;
    t->grp = g;
    c->transport = t;
    c->myRank = r; c->nRanks = n; c->exch_dirty = true;
  }
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// lists, buffers
// ---------------------------------------------------------------------------
static int exchange_incident_tests(umt_ctx *ctx) {
  // every shared boundary without a neighbour test yet gets it through the transport
  bool missing = false;
  for (auto &s : ctx->shared) missing = missing || s.incTestR.size() != (size_t)s.n * ctx->NA;
  if (!missing) return UMT_OK;
  if (!ctx->transport) UMT_FAIL(ctx, UMT_ERR_STATE, "shared boundaries need umt_set_comm / umt_connect_local, or umt_set_incident_test from the caller");
  const size_t nS = ctx->shared.size();
  std::vector<signed char *> dS(nS, nullptr), dR(nS, nullptr);
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  int rc = UMT_OK;
  for (size_t k = 0; k < nS && !rc; k++) {
    SharedBdy &s = ctx->shared[k];
    const size_t n = (size_t)s.n * ctx->NA;
    std::vector<signed char> mine(n);
    rc = umt_get_incident_test(ctx, (int)k, mine.data());
    if (!rc && (cudaMalloc((void **)&dS[k], n) != cudaSuccess || cudaMalloc((void **)&dR[k], n) != cudaSuccess)) { ctx->err = "cudaMalloc (incident test)"; rc = UMT_ERR_CUDA; }
    if (!rc) umt_memcpy(ctx, dS[k], mine.data(), n, cudaMemcpyHostToDevice);
    sp[k] = dS[k]; rp[k] = dR[k]; sb[k] = rb[k] = n;
  }
  if (!rc) rc = ctx->transport->exchange(ctx, sp, sb, rp, rb);
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "incident test exchange failed"; rc = UMT_ERR_CUDA; }
  for (size_t k = 0; k < nS; k++) {
    if (!rc) {
      SharedBdy &s = ctx->shared[k];
      s.incTestR.resize((size_t)s.n * ctx->NA);
      umt_memcpy(ctx, s.incTestR.data(), dR[k], s.incTestR.size(), cudaMemcpyDeviceToHost);
    }
    if (dS[k]) cudaFree(dS[k]);
    if (dR[k]) cudaFree(dR[k]);
  }
  return rc;
}

// ListSend / ListRecv per (shared boundary, angle): findexit.F90:193-287
static int build_lists(umt_ctx *ctx) {
  int r = need_abdy(ctx);
  if (r) return r;
  const int NA = ctx->NA;
  for (size_t k = 0; k < ctx->shared.size(); k++) {
    SharedBdy &s = ctx->shared[k];
    if (s.incTestR.size() != (size_t)s.n * NA) UMT_FAIL(ctx, UMT_ERR_STATE, "shared boundary %zu: neighbour's incident test missing", k);
    std::vector<signed char> mine((size_t)s.n * NA);
    r = umt_get_incident_test(ctx, (int)k, mine.data());
    if (r) return r;
    s.send_b.assign(NA, {}); s.recv_b.assign(NA, {});
    for (int a = 0; a < NA; a++) {
      const bool me = i_decide(ctx, s.neighbor, a);
      for (int b = 0; b < s.n; b++) {
        const int t = me ? mine[(size_t)a * s.n + b] : -s.incTestR[(size_t)a * s.n + b];
        if (t < 0) s.recv_b[a].push_back(s.first + b);
        else if (t > 0) s.send_b[a].push_back(s.first + b);
      }
    }
  }
  return UMT_OK;
}

extern "C" int umt_get_exchange_counts(umt_ctx *ctx, int sharedIndex, int *nSend /* (NA) */, int *nRecv /* (NA) */) {
  if (!ctx || sharedIndex < 0 || sharedIndex >= (int)ctx->shared.size()) return UMT_ERR_ARG;
  const SharedBdy &s = ctx->shared[sharedIndex];
  if ((int)s.send_b.size() != ctx->NA) UMT_FAIL(ctx, UMT_ERR_STATE, "exchange lists not built (umt_build_exchange)");
  for (int a = 0; a < ctx->NA; a++) {
    if (nSend) nSend[a] = (int)s.send_b[a].size();
    if (nRecv) nRecv[a] = (int)s.recv_b[a].size();
  }
  return UMT_OK;
}

// ListSend(1,:) / ListRecv of one angle (1-based angle, 1-based boundary elements like the reference)
extern "C" int umt_get_exchange_lists(umt_ctx *ctx, int sharedIndex, int angle, int *listSend, int *listRecv) {
  if (!ctx || sharedIndex < 0 || sharedIndex >= (int)ctx->shared.size() || angle < 1 || angle > ctx->NA) return UMT_ERR_ARG;
  const SharedBdy &s = ctx->shared[sharedIndex];
  if ((int)s.send_b.size() != ctx->NA) UMT_FAIL(ctx, UMT_ERR_STATE, "exchange lists not built (umt_build_exchange)");
  if (listSend) for (size_t i = 0; i < s.send_b[angle - 1].size(); i++) listSend[i] = s.send_b[angle - 1][i] + 1;
  if (listRecv) for (size_t i = 0; i < s.recv_b[angle - 1].size(); i++) listRecv[i] = s.recv_b[angle - 1][i] + 1;
  return UMT_OK;
}

// Put path: peer pointers to the neighbours' receive buffers (CUDA IPC between ranks, plain pointers between the domains of one
// process).  pack_tally_kernel then stores the exiting rows straight into the neighbour's buffer -- pack and transfer in one kernel, NVLink
// stores when the neighbour sits on another GPU -- instead of packing a local send buffer that ncclSend/ncclRecv copy across; the
// neighbour's rows and mine are in the same order (findexit.F90 matches the shared elements one by one).  Collective (the handles
// travel through the communicator).  Any failure on any rank leaves put_ready false everywhere and the NCCL path in force.
// Default inside a process; between ranks only with UMT_EXCHANGE_PUT=1 (verified at 2 ranks; a 4-rank run stalled in the bench's
// host-buffer loop and was not resolved).  UMT_EXCHANGE_PUT=0 switches it off everywhere.  (Storing the rows from inside the sweep kernel was measured and dropped: the extra code in the
// zone solve cost 10 % of the sweep, 36.7 -> 40.4 ms at -d 20 -G 128, to save a 0.2 ms pack kernel.)
static int setup_put(umt_ctx *ctx) {
  ctx->put_ready = false;
  const size_t nS = ctx->shared.size();
  bool want = nS > 0 && ctx->transport != nullptr;
  const bool local = dynamic_cast<LocalTransport *>(ctx->transport) != nullptr;
  // between the domains of one process the put path is the default; between ranks (CUDA IPC) it is opt-in, UMT_EXCHANGE_PUT=1
  if (const char *e = getenv("UMT_EXCHANGE_PUT")) want = want && atoi(e) != 0;
  else want = want && local;
  UMT_TRACE(ctx, "setup_put: want %d local %d boundaries %zu", (int)want, (int)local, nS);
  // every rank must take the same decision: the handle exchange below is collective, and `want` only depends on things all ranks share
  if (!want) return UMT_OK;
  bool ok = true;
  if (auto *lt = dynamic_cast<LocalTransport *>(ctx->transport)) {
    lt->grp->barrier();   // every member has allocated its buffers
    for (size_t k = 0; k < nS; k++) {
      umt_ctx *peer = lt->grp->members[ctx->shared[k].neighbor];
      int t = -1;
      for (size_t j = 0; j < peer->shared.size(); j++) if (peer->shared[j].neighbor == ctx->myRank) t = (int)j;
      if (t < 0 || peer->device != ctx->device) { ok = false; continue; }   // (in-process domains on different GPUs keep the copy path)
      ctx->shared[k].peer_recv[0] = peer->shared[t].d_recvbuf; ctx->shared[k].peer_recv[1] = peer->shared[t].d_recvbuf2;
      ctx->shared[k].peer_is_ipc = false;
    }
    lt->grp->barrier();
  } else {
    // CUDA IPC handles of my two receive buffers go to the neighbour that will write them
    std::vector<cudaIpcMemHandle_t> mine(2 * nS), theirs(2 * nS);
    unsigned char *dS = nullptr, *dR = nullptr;
    const size_t hb = sizeof(cudaIpcMemHandle_t);
    if (cudaMalloc((void **)&dS, 2 * nS * hb) != cudaSuccess || cudaMalloc((void **)&dR, 2 * nS * hb) != cudaSuccess) { cudaGetLastError(); ok = false; }
    for (size_t k = 0; k < nS && ok; k++) {
      if (cudaIpcGetMemHandle(&mine[2 * k], ctx->shared[k].d_recvbuf) != cudaSuccess || cudaIpcGetMemHandle(&mine[2 * k + 1], ctx->shared[k].d_recvbuf2) != cudaSuccess) { cudaGetLastError(); ok = false; }
    }
    if (!ok) std::memset(mine.data(), 0, mine.size() * hb);   // still take part in the collective below
    if (dS && dR) {
      umt_memcpy(ctx, dS, mine.data(), 2 * nS * hb, cudaMemcpyHostToDevice);
      std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS, 2 * hb), rb(nS, 2 * hb);
      for (size_t k = 0; k < nS; k++) { sp[k] = dS + 2 * k * hb; rp[k] = dR + 2 * k * hb; }
      UMT_TRACE(ctx, "setup_put: trading %zu handles", 2 * nS);
      if (ctx->transport->exchange(ctx, sp, sb, rp, rb) != UMT_OK || cudaStreamSynchronize(ctx->stream) != cudaSuccess) ok = false;
      else umt_memcpy(ctx, theirs.data(), dR, 2 * nS * hb, cudaMemcpyDeviceToHost);
      UMT_TRACE(ctx, "setup_put: handles traded ok %d", (int)ok);
    }
    if (dS) cudaFree(dS);
    if (dR) cudaFree(dR);
    const cudaIpcMemHandle_t zero{};
    for (size_t k = 0; k < nS && ok; k++)
      for (int j = 0; j < 2; j++) {
        void *q = nullptr;
        if (!std::memcmp(&theirs[2 * k + j], &zero, hb) || cudaIpcOpenMemHandle(&q, theirs[2 * k + j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        ctx->shared[k].peer_recv[j] = static_cast<double *>(q);
        ctx->shared[k].peer_is_ipc = true;
      }
    UMT_TRACE(ctx, "setup_put: handles opened ok %d", (int)ok);
    // all ranks use the put path or none does (a rank that failed would otherwise wait for rows nobody sends)
    int *d_ok = nullptr;
    int h_ok = ok ? 0 : 1;
    if (cudaMalloc((void **)&d_ok, sizeof(int)) == cudaSuccess) {
      umt_memcpy(ctx, d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice);
      if (ctx->transport->allreduce_max(ctx, d_ok) == UMT_OK) umt_memcpy(ctx, &h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
      else h_ok = 1;
      cudaFree(d_ok);
    } else h_ok = 1;
    ok = h_ok == 0;
  }
  if (auto *lt = dynamic_cast<LocalTransport *>(ctx->transport)) {   // same all-or-none rule inside a process
    { std::lock_guard<std::mutex> lk(lt->grp->m); if (!ok) lt->grp->redux = 1; }
    lt->grp->barrier();
    { std::lock_guard<std::mutex> lk(lt->grp->m); ok = lt->grp->redux == 0; }
    lt->grp->barrier();
    if (ctx->myRank == 0) { std::lock_guard<std::mutex> lk(lt->grp->m); lt->grp->redux = 0; }
    lt->grp->barrier();
  }
  ctx->put_ready = ok;
  UMT_TRACE(ctx, "setup_put: put_ready %d", (int)ok);
  return UMT_OK;
}

extern "C" int umt_build_exchange(umt_ctx *ctx) {
  if (!ctx) return UMT_ERR_ARG;
  if (!ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_build_exchange: quadrature not set");
  if (ctx->shared.empty()) { ctx->exch_dirty = false; return UMT_OK; }
  if (ctx->device >= 0) UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  int r = UMT_OK;
  if (ctx->device >= 0) r = exchange_incident_tests(ctx);
  if (!r) r = build_lists(ctx);
  if (r) return r;
  if (ctx->device < 0) { ctx->exch_dirty = false; return UMT_OK; }   // host-only: lists only
  const int NA = ctx->NA, nd = ctx->ndim, G = ctx->G;
  const int ROWS_PER_CHUNK = std::max(1, 4096 / G);
  for (auto &s : ctx->shared) {
    std::vector<long long> srow, rrow;
    std::vector<double> coef;
    std::vector<PackChunk> chunks;
    std::vector<int> nChunksOfAngle(NA, 0);
    int maxChunks = 1;
    for (int a = 0; a < NA; a++) maxChunks = std::max(maxChunks, ((int)s.send_b[a].size() + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK);
    s.send_off.assign(NA + 1, 0); s.recv_off.assign(NA + 1, 0);
    for (int a = 0; a < NA; a++) {
      const double *om = &ctx->h_omega[(size_t)a * nd];
      const int beg = (int)srow.size();
      for (int b : s.send_b[a]) {
        double dot = 0.0;
        for (int d = 0; d < nd; d++) dot += om[d] * ctx->h_Abdy[(size_t)b * nd + d];
        srow.push_back((long long)a * ctx->rows_total() + ctx->nc + b);
        coef.push_back(ctx->h_weight[a] * dot);
      }
      for (int b : s.recv_b[a]) rrow.push_back((long long)a * ctx->rows_total() + ctx->nc + b);
      const int end = (int)srow.size();
      for (int c0 = beg, k = 0; c0 < end; c0 += ROWS_PER_CHUNK, k++) {
        chunks.push_back({c0, std::min(end, c0 + ROWS_PER_CHUNK), a * maxChunks + k, 0});
        nChunksOfAngle[a] = k + 1;
      }
      s.send_off[a + 1] = (int)srow.size(); s.recv_off[a + 1] = (int)rrow.size();
    }
    s.send_rows = srow.size(); s.recv_rows = rrow.size();
    s.nChunks = (int)chunks.size(); s.maxChunks = maxChunks;
    if ((r = upload(ctx, &s.d_send_row, srow))) return r;
    if ((r = upload(ctx, &s.d_recv_row, rrow))) return r;
    if ((r = upload(ctx, &s.d_send_coef, coef))) return r;
    if ((r = upload(ctx, &s.d_chunks, chunks))) return r;
    if ((r = upload(ctx, &s.d_nChunksOfAngle, nChunksOfAngle))) return r;
    if (s.d_partial) cudaFree(s.d_partial);
    if (s.d_sendbuf) cudaFree(s.d_sendbuf);
    if (s.d_recvbuf) cudaFree(s.d_recvbuf);
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_partial, sizeof(double) * (size_t)NA * maxChunks));
    UMT_CUDA(ctx, cudaMemsetAsync(s.d_partial, 0, sizeof(double) * (size_t)NA * maxChunks, ctx->stream));
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_sendbuf, sizeof(double) * std::max<size_t>(s.send_rows * G, 1)));
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_recvbuf, sizeof(double) * std::max<size_t>(s.recv_rows * G, 1)));
    if (s.peer_is_ipc) for (double *q : s.peer_recv) if (q) cudaIpcCloseMemHandle(q);
    s.peer_recv[0] = s.peer_recv[1] = nullptr; s.peer_is_ipc = false;
    if (s.d_recvbuf2) { cudaFree(s.d_recvbuf2); s.d_recvbuf2 = nullptr; }
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_recvbuf2, sizeof(double) * std::max<size_t>(s.recv_rows * G, 1)));
  }
  const size_t nS = ctx->shared.size();
  // flux-convergence bins: one per comm set (3-D: angle, 2-D: xi-level)
  std::vector<int> bin(NA);
  for (int a = 0; a < NA; a++) bin[a] = nd == 3 ? a : ctx->h_level[a];
  ctx->nBins = nd == 3 ? NA : ctx->nLevels;
  if ((r = upload(ctx, &ctx->d_binOfAngle, bin))) return r;
  void **arrs[] = {(void **)&ctx->d_exitFlux, (void **)&ctx->d_incRecv, (void **)&ctx->d_incFlux, (void **)&ctx->d_incFluxOld};
  const size_t sizes[] = {nS * NA, nS * NA, (size_t)ctx->nBins, (size_t)ctx->nBins};
  for (int i = 0; i < 4; i++) {
    if (*arrs[i]) cudaFree(*arrs[i]);
    UMT_CUDA(ctx, cudaMalloc(arrs[i], sizeof(double) * std::max<size_t>(sizes[i], 1)));
    UMT_CUDA(ctx, cudaMemsetAsync(*arrs[i], 0, sizeof(double) * std::max<size_t>(sizes[i], 1), ctx->stream));
  }
  if (!ctx->d_nNotConv) UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_nNotConv, sizeof(int)));
  ctx->exch_dirty = false;
  ctx->pack_valid = ctx->recv_valid = false;
  ctx->passCount = 0;
  return setup_put(ctx);
}

// ---------------------------------------------------------------------------
// per-pass pieces used by umt_sweep
// ---------------------------------------------------------------------------
static int ready(umt_ctx *ctx) {
  if (ctx->exch_dirty) { int r = umt_build_exchange(ctx); if (r) return r; }
  if (!ctx->transport) UMT_FAIL(ctx, UMT_ERR_STATE, "domain has shared boundaries but no communicator (umt_set_comm / umt_connect_local)");
  return UMT_OK;
}

// pack the exiting rows of every angle for every neighbour and tally the exit currents; then trade the
// currents with the neighbours and update IncFlux / IncFluxOld (setIncidentFlux)
int umt_exchange_tally(umt_ctx *ctx, double tol) {
  int r = ready(ctx);
  if (r) return r;
  const int NA = ctx->NA, G = ctx->G;
  const size_t nS = ctx->shared.size();
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    if (s.nChunks > 0) {
      // put path: the rows the neighbour's NEXT pass to be swept starts from go straight into the buffer it will unpack then
      double *dst = ctx->put_now ? s.peer_recv[(ctx->passCount + 1) & 1] : s.d_sendbuf;
      pack_tally_kernel<<<s.nChunks, 256, 0, xstream(ctx)>>>(ctx->psib_buf(), s.d_send_row, s.d_send_coef, s.d_chunks, dst, s.d_partial, G);
      ctx->last_launches++;
    }
    tally_finish_kernel<<<(NA + 127) / 128, 128, 0, xstream(ctx)>>>(s.d_partial, s.d_nChunksOfAngle, s.maxChunks, ctx->d_exitFlux + k * NA, NA);
    ctx->last_launches++;
    sp[k] = ctx->d_exitFlux + k * NA; rp[k] = ctx->d_incRecv + k * NA; sb[k] = rb[k] = sizeof(double) * NA;
  }
  UMT_CUDA(ctx, cudaGetLastError());
  UMT_TRACE(ctx, "tally: packed (put %d, passCount %lld), trading currents with %zu neighbours", (int)ctx->put_now, ctx->passCount, nS);
  r = ctx->transport->exchange(ctx, sp, sb, rp, rb);
  UMT_TRACE(ctx, "tally: trade enqueued rc %d", r);
  if (r) return r;
  const int binsPerSet = ctx->nCommSets > 0 ? ctx->nBins / ctx->nCommSets : 1;
  flux_conv_kernel<<<1, 32, 0, xstream(ctx)>>>(ctx->d_incRecv, (int)ctx->shared.size(), ctx->NA, ctx->d_binOfAngle, ctx->nBins, binsPerSet, ctx->d_incFlux,
                                              ctx->d_incFluxOld, tol, ctx->fluxFloor, ctx->d_nNotConv);
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches++;
  return UMT_OK;
}

// The receive buffer that holds (or will hold) the rows the NEXT pass to be swept starts from: pass n (= passCount) reads buffer
// (n + 1) & 1, and while it is swept the neighbours' put stores fill buffer n & 1 for pass n + 1.
double *umt_recv_buffer(const umt_ctx *ctx, const SharedBdy &s) { return ((ctx->passCount + 1) & 1) ? s.d_recvbuf2 : s.d_recvbuf; }

// SendFlux / RecvFlux for every angle, first half: the rows every domain packed after its last sweep travel to the neighbours'
// receive buffers (collective).  umt_sweep issues this right after the post-sweep tally, on the second stream, so that the
// transfer of the NEXT pass's incident rows overlaps the phi tally of this one (the exchange is lagged one pass anyway).
int umt_exchange_rows(umt_ctx *ctx) {
  int r = ready(ctx);
  if (r) return r;
  const int G = ctx->G;
  const size_t nS = ctx->shared.size();
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    sp[k] = s.d_sendbuf; rp[k] = umt_recv_buffer(ctx, s);
    sb[k] = sizeof(double) * s.send_rows * G; rb[k] = sizeof(double) * s.recv_rows * G;
  }
  return ctx->transport->exchange(ctx, sp, sb, rp, rb);
}

// second half: the received rows land in my incident PsiB rows (RecvFlux.F90:68-78)
int umt_exchange_unpack(umt_ctx *ctx) {
  const int G = ctx->G;
  const size_t nS = ctx->shared.size();
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    const long long n = (long long)s.recv_rows * G;
    if (n > 0) {
      unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, xstream(ctx)>>>(ctx->psib_buf(), s.d_recv_row, umt_recv_buffer(ctx, s), n, G);
      ctx->last_launches++;
    }
  }
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}

// Allreduce(max) of the nNotConv the last umt_exchange_tally left on the device (SetSweep.F90:189-199)
int umt_exchange_test_convergence(umt_ctx *ctx, int *nNotConv) {
  int r = ctx->transport->allreduce_max(ctx, ctx->d_nNotConv);
  if (r) return r;
  UMT_CUDA(ctx, cudaMemcpyAsync(nNotConv, ctx->d_nNotConv, sizeof(int), cudaMemcpyDeviceToHost, xstream(ctx)));
  UMT_CUDA(ctx, cudaStreamSynchronize(xstream(ctx)));
  return UMT_OK;
}

extern "C" int umt_get_incident_flux(umt_ctx *ctx, double *incFlux, double *incFluxOld) {
  if (!ctx) return UMT_ERR_ARG;
  if (!ctx->d_incFlux) UMT_FAIL(ctx, UMT_ERR_STATE, "no exchange state");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (incFlux) UMT_CUDA(ctx, umt_memcpy(ctx, incFlux, ctx->d_incFlux, sizeof(double) * ctx->nBins, cudaMemcpyDeviceToHost));
  if (incFluxOld) UMT_CUDA(ctx, umt_memcpy(ctx, incFluxOld, ctx->d_incFluxOld, sizeof(double) * ctx->nBins, cudaMemcpyDeviceToHost));
  return UMT_OK;
}

extern "C" int umt_set_flux_floor(umt_ctx *ctx, double floorFlux) {
  if (!ctx) return UMT_ERR_ARG;
  ctx->fluxFloor = floorFlux;
  return UMT_OK;
}

// ---------------------------------------------------------------------------
// grey (GTA) exchange: GTASweep.F90:139-146 with the restored communication order -- every angle's exiting PsiB
// elements go to the neighbour's incident elements before the grey sweeps (lagged one grey sweep, one double per row)
// ---------------------------------------------------------------------------
namespace {
__global__ void gather_kernel(const double *__restrict__ src, const int *__restrict__ idx, double *__restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[idx[i]];
}
__global__ void scatter_kernel(double *__restrict__ dst, const int *__restrict__ idx, const double *__restrict__ in, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = in[i];
}
}  // namespace

// findexit.F90:128-287 for the GTA angle set (every S2 angle its own angle set without reflecting boundaries, so of each
// pair of ranks the higher one classifies and the lower one negates what it receives).  Collective over the domains.
int umt_gta_build_exchange(umt_ctx *ctx) {
  if (ctx->shared.empty()) return UMT_OK;
  if (!ctx->transport) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA on a decomposed mesh needs a communicator (umt_set_comm / umt_connect_local) before umt_gta_setup");
  int r = need_abdy(ctx);
  if (r) return r;
  const GtaState &g = ctx->gta;
  const int nA = g.nAng, nd = ctx->ndim, nb = ctx->nb;
  const size_t nS = ctx->shared.size();
  std::vector<std::vector<signed char>> mine(nS), theirs(nS);
  std::vector<signed char *> dS(nS, nullptr), dR(nS, nullptr);
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  int rc = UMT_OK;
  for (size_t k = 0; k < nS && !rc; k++) {
    const SharedBdy &s = ctx->shared[k];
    const size_t n = (size_t)s.n * nA;
    mine[k].assign(n, 0);
    if (ctx->myRank > s.neighbor)
      for (int a = 0; a < nA; a++)
        for (int b = 0; b < s.n; b++) {
          double dot = 0.0;
          for (int d = 0; d < nd; d++) dot += g.omega[(size_t)a * nd + d] * ctx->h_Abdy[(size_t)(s.first + b) * nd + d];
          mine[k][(size_t)a * s.n + b] = dot < 0.0 ? -1 : (dot > 0.0 ? 1 : 0);
        }
    if (cudaMalloc((void **)&dS[k], n) != cudaSuccess || cudaMalloc((void **)&dR[k], n) != cudaSuccess) { ctx->err = "cudaMalloc (GTA incident test)"; rc = UMT_ERR_CUDA; }
    if (!rc) umt_memcpy(ctx, dS[k], mine[k].data(), n, cudaMemcpyHostToDevice);
    sp[k] = dS[k]; rp[k] = dR[k]; sb[k] = rb[k] = n;
  }
  if (!rc) rc = ctx->transport->exchange(ctx, sp, sb, rp, rb);
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "GTA incident test exchange failed"; rc = UMT_ERR_CUDA; }
  for (size_t k = 0; k < nS; k++) {
    if (!rc) { theirs[k].resize(mine[k].size()); umt_memcpy(ctx, theirs[k].data(), dR[k], theirs[k].size(), cudaMemcpyDeviceToHost); }
    if (dS[k]) cudaFree(dS[k]);
    if (dR[k]) cudaFree(dR[k]);
  }
  if (rc) return rc;
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    const bool me = ctx->myRank > s.neighbor;
    std::vector<int> snd, rcv;
    for (int a = 0; a < nA; a++)
      for (int b = 0; b < s.n; b++) {
        const int t = me ? mine[k][(size_t)a * s.n + b] : -theirs[k][(size_t)a * s.n + b];
        if (t < 0) rcv.push_back(a * nb + s.first + b);
        else if (t > 0) snd.push_back(a * nb + s.first + b);
      }
    s.gsend_n = snd.size(); s.grecv_n = rcv.size();
    if ((r = upload(ctx, &s.d_gsend, snd))) return r;
    if ((r = upload(ctx, &s.d_grecv, rcv))) return r;
    if (s.d_gsendbuf) cudaFree(s.d_gsendbuf);
    if (s.d_grecvbuf) cudaFree(s.d_grecvbuf);
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_gsendbuf, sizeof(double) * std::max<size_t>(s.gsend_n, 1)));
    UMT_CUDA(ctx, cudaMalloc((void **)&s.d_grecvbuf, sizeof(double) * std::max<size_t>(s.grecv_n, 1)));
  }
  return UMT_OK;
}

int umt_gta_exchange(umt_ctx *ctx, double *d_PsiB) {
  if (ctx->shared.empty()) return UMT_OK;
  const size_t nS = ctx->shared.size();
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    if (!s.d_gsend) UMT_FAIL(ctx, UMT_ERR_STATE, "GTA exchange lists not built (umt_gta_setup after the shared boundaries and the communicator)");
    if (s.gsend_n) gather_kernel<<<(unsigned)((s.gsend_n + 255) / 256), 256, 0, ctx->stream>>>(d_PsiB, s.d_gsend, s.d_gsendbuf, (int)s.gsend_n);
    sp[k] = s.d_gsendbuf; rp[k] = s.d_grecvbuf; sb[k] = sizeof(double) * s.gsend_n; rb[k] = sizeof(double) * s.grecv_n;
  }
  UMT_CUDA(ctx, cudaGetLastError());
  int r = ctx->transport->exchange(ctx, sp, sb, rp, rb);
  if (r) return r;
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    if (s.grecv_n) scatter_kernel<<<(unsigned)((s.grecv_n + 255) / 256), 256, 0, ctx->stream>>>(d_PsiB, s.d_grecv, s.d_grecvbuf, (int)s.grecv_n);
  }
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}

int umt_allreduce_f64(umt_ctx *ctx, double *d_vals, int n, int op) {
  if (ctx->nRanks <= 1 || !ctx->transport) return UMT_OK;
  return ctx->transport->allreduce_f64(ctx, d_vals, n, op);
}

// ---------------------------------------------------------------------------
// SweepScheduler (rt/SweepScheduler.F90:32-313) + setNetFlux (rt/setNetFlux.F90:9-143) and the per-step exchange of
// snac/SetSweep.F90:113-170 for comm sets that hold several angle bins
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) gather_rows_kernel(const double *__restrict__ psi1, const long long *__restrict__ row, double *__restrict__ out,
                                                          long long n, int G) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / G;
  out[i] = __ldcg(&psi1[row[r] * G + (i - r * G)]);
}
}  // namespace

// Angle bins of the scheduler (rt/SweepScheduler.F90:110-117: "an angle-bin is a xi-level in 2d and an angle in 3D"): bin of every
// angle and the angles of every bin in sweep order.
static void angle_bins(const umt_ctx *ctx, int &nBins, std::vector<int> &binOf, std::vector<std::vector<int>> &anglesOf) {
  const int NA = ctx->NA;
  binOf.assign(NA, 0);
  if (ctx->ndim == 3) { nBins = NA; for (int a = 0; a < NA; a++) binOf[a] = a; }
  else { nBins = std::max(ctx->nLevels, 1); for (int a = 0; a < NA; a++) binOf[a] = ctx->h_level[a]; }
  anglesOf.assign(nBins, {});
  for (int a = 0; a < NA; a++) anglesOf[binOf[a]].push_back(a);
}

// nCommSets consecutive groups of angle bins (3-D: bins are angles, r-z: xi-levels); 0 restores the finest decomposition
extern "C" int umt_set_comm_sets(umt_ctx *ctx, int nCommSets) {
  if (!ctx) return UMT_ERR_ARG;
  if (!ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_set_comm_sets: quadrature not set");
  const int nBinsAll = ctx->ndim == 3 ? ctx->NA : std::max(ctx->nLevels, 1);
  if (nCommSets < 0 || (nCommSets > 0 && nBinsAll % nCommSets != 0)) UMT_FAIL(ctx, UMT_ERR_ARG, "umt_set_comm_sets: %d sets do not divide %d angle bins", nCommSets, nBinsAll);
  if (nCommSets > 0 && nCommSets < nBinsAll && ctx->ndim != 3 && !ctx->refl.empty())
    UMT_FAIL(ctx, UMT_ERR_STATE, "umt_set_comm_sets: r-z comm sets of several xi-levels are not supported together with reflecting boundaries");
  ctx->nCommSets = nCommSets == nBinsAll ? 0 : nCommSets;
  ctx->have_comm_order = false;
  ctx->sched_dirty = true;
  return UMT_OK;
}

// The order in which every comm set sweeps its bins and the order in which the neighbours sweep theirs.  Collective.
// netFlux (nShared, NA) = exiting minus incident current per shared boundary and angle; NULL: tallied from the PsiB on the
// device (setNetFlux).  Dependency weight of a bin = sum over the shared boundaries of its net flux, minus what neighbours
// have already swept; bins whose mirror images (reflecting boundaries) are still to come wait; largest weight first.
extern "C" int umt_sweep_scheduler(umt_ctx *ctx, const double *netFlux) {
  if (!ctx) return UMT_ERR_ARG;
  if (ctx->nCommSets <= 0) { ctx->have_comm_order = false; return UMT_OK; }   // one bin per comm set: identity
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_sweep_scheduler: host-only context");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int NA = ctx->NA, nC = ctx->nCommSets;
  int nBins = 0;
  std::vector<int> binOf;
  std::vector<std::vector<int>> anglesOf;
  angle_bins(ctx, nBins, binOf, anglesOf);
  const int bps = nBins / nC;           // bins per comm set
  const size_t nS = ctx->shared.size();
  if (nS > 0) { int r = ready(ctx); if (r) return r; }
  int r = umt_reflect_stages(ctx);   // mirror angles
  if (r) return r;
  // weightComm(shared, bin): rows of NA entries, the first nBins used (3-D: nBins == NA)
  std::vector<double> w(std::max<size_t>(nS, 1) * NA, 1.0);
  if (nS > 0) {
    if (netFlux) std::copy(netFlux, netFlux + nS * NA, w.begin());
    else {
      if (!ctx->d_psi) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_sweep_scheduler: no state on the device to tally the net flux from");
      r = umt_exchange_tally(ctx, 0.0);
      if (r) return r;
      ctx->pack_valid = true; ctx->recv_valid = false;
      std::vector<double> ex(nS * NA), in(nS * NA);
      UMT_CUDA(ctx, cudaMemcpyAsync(ex.data(), ctx->d_exitFlux, sizeof(double) * nS * NA, cudaMemcpyDeviceToHost, ctx->stream));
      UMT_CUDA(ctx, cudaMemcpyAsync(in.data(), ctx->d_incRecv, sizeof(double) * nS * NA, cudaMemcpyDeviceToHost, ctx->stream));
      UMT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      std::fill(w.begin(), w.end(), 0.0);
      for (size_t k = 0; k < nS; k++)   // setNetFlux.F90:61-141: exit minus incident current per shared boundary and bin
        for (int a = 0; a < NA; a++) w[k * NA + binOf[a]] += ex[k * NA + a] - in[k * NA + a];
    }
  }
  ctx->netFlux.assign(w.begin(), w.begin() + nS * NA);
  std::vector<double> depend(nBins, 0.0);
  for (int b = 0; b < nBins; b++)
    for (size_t k = 0; k < nS; k++) depend[b] += w[k * NA + b];
  const int nR = (int)ctx->refl.size();
  std::vector<int> nRefl(nBins, 0), depAngle((size_t)std::max(nR, 1) * NA, -1);
  for (int n = 0; n < nR; n++)
    for (int a = 0; a < NA; a++) {
      const int m = ctx->refl[n].mref[a];
      if (m >= 0) { depAngle[(size_t)n * NA + m] = a; nRefl[binOf[a]]++; }
    }
  std::vector<unsigned char> notDone(nBins, 1);
  std::vector<int> binOrder(nBins, 0);                                // per comm set concatenated: bin swept at each step
  std::vector<std::vector<int>> binRecvOrder(nS, std::vector<int>(nBins, 0));
  int *d_s = nullptr, *d_r = nullptr;
  UMT_CUDA(ctx, cudaMalloc((void **)&d_s, sizeof(int) * nC));
  UMT_CUDA(ctx, cudaMalloc((void **)&d_r, sizeof(int) * nC * std::max<size_t>(nS, 1)));
  std::vector<int> newbin(nC), binRecv(nC * std::max<size_t>(nS, 1));
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS, sizeof(int) * nC), rb(nS, sizeof(int) * nC);
  int rc = UMT_OK;
  for (int i = 0; i < bps && !rc; i++) {
    for (int c = 0; c < nC; c++) {
      const int b0 = c * bps, b1 = b0 + bps;
      int imin = -1;
      for (int b = b0; b < b1; b++) if (notDone[b] && (imin < 0 || nRefl[b] < nRefl[imin])) imin = b;   // minloc(nRefl, notDone)
      if (nRefl[imin] != 0) nRefl[imin] = 0;
      int best = -1;
      for (int b = b0; b < b1; b++) if (notDone[b] && nRefl[b] == 0 && (best < 0 || depend[b] > depend[best])) best = b;   // maxloc(depend, Ready)
      newbin[c] = best;
    }
    if (nS > 0) {   // every neighbour learns my bins of this step, I learn theirs
      if (cudaMemcpyAsync(d_s, newbin.data(), sizeof(int) * nC, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { rc = UMT_ERR_CUDA; break; }
      for (size_t k = 0; k < nS; k++) { sp[k] = d_s; rp[k] = d_r + k * nC; }
      rc = ctx->transport->exchange(ctx, sp, sb, rp, rb);
      if (rc) break;
      if (cudaMemcpyAsync(binRecv.data(), d_r, sizeof(int) * nC * nS, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize(ctx->stream) != cudaSuccess) { rc = UMT_ERR_CUDA; break; }
    }
    for (int c = 0; c < nC; c++) {
      const int nb_ = newbin[c];
      binOrder[c * bps + i] = nb_;
      for (int n = 0; n < nR; n++)
        for (int a : anglesOf[nb_]) {
          const int aRef = depAngle[(size_t)n * NA + a];
          if (aRef >= 0 && notDone[binOf[aRef]]) nRefl[binOf[aRef]]--;
        }
      notDone[nb_] = 0;
    }
    for (size_t k = 0; k < nS; k++)
      for (int c = 0; c < nC; c++) {
        const int b = binRecv[k * nC + c];
        if (b < c * bps || b >= (c + 1) * bps) { ctx->err = "umt_sweep_scheduler: neighbour sent a bin outside the comm set"; rc = UMT_ERR_STATE; break; }
        binRecvOrder[k][c * bps + i] = b;
        if (notDone[b]) depend[b] -= w[k * NA + b];
      }
    // (a failure leaves the step loop through its !rc condition; the neighbours fail on their next exchange with this rank)
  }
  cudaFree(d_s); cudaFree(d_r);
  if (rc) { if (rc == UMT_ERR_CUDA) ctx->err = "umt_sweep_scheduler: CUDA copy failed"; return rc; }
  // CSet%AngleOrder / RecvOrder: the angles of the chosen bins in bin order (SweepScheduler.F90:269-300), and the step of every angle
  ctx->angleOrder.clear();
  ctx->recvOrder.assign(nS, {});
  ctx->commStageOf.assign(NA, 0);
  ctx->binOrder = binOrder; ctx->binRecvOrder = binRecvOrder; ctx->nSchedBins = nBins; ctx->binsPerSet = bps;
  for (int c = 0; c < nC; c++)
    for (int i = 0; i < bps; i++) {
      for (int a : anglesOf[binOrder[c * bps + i]]) { ctx->angleOrder.push_back(a); ctx->commStageOf[a] = i; }
      for (size_t k = 0; k < nS; k++) for (int a : anglesOf[binRecvOrder[k][c * bps + i]]) ctx->recvOrder[k].push_back(a);
    }
  ctx->have_comm_order = true;
  ctx->sched_dirty = true;
  return nS > 0 ? umt_exchange_build_stages(ctx) : UMT_OK;
}

// CSet%NetFlux(shared, bin) the last umt_sweep_scheduler call worked with
extern "C" int umt_get_net_flux(umt_ctx *ctx, double *netFlux /* (nShared, NA) */) {
  if (!ctx || !netFlux) return UMT_ERR_ARG;
  if (ctx->netFlux.size() != ctx->shared.size() * (size_t)ctx->NA) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_get_net_flux: no scheduler run yet");
  std::copy(ctx->netFlux.begin(), ctx->netFlux.end(), netFlux);
  return UMT_OK;
}

// CSet%AngleOrder of every comm set (concatenated, 1-based angles) and CSet%RecvOrder(:, shared) likewise
extern "C" int umt_get_angle_order(umt_ctx *ctx, int *angleOrder /* (NA) */, int *recvOrder /* (nShared, NA) or NULL */) {
  if (!ctx || !angleOrder) return UMT_ERR_ARG;
  const int NA = ctx->NA;
  if (!ctx->have_comm_order) {   // identity
    for (int a = 0; a < NA; a++) angleOrder[a] = a + 1;
    if (recvOrder) for (size_t k = 0; k < ctx->shared.size(); k++) for (int a = 0; a < NA; a++) recvOrder[k * NA + a] = a + 1;
    return UMT_OK;
  }
  for (int a = 0; a < NA; a++) angleOrder[a] = ctx->angleOrder[a] + 1;
  if (recvOrder) for (size_t k = 0; k < ctx->shared.size(); k++) for (int a = 0; a < NA; a++) recvOrder[k * NA + a] = ctx->recvOrder[k][a] + 1;
  return UMT_OK;
}

// rows each neighbour needs from me at every step (the angles it sweeps then: RecvOrder) and rows I receive (AngleOrder)
int umt_exchange_build_stages(umt_ctx *ctx) {
  const int NA = ctx->NA, nC = ctx->nCommSets, bps = ctx->binsPerSet;
  int nBins = 0;
  std::vector<int> binOf;
  std::vector<std::vector<int>> anglesOf;
  angle_bins(ctx, nBins, binOf, anglesOf);
  for (size_t k = 0; k < ctx->shared.size(); k++) {
    SharedBdy &s = ctx->shared[k];
    if ((int)s.send_b.size() != NA) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_exchange_build_stages: exchange lists not built");
    std::vector<long long> snd, rcv;
    s.stage_send_off.assign(bps + 1, 0); s.stage_recv_off.assign(bps + 1, 0);
    for (int i = 0; i < bps; i++) {
      for (int c = 0; c < nC; c++) {   // the neighbour gets the angles of the bin IT sweeps at this step, I receive those of mine
        for (int as : anglesOf[ctx->binRecvOrder[k][c * bps + i]])
          for (int b : s.send_b[as]) snd.push_back((long long)as * ctx->rows_total() + ctx->nc + b);
        for (int ar : anglesOf[ctx->binOrder[c * bps + i]])
          for (int b : s.recv_b[ar]) rcv.push_back((long long)ar * ctx->rows_total() + ctx->nc + b);
      }
      s.stage_send_off[i + 1] = snd.size(); s.stage_recv_off[i + 1] = rcv.size();
    }
    if (snd.size() != s.send_rows || rcv.size() != s.recv_rows) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_exchange_build_stages: row count mismatch");
    int r;
    if ((r = upload(ctx, &s.d_stage_send, snd))) return r;
    if ((r = upload(ctx, &s.d_stage_recv, rcv))) return r;
  }
  return UMT_OK;
}

// SendFlux / TestSend / RecvFlux of sweep step `step` for every comm set (SetSweep.F90:126-132): what the neighbours sweep at
// this step gets my *current* exiting rows (fresh where I swept that angle at an earlier step of this pass)
int umt_exchange_stage(umt_ctx *ctx, int step) {
  const int G = ctx->G;
  const size_t nS = ctx->shared.size();
  std::vector<const void *> sp(nS); std::vector<void *> rp(nS); std::vector<size_t> sb(nS), rb(nS);
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    if (!s.d_stage_send) UMT_FAIL(ctx, UMT_ERR_STATE, "staged exchange not built (umt_sweep_scheduler)");
    const size_t o = s.stage_send_off[step];
    const long long n = (long long)(s.stage_send_off[step + 1] - o) * G;
    if (n > 0) {
      gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->psib_buf(), s.d_stage_send + o, s.d_sendbuf + o * G, n, G);
      ctx->last_launches++;
    }
    sp[k] = s.d_sendbuf + o * G; sb[k] = sizeof(double) * (size_t)n;
    const size_t ro = s.stage_recv_off[step];
    rp[k] = s.d_recvbuf + ro * G; rb[k] = sizeof(double) * (s.stage_recv_off[step + 1] - ro) * G;
  }
  UMT_CUDA(ctx, cudaGetLastError());
  int r = ctx->transport->exchange(ctx, sp, sb, rp, rb);
  if (r) return r;
  for (size_t k = 0; k < nS; k++) {
    SharedBdy &s = ctx->shared[k];
    const size_t ro = s.stage_recv_off[step];
    const long long n = (long long)(s.stage_recv_off[step + 1] - ro) * G;
    if (n > 0) {
      unpack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->psib_buf(), s.d_stage_recv + ro, s.d_recvbuf + ro * G, n, G);
      ctx->last_launches++;
    }
  }
  UMT_CUDA(ctx, cudaGetLastError());
  return UMT_OK;
}
