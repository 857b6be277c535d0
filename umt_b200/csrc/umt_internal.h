// Internal state of libumtsweep.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/umt_sweep.h"

struct WorkItem {       // one chunk of one hyperplane of one angle
  int angle;            // 0-based angle; phi-tally items of the 3-D plan kernel: -1 - (first angle of the batch)
  int zbeg, zend;       // range in nextZ(:,angle); phi-tally items: range of 16-byte columns of PhiTotal
  int wait_idx;         // counter to wait on (-1: none)
  int wait_count;       // value it must reach
  int signal_idx;       // counter to bump when done
  int pad0, pad1;       // r-z / grey kernels: second dependency (counter, value), pad0 < 0: none.
                        // 3-D plan kernel: pad0 = second counter to bump (-1: none; the batch's gate counter in ring mode),
                        // pad1 = slab of the Psi1 workspace the item's angle writes (ring slot; legacy / in place: the angle);
                        // phi-tally items: pad1 = first slot | angles in the batch << 16 | first tally of the sweep << 30
};

// Plan record of one (zone, angle) for the 3-D plan kernel (sweep3d.cu): the zone's corners
// relabelled into solve order ("positions"), omega.A products, upstream rows and the
// group-independent pieces of the corner-balance closure, as compact lists in position order.
enum { ZREC_SLOW = 16u, ZREC_HAS_EXIT = 32u, ZREC_CANON = 64u };   // flags bits above the corner count (bits 0..3); CANON: see sweep3d.cu
struct ZoneEdge {                // EZ face carrying flux from position p into a later position
  double ainv, cp, ha;           // 1/aez, coefpsi (SweepUCBxyz.F90:168-179), aez/2
  int qoff, hasOpp;              // byte offset of the downstream corner's column in the landing area; 1 if the opposite
};                               // FP face is incident (it is then slot k of p's incident list for p's k-th edge)
struct alignas(16) ZoneRec {
  int c0, zone0;                 // first corner row of the zone; signed 1-based zone id from nextZ
  unsigned flags, exitMask;      // NC | ZREC_*; bit p*3+f: FP face f of position p exits through the boundary
  unsigned char nIn[8], nOut[8]; // highest used incident slot + 1 / outgoing EZ faces of position p
  int crow[8];                   // (c0 + local corner of position p) * G: element offset of its Psi/Psi1 row in the angle slab
  int coff[8];                   // byte offset of that corner's column in the landing area (local corner * G/2 * 16)
  double vol[8], sumArea[8];     // by position
  int inOff[8][3];               // element offset (row * G) of the Psi1 row behind incident slot k (rows >= ncornr: boundary elements)
  double inAfp[8][3];            // omega . A_fp (< 0) of that face; 0 in unused slots
  ZoneEdge edge[12];
  int exitOff[8][3];             // element offset of the boundary-element row behind exiting FP face f of position p
};
static_assert(sizeof(ZoneEdge) == 32 && sizeof(ZoneRec) == 992, "ZoneRec layout");

struct PackChunk { int rowBeg, rowEnd, slot, pad; };   // send rows of one (neighbour, angle) handled by one CTA; slot = angle * maxChunks + chunk
struct UmtTransport;    // exchange.cu: NCCL or in-process
struct SharedBdy {      // one neighbour (rt/findexit.F90:102-294)
  int neighbor;
  int first;            // 0-based first boundary element
  int n;
  std::vector<signed char> incTestR;                       // the neighbour's sign(omega.A_bdy) for the angles it decides, (n, NA)
  // per angle: send rows (boundary element, 0-based) and recv rows
  std::vector<std::vector<int>> send_b, recv_b;
  long long *d_send_row = nullptr, *d_recv_row = nullptr;  // row index into the (NA, nc+nb) slab array, concatenated over angles
  double *d_send_coef = nullptr;                           // w_a (omega_a . A_bdy) per send row
  PackChunk *d_chunks = nullptr;
  int *d_nChunksOfAngle = nullptr;
  double *d_partial = nullptr;                             // (NA, maxChunks) partial exit currents
  int nChunks = 0, maxChunks = 1;
  std::vector<int> send_off, recv_off;                     // per-angle offsets (NA+1)
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
  // put path (exchange.cu): pack_tally_kernel stores the exiting rows straight into the neighbour's receive buffer over peer memory
  // (NVLink stores when the neighbour is another GPU) instead of packing a send buffer for ncclSend/ncclRecv.  Two receive buffers
  // alternate by pass (d_recvbuf = [0], d_recvbuf2 = [1]): a neighbour may already be writing the rows of its next pass while this
  // domain still unpacks the previous ones.  peer_recv: the NEIGHBOUR's two buffers as seen from this process.
  double *d_recvbuf2 = nullptr;
  double *peer_recv[2] = {nullptr, nullptr};
  bool peer_is_ipc = false;
  size_t send_rows = 0, recv_rows = 0;
  // staged exchange (comm sets with several bins): rows to send / receive at each step, grouped by step
  long long *d_stage_send = nullptr, *d_stage_recv = nullptr;
  std::vector<size_t> stage_send_off, stage_recv_off;   // (nSteps + 1) row offsets
  // grey (GTA) exchange: positions in the (8, nbelem) PsiB array of the exiting / incident elements, all angles concatenated
  int *d_gsend = nullptr, *d_grecv = nullptr;
  double *d_gsendbuf = nullptr, *d_grecvbuf = nullptr;
  size_t gsend_n = 0, grecv_n = 0;
};

// grey transport acceleration state (gta.cu): 3-D, "new" GTA solver
struct GtaState {
  bool ready = false;
  int nAng = 0;
  std::vector<double> omega, weight;
  std::vector<int> nHyp;
  std::vector<std::vector<int>> zonesInPlane, nextZ, nextC;
  double *d_omega = nullptr, *d_weight = nullptr;
  int *d_nextZ = nullptr;              // (nAng, nz) signed 1-based
  unsigned char *d_nextC = nullptr;    // (nAng, nc) 0-based local corner
  WorkItem *d_items = nullptr;
  int nItems = 0, nCounters = 0, maxHyp = 0;
  int *d_counters = nullptr;
  // opacities and sources (nc)
  double *d_sigTotal = nullptr, *d_sigtInv = nullptr, *d_sigScat = nullptr, *d_sigScatVol = nullptr;
  double *d_greySource = nullptr, *d_tsaSource = nullptr, *d_phiInc = nullptr, *d_correction = nullptr;
  double *d_chi = nullptr;             // (nc, G)
  double *d_TT = nullptr;              // (nc, maxCorner, maxCorner): TT(cc, c0+c) at [(c0+c)*mC + cc]
  bool have_opacity = false, tt_decomposed = false;
  double *d_tpsi = nullptr;            // (nAng, nc+nb): tPsi per angle; the nb tail rows are PsiB(:, angle)
  double *d_pinc = nullptr;            // (nAng, nc)
  double *d_vec[4] = {nullptr, nullptr, nullptr, nullptr};    // BiCGSTAB: residual, direction, action, actionS (nc)
  double *d_vecB[4] = {nullptr, nullptr, nullptr, nullptr};   // their boundary parts (nAng, nb)
  double *d_radEnergy = nullptr, *d_pzOld = nullptr, *d_volZone = nullptr;   // (nz)
  double *d_red = nullptr;             // reduction scratch
  double *d_P = nullptr, *d_PB = nullptr;   // staging for the host-facing sweep calls
  // reflecting boundaries (3-D): the angles are swept in stages, mirror images first; the PsiB copies of snreflect before each stage
  int nStagesR = 1;
  std::vector<int> stageOf, stageItemBegin, reflOpBegin;
  std::vector<std::vector<int>> mref;
  int4 *d_reflOps = nullptr;           // (minc, mref, first, n) grouped by stage
  // r-z: xi-levels chained through the half-angle values tPsiM / tInc (SweepGreyUCBrz.F90)
  std::vector<unsigned char> start, finish;
  std::vector<double> angDerivFac, tauW1, tauW2;
  std::vector<int> level;
  int nLevels = 0;
  unsigned char *d_start = nullptr;
  int *d_level = nullptr;
  double *d_fac = nullptr, *d_w1 = nullptr, *d_w2 = nullptr, *d_psim = nullptr, *d_tinc = nullptr;   // psim/tinc: (nLevels, nc)
  unsigned char *d_finish = nullptr;
  bool hexTT = false;                  // InitGreySweep by gta_init_tt_hex_kernel (all zones hexahedra)
  bool flow3d = false;                 // 3-D dataflow kernel (gta_sweep_flow_kernel): values as their own completion flags
  unsigned *d_pollMask = nullptr;      // (nAng, nz) which incident faces of a zone the schedule orders (see gta.cu)
  bool rz_chain = false;               // one CTA per xi-level (gta_sweep_rz_chain_kernel)
  bool rz_flow = false;                // dataflow kernel (gta_sweep_rz_flow_kernel): values as their own completion flags
  int *d_prevAngle = nullptr;          // (nAng) previous swept angle of the xi-level, -1: none
  double *d_psimA = nullptr, *d_tincA = nullptr;   // (nAng, nc) half-angle values as written by each angle
  int rz_maxAngLevel = 0, rz_threads = 64;
  int *d_levelAngles = nullptr, *d_planeOff = nullptr, *d_nHyp = nullptr;
};

struct umt_ctx {
  int device = 0;
  int ndim = 0, nz = 0, nc = 0, nb = 0, maxcf = 0, maxCorner = 0, maxFaces = 0, G = 0;
  int NA = 0;
  double tau = 0.0;
  std::string err;
  int sm_count = 0;
  bool l2_persist = false;             // the 3-D sweep runs with an L2 set-aside for its evict_last Psi1 lines
  size_t l2_persist_bytes = 0, l2_persist_before = 0;   // before: the device limit found at creation, restored at destroy
  cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;   // main; phi chunks -> host; overlapped exchange
  cudaStream_t xstream = nullptr;      // where exchange work goes right now (nullptr: the main stream), see exchange.cu
  cudaEvent_t ev[8] = {}, evx[3] = {};
  // lagged exchange state: the send buffers hold the exiting rows of the current PsiB / the receive buffers already hold what
  // the neighbours packed after their last sweep (the transfer was overlapped with the phi tally of that sweep)
  bool pack_valid = false, recv_valid = false;
  bool put_ready = false;              // peer pointers to every neighbour's receive buffers are open
  bool put_now = false;                // umt_exchange_tally packs into the neighbours' buffers (set by the controller)
  long long passCount = 0;             // passes swept since the exchange was built (parity of the receive buffers)

  // host copies (for schedule builder, exit lists, tallies)
  std::vector<int> h_numCorner, h_cOffSet, h_nCFaces, h_cFP, h_cEZ, h_zoneFaces, h_zoneOpp, h_faceOpp, h_CToFace, h_BdyToC;
  std::vector<unsigned char> h_BoundaryZone;
  std::vector<double> h_Volume, h_Afp, h_Aez, h_Abdy, h_Area, h_RadiusFP, h_RadiusEZ, h_VolumeZone;
  std::vector<double> h_omega, h_weight, h_angDerivFac, h_tauW1, h_tauW2;
  std::vector<unsigned char> h_start, h_finish;
  bool have_conn = false, have_geom = false, have_quad = false, have_abdy = false;

  // schedule (host)
  std::vector<int> nHyp, numCycles, cycleOffSet, nBad;
  std::vector<std::vector<int>> zonesInPlane, nextZ, nextC, cycleList, bdyList;
  bool sched_dirty = true;

  // device: connectivity / geometry
  int *d_numCorner = nullptr, *d_cOffSet = nullptr, *d_nCFaces = nullptr, *d_cFP = nullptr, *d_cEZ = nullptr;
  double *d_Volume = nullptr, *d_Afp = nullptr, *d_Aez = nullptr, *d_Area = nullptr, *d_RadiusFP = nullptr, *d_RadiusEZ = nullptr;
  double *d_omega = nullptr, *d_weight = nullptr;
  // RZ only: angular-derivative coefficients, starting flags, FinishingDirection(a+1), xi-level of each angle
  double *d_angDerivFac = nullptr, *d_tauW1 = nullptr, *d_tauW2 = nullptr;
  unsigned char *d_start = nullptr, *d_finishNext = nullptr;
  int *d_level = nullptr;
  std::vector<int> h_level;
  int nLevels = 0;
  // r-z chain kernel (sweeprz.cu): one CTA per (xi-level, group block)
  bool rz_chain = false;
  int rz_gb = 4, rz_maxAngLevel = 0, rz_threads = 256;
  int *d_rzLevelAngles = nullptr, *d_rzPlaneOff = nullptr, *d_rzNHyp = nullptr;
  // r-z record kernel (sweeprz.cu): group-independent half of the zone solve precomputed per (angle, zone), 384 B each
  void *d_rzRecs = nullptr;            // (NA, nz) RZRec in sweep order
  int *d_rzBad = nullptr;              // number of zones that do not fit the canonical labelling
  bool rz_rec = false, rz_recs_valid = false, rz_canon = false;
  // r-z level-chain kernel (sweeprz.cu): steps (plane chunks) of every xi-level
  void *d_rzSteps = nullptr; int *d_rzNSteps = nullptr; int rz_lc_zch = 0, rz_lc_maxSteps = 0;
  // r-z dataflow kernel (sweeprz.cu): no counters, the angular fluxes themselves are the completion flags
  bool rz_flow = false;
  int *d_rzPrev = nullptr;             // (NA) previous swept angle of the xi-level, -1 for the first
  double *d_rzPsimA = nullptr;         // (NA, nc, G) half-angle intensity written by each angle (read by the next one of its level)
  // device: schedule
  int *d_nextZ = nullptr;              // (NA, nz) signed 1-based
  unsigned char *d_nextC = nullptr;    // (NA, nc) 0-based local corner
  WorkItem *d_items = nullptr;
  int nItems = 0, nCounters = 0, maxHyp = 0;
  WorkItem *d_itemsRing = nullptr;     // single-psi layout, non-final sweeps: items with ring slots and the in-kernel phi-tally items
  int nItemsRing = 0;
  int *d_counters = nullptr;           // [0]=ticket, [1..] per (angle,plane), then per batch: gate, tally done (ring mode)
  ZoneRec *d_recs = nullptr;           // (NA, nz) plan records in sweep order
  int2 *d_zinfo = nullptr;             // (NA, nz) first corner row, zone | numCorner << 28 (what the TMA producer needs)
  bool use_plan = false;
  int plan_ncw = 4, plan_slow_zones = 0, plan_canon_zones = 0, zones_per_item = 1;
  int plan_nh = 1;                     // 16-byte columns (pairs of groups) per lane in the plan kernel: 1 or 2
  int *d_cycleList = nullptr, *d_cycleAngle = nullptr;   // flattened (totalCycles): corner (0-based), angle
  int totalCycles = 0;
  double *d_cyclePsi = nullptr;
  int *d_exitB = nullptr, *d_exitC = nullptr, *d_exitA = nullptr; int nExit = 0;  // flattened bdyList over angles
  std::vector<int> exitOff;            // per-angle offsets into d_exit*
  // device: state
  // d_psi: (NA, rows = nc+nb, G) = Set%Psi (corner rows; the nb tail rows of each slab are padding or Set%PsiB, see below).
  // d_psi1: (psi1Slots, rows, G) workspace for Set%Psi1.  Two layouts (umt_api.cu ensure_layout):
  //   legacy  psi1Slots = NA, the tails of d_psi1 are Set%PsiB, a savePsi sweep ends with the two buffers trading roles.  Every
  //           r-z problem, and 3-D problems with cycle lists, direct-solve zones, reflecting boundaries or staged comm sets.
  //   single  the tails of d_psi are Set%PsiB; a savePsi sweep writes Psi in place (a zone reads its own Psi^n rows before it
  //           writes them and only ever reads new upstream rows, SweepUCBxyz.F90:119-126,149-158,314-318); the other sweeps
  //           keep Psi1 in a ring of psi1Slots <= NA slabs whose angle batches are tallied into PhiTotal as they retire.
  double *d_psi = nullptr, *d_psi1 = nullptr, *d_stotal = nullptr, *d_sigt = nullptr, *d_phi = nullptr;
  int rows = 0;
  bool single_psi = false;             // layout in force (valid once d_psi1 exists)
  bool psib_in_psi = true;             // Set%PsiB = tails of d_psi (true until a legacy workspace exists)
  bool force_legacy = false;           // compat.cu pokes Psi1 / PsiB directly
  int psi1Slots = 0;                   // slabs allocated in d_psi1
  bool single_psi_wanted = false;      // what the schedule allows (finalize_schedule)
  int ringBatchesAuto = 0;             // ring size the free HBM allows, in batches
  std::vector<int> h_cycleFlat, h_cycleAngleFlat;   // the cycle list Set%cyclePsi on the device belongs to
  int ringBatchesWanted = 0;           // umt_set_psi1_ring: angle batches the ring should hold (0 = as memory allows)
  int angleBatch = 0, ringBatches = 0, nBatches = 0;   // K, ring size in batches (>= nBatches: no in-kernel tally), batches
  std::vector<int> slotOfAngle;        // (NA) ring slot of each angle's Psi1 in a non-final sweep
  int nTallied = 0;                    // angles [0, nTallied) are tallied into PhiTotal by the sweep kernel itself
  int *d_tailSlot = nullptr; double *d_tailW = nullptr;   // slots / weights of the angles the post-sweep tally still has to add
  double *psib_buf() const { return psib_in_psi ? d_psi : d_psi1; }   // buffer whose slab tails are Set%PsiB
  double *d_psim = nullptr;            // RZ half-angle intensity (G,nc) per xi-level
  size_t psi_elems = 0;

  // reflecting boundaries (snac/snreflect.F90, rt/findReflectedAngles.F90): per boundary the mirror angle of every
  // incident angle (-1: not incident); stageOf[a] orders the angles so that a mirror image is swept first
  struct ReflBdy { int first, n; std::vector<int> mref; };
  std::vector<ReflBdy> refl;
  std::vector<int> stageOf, stageItemBegin;
  int nStages = 1;
  // SweepScheduler (rt/SweepScheduler.F90): comm sets of several angle bins swept in AngleOrder, all comm sets concurrently.
  // nCommSets == 0: the finest decomposition (one bin per comm set, scheduler = identity, exchange lagged a whole pass).
  int nCommSets = 0;
  bool have_comm_order = false;
  std::vector<int> angleOrder;                   // (NA) per comm set concatenated: 0-based angle swept at each step
  std::vector<std::vector<int>> recvOrder;       // [shared][NA] the neighbour's AngleOrder (0-based), same layout
  std::vector<int> commStageOf;                  // (NA) step at which angle a is swept
  std::vector<int> binOrder;                     // (nSchedBins) per comm set concatenated: bin swept at each step (3-D: bin = angle, r-z: xi-level)
  std::vector<std::vector<int>> binRecvOrder;    // [shared][nSchedBins] the neighbour's bin order
  int nSchedBins = 0, binsPerSet = 1;
  std::vector<double> netFlux;                   // (nShared, NA) CSet%NetFlux of the last scheduler run
  int4 *d_reflOps = nullptr;           // (minc, mref, first, n) grouped by stage
  std::vector<int> reflOpBegin;        // per-stage offsets into d_reflOps

  // exchange
  std::vector<SharedBdy> shared;
  int myRank = 0, nRanks = 1;
  void *nccl_comm = nullptr;
  UmtTransport *transport = nullptr;
  bool exch_dirty = true;
  double *d_exitFlux = nullptr, *d_incRecv = nullptr;      // (nShared, NA) exit currents sent / received
  double *d_incFlux = nullptr, *d_incFluxOld = nullptr;    // (nBins) CSet%IncFlux, IncFluxOld
  int *d_binOfAngle = nullptr, *d_nNotConv = nullptr;
  int nBins = 0;
  double fluxFloor = 0.0;
  int rows_total() const { return nc + nb; }

  GtaState gta;
  // watchdog of the dataflow kernels: device flag, its page-locked host copy (fetched after every dataflow launch), poll budget
  int *d_abort = nullptr, *h_abort = nullptr;
  unsigned spinLimit = 1u << 21;       // polls of ~1 us each before a thread gives up (UMT_SPIN_LIMIT)
  std::map<void *, size_t> host_blocks;   // umt_host_alloc: live blocks and their mapped length

  // stats
  double last_ms[4] = {0, 0, 0, 0};
  int last_launches = 0;
};

#define UMT_FAIL(ctx, code, ...)                         \
  do {                                                   \
    char _b[512];                                        \
    snprintf(_b, sizeof(_b), __VA_ARGS__);               \
    (ctx)->err = _b;                                     \
    return (code);                                       \
  } while (0)

#define UMT_CUDA(ctx, call)                                                            \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess)                                                             \
      UMT_FAIL(ctx, UMT_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, \
               cudaGetErrorString(_e));                                                \
  } while (0)

// UMT_TRACE=1 in the environment: progress lines of the collective parts (exchange set-up, flux passes) on stderr, one per rank
#define UMT_TRACE(ctx, ...)                                                                                     \
  do {                                                                                                          \
    static const bool umt_trace_on_ = getenv("UMT_TRACE") && atoi(getenv("UMT_TRACE")) != 0;                    \
    if (umt_trace_on_) { fprintf(stderr, "[umt rank %d] ", (ctx)->myRank); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } \
  } while (0)

// Blocking copy ordered with the context's stream.  (cudaMemcpy runs on the legacy default stream, which the context's non-blocking
// streams do not synchronise with: a pageable host-to-device copy may still be in flight when it returns, and a device-to-host copy
// does not wait for kernels queued on the context's stream.)
static inline cudaError_t umt_memcpy(const umt_ctx *ctx, void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
  cudaError_t e = cudaMemcpyAsync(dst, src, n, kind, ctx->stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(ctx->stream);
}

// kernels / host pieces implemented in other translation units
int umt_check_abort(umt_ctx *ctx, const char *what);   // after a sync: UMT_ERR_STATE if a dataflow kernel gave up waiting
int umt_launch_sweep3d(umt_ctx *ctx, int savePsi);
int umt_build_plan3d(umt_ctx *ctx);
int umt_sweep3d_zones_per_item(const umt_ctx *ctx);
int umt_launch_sweeprz(umt_ctx *ctx, int savePsi);
int umt_build_items_rz(umt_ctx *ctx, std::vector<WorkItem> &items, int zonesPerItem);
int umt_build_items_rz_set(int nz, int NA, const std::vector<int> &nHyp, const std::vector<std::vector<int>> &zonesInPlane,
                           const std::vector<std::vector<int>> &nextZ, const std::vector<unsigned char> &start, int zonesPerItem,
                           std::vector<WorkItem> &items, std::vector<int> &level, int &nLevels, int &maxHyp);
int umt_host_gta_quadrature_rz(std::vector<double> &omega, std::vector<double> &weight, std::vector<unsigned char> &start,
                               std::vector<unsigned char> &finish, std::vector<double> &angDerivFac, std::vector<double> &w1,
                               std::vector<double> &w2);
int umt_gta_setup_rz(umt_ctx *ctx);          // gta_rz.cu
int umt_gta_finish_setup_rz(umt_ctx *ctx, std::vector<WorkItem> &items);
int umt_gta_launch_sweep_rz(umt_ctx *ctx);
int umt_gta_launch_init_tt_rz(umt_ctx *ctx);
int umt_host_build_schedule(umt_ctx *ctx);
int umt_host_product_quadrature(int ndim, int npolar, int nazimuthal, int polaraxis,
                                std::vector<double> &omega, std::vector<double> &weight,
                                std::vector<unsigned char> &start, std::vector<unsigned char> &finish,
                                std::vector<double> &angDerivFac, std::vector<double> &w1, std::vector<double> &w2);
int umt_device_geometry(umt_ctx *ctx, const double *d_px);
void umt_exchange_release(umt_ctx *ctx);
int umt_reflect_stages(umt_ctx *ctx);
void umt_level_stages(int nL, const std::vector<std::vector<int>> &ldeps, std::vector<int> &lstage);
int umt_reflect_analyze(umt_ctx *ctx, const double *omegas, int NA, std::vector<std::vector<int>> &mref, std::vector<int> &stageOf);
int umt_launch_reflect(umt_ctx *ctx, int stage);
void umt_gta_release(umt_ctx *ctx);
int umt_host_build_order(umt_ctx *ctx, const double *omegas, int nAng, std::vector<int> &nHyp, std::vector<std::vector<int>> &zonesInPlane,
                         std::vector<std::vector<int>> &nextZ, std::vector<std::vector<int>> &nextC);
int umt_exchange_tally(umt_ctx *ctx, double tol);
int umt_exchange_rows(umt_ctx *ctx);      // collective: packed exiting rows -> the neighbours' receive buffers
double *umt_recv_buffer(const umt_ctx *ctx, const SharedBdy &s);
int umt_exchange_unpack(umt_ctx *ctx);    // receive buffers -> incident PsiB rows
int umt_exchange_test_convergence(umt_ctx *ctx, int *nNotConv);
int umt_exchange_stage(umt_ctx *ctx, int step);               // SendFlux / RecvFlux of one sweep step (staged comm sets)
int umt_exchange_build_stages(umt_ctx *ctx);
int umt_gta_build_exchange(umt_ctx *ctx);                    // collective: GTA ListSend / ListRecv (findexit.F90 on the GTA angle set)
int umt_gta_exchange(umt_ctx *ctx, double *d_PsiB);          // SendFlux / RecvFlux of GTASweep.F90:139-146 for all 8 angles
int umt_allreduce_f64(umt_ctx *ctx, double *d_vals, int n, int op /* 0 sum, 1 max */);   // MPIAllReduce over the domains
