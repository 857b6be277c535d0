// Reflecting boundaries: snac/snreflect.F90:10-83 (PsiB(:,b,Minc) <- PsiB(:,b,Mref)) with the mirror angles of
// rt/findReflectedAngles.F90:14-154 + snac/reflectAxis.F90 (axis-aligned planes: the "90 degree" branch).
//
// In the reference the copy happens inside the angle loop right before angle Minc is swept, so it sees the exiting
// flux of Mref from the current pass whenever Mref was swept earlier.  Here all angles of a *stage* are swept by one
// persistent launch: stage(a) = 1 + max stage of its mirror images (0 when a is not incident on a reflecting
// boundary), and the copies of a stage run right before its launch.  Where reflecting planes face each other the
// dependency is cyclic; the edge from the lower to the higher angle index is then lagged one sweep (the reference
// lags whichever angle its scheduler happens to sweep first).
#include <algorithm>
#include <cmath>

#include "umt_internal.h"

namespace {

__global__ void __launch_bounds__(256) snreflect_kernel(double *psi1, const int4 *ops, int nOps, int rows, int nc, int G) {
  const int op = blockIdx.y;
  if (op >= nOps) return;
  const int4 o = ops[op];   // x = Minc, y = Mref, z = first boundary element, w = count
  const size_t n = (size_t)o.w * G;
  const double *src = psi1 + ((size_t)o.y * rows + nc + o.z) * G;
  double *dst = psi1 + ((size_t)o.x * rows + nc + o.z) * G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// reflectAxis.F90:74-123: the boundary normal has exactly one non-zero component; the mirror angle flips it
int mirror_angle(const umt_ctx *ctx, const double *omegas, int NA, int inc, const double *Area, std::string &why) {
  const int nd = ctx->ndim;
  const double fuz = 1.0e-6, tol = 1.0e-10;
  double mag = 0.0;
  for (int d = 0; d < nd; d++) mag += Area[d] * Area[d];
  int nzero = 0, nmax = -1;
  for (int d = 0; d < nd; d++) {
    if (std::fabs(Area[d] * Area[d] / mag) < tol) nzero++;
    else nmax = d;
  }
  if (nzero != nd - 1) { why = "only axis-aligned reflecting planes are supported"; return -2; }
  const double *oi = &omegas[(size_t)inc * nd];
  int mref = -1;
  for (int ia = 0; ia < NA; ia++) {
    const double *o = &omegas[(size_t)ia * nd];
    if (std::fabs(o[nmax] + oi[nmax]) >= fuz) continue;
    bool same = true;
    for (int d = 0; d < nd; d++)
      if (d != nmax && std::fabs(o[d] - oi[d]) >= fuz) same = false;
    if (same) mref = ia;   // last match wins, as in the reference's loop
  }
  if (mref < 0) why = "no reflected angle found in the quadrature set";
  return mref;
}

}  // namespace

extern "C" int umt_add_reflecting_boundary(umt_ctx *ctx, int firstBdyElem, int nBdyElem) {
  if (!ctx) return UMT_ERR_ARG;
  if (firstBdyElem < 1 || nBdyElem < 1 || firstBdyElem - 1 + nBdyElem > ctx->nb)
    UMT_FAIL(ctx, UMT_ERR_ARG, "umt_add_reflecting_boundary: elements %d..%d outside 1..%d", firstBdyElem, firstBdyElem + nBdyElem - 1, ctx->nb);
  umt_ctx::ReflBdy r;
  r.first = firstBdyElem - 1; r.n = nBdyElem;
  ctx->refl.push_back(r);
  ctx->sched_dirty = true;
  return UMT_OK;
}

// Mref of every angle on one reflecting boundary (1-based, -1 when the angle is not incident): AngleSet getReflectedAngle
extern "C" int umt_get_reflected_angles(umt_ctx *ctx, int reflIndex, int *mref) {
  if (!ctx || !mref || reflIndex < 0 || reflIndex >= (int)ctx->refl.size()) return UMT_ERR_ARG;
  int r = umt_reflect_stages(ctx);
  if (r) return r;
  for (int a = 0; a < ctx->NA; a++) mref[a] = ctx->refl[reflIndex].mref[a] >= 0 ? ctx->refl[reflIndex].mref[a] + 1 : -1;
  return UMT_OK;
}

extern "C" int umt_get_reflect_stages(umt_ctx *ctx, int *stageOf) {
  if (!ctx || !stageOf) return UMT_ERR_ARG;
  int r = umt_reflect_stages(ctx);
  if (r) return r;
  std::copy(ctx->stageOf.begin(), ctx->stageOf.end(), stageOf);
  return UMT_OK;
}

// mirror angles (per reflecting boundary) and sweep stages of an arbitrary ordinate set: the Sn set below, the GTA set in gta.cu
int umt_reflect_analyze(umt_ctx *ctx, const double *omegas, int NA, std::vector<std::vector<int>> &mrefOut, std::vector<int> &stageOut) {
  const int nd = ctx->ndim;
  mrefOut.assign(ctx->refl.size(), std::vector<int>(NA, -1));
  stageOut.assign(NA, 0);
  if (ctx->refl.empty()) return UMT_OK;
  if (!ctx->have_geom) UMT_FAIL(ctx, UMT_ERR_STATE, "reflecting boundaries need geometry");
  std::vector<double> Abdy((size_t)nd * std::max(ctx->nb, 1), 0.0);
  for (int c = 0; c < ctx->nc; c++)
    for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
      const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
      if (v > ctx->nc)
        for (int d = 0; d < nd; d++) Abdy[(size_t)(v - ctx->nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * ctx->maxcf + f) * nd + d];
    }
  const double eps = 1.0e-15;
  for (size_t k = 0; k < ctx->refl.size(); k++) {
    const double *A0 = &Abdy[(size_t)ctx->refl[k].first * nd];
    for (int a = 0; a < NA; a++) {
      double dot = 0.0;
      for (int d = 0; d < nd; d++) dot += omegas[(size_t)a * nd + d] * A0[d];
      if (dot < -eps) {
        std::string why;
        const int m = mirror_angle(ctx, omegas, NA, a, A0, why);
        if (m < 0) UMT_FAIL(ctx, UMT_ERR_ARG, "reflecting boundary %zu, angle %d: %s", k, a + 1, why.c_str());
        mrefOut[k][a] = m;
      }
    }
  }
  std::vector<int> state(NA, 0);
  std::vector<std::vector<int>> deps(NA);
  for (const auto &mr : mrefOut)
    for (int a = 0; a < NA; a++) if (mr[a] >= 0) deps[a].push_back(mr[a]);
  struct Frame { int a; size_t i; };
  for (int root = 0; root < NA; root++) {
    if (state[root]) continue;
    std::vector<Frame> st{{root, 0}};
    state[root] = 1;
    while (!st.empty()) {
      Frame &f = st.back();
      if (f.i < deps[f.a].size()) {
        const int m = deps[f.a][f.i++];
        if (state[m] == 0) { state[m] = 1; st.push_back({m, 0}); }
        else if (state[m] == 2) stageOut[f.a] = std::max(stageOut[f.a], stageOut[m] + 1);
      } else {
        state[f.a] = 2;
        const int done = f.a;
        st.pop_back();
        if (!st.empty()) stageOut[st.back().a] = std::max(stageOut[st.back().a], stageOut[done] + 1);
      }
    }
  }
  return UMT_OK;
}

// stage of every node of a dependency graph = longest chain of dependencies below it; an edge that would close a cycle is lagged
void umt_level_stages(int nL, const std::vector<std::vector<int>> &ldeps, std::vector<int> &lstage) {
  struct Frame { int a; size_t i; };
  lstage.assign(nL, 0);
  std::vector<int> lstate(nL, 0);
  for (int root = 0; root < nL; root++) {
    if (lstate[root]) continue;
    std::vector<Frame> st{{root, 0}};
    lstate[root] = 1;
    while (!st.empty()) {
      Frame &f = st.back();
      if (f.i < ldeps[f.a].size()) {
        const int m = ldeps[f.a][f.i++];
        if (lstate[m] == 0) { lstate[m] = 1; st.push_back({m, 0}); }
        else if (lstate[m] == 2) lstage[f.a] = std::max(lstage[f.a], lstage[m] + 1);
      } else {
        lstate[f.a] = 2;
        const int done = f.a;
        st.pop_back();
        if (!st.empty()) lstage[st.back().a] = std::max(lstage[st.back().a], lstage[done] + 1);
      }
    }
  }
}

int umt_reflect_stages(umt_ctx *ctx) {
  const int NA = ctx->NA, nd = ctx->ndim;
  ctx->stageOf.assign(NA, 0);
  ctx->nStages = 1;
  ctx->reflOpBegin.assign(2, 0);
  if (ctx->refl.empty()) {
    if (ctx->have_comm_order) {
      ctx->stageOf = ctx->commStageOf;
      ctx->nStages = 1 + *std::max_element(ctx->stageOf.begin(), ctx->stageOf.end());
      ctx->reflOpBegin.assign(ctx->nStages + 1, 0);
    }
    return UMT_OK;
  }
  if (!ctx->have_geom || !ctx->have_quad) UMT_FAIL(ctx, UMT_ERR_STATE, "reflecting boundaries need geometry and quadrature");
  // boundary-element area vectors = A_fp of the corner face they sit on
  std::vector<double> Abdy((size_t)nd * std::max(ctx->nb, 1), 0.0);
  for (int c = 0; c < ctx->nc; c++)
    for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
      const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
      if (v > ctx->nc)
        for (int d = 0; d < nd; d++) Abdy[(size_t)(v - ctx->nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * ctx->maxcf + f) * nd + d];
    }
  const double eps = 1.0e-15, tolPlane = 1.0e-6;
  for (size_t k = 0; k < ctx->refl.size(); k++) {
    auto &R = ctx->refl[k];
    const double *A0 = &Abdy[(size_t)R.first * nd];
    double m0 = 0.0;
    for (int d = 0; d < nd; d++) m0 += A0[d] * A0[d];
    for (int b = 0; b < R.n; b++) {   // findReflectedAngles.F90:64-110: every element must lie in one plane
      const double *A = &Abdy[(size_t)(R.first + b) * nd];
      double m = 0.0, delta = 0.0;
      for (int d = 0; d < nd; d++) m += A[d] * A[d];
      for (int d = 0; d < nd; d++) delta += std::fabs(A0[d] / std::sqrt(m0) - A[d] / std::sqrt(m));
      if (delta > tolPlane) UMT_FAIL(ctx, UMT_ERR_ARG, "reflecting boundary %zu: not all faces lie in one plane (each plane of reflection needs its own boundary)", k);
    }
    R.mref.assign(NA, -1);
    for (int a = 0; a < NA; a++) {
      double dot = 0.0;
      for (int d = 0; d < nd; d++) dot += ctx->h_omega[(size_t)a * nd + d] * A0[d];
      if (dot < -eps) {
        std::string why;
        const int m = mirror_angle(ctx, ctx->h_omega.data(), NA, a, A0, why);
        if (m < 0) UMT_FAIL(ctx, UMT_ERR_ARG, "reflecting boundary %zu, angle %d: %s", k, a + 1, why.c_str());
        R.mref[a] = m;
      }
    }
  }
  // stages: longest chain of mirror dependencies; an edge that would close a cycle (facing planes) is lagged
  std::vector<int> state(NA, 0);   // 0 new, 1 on stack, 2 done
  std::vector<std::vector<int>> deps(NA);
  for (const auto &R : ctx->refl)
    for (int a = 0; a < NA; a++) if (R.mref[a] >= 0) deps[a].push_back(R.mref[a]);
  struct Frame { int a; size_t i; };
  for (int root = 0; root < NA; root++) {
    if (state[root]) continue;
    std::vector<Frame> st{{root, 0}};
    state[root] = 1;
    while (!st.empty()) {
      Frame &f = st.back();
      if (f.i < deps[f.a].size()) {
        const int m = deps[f.a][f.i++];
        if (state[m] == 0) { state[m] = 1; st.push_back({m, 0}); }
        else if (state[m] == 2) ctx->stageOf[f.a] = std::max(ctx->stageOf[f.a], ctx->stageOf[m] + 1);
        // state 1: back edge -> lagged
      } else {
        state[f.a] = 2;
        const int done = f.a;
        st.pop_back();
        if (!st.empty()) ctx->stageOf[st.back().a] = std::max(ctx->stageOf[st.back().a], ctx->stageOf[done] + 1);
      }
    }
  }
  if (ctx->have_comm_order) ctx->stageOf = ctx->commStageOf;   // SweepScheduler's order already honours the mirror dependencies
  if (nd == 2) {
    // r-z: the angles of a xi-level are a chain (PsiM), so stages are assigned per level from the mirror dependencies *between*
    // levels (z-normal planes: the partner level), and inside a level every angle gets its own step: snreflect then runs right
    // before each angle exactly as in the reference (a mirror image earlier in the level is fresh, a later one lagged one pass).
    // Finishing directions are not swept and get no copy (SetSweep.F90:141-143).
    const int nL = std::max(ctx->nLevels, 1);
    std::vector<std::vector<int>> ldeps(nL);
    for (const auto &R : ctx->refl)
      for (int a = 0; a < NA; a++)
        if (R.mref[a] >= 0 && ctx->h_level[R.mref[a]] != ctx->h_level[a]) ldeps[ctx->h_level[a]].push_back(ctx->h_level[R.mref[a]]);
    std::vector<int> lstage;
    umt_level_stages(nL, ldeps, lstage);
    int maxPos = 1;
    std::vector<int> pos(NA, 0), cnt(nL, 0);
    for (int a = 0; a < NA; a++) { pos[a] = cnt[ctx->h_level[a]]++; maxPos = std::max(maxPos, cnt[ctx->h_level[a]]); }
    for (int a = 0; a < NA; a++) ctx->stageOf[a] = lstage[ctx->h_level[a]] * maxPos + pos[a];
    for (auto &R : ctx->refl)
      for (int a = 0; a < NA; a++) if (ctx->h_finish[a]) R.mref[a] = -1;
  }
  ctx->nStages = 1 + *std::max_element(ctx->stageOf.begin(), ctx->stageOf.end());
  if (ctx->device < 0) return UMT_OK;
  std::vector<int4> ops;
  ctx->reflOpBegin.assign(ctx->nStages + 1, 0);
  for (int s = 0; s < ctx->nStages; s++) {
    for (const auto &R : ctx->refl)
      for (int a = 0; a < NA; a++)
        if (ctx->stageOf[a] == s && R.mref[a] >= 0) ops.push_back(make_int4(a, R.mref[a], R.first, R.n));
    ctx->reflOpBegin[s + 1] = (int)ops.size();
  }
  if (ctx->d_reflOps) { cudaFree(ctx->d_reflOps); ctx->d_reflOps = nullptr; }
  UMT_CUDA(ctx, cudaMalloc((void **)&ctx->d_reflOps, sizeof(int4) * std::max<size_t>(ops.size(), 1)));
  if (!ops.empty()) UMT_CUDA(ctx, umt_memcpy(ctx, ctx->d_reflOps, ops.data(), sizeof(int4) * ops.size(), cudaMemcpyHostToDevice));
  return UMT_OK;
}

int umt_launch_reflect(umt_ctx *ctx, int stage) {
  if (ctx->refl.empty()) return UMT_OK;
  const int begin = ctx->reflOpBegin[stage], end = ctx->reflOpBegin[stage + 1];
  if (end == begin) return UMT_OK;
  int maxN = 1;
  for (const auto &R : ctx->refl) maxN = std::max(maxN, R.n);
  const unsigned gx = (unsigned)std::min<size_t>(((size_t)maxN * ctx->G + 255) / 256, 1024);
  snreflect_kernel<<<dim3(gx, end - begin), 256, 0, ctx->stream>>>(ctx->psib_buf(), ctx->d_reflOps + begin, end - begin, ctx->rows, ctx->nc, ctx->G);
  UMT_CUDA(ctx, cudaGetLastError());
  ctx->last_launches += 1;
  return UMT_OK;
}
