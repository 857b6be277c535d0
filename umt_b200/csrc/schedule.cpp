// Host-side sweep-order construction for every angle of the quadrature
// (what rt/rtorder.F90 -> snac/snnext.F90 does per cycle in the reference, plus the
// non-shared part of rt/findexit.F90:296-349).  Semantics that influence results
// are kept exactly: which zones land on the cycle list (snneed.F90:156-183,
// findseeds.F90:72-104, sccsearch.F90:137-166, fixZone.F90), the signed nextZ for
// zones with an intra-zone cycle and the minloc corner order
// (getDownStreamData.F90:124-149).  Angles are independent, so they are built on
// a pool of host threads; on a static mesh the result is cached by the context
// until geometry or quadrature change.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>

#include "umt_internal.h"

namespace {

struct MeshView {
  int ndim, nz, nc, nb, mcf, mf, maxCorner;
  const int *numCorner, *cOffSet, *nCFaces, *cFP, *cEZ, *zoneFaces, *zoneOpp, *faceOpp, *CToFace;
  const unsigned char *bzone;
  const double *Afp, *Aez;
  // 1-based accessors mirroring the Fortran arrays
  int zopp(int face, int zone) const { return zoneOpp[(face - 1) + (size_t)mf * (zone - 1)]; }
  int fopp(int face, int zone) const { return faceOpp[(face - 1) + (size_t)mf * (zone - 1)]; }
  int ctf(int cf, int c) const { return CToFace[(cf - 1) + (size_t)mcf * (c - 1)]; }
  int cez(int cf, int c) const { return cEZ[(cf - 1) + (size_t)mcf * (c - 1)]; }
  double dotA(const double *A, int cf, int c, const double *om) const {
    const double *v = A + ((size_t)(c - 1) * mcf + (cf - 1)) * ndim;
    double s = 0.0;
    for (int d = 0; d < ndim; d++) s += om[d] * v[d];
    return s;
  }
};

struct AngleSchedule {
  int nHyp = 0;
  std::vector<int> zonesInPlane, nextZ, nextC, cycleList;
  int nBad = 0;
  std::string error;
};

class OrderBuilder {
 public:
  OrderBuilder(const MeshView &m, const double *om) : M(m), omega(om) {}

  void run(AngleSchedule &out) {
    const int nz = M.nz;
    need.assign(nz + 1, 0);
    exitFace.assign((size_t)M.mf * nz, 0);
    onCycle.assign(nz + 1, 0);
    bad.assign(nz + 1, 0);
    done.assign(nz + 1, 0);
    order.assign(nz, 0);
    out.nextC.assign(M.nc, 0);
    out.nextZ.assign(nz, 0);
    cycles = &out.cycleList;
    cycles->clear();
    count_upstream_faces();
    int fresh = seed();
    if (!err.empty()) { out.error = err; return; }
    corner_orders(out.nextC);
    int ndone = 0, filled = 0, last = 0;
    for (;;) {
      out.zonesInPlane.push_back(fresh);
      filled = last + fresh;
      int added = 0;
      for (int k = 0; k < fresh; k++) {
        const int zone = order[last + k];
        done[zone] = 1;
        for (int face = 1; face <= M.zoneFaces[zone - 1]; face++) {
          if (!xf(face, zone)) continue;
          const int zex = M.zopp(face, zone);
          if (zex > 0 && !done[zex]) {
            if (--need[zex] == 0) { order[filled++] = zex; added++; }
            else if (need[zex] < 0) { out.error = "needZ < 0 while ordering zones"; return; }
          }
        }
        out.nextZ[ndone++] = bad[zone] ? -zone : zone;
        out.nBad += bad[zone];
      }
      last += fresh;
      if (last == nz) break;
      if (added == 0) {
        added = break_cycles(ndone, filled);
        if (!err.empty()) { out.error = err; return; }
      }
      fresh = added;
    }
    out.nHyp = (int)out.zonesInPlane.size();
    if ((int)cycles->size() > M.nc) out.error = "mesh cycles exceed the number of corners";
  }

 private:
  const MeshView &M;
  const double *omega;
  std::vector<int> need, order;
  std::vector<unsigned char> exitFace, onCycle, bad, done;
  std::vector<int> *cycles = nullptr;
  std::string err;

  unsigned char &xf(int face, int zone) { return exitFace[(face - 1) + (size_t)M.mf * (zone - 1)]; }
  void lag_zone(int zone) {
    for (int c = 1; c <= M.numCorner[zone - 1]; c++) cycles->push_back(M.cOffSet[zone - 1] + c);
  }

  void count_upstream_faces() {
    if (M.ndim == 2) {
      for (int zone = 1; zone <= M.nz; zone++) {
        const int c0 = M.cOffSet[zone - 1];
        for (int c = 1; c <= M.numCorner[zone - 1]; c++) {
          const int face = M.ctf(1, c0 + c), zo = M.zopp(face, zone);
          if (zone < zo) {
            const double a = M.dotA(M.Afp, 1, c0 + c, omega);
            if (a < 0.0) { need[zone]++; xf(M.fopp(face, zone), zo) = 1; }
            else if (a > 0.0) { need[zo]++; xf(face, zone) = 1; }
          }
        }
      }
      return;
    }
    std::vector<double> fsum(M.mf);
    std::vector<int> nin(M.mf), nout(M.mf);
    for (int zone = 1; zone <= M.nz; zone++) {
      const int c0 = M.cOffSet[zone - 1], nF = M.zoneFaces[zone - 1];
      std::fill(fsum.begin(), fsum.end(), 0.0);
      std::fill(nin.begin(), nin.end(), 0);
      std::fill(nout.begin(), nout.end(), 0);
      for (int c = 1; c <= M.numCorner[zone - 1]; c++)
        for (int cf = 1; cf <= M.nCFaces[c0 + c - 1]; cf++) {
          const int face = M.ctf(cf, c0 + c);
          if (M.zopp(face, zone) > zone) {
            const double a = M.dotA(M.Afp, cf, c0 + c, omega);
            fsum[face - 1] += a;
            if (a < 0.0) nin[face - 1]++;
            else if (a > 0.0) nout[face - 1]++;
          }
        }
      for (int face = 1; face <= nF; face++) {
        const int zo = M.zopp(face, zone);
        if (zo <= zone) continue;
        if (fsum[face - 1] < 0.0) {
          need[zone]++;
          xf(M.fopp(face, zone), zo) = 1;
          if (nout[face - 1] > 0 && !onCycle[zone]) { lag_zone(zone); onCycle[zone] = 1; }   // mixed-sign face
        } else if (fsum[face - 1] > 0.0) {
          need[zo]++;
          xf(face, zone) = 1;
          if (nin[face - 1] > 0 && !onCycle[zo]) { lag_zone(zo); onCycle[zo] = 1; }
        }
      }
    }
  }

  int seed() {
    int n = 0;
    for (int zone = 1; zone <= M.nz; zone++)
      if (need[zone] == 0) order[n++] = zone;
    if (n > 0) return n;
    // no zone is free of upstream neighbours: start from the boundary zone that needs the fewest
    int best = 0, bestNeed = M.nz;
    for (int zone = 1; zone <= M.nz; zone++)
      if (M.bzone && M.bzone[zone - 1] && need[zone] < bestNeed) { best = zone; bestNeed = need[zone]; }
    if (best == 0) { err = "no seed zone found for the sweep"; return 0; }
    order[0] = best;
    need[best] = 0;
    for (int face = 1; face <= M.zoneFaces[best - 1]; face++) {
      if (xf(face, best)) continue;
      const int zo = M.zopp(face, best);
      if (zo > 0) { lag_zone(zo); xf(M.fopp(face, best), zo) = 0; onCycle[zo] = 1; }
    }
    return 1;
  }

  void corner_orders(std::vector<int> &nextC) {
    const int mc = M.maxCorner;
    std::vector<int> cneed(mc), nds(mc), ds((size_t)mc * 8);
    for (int zone = 1; zone <= M.nz; zone++) {
      const int nC = M.numCorner[zone - 1], c0 = M.cOffSet[zone - 1];
      std::fill(cneed.begin(), cneed.end(), 0);
      std::fill(nds.begin(), nds.end(), 0);
      for (int c = 1; c <= nC; c++) {
        const int ncf = M.ndim == 2 ? 2 : M.nCFaces[c0 + c - 1];
        for (int cf = 1; cf <= ncf; cf++) {
          const int ce = M.cez(cf, c0 + c);
          if (ce <= c) continue;
          const double a = M.dotA(M.Aez, cf, c0 + c, omega);
          if (a < 0.0) { cneed[c - 1]++; ds[(size_t)(ce - 1) * 8 + nds[ce - 1]++] = c; }
          else if (a > 0.0) { cneed[ce - 1]++; ds[(size_t)(c - 1) * 8 + nds[c - 1]++] = ce; }
        }
      }
      bool cyc = false;
      for (int i = 1; i <= nC; i++) {
        int c = 1;
        for (int k = 2; k <= nC; k++) if (cneed[k - 1] < cneed[c - 1]) c = k;   // first minimum
        if (cneed[c - 1] != 0) cyc = true;
        nextC[c0 + i - 1] = c;
        for (int k = 0; k < nds[c - 1]; k++) cneed[ds[(size_t)(c - 1) * 8 + k] - 1]--;
        cneed[c - 1] = 99;
      }
      bad[zone] = cyc;
      if (cyc) {
        for (int i = 1; i <= nC; i++) nextC[c0 + i - 1] = i;
        lag_zone(zone);
      }
    }
  }

  // Tarjan SCC over the zones that are still waiting; every non-trivial component
  // has the links into its root cut (the upstream zones go on the cycle list).
  int break_cycles(int ndone, int &filled) {
    const int nz = M.nz, ngraph = nz - ndone;
    std::vector<int> waiting, released, dfn(nz + 1, 0), low(nz + 1, 0), stk, comp;
    std::vector<unsigned char> fresh(nz + 1, 1), onstk(nz + 1, 0);
    struct Frame { int zone, face, child; };
    std::vector<Frame> fr;
    for (int zone = 1; zone <= nz; zone++) {
      if (need[zone] == 0) fresh[zone] = 0;
      else waiting.push_back(zone);
    }
    if ((int)waiting.size() != ngraph) { err = "miscount of remaining zones while breaking cycles"; return 0; }
    int counter = 0;
    for (int root : waiting) {
      if (!fresh[root]) continue;
      auto enter = [&](int z) {
        dfn[z] = low[z] = ++counter;
        fresh[z] = 0;
        stk.push_back(z);
        onstk[z] = 1;
        fr.push_back({z, 1, 0});
      };
      enter(root);
      while (!fr.empty()) {
        const size_t top = fr.size() - 1;
        const int zone = fr[top].zone;
        if (fr[top].child) {
          low[zone] = std::min(low[zone], low[fr[top].child]);
          fr[top].child = 0;
        }
        bool down = false;
        const int nF = M.zoneFaces[zone - 1];
        while (fr[top].face <= nF) {
          const int face = fr[top].face++;
          if (!xf(face, zone)) continue;
          const int z2 = M.zopp(face, zone);
          if (z2 <= 0) continue;
          if (fresh[z2]) { fr[top].child = z2; enter(z2); down = true; break; }
          if (dfn[z2] < dfn[zone] && onstk[z2] && low[z2] < low[zone]) low[zone] = low[z2];
        }
        if (down) continue;
        if (low[zone] == dfn[zone]) {
          int z2 = stk.back(); stk.pop_back();
          onstk[z2] = 0;
          if (z2 != zone) {
            comp.clear();
            while (z2 != zone) { comp.push_back(z2); z2 = stk.back(); stk.pop_back(); }
            comp.push_back(z2);
            onstk[comp.front()] = 1;
            const int rootZ = comp.back();
            for (int face = 1; face <= M.zoneFaces[rootZ - 1]; face++) {
              const int zb = M.zopp(face, rootZ), fb = M.fopp(face, rootZ);
              if (zb > 0 && onstk[zb] && xf(fb, zb)) {
                if (!onCycle[zb]) { lag_zone(zb); onCycle[zb] = 1; }
                need[rootZ]--;
                xf(fb, zb) = 0;
                if (need[rootZ] == 0) released.push_back(rootZ);
              }
            }
            for (int z : comp) onstk[z] = 0;
          }
        }
        fr.pop_back();
      }
    }
    if (released.empty()) { err = "cycle detection failed, no dependencies broken"; return 0; }
    int added = 0;
    for (int zone : released) {
      if (need[zone] == 0) { order[filled++] = zone; added++; }
      else if (need[zone] < 0) { err = "needZ < 0 after breaking cycles"; return 0; }
    }
    if (added == 0) err = "cycles found, but not broken";
    return added;
  }
};

}  // namespace

static void fill_view(umt_ctx *ctx, MeshView &M) {
  M.ndim = ctx->ndim; M.nz = ctx->nz; M.nc = ctx->nc; M.nb = ctx->nb; M.mcf = ctx->maxcf; M.mf = ctx->maxFaces; M.maxCorner = ctx->maxCorner;
  M.numCorner = ctx->h_numCorner.data(); M.cOffSet = ctx->h_cOffSet.data(); M.nCFaces = ctx->h_nCFaces.data();
  M.cFP = ctx->h_cFP.data(); M.cEZ = ctx->h_cEZ.data(); M.zoneFaces = ctx->h_zoneFaces.data();
  M.zoneOpp = ctx->h_zoneOpp.data(); M.faceOpp = ctx->h_faceOpp.data(); M.CToFace = ctx->h_CToFace.data();
  M.bzone = ctx->h_BoundaryZone.empty() ? nullptr : ctx->h_BoundaryZone.data();
  M.Afp = ctx->h_Afp.data(); M.Aez = ctx->h_Aez.data();
}

// sweep order of an arbitrary set of ordinates (the GTA angle set: rtorder.F90 runs snnext for it as well)
int umt_host_build_order(umt_ctx *ctx, const double *omegas, int nAng, std::vector<int> &nHyp, std::vector<std::vector<int>> &zonesInPlane,
                         std::vector<std::vector<int>> &nextZ, std::vector<std::vector<int>> &nextC) {
  MeshView M;
  fill_view(ctx, M);
  nHyp.assign(nAng, 0); zonesInPlane.assign(nAng, {}); nextZ.assign(nAng, {}); nextC.assign(nAng, {});
  std::vector<AngleSchedule> res(nAng);
  std::vector<std::thread> pool;
  for (int a = 0; a < nAng; a++)
    pool.emplace_back([&, a]() { OrderBuilder ob(M, omegas + (size_t)a * ctx->ndim); ob.run(res[a]); });
  for (auto &t : pool) t.join();
  for (int a = 0; a < nAng; a++) {
    if (!res[a].error.empty()) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "GTA sweep order, angle %d: %s", a + 1, res[a].error.c_str());
    nHyp[a] = res[a].nHyp;
    zonesInPlane[a] = std::move(res[a].zonesInPlane);
    nextZ[a] = std::move(res[a].nextZ);
    nextC[a] = std::move(res[a].nextC);
  }
  return UMT_OK;
}

int umt_host_build_schedule(umt_ctx *ctx) {
  const int NA = ctx->NA, nd = ctx->ndim;
  MeshView M;
  M.ndim = nd; M.nz = ctx->nz; M.nc = ctx->nc; M.nb = ctx->nb; M.mcf = ctx->maxcf; M.mf = ctx->maxFaces; M.maxCorner = ctx->maxCorner;
  M.numCorner = ctx->h_numCorner.data(); M.cOffSet = ctx->h_cOffSet.data(); M.nCFaces = ctx->h_nCFaces.data();
  M.cFP = ctx->h_cFP.data(); M.cEZ = ctx->h_cEZ.data(); M.zoneFaces = ctx->h_zoneFaces.data();
  M.zoneOpp = ctx->h_zoneOpp.data(); M.faceOpp = ctx->h_faceOpp.data(); M.CToFace = ctx->h_CToFace.data();
  M.bzone = ctx->h_BoundaryZone.empty() ? nullptr : ctx->h_BoundaryZone.data();
  M.Afp = ctx->h_Afp.data(); M.Aez = ctx->h_Aez.data();

  // boundary-element area vectors are the A_fp of the corner face they sit on
  // (volumeUCBxyz.F90 / volumeUCBrz.F90 compute the same expression)
  if (!ctx->have_abdy) {
    ctx->h_Abdy.assign((size_t)nd * std::max(ctx->nb, 1), 0.0);
    for (int c = 0; c < ctx->nc; c++)
      for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
        const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
        if (v > ctx->nc)
          for (int d = 0; d < nd; d++) ctx->h_Abdy[(size_t)(v - ctx->nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * ctx->maxcf + f) * nd + d];
      }
    ctx->have_abdy = true;
  }
  if ((int)ctx->h_BdyToC.size() != ctx->nb) {
    ctx->h_BdyToC.assign(ctx->nb, 0);
    for (int c = 0; c < ctx->nc; c++)
      for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
        const int v = ctx->h_cFP[(size_t)c * ctx->maxcf + f];
        if (v > ctx->nc) ctx->h_BdyToC[v - ctx->nc - 1] = c + 1;
      }
  }
  std::vector<unsigned char> isShared(std::max(ctx->nb, 1), 0);
  for (const auto &s : ctx->shared)
    for (int b = s.first; b < s.first + s.n; b++) isShared[b] = 1;

  std::vector<AngleSchedule> res(NA);
  std::atomic<int> next(0);
  unsigned nthr = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if ((unsigned)NA < nthr) nthr = NA;
  auto work = [&]() {
    for (;;) {
      const int a = next.fetch_add(1);
      if (a >= NA) break;
      if (nd == 2 && ctx->h_finish[a]) continue;   // finishing directions are not swept (rtorder.F90)
      OrderBuilder ob(M, &ctx->h_omega[(size_t)a * nd]);
      ob.run(res[a]);
    }
  };
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < nthr; t++) pool.emplace_back(work);
  work();
  for (auto &t : pool) t.join();

  for (int a = 0; a < NA; a++) {
    if (!res[a].error.empty()) UMT_FAIL(ctx, UMT_ERR_SCHEDULE, "umt_build_schedule: angle %d: %s", a + 1, res[a].error.c_str());
    ctx->nHyp[a] = res[a].nHyp;
    ctx->zonesInPlane[a] = std::move(res[a].zonesInPlane);
    ctx->nextZ[a] = std::move(res[a].nextZ);
    ctx->nextC[a] = std::move(res[a].nextC);
    ctx->cycleList[a] = std::move(res[a].cycleList);
    ctx->numCycles[a] = (int)ctx->cycleList[a].size();
    ctx->nBad[a] = res[a].nBad;
    // exit list: non-shared boundaries in boundary order, then shared send lists
    auto &bl = ctx->bdyList[a];
    bl.clear();
    const double *om = &ctx->h_omega[(size_t)a * nd];
    for (int pass = 0; pass < 2; pass++)
      for (int b = 0; b < ctx->nb; b++) {
        if (isShared[b] != pass) continue;
        double dot = 0.0;
        for (int d = 0; d < nd; d++) dot += om[d] * ctx->h_Abdy[(size_t)b * nd + d];
        if (dot > 0.0) { bl.push_back(b + 1); bl.push_back(ctx->h_BdyToC[b]); }
      }
  }
  ctx->sched_dirty = true;
  ctx->exch_dirty = true;
  return UMT_OK;
}
