// Corner geometry on the device (Y1): A_fp, A_ez, Volume (+ Area, RadiusFP, RadiusEZ
// in RZ) from the corner coordinates px(ndim,nc).  Same quantities as
// rt/geometryUCBxyz.F90:87-179 / rt/geometryUCBrz.F90, but gathered: one thread
// per corner recomputes whatever neighbour value the reference's zone-serial
// "equal and opposite" fix-ups would have stored, so no write ordering is needed.
#include "umt_internal.h"

namespace {

struct GeomParams {
  int ndim, nz, nc, nb, mcf, mf;
  const int *numCorner, *cOffSet, *nCFaces, *cFP /*0-based row*/, *cEZ /*0-based*/, *CToFace /*1-based*/, *zoneOpp /*1-based*/, *c2z;
  const double *px;
  double *Volume, *Afp, *Aez, *Area, *RadiusFP, *RadiusEZ;
};

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 ld3(const double *p, int c) { return {p[3 * (size_t)c], p[3 * (size_t)c + 1], p[3 * (size_t)c + 2]}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }

__device__ V3 zone_center3(const GeomParams &P, int zone) {
  const int n = P.numCorner[zone], c0 = P.cOffSet[zone];
  V3 s = {0, 0, 0};
  for (int c = 0; c < n; c++) s = s + ld3(P.px, c0 + c);
  const double inv = (double)n;
  return {s.x / inv, s.y / inv, s.z / inv};
}
__device__ V3 face_center3(const GeomParams &P, int zone, int face1) {
  const int n = P.numCorner[zone], c0 = P.cOffSet[zone];
  V3 s = {0, 0, 0};
  int k = 0;
  for (int c = 0; c < n; c++)
    for (int f = 0; f < P.nCFaces[c0 + c]; f++)
      if (P.CToFace[(size_t)(c0 + c) * P.mcf + f] == face1) { s = s + ld3(P.px, c0 + c); k++; }
  const double inv = (double)k;
  return {s.x / inv, s.y / inv, s.z / inv};
}
// outward area vector of FP corner face f of (zone-local) corner c, from this zone's own points
__device__ V3 fep3(const GeomParams &P, int zone, int c, int f) {
  const int c0 = P.cOffSet[zone], cc = c0 + c, n = P.nCFaces[cc];
  const int f2 = (f + 1) % n;
  const int e1 = P.cEZ[(size_t)cc * P.mcf + f], e2 = P.cEZ[(size_t)cc * P.mcf + f2];
  const V3 fc = face_center3(P, zone, P.CToFace[(size_t)cc * P.mcf + f]);
  const V3 p0 = ld3(P.px, cc), p1 = ld3(P.px, c0 + e1), p2 = ld3(P.px, c0 + e2);
  const V3 tdl = 0.5 * (p1 - p2), tfl = fc - p0;
  return {0.5 * (tfl.z * tdl.y - tfl.y * tdl.z), 0.5 * (tfl.x * tdl.z - tfl.z * tdl.x), 0.5 * (tfl.y * tdl.x - tfl.x * tdl.y)};
}
// un-symmetrised A_ez of EZ face f of corner c: z1(f) + z2(previous face)
__device__ V3 ez3(const GeomParams &P, int zone, int c, int f, V3 zc) {
  const int c0 = P.cOffSet[zone], cc = c0 + c, n = P.nCFaces[cc];
  const V3 p0 = ld3(P.px, cc);
  V3 acc = {0, 0, 0};
  {
    const int e1 = P.cEZ[(size_t)cc * P.mcf + f];
    const V3 fc = face_center3(P, zone, P.CToFace[(size_t)cc * P.mcf + f]);
    const V3 tfz = fc - zc, tfe1 = fc - 0.5 * (ld3(P.px, c0 + e1) + p0);
    acc = {0.5 * (tfz.z * tfe1.y - tfz.y * tfe1.z), 0.5 * (tfz.x * tfe1.z - tfz.z * tfe1.x), 0.5 * (tfz.y * tfe1.x - tfz.x * tfe1.y)};
  }
  {
    const int fp = (f + n - 1) % n;                 // the face whose "cface2" is f
    const int e2 = P.cEZ[(size_t)cc * P.mcf + f];   // cEZ(cface2) of that face
    const V3 fc = face_center3(P, zone, P.CToFace[(size_t)cc * P.mcf + fp]);
    const V3 tfz = fc - zc, tfe2 = fc - 0.5 * (ld3(P.px, c0 + e2) + p0);
    const V3 z2 = {0.5 * (tfz.y * tfe2.z - tfz.z * tfe2.y), 0.5 * (tfz.z * tfe2.x - tfz.x * tfe2.z), 0.5 * (tfz.x * tfe2.y - tfz.y * tfe2.x)};
    acc = n == 1 ? acc + z2 : (f == 0 ? acc + z2 : z2 + acc);
  }
  return acc;
}

__global__ void geometry_xyz_kernel(GeomParams P) {
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= P.nc) return;
  const int zone = P.c2z[cc], c0 = P.cOffSet[zone], c = cc - c0, n = P.nCFaces[cc];
  const V3 zc = zone_center3(P, zone);
  const V3 tzl = zc - ld3(P.px, cc);
  double vol = 0.0;
  for (int f = 0; f < n; f++) {
    const int face1 = P.CToFace[(size_t)cc * P.mcf + f];
    const int zo = P.zoneOpp[(size_t)zone * P.mf + face1 - 1];
    V3 A;
    if (zo > 0 && zo - 1 < zone) {
      // the lower-numbered zone's vector wins, negated (geometryUCBxyz.F90:114-127)
      const int nb = P.cFP[(size_t)cc * P.mcf + f];
      const int zn = P.c2z[nb];
      A = {0, 0, 0};
      for (int g = 0; g < P.nCFaces[nb]; g++)
        if (P.cFP[(size_t)nb * P.mcf + g] == cc) { const V3 t = fep3(P, zn, nb - P.cOffSet[zn], g); A = {-t.x, -t.y, -t.z}; }
    } else {
      A = fep3(P, zone, c, f);
    }
    double *o = P.Afp + ((size_t)cc * P.mcf + f) * 3;
    o[0] = A.x; o[1] = A.y; o[2] = A.z;
    vol += (1.0 / 3.0) * fabs(tzl.x * A.x + tzl.y * A.y + tzl.z * A.z);
    // EZ faces: the lower zone-local corner's vector wins (geometryUCBxyz.F90:167-179)
    const int e = P.cEZ[(size_t)cc * P.mcf + f];
    V3 E;
    if (e > c) E = ez3(P, zone, c, f, zc);
    else {
      E = {0, 0, 0};
      for (int g = 0; g < P.nCFaces[c0 + e]; g++)
        if (P.cEZ[(size_t)(c0 + e) * P.mcf + g] == c) { const V3 t = ez3(P, zone, e, g, zc); E = {-t.x, -t.y, -t.z}; }
    }
    double *q = P.Aez + ((size_t)cc * P.mcf + f) * 3;
    q[0] = E.x; q[1] = E.y; q[2] = E.z;
  }
  P.Volume[cc] = vol;
}

// ---- RZ ---------------------------------------------------------------------
struct RZown { double fp[2][2], ez[2][2]; };
__device__ void rz_own(const GeomParams &P, int zone, int c, RZown &o, double *area, double *vol, double *rfp, double *rez) {
  const int c0 = P.cOffSet[zone], n = P.numCorner[zone];
  double rz = 0, zz = 0;
  for (int k = 0; k < n; k++) { rz += P.px[2 * (size_t)(c0 + k)]; zz += P.px[2 * (size_t)(c0 + k) + 1]; }
  rz /= (double)n; zz /= (double)n;
  const int c1 = P.cEZ[(size_t)(c0 + c) * 2], c2 = P.cEZ[(size_t)(c0 + c) * 2 + 1];
  const double rp = P.px[2 * (size_t)(c0 + c)], zp = P.px[2 * (size_t)(c0 + c) + 1];
  const double r1 = P.px[2 * (size_t)(c0 + c1)], z1 = P.px[2 * (size_t)(c0 + c1) + 1];
  const double r2 = P.px[2 * (size_t)(c0 + c2)], z2 = P.px[2 * (size_t)(c0 + c2) + 1];
  const double re1 = 0.5 * (rp + r1), ze1 = 0.5 * (zp + z1), re2 = 0.5 * (rp + r2), ze2 = 0.5 * (zp + z2);
  o.fp[0][0] = 0.5 * (z2 - zp); o.fp[0][1] = 0.5 * (rp - r2);
  o.fp[1][0] = 0.5 * (zp - z1); o.fp[1][1] = 0.5 * (r1 - rp);
  o.ez[0][0] = ze1 - zz; o.ez[0][1] = rz - re1;
  o.ez[1][0] = zz - ze2; o.ez[1][1] = re2 - rz;
  if (area) {
    rfp[0] = 0.5 * (rp + re2); rfp[1] = 0.5 * (rp + re1);
    rez[0] = 0.5 * (rz + re1); rez[1] = 0.5 * (rz + re2);
    const double a1 = fabs((re2 - rp) * (zz - zp) - (ze2 - zp) * (rz - rp));
    const double a2 = fabs((rz - rp) * (ze1 - zp) - (zz - zp) * (re1 - rp));
    *area = 0.5 * (a1 + a2);
    const double rb1 = (1.0 / 3.0) * (rp + re2 + rz), rb2 = (1.0 / 3.0) * (rp + re1 + rz);
    *vol = 0.5 * (rb1 * a1 + rb2 * a2);
  }
}

__global__ void geometry_rz_kernel(GeomParams P) {
  const int cc = blockIdx.x * blockDim.x + threadIdx.x;
  if (cc >= P.nc) return;
  const int zone = P.c2z[cc], c0 = P.cOffSet[zone], c = cc - c0;
  RZown me;
  double area, vol, rfp[2], rez[2];
  rz_own(P, zone, c, me, &area, &vol, rfp, rez);
  P.Area[cc] = area; P.Volume[cc] = vol;
  P.RadiusFP[2 * (size_t)cc] = rfp[0]; P.RadiusFP[2 * (size_t)cc + 1] = rfp[1];
  P.RadiusEZ[2 * (size_t)cc] = rez[0]; P.RadiusEZ[2 * (size_t)cc + 1] = rez[1];
  for (int s = 0; s < 2; s++) {
    // FP: the reference's zone-serial fix-up (geometryUCBrz.F90:92-109) leaves the value
    // written last; candidates are my own vector (written when local id < partner's global
    // id, 1-based) and the partner's negated one (same test on its side).
    const int nb = P.cFP[(size_t)cc * 2 + s];
    double A0 = me.fp[s][0], A1 = me.fp[s][1];
    if (nb < P.nc) {
      const int zn = P.c2z[nb], ln = nb - P.cOffSet[zn];
      const bool mineValid = (c + 1) < (nb + 1), theirsValid = (ln + 1) < (cc + 1);
      if (theirsValid && (!mineValid || nb > cc)) {
        RZown o;
        rz_own(P, zn, ln, o, nullptr, nullptr, nullptr, nullptr);
        A0 = -o.fp[1 - s][0]; A1 = -o.fp[1 - s][1];
      }
    }
    P.Afp[((size_t)cc * 2 + s) * 2] = A0; P.Afp[((size_t)cc * 2 + s) * 2 + 1] = A1;
    // EZ: lower zone-local corner wins (geometryUCBrz.F90:113-125)
    const int e = P.cEZ[(size_t)cc * 2 + s];
    double E0 = me.ez[s][0], E1 = me.ez[s][1];
    if (e < c) {
      RZown o;
      rz_own(P, zone, e, o, nullptr, nullptr, nullptr, nullptr);
      E0 = -o.ez[1 - s][0]; E1 = -o.ez[1 - s][1];
    }
    P.Aez[((size_t)cc * 2 + s) * 2] = E0; P.Aez[((size_t)cc * 2 + s) * 2 + 1] = E1;
  }
}

template <class T>
int up(umt_ctx *ctx, T **d, const std::vector<T> &h) {
  UMT_CUDA(ctx, cudaMalloc((void **)d, sizeof(T) * std::max<size_t>(h.size(), 1)));
  UMT_CUDA(ctx, umt_memcpy(ctx, *d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
  return UMT_OK;
}
}  // namespace

extern "C" int umt_compute_geometry(umt_ctx *ctx, const double *px) {
  if (!ctx || !px) return UMT_ERR_ARG;
  if (!ctx->have_conn || ctx->h_CToFace.empty()) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_compute_geometry: needs full connectivity (CToFace, zoneOpp)");
  if (ctx->device < 0) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_compute_geometry: host-only context (device -1) cannot run kernels");
  UMT_CUDA(ctx, cudaSetDevice(ctx->device));
  const int nd = ctx->ndim, nc = ctx->nc, nz = ctx->nz, mcf = ctx->maxcf;
  std::vector<int> c2z(nc, 0);
  for (int z = 0; z < nz; z++)
    for (int c = 0; c < ctx->h_numCorner[z]; c++) c2z[ctx->h_cOffSet[z] + c] = z;
  std::vector<double> hpx(px, px + (size_t)nd * nc);
  int *d_c2z = nullptr, *d_ctf = nullptr, *d_zopp = nullptr;
  double *d_px = nullptr;
  int rc = up(ctx, &d_c2z, c2z);
  if (!rc) rc = up(ctx, &d_ctf, ctx->h_CToFace);
  if (!rc) rc = up(ctx, &d_zopp, ctx->h_zoneOpp);
  if (!rc) rc = up(ctx, &d_px, hpx);
  auto alloc = [&](double **p, size_t n) { if (!*p) return cudaMalloc((void **)p, sizeof(double) * n); return cudaSuccess; };
  if (!rc) {
    cudaError_t e = alloc(&ctx->d_Volume, nc);
    if (e == cudaSuccess) e = alloc(&ctx->d_Afp, (size_t)nd * mcf * nc);
    if (e == cudaSuccess) e = alloc(&ctx->d_Aez, (size_t)nd * mcf * nc);
    if (e == cudaSuccess && nd == 2) e = alloc(&ctx->d_Area, nc);
    if (e == cudaSuccess && nd == 2) e = alloc(&ctx->d_RadiusFP, 2 * (size_t)nc);
    if (e == cudaSuccess && nd == 2) e = alloc(&ctx->d_RadiusEZ, 2 * (size_t)nc);
    if (e != cudaSuccess) { ctx->err = std::string("umt_compute_geometry: ") + cudaGetErrorString(e); rc = UMT_ERR_CUDA; }
  }
  if (!rc) {
    GeomParams P;
    P.ndim = nd; P.nz = nz; P.nc = nc; P.nb = ctx->nb; P.mcf = mcf; P.mf = ctx->maxFaces;
    P.numCorner = ctx->d_numCorner; P.cOffSet = ctx->d_cOffSet; P.nCFaces = ctx->d_nCFaces; P.cFP = ctx->d_cFP; P.cEZ = ctx->d_cEZ;
    P.CToFace = d_ctf; P.zoneOpp = d_zopp; P.c2z = d_c2z; P.px = d_px;
    P.Volume = ctx->d_Volume; P.Afp = ctx->d_Afp; P.Aez = ctx->d_Aez; P.Area = ctx->d_Area; P.RadiusFP = ctx->d_RadiusFP; P.RadiusEZ = ctx->d_RadiusEZ;
    const int grid = (nc + 127) / 128;
    if (nd == 3) geometry_xyz_kernel<<<grid, 128, 0, ctx->stream>>>(P);
    else geometry_rz_kernel<<<grid, 128, 0, ctx->stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = std::string("geometry kernel: ") + cudaGetErrorString(e); rc = UMT_ERR_CUDA; }
  }
  if (!rc) {   // host mirrors for the schedule builder
    ctx->h_Volume.resize(nc); ctx->h_Afp.resize((size_t)nd * mcf * nc); ctx->h_Aez.resize((size_t)nd * mcf * nc);
    umt_memcpy(ctx, ctx->h_Volume.data(), ctx->d_Volume, sizeof(double) * nc, cudaMemcpyDeviceToHost);
    umt_memcpy(ctx, ctx->h_Afp.data(), ctx->d_Afp, sizeof(double) * ctx->h_Afp.size(), cudaMemcpyDeviceToHost);
    umt_memcpy(ctx, ctx->h_Aez.data(), ctx->d_Aez, sizeof(double) * ctx->h_Aez.size(), cudaMemcpyDeviceToHost);
    if (nd == 2) {
      ctx->h_Area.resize(nc); ctx->h_RadiusFP.resize(2 * (size_t)nc); ctx->h_RadiusEZ.resize(2 * (size_t)nc);
      umt_memcpy(ctx, ctx->h_Area.data(), ctx->d_Area, sizeof(double) * nc, cudaMemcpyDeviceToHost);
      umt_memcpy(ctx, ctx->h_RadiusFP.data(), ctx->d_RadiusFP, sizeof(double) * 2 * nc, cudaMemcpyDeviceToHost);
      umt_memcpy(ctx, ctx->h_RadiusEZ.data(), ctx->d_RadiusEZ, sizeof(double) * 2 * nc, cudaMemcpyDeviceToHost);
    }
    ctx->have_geom = true; ctx->have_abdy = false; ctx->sched_dirty = true;
  }
  cudaFree(d_c2z); cudaFree(d_ctf); cudaFree(d_zopp); cudaFree(d_px);
  return rc;
}

extern "C" int umt_download_geometry(umt_ctx *ctx, double *Volume, double *A_fp, double *A_ez, double *Area, double *RadiusFP,
                                     double *RadiusEZ, double *A_bdy, double *VolumeZone) {
  if (!ctx) return UMT_ERR_ARG;
  if (!ctx->have_geom) UMT_FAIL(ctx, UMT_ERR_STATE, "umt_download_geometry: no geometry");
  const int nd = ctx->ndim, nc = ctx->nc, mcf = ctx->maxcf;
  if (Volume) std::copy(ctx->h_Volume.begin(), ctx->h_Volume.end(), Volume);
  if (A_fp) std::copy(ctx->h_Afp.begin(), ctx->h_Afp.end(), A_fp);
  if (A_ez) std::copy(ctx->h_Aez.begin(), ctx->h_Aez.end(), A_ez);
  if (nd == 2) {
    if (Area) std::copy(ctx->h_Area.begin(), ctx->h_Area.end(), Area);
    if (RadiusFP) std::copy(ctx->h_RadiusFP.begin(), ctx->h_RadiusFP.end(), RadiusFP);
    if (RadiusEZ) std::copy(ctx->h_RadiusEZ.begin(), ctx->h_RadiusEZ.end(), RadiusEZ);
  }
  if (A_bdy)   // boundary elements carry the A_fp of their corner face (volumeUCBxyz.F90 / volumeUCBrz.F90)
    for (int c = 0; c < nc; c++)
      for (int f = 0; f < ctx->h_nCFaces[c]; f++) {
        const int v = ctx->h_cFP[(size_t)c * mcf + f];
        if (v > nc) for (int d = 0; d < nd; d++) A_bdy[(size_t)(v - nc - 1) * nd + d] = ctx->h_Afp[((size_t)c * mcf + f) * nd + d];
      }
  if (VolumeZone)
    for (int z = 0; z < ctx->nz; z++) {
      double s = 0.0;
      for (int c = 0; c < ctx->h_numCorner[z]; c++) s += ctx->h_Volume[ctx->h_cOffSet[z] + c];
      VolumeZone[z] = s;
    }
  return UMT_OK;
}
