/*
 * umt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, line-by-line CPU restatement of the reference's (LLNL/UMT, Teton
 * 5.3.0) discrete-ordinates sweep hot path.  It exists so that the CUDA path in
 * umt_b200/csrc can be checked against "what Teton's CPU build computes" inside a
 * container that has no Fortran compiler, MPI or Conduit.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product never does.
 *
 * PARITY STATUS.  Pinned to real reference code: orc_sweep_xyz (SweepUCBxyz, source term,
 * Phi accumulation, exiting PsiB, cycle lists, the direct-solve branch of zones with an
 * intra-zone cycle) is checked on the GPU box against the reference's OWN implementation of
 * the same routine, gpu/GPU_SweepUCBxyz.cu, compiled unmodified by oracle/Makefile into
 * oracle/_ref/libgpu_sweepucbxyz_ref.so (tests/test_gpu_reference_cuda.py, 1e-12); the
 * Planck-group integrals are checked against misc/NormalizedBlackBody.cc compiled the same
 * way (oracle/_ref/libnbb_ref.so).  "parity unpinned" for everything else (quadrature,
 * geometry, snnext schedules, the r-z sweep, exchange, GTA): the Fortran reference cannot be
 * compiled here and ships no golden vectors, known-answer tests or fixtures for this path
 * (SURVEY.md section 4, section 8c); those parts are pinned by (a) following the cited
 * Fortran line by line, (b) the analytic invariants of SURVEY.md section 4
 * (tests/test_oracle_invariants.py) and (c) the committed fixtures in tests/golden/ that
 * freeze the restatement's own output.
 *
 * Conventions: every integer id stored in an array is 1-based exactly as
 * Teton's Fortran holds it; arrays are the Fortran column-major memory image
 * (first index fastest).  The F2/F3 macros below index them with 1-based
 * subscripts so the code reads like the cited source.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../umt_b200/csrc/quad_tables.inc"

#define F2(a, i, j, n1) ((a)[((i) - 1) + (size_t)(n1) * ((j) - 1)])
#define F3(a, i, j, k, n1, n2) ((a)[((i) - 1) + (size_t)(n1) * (((j) - 1) + (size_t)(n2) * ((k) - 1))])

typedef struct {
  int ndim, nzones, ncornr, nbelem, maxcf, maxCorner, maxFaces;
  const int *numCorner;   /* (nz) */
  const int *cOffSet;     /* (nz) 0-based offset: corners of zone are cOffSet+1..cOffSet+numCorner */
  const int *zoneFaces;   /* (nz) */
  const int *zoneOpp;     /* (maxFaces,nz) */
  const int *faceOpp;     /* (maxFaces,nz) */
  const int *nCFaces;     /* (nc) */
  const int *cFP;         /* (maxcf,nc) */
  const int *cEZ;         /* (maxcf,nc) */
  const int *CToFace;     /* (maxcf,nc) */
  const unsigned char *BoundaryZone; /* (nz) */
  const double *px;       /* (ndim,nc) */
} orc_mesh;

static void die(const char *msg) {
  fprintf(stderr, "umt_oracle fatal: %s\n", msg);
  abort();
}

/* ------------------------------------------------------------------ */
/* Quadrature: rt/quadProduct.F90:21-186, rt/quadrz.F90 (product       */
/* branch), rt/rtquad.F90:95-127, rt/AngleCoef2D.F90,                  */
/* mods/AngleSet_mod.F90:310-347                                        */
/* ------------------------------------------------------------------ */
static int tfirst(int n) { return n * (n - 1) / 2 + 1; } /* QuadratureData_mod.F90:1144 */
static int tlast(int n) { return n * (n + 1) / 2; }      /* :1145 */

static void rtquad_normalize(int NumAngles, double wtiso, double *weight, unsigned char *start, unsigned char *finish) {
  /* rtquad.F90:95-127 */
  double sumwgt = 0.0;
  for (int ia = 0; ia < NumAngles; ia++) sumwgt += weight[ia];
  double fac = 1.0 / (wtiso * sumwgt);
  for (int ia = 0; ia < NumAngles; ia++) weight[ia] = fac * weight[ia];
  int iang = -1;
  const double eps = 2.220446049250313e-16; /* adqtEpsilon = epsilon(one) */
  for (int ia = 0; ia < NumAngles; ia++) {
    if (start) start[ia] = 0;
    if (finish) finish[ia] = 0;
    if (weight[ia] < eps) {
      if (iang == -1) { if (start) start[ia] = 1; iang = -iang; }
      else if (iang == 1) { if (finish) finish[ia] = 1; iang = -iang; }
    }
  }
}

int orc_quad_xyz(int npolar, int nazimuthal, int polaraxis, double *omega /* (3,NA) */, double *weight /* (NA) */) {
  const double pi = 3.14159265358979323846;
  if (npolar < 1 || npolar > 32 || nazimuthal < 1 || nazimuthal > 32 || polaraxis < 1 || polaraxis > 3) return -1;
  int nangoct = npolar * nazimuthal;
  double *ox = malloc(sizeof(double) * nangoct), *oy = malloc(sizeof(double) * nangoct),
         *oz = malloc(sizeof(double) * nangoct), *qw = malloc(sizeof(double) * nangoct);
  int m = 0;
  for (int iPhi = tfirst(nazimuthal); iPhi <= tlast(nazimuthal); iPhi++) { /* quadProduct.F90:97-111 */
    double cosinePhi = UMT_QT_cosPhiXYZ[iPhi - 1];
    for (int jTheta = tlast(npolar); jTheta >= tfirst(npolar); jTheta--) {
      double cosineTheta = UMT_QT_cosTheta[jTheta - 1];
      double sineTheta = sqrt(1.0 - cosineTheta * cosineTheta);
      if (polaraxis == 1) {
        ox[m] = cosineTheta; oy[m] = sineTheta * cosinePhi;
        oz[m] = sqrt(1.0 - ox[m] * ox[m] - oy[m] * oy[m]);
      } else if (polaraxis == 2) {
        oy[m] = cosineTheta; oz[m] = sineTheta * cosinePhi;
        ox[m] = sqrt(1.0 - oy[m] * oy[m] - oz[m] * oz[m]);
      } else {
        oz[m] = cosineTheta; ox[m] = sineTheta * cosinePhi;
        oy[m] = sqrt(1.0 - ox[m] * ox[m] - oz[m] * oz[m]);
      }
      qw[m] = UMT_QT_weightTheta[jTheta - 1] * UMT_QT_weightPhiXYZ[iPhi - 1];
      m++;
    }
  }
  static const int sx[8] = {1, -1, -1, 1, 1, -1, -1, 1};  /* quadProduct.F90:122-182 */
  static const int sy[8] = {1, 1, -1, -1, 1, 1, -1, -1};
  static const int sz[8] = {1, 1, 1, 1, -1, -1, -1, -1};
  int nn = 0;
  for (int i = 0; i < nangoct; i++) {
    for (int o = 0; o < 8; o++) {
      omega[3 * (nn + o) + 0] = sx[o] * ox[i];
      omega[3 * (nn + o) + 1] = sy[o] * oy[i];
      omega[3 * (nn + o) + 2] = sz[o] * oz[i];
      weight[nn + o] = qw[i];
    }
    nn += 8;
  }
  free(ox); free(oy); free(oz); free(qw);
  rtquad_normalize(8 * nangoct, 1.0 / (4.0 * pi), weight, NULL, NULL); /* Size_mod.F90:281 */
  return 8 * nangoct;
}

int orc_quad_rz(int npolar, int nazimuthal, double *omega /* (2,NA) */, double *weight,
                unsigned char *start, unsigned char *finish, int *angleToLevel /* 1-based xi-level */,
                double *alpha, double *tauc, double *angDerivFac, double *quadTauW1, double *quadTauW2) {
  const double pi = 3.14159265358979323846;
  if (npolar < 1 || npolar > 32 || nazimuthal < 1 || nazimuthal > 32) return -1;
  int m = 0;
  for (int jTheta = tlast(npolar); jTheta >= tfirst(npolar); jTheta--) { /* quadrz.F90 product branch */
    double cosineTheta = UMT_QT_cosTheta[jTheta - 1];
    double sineTheta = sqrt(1.0 - cosineTheta * cosineTheta);
    double xilev = cosineTheta;
    int Phi1 = tfirst(nazimuthal), Phi2 = tlast(nazimuthal);
    for (int half = 0; half < 2; half++) {
      double sgn = half == 0 ? -1.0 : 1.0;
      omega[2 * m] = -sqrt(1.0 - xilev * xilev); omega[2 * m + 1] = sgn * xilev; weight[m] = 0.0; m++;
      for (int iPhi = Phi2; iPhi >= Phi1; iPhi--) {
        omega[2 * m] = -sineTheta * UMT_QT_cosPhiRZ[iPhi - 1]; omega[2 * m + 1] = sgn * xilev;
        weight[m] = UMT_QT_weightTheta[jTheta - 1] * UMT_QT_weightPhiRZ[iPhi - 1]; m++;
      }
      for (int iPhi = Phi1; iPhi <= Phi2; iPhi++) {
        omega[2 * m] = sineTheta * UMT_QT_cosPhiRZ[iPhi - 1]; omega[2 * m + 1] = sgn * xilev;
        weight[m] = UMT_QT_weightTheta[jTheta - 1] * UMT_QT_weightPhiRZ[iPhi - 1]; m++;
      }
      omega[2 * m] = sqrt(1.0 - xilev * xilev); omega[2 * m + 1] = sgn * xilev; weight[m] = 0.0; m++;
    }
  }
  int NA = m;
  rtquad_normalize(NA, 1.0 / (2.0 * pi), weight, start, finish); /* Size_mod.F90:278 */
  /* AngleSet_mod.F90:325-335: xi-levels begin at each starting direction */
  int nLevels = 0;
  for (int n = 0; n < NA; n++) { if (start[n]) nLevels++; angleToLevel[n] = nLevels; }
  /* AngleCoef2D.F90 */
  int a = 0;
  while (a < NA) {
    int a1 = a, lev = angleToLevel[a];
    int a2 = a1; while (a2 + 1 < NA && angleToLevel[a2 + 1] == lev) a2++;
    double weightLevel = 0.0, Phimh = pi, Mumh = omega[0];
    for (int k = a1; k <= a2; k++) weightLevel += weight[k];
    for (int k = a1; k <= a2; k++) {
      if (start[k]) {
        alpha[k] = 0.0; tauc[k] = 0.0;
        if (k != a1) die("Mu not increasing in xi-level, AngleCoef2D");
        Phimh = pi; Mumh = omega[2 * k];
      } else if (finish[k]) {
        alpha[k] = 0.0; tauc[k] = 0.0;
      } else {
        alpha[k] = alpha[k - 1] - weight[k] * omega[2 * k];
        double Phiph = Phimh - weight[k] * pi / weightLevel;
        double Muph = sqrt(1.0 - omega[2 * k + 1] * omega[2 * k + 1]) * cos(Phiph);
        if (omega[2 * k] < Mumh || omega[2 * k] > Muph) die("Mu not between limits, AngleCoef2D");
        tauc[k] = (omega[2 * k] - Mumh) / (Muph - Mumh);
        Phimh = Phiph; Mumh = Muph;
      }
    }
    a = a2 + 1;
  }
  for (int n = 0; n < NA; n++) { /* AngleSet_mod.F90:337-347 */
    if (start[n] || finish[n]) { angDerivFac[n] = 0.0; quadTauW1[n] = 1.0; quadTauW2[n] = 0.0; }
    else {
      angDerivFac[n] = omega[2 * n] + alpha[n] / (weight[n] * tauc[n]);
      quadTauW1[n] = 1.0 / tauc[n];
      quadTauW2[n] = (1.0 - tauc[n]) / tauc[n];
    }
  }
  return NA;
}

/* rtquad.F90:95-127 normalisation and start/finish flags, AngleSet_mod.F90:325-335 xi-levels, AngleCoef2D.F90 and
   AngleSet_mod.F90:337-347 for any r-z ordinate set given level by level (same statements as the tail of orc_quad_rz; used for
   the level-symmetric GTA set) */
int orc_rz_angle_coefs(int NA, const double *omega, double *weight, unsigned char *start, unsigned char *finish, int *angleToLevel,
                       double *alpha, double *tauc, double *angDerivFac, double *quadTauW1, double *quadTauW2) {
  const double pi = 3.14159265358979323846;
  rtquad_normalize(NA, 1.0 / (2.0 * pi), weight, start, finish);
  int nLevels = 0;
  for (int n = 0; n < NA; n++) { if (start[n]) nLevels++; angleToLevel[n] = nLevels; }
  int a = 0;
  while (a < NA) {
    int a1 = a, lev = angleToLevel[a];
    int a2 = a1; while (a2 + 1 < NA && angleToLevel[a2 + 1] == lev) a2++;
    double weightLevel = 0.0, Phimh = pi, Mumh = omega[0];
    for (int k = a1; k <= a2; k++) weightLevel += weight[k];
    for (int k = a1; k <= a2; k++) {
      if (start[k]) {
        alpha[k] = 0.0; tauc[k] = 0.0;
        if (k != a1) die("Mu not increasing in xi-level, AngleCoef2D");
        Phimh = pi; Mumh = omega[2 * k];
      } else if (finish[k]) {
        alpha[k] = 0.0; tauc[k] = 0.0;
      } else {
        alpha[k] = alpha[k - 1] - weight[k] * omega[2 * k];
        double Phiph = Phimh - weight[k] * pi / weightLevel;
        double Muph = sqrt(1.0 - omega[2 * k + 1] * omega[2 * k + 1]) * cos(Phiph);
        if (omega[2 * k] < Mumh || omega[2 * k] > Muph) die("Mu not between limits, AngleCoef2D");
        tauc[k] = (omega[2 * k] - Mumh) / (Muph - Mumh);
        Phimh = Phiph; Mumh = Muph;
      }
    }
    a = a2 + 1;
  }
  for (int n = 0; n < NA; n++) {
    if (start[n] || finish[n]) { angDerivFac[n] = 0.0; quadTauW1[n] = 1.0; quadTauW2[n] = 0.0; }
    else {
      angDerivFac[n] = omega[2 * n] + alpha[n] / (weight[n] * tauc[n]);
      quadTauW1[n] = 1.0 / tauc[n];
      quadTauW2[n] = (1.0 - tauc[n]) / tauc[n];
    }
  }
  return NA;
}

/* ------------------------------------------------------------------ */
/* Geometry: mods/Geometry_mod.F90:344-441, rt/geometryUCBxyz.F90,     */
/* rt/volumeUCBxyz.F90, rt/geometryUCBrz.F90, rt/volumeUCBrz.F90        */
/* ------------------------------------------------------------------ */
static void zone_center(const orc_mesh *M, int zone, double *zc) {
  int nd = M->ndim, nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
  for (int d = 0; d < nd; d++) zc[d] = 0.0;
  for (int c = 1; c <= nCorner; c++)
    for (int d = 1; d <= nd; d++) zc[d - 1] += F2(M->px, d, c0 + c, nd);
  for (int d = 0; d < nd; d++) zc[d] = zc[d] / (double)nCorner;
}

static void face_centers(const orc_mesh *M, int zone, int nFaces, double *fc /* (3,nFaces) */) {
  int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
  int nc_face[64];
  for (int f = 0; f < nFaces; f++) { nc_face[f] = 0; fc[3 * f] = fc[3 * f + 1] = fc[3 * f + 2] = 0.0; }
  for (int c = 1; c <= nCorner; c++) {
    int ncf = M->nCFaces[c0 + c - 1];
    for (int cface = 1; cface <= ncf; cface++) {
      int face = F2(M->CToFace, cface, c0 + c, M->maxcf);
      nc_face[face - 1]++;
      for (int d = 1; d <= 3; d++) fc[3 * (face - 1) + d - 1] += F2(M->px, d, c0 + c, 3);
    }
  }
  for (int f = 0; f < nFaces; f++)
    for (int d = 0; d < 3; d++) fc[3 * f + d] = nc_face[f] != 0 ? fc[3 * f + d] / (double)nc_face[f] : 0.0;
}

/* which = 0: geometryUCBxyz (A_fp, A_ez, Volume); which = 1: volumeUCBxyz (Volume, VolumeZone, A_bdy) */
static void geom_xyz(const orc_mesh *M, int which, double *A_fp, double *A_ez, double *Volume,
                     double *VolumeZone, double *A_bdy) {
  const int mcf = M->maxcf;
  double fc[3 * 64];
  double *Afp_tmp = NULL;
  if (which == 1) { Afp_tmp = calloc((size_t)3 * mcf * M->ncornr, sizeof(double)); A_fp = Afp_tmp; }
  for (int zone = 1; zone <= M->nzones; zone++) {
    int nFaces = M->zoneFaces[zone - 1];
    double zc[3];
    zone_center(M, zone, zc);
    face_centers(M, zone, nFaces, fc);
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    if (which == 1) VolumeZone[zone - 1] = 0.0;
    for (int c = 1; c <= nCorner; c++) {
      Volume[c0 + c - 1] = 0.0;
      if (which == 0)
        for (int f = 1; f <= mcf; f++) for (int d = 1; d <= 3; d++) F3(A_ez, d, f, c0 + c, 3, mcf) = 0.0;
    }
    for (int c = 1; c <= nCorner; c++) {
      int cc = c0 + c, nCFaces = M->nCFaces[cc - 1];
      for (int cface1 = 1; cface1 <= nCFaces; cface1++) {
        int cface2 = cface1 % nCFaces + 1;
        int cfp = F2(M->cFP, cface1, cc, mcf);
        int cez1 = F2(M->cEZ, cface1, cc, mcf), cez2 = F2(M->cEZ, cface2, cc, mcf);
        int face = F2(M->CToFace, cface1, cc, mcf);
        double tdl[3], tfl[3], tzl[3], tfz[3], tfe1[3], tfe2[3], A_fep[3];
        for (int d = 1; d <= 3; d++) {
          double p0 = F2(M->px, d, c0 + c, 3), p1 = F2(M->px, d, c0 + cez1, 3), p2 = F2(M->px, d, c0 + cez2, 3);
          double fcd = fc[3 * (face - 1) + d - 1];
          tdl[d - 1] = 0.5 * (p1 - p2);
          tfl[d - 1] = fcd - p0;
          tzl[d - 1] = zc[d - 1] - p0;
          tfz[d - 1] = fcd - zc[d - 1];
          tfe1[d - 1] = fcd - 0.5 * (p1 + p0);
          tfe2[d - 1] = fcd - 0.5 * (p2 + p0);
        }
        A_fep[0] = 0.5 * (tfl[2] * tdl[1] - tfl[1] * tdl[2]);
        A_fep[1] = 0.5 * (tfl[0] * tdl[2] - tfl[2] * tdl[0]);
        A_fep[2] = 0.5 * (tfl[1] * tdl[0] - tfl[0] * tdl[1]);
        int zoneOpp = F2(M->zoneOpp, face, zone, M->maxFaces);
        if (zoneOpp > 0) {
          if (zoneOpp > zone) { /* geometryUCBxyz.F90:114-127 */
            for (int d = 1; d <= 3; d++) F3(A_fp, d, cface1, cc, 3, mcf) = A_fep[d - 1];
            for (int cf = 1; cf <= M->nCFaces[cfp - 1]; cf++)
              if (F2(M->cFP, cf, cfp, mcf) == cc)
                for (int d = 1; d <= 3; d++) F3(A_fp, d, cf, cfp, 3, mcf) = -A_fep[d - 1];
          } else {
            for (int d = 1; d <= 3; d++) A_fep[d - 1] = F3(A_fp, d, cface1, cc, 3, mcf);
          }
        } else if (zoneOpp < 0) {
          if (which == 0) for (int d = 1; d <= 3; d++) F3(A_fp, d, cface1, cc, 3, mcf) = A_fep[d - 1];
          else for (int d = 1; d <= 3; d++) F2(A_bdy, d, cfp - M->ncornr, 3) = A_fep[d - 1]; /* volumeUCBxyz */
        }
        if (which == 0) { /* geometryUCBxyz.F90:141-152 */
          double z1[3], z2[3];
          z1[0] = 0.5 * (tfz[2] * tfe1[1] - tfz[1] * tfe1[2]);
          z1[1] = 0.5 * (tfz[0] * tfe1[2] - tfz[2] * tfe1[0]);
          z1[2] = 0.5 * (tfz[1] * tfe1[0] - tfz[0] * tfe1[1]);
          z2[0] = 0.5 * (tfz[1] * tfe2[2] - tfz[2] * tfe2[1]);
          z2[1] = 0.5 * (tfz[2] * tfe2[0] - tfz[0] * tfe2[2]);
          z2[2] = 0.5 * (tfz[0] * tfe2[1] - tfz[1] * tfe2[0]);
          for (int d = 1; d <= 3; d++) {
            F3(A_ez, d, cface1, cc, 3, mcf) += z1[d - 1];
            F3(A_ez, d, cface2, cc, 3, mcf) += z2[d - 1];
          }
        }
        Volume[cc - 1] += (1.0 / 3.0) * fabs(tzl[0] * A_fep[0] + tzl[1] * A_fep[1] + tzl[2] * A_fep[2]);
      }
      if (which == 1) VolumeZone[zone - 1] += Volume[cc - 1];
    }
    if (which == 0) { /* geometryUCBxyz.F90:167-179 */
      for (int c = 1; c <= nCorner; c++) {
        int cc = c0 + c;
        for (int cface1 = 1; cface1 <= M->nCFaces[cc - 1]; cface1++) {
          int cez = F2(M->cEZ, cface1, cc, mcf);
          if (cez > c)
            for (int cface2 = 1; cface2 <= M->nCFaces[c0 + cez - 1]; cface2++)
              if (F2(M->cEZ, cface2, c0 + cez, mcf) == c)
                for (int d = 1; d <= 3; d++)
                  F3(A_ez, d, cface2, c0 + cez, 3, mcf) = -F3(A_ez, d, cface1, cc, 3, mcf);
        }
      }
    }
  }
  free(Afp_tmp);
}

void orc_geometry_xyz(const orc_mesh *M, double *A_fp, double *A_ez, double *Volume) {
  geom_xyz(M, 0, A_fp, A_ez, Volume, NULL, NULL);
}
void orc_volume_xyz(const orc_mesh *M, double *Volume, double *VolumeZone, double *A_bdy) {
  geom_xyz(M, 1, NULL, NULL, Volume, VolumeZone, A_bdy);
}

void orc_geometry_rz(const orc_mesh *M, double *A_fp, double *A_ez, double *Area, double *Volume,
                     double *RadiusFP, double *RadiusEZ, double *VolumeZone, double *A_bdy, double *RadiusB) {
  /* geometryUCBrz.F90 + volumeUCBrz.F90 (same arithmetic; the boundary parts come from the latter) */
  const int nc = M->ncornr;
  for (int zone = 1; zone <= M->nzones; zone++) {
    double zc[2];
    zone_center(M, zone, zc);
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    VolumeZone[zone - 1] = 0.0;
    for (int c = 1; c <= nCorner; c++) {
      int c1 = F2(M->cEZ, 1, c0 + c, 2), c2 = F2(M->cEZ, 2, c0 + c, 2);
      int cfp1 = F2(M->cFP, 1, c0 + c, 2), cfp2 = F2(M->cFP, 2, c0 + c, 2);
      double r_zone = zc[0], z_zone = zc[1];
      double r_point = F2(M->px, 1, c0 + c, 2), z_point = F2(M->px, 2, c0 + c, 2);
      double r_point1 = F2(M->px, 1, c0 + c1, 2), z_point1 = F2(M->px, 2, c0 + c1, 2);
      double r_point2 = F2(M->px, 1, c0 + c2, 2), z_point2 = F2(M->px, 2, c0 + c2, 2);
      double r_edge1 = 0.5 * (r_point + r_point1), z_edge1 = 0.5 * (z_point + z_point1);
      double r_edge2 = 0.5 * (r_point + r_point2), z_edge2 = 0.5 * (z_point + z_point2);
      F2(RadiusFP, 1, c0 + c, 2) = 0.5 * (r_point + r_edge2);
      F2(RadiusFP, 2, c0 + c, 2) = 0.5 * (r_point + r_edge1);
      F2(RadiusEZ, 1, c0 + c, 2) = 0.5 * (r_zone + r_edge1);
      F2(RadiusEZ, 2, c0 + c, 2) = 0.5 * (r_zone + r_edge2);
      /* note: the reference compares the zone-local c with the global cfp (geometryUCBrz.F90:92,100) */
      if (c < cfp1) {
        F3(A_fp, 1, 1, c0 + c, 2, 2) = 0.5 * (z_point2 - z_point);
        F3(A_fp, 2, 1, c0 + c, 2, 2) = 0.5 * (r_point - r_point2);
        if (cfp1 <= nc) {
          F3(A_fp, 1, 2, cfp1, 2, 2) = -F3(A_fp, 1, 1, c0 + c, 2, 2);
          F3(A_fp, 2, 2, cfp1, 2, 2) = -F3(A_fp, 2, 1, c0 + c, 2, 2);
        }
      }
      if (c < cfp2) {
        F3(A_fp, 1, 2, c0 + c, 2, 2) = 0.5 * (z_point - z_point1);
        F3(A_fp, 2, 2, c0 + c, 2, 2) = 0.5 * (r_point1 - r_point);
        if (cfp2 <= nc) {
          F3(A_fp, 1, 1, cfp2, 2, 2) = -F3(A_fp, 1, 2, c0 + c, 2, 2);
          F3(A_fp, 2, 1, cfp2, 2, 2) = -F3(A_fp, 2, 2, c0 + c, 2, 2);
        }
      }
      if (c < c1) {
        F3(A_ez, 1, 1, c0 + c, 2, 2) = z_edge1 - z_zone;
        F3(A_ez, 2, 1, c0 + c, 2, 2) = r_zone - r_edge1;
        F3(A_ez, 1, 2, c0 + c1, 2, 2) = -F3(A_ez, 1, 1, c0 + c, 2, 2);
        F3(A_ez, 2, 2, c0 + c1, 2, 2) = -F3(A_ez, 2, 1, c0 + c, 2, 2);
      }
      if (c < c2) {
        F3(A_ez, 1, 2, c0 + c, 2, 2) = z_zone - z_edge2;
        F3(A_ez, 2, 2, c0 + c, 2, 2) = r_edge2 - r_zone;
        F3(A_ez, 1, 1, c0 + c2, 2, 2) = -F3(A_ez, 1, 2, c0 + c, 2, 2);
        F3(A_ez, 2, 1, c0 + c2, 2, 2) = -F3(A_ez, 2, 2, c0 + c, 2, 2);
      }
      if (cfp1 > nc) { /* volumeUCBrz.F90 */
        int b = cfp1 - nc;
        F2(A_bdy, 1, b, 2) = 0.5 * (z_point2 - z_point);
        F2(A_bdy, 2, b, 2) = 0.5 * (r_point - r_point2);
        RadiusB[b - 1] = 0.5 * (r_point + r_edge2);
      }
      if (cfp2 > nc) {
        int b = cfp2 - nc;
        F2(A_bdy, 1, b, 2) = 0.5 * (z_point - z_point1);
        F2(A_bdy, 2, b, 2) = 0.5 * (r_point1 - r_point);
        RadiusB[b - 1] = 0.5 * (r_point + r_edge1);
      }
      double area1 = fabs((r_edge2 - r_point) * (z_zone - z_point) - (z_edge2 - z_point) * (r_zone - r_point));
      double area2 = fabs((r_zone - r_point) * (z_edge1 - z_point) - (z_zone - z_point) * (r_edge1 - r_point));
      Area[c0 + c - 1] = 0.5 * (area1 + area2);
      double rbar1 = (1.0 / 3.0) * (r_point + r_edge2 + r_zone);
      double rbar2 = (1.0 / 3.0) * (r_point + r_edge1 + r_zone);
      Volume[c0 + c - 1] = 0.5 * (rbar1 * area1 + rbar2 * area2);
      VolumeZone[zone - 1] += Volume[c0 + c - 1];
    }
  }
}

/* ------------------------------------------------------------------ */
/* Sweep ordering: snac/snnext.F90, snneed.F90, findseeds.F90,         */
/* getDownStreamData.F90, fixZone.F90, cyclebreaker.F90, sccsearch.F90 */
/* ------------------------------------------------------------------ */
static double dotn(const double *a, const double *b, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

typedef struct {
  const orc_mesh *M;
  int *needZ, *listZone, *cycleList;
  unsigned char *exitFace /* (maxFaces,nz) */, *onCycleList, *badZone, *doneZ;
  int meshCycles;
} sched_t;

static void add_zone_to_cycle_list(sched_t *S, int zone) {
  const orc_mesh *M = S->M;
  for (int c = 1; c <= M->numCorner[zone - 1]; c++) {
    if (S->meshCycles >= M->ncornr) die("MeshCycles exceeds the number of corners in SNNEXT!");
    S->cycleList[S->meshCycles++] = M->cOffSet[zone - 1] + c;
  }
}

static void snneed(sched_t *S, const double *omega, const double *A_fp) {
  const orc_mesh *M = S->M;
  const int nd = M->ndim, mcf = M->maxcf, mf = M->maxFaces;
  S->meshCycles = 0;
  memset(S->needZ, 0, sizeof(int) * M->nzones);
  memset(S->exitFace, 0, (size_t)mf * M->nzones);
  if (nd == 2) { /* snneed.F90:68-95 */
    for (int zone = 1; zone <= M->nzones; zone++) {
      int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      for (int c = 1; c <= nCorner; c++) {
        int face = F2(M->CToFace, 1, c0 + c, mcf);
        int zoneOpp = F2(M->zoneOpp, face, zone, mf);
        if (zone < zoneOpp) {
          int faceOpp = F2(M->faceOpp, face, zone, mf);
          double afpm = dotn(omega, &F3(A_fp, 1, 1, c0 + c, nd, mcf), nd);
          if (afpm < 0.0) { S->needZ[zone - 1]++; F2(S->exitFace, faceOpp, zoneOpp, mf) = 1; }
          else if (afpm > 0.0) { S->needZ[zoneOpp - 1]++; F2(S->exitFace, face, zone, mf) = 1; }
        }
      }
    }
    return;
  }
  double afpm_Face[64]; int nInc[64], nExit[64];
  for (int zone = 1; zone <= M->nzones; zone++) { /* snneed.F90:97-191 */
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1], nFaces = M->zoneFaces[zone - 1];
    for (int f = 0; f < 64; f++) { afpm_Face[f] = 0.0; nInc[f] = 0; nExit[f] = 0; }
    for (int c = 1; c <= nCorner; c++) {
      int cc = c0 + c;
      for (int cface = 1; cface <= M->nCFaces[cc - 1]; cface++) {
        int face = F2(M->CToFace, cface, cc, mcf);
        int zoneOpp = F2(M->zoneOpp, face, zone, mf);
        if (zoneOpp > zone) {
          double afpm = dotn(omega, &F3(A_fp, 1, cface, cc, 3, mcf), 3);
          afpm_Face[face - 1] += afpm;
          if (afpm < 0.0) nInc[face - 1]++;
          else if (afpm > 0.0) nExit[face - 1]++;
        }
      }
    }
    for (int face = 1; face <= nFaces; face++) {
      int zoneOpp = F2(M->zoneOpp, face, zone, mf);
      if (zoneOpp > zone) {
        int faceOpp = F2(M->faceOpp, face, zone, mf);
        if (afpm_Face[face - 1] < 0.0) {
          S->needZ[zone - 1]++;
          F2(S->exitFace, faceOpp, zoneOpp, mf) = 1;
          if (nExit[face - 1] > 0 && !S->onCycleList[zone - 1]) {
            add_zone_to_cycle_list(S, zone);
            S->onCycleList[zone - 1] = 1;
          }
        } else if (afpm_Face[face - 1] > 0.0) {
          S->needZ[zoneOpp - 1]++;
          F2(S->exitFace, face, zone, mf) = 1;
          if (nInc[face - 1] > 0 && !S->onCycleList[zoneOpp - 1]) {
            add_zone_to_cycle_list(S, zoneOpp);
            S->onCycleList[zoneOpp - 1] = 1;
          }
        }
      }
    }
  }
}

static int findseeds(sched_t *S) {
  const orc_mesh *M = S->M;
  const int mf = M->maxFaces;
  int nseed = 0;
  for (int zone = 1; zone <= M->nzones; zone++)
    if (S->needZ[zone - 1] == 0) S->listZone[nseed++] = zone;
  if (nseed == 0) { /* findseeds.F90:72-104 */
    int minNeed = M->nzones, zoneID = 0;
    for (int zone = 1; zone <= M->nzones; zone++)
      if (M->BoundaryZone[zone - 1] && S->needZ[zone - 1] < minNeed) { zoneID = zone; minNeed = S->needZ[zone - 1]; }
    if (zoneID == 0) die("No seeds found in FINDSEEDS!");
    nseed = 1;
    S->listZone[0] = zoneID;
    S->needZ[zoneID - 1] = 0;
    for (int face = 1; face <= M->zoneFaces[zoneID - 1]; face++) {
      if (!F2(S->exitFace, face, zoneID, mf)) {
        int zoneOpp = F2(M->zoneOpp, face, zoneID, mf), faceOpp = F2(M->faceOpp, face, zoneID, mf);
        if (zoneOpp > 0) {
          add_zone_to_cycle_list(S, zoneOpp);
          F2(S->exitFace, faceOpp, zoneOpp, mf) = 0;
          S->onCycleList[zoneOpp - 1] = 1;
        }
      }
    }
  }
  return nseed;
}

static void getDownStreamData(sched_t *S, const double *omega, const double *A_ez, int *nextC /* (nc) */) {
  const orc_mesh *M = S->M;
  const int nd = M->ndim, mcf = M->maxcf;
  int need[64], nDSC[64], DownStreamC[64][8];
  for (int zone = 1; zone <= M->nzones; zone++) {
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    for (int c = 0; c < M->maxCorner; c++) { nDSC[c] = 0; need[c] = 0; }
    for (int c = 1; c <= nCorner; c++) {
      int cc = c0 + c;
      int nCFaces = nd == 2 ? 2 : M->nCFaces[cc - 1];
      for (int cface = 1; cface <= nCFaces; cface++) {
        int cez = F2(M->cEZ, cface, cc, mcf);
        if (cez > c) {
          double aez = dotn(omega, &F3(A_ez, 1, cface, cc, nd, mcf), nd);
          if (aez < 0.0) { need[c - 1]++; DownStreamC[cez - 1][nDSC[cez - 1]++] = c; }
          else if (aez > 0.0) { need[cez - 1]++; DownStreamC[c - 1][nDSC[c - 1]++] = cez; }
        }
      }
    }
    S->badZone[zone - 1] = 0;
    for (int i = 1; i <= nCorner; i++) {
      int c = 1, minNeed = need[0];
      for (int k = 2; k <= nCorner; k++) if (need[k - 1] < minNeed) { minNeed = need[k - 1]; c = k; } /* minloc: first minimum */
      nextC[c0 + i - 1] = c;
      if (minNeed != 0) S->badZone[zone - 1] = 1;
      for (int k = 0; k < nDSC[c - 1]; k++) need[DownStreamC[c - 1][k] - 1]--;
      need[c - 1] = 99;
    }
    if (S->badZone[zone - 1]) {
      for (int i = 1; i <= nCorner; i++) nextC[c0 + i - 1] = i;
      add_zone_to_cycle_list(S, zone); /* fixZone.F90 */
    }
  }
}

typedef struct { int zone, face, child; } scc_frame;

static void sccsearch(sched_t *S, int zone0, int *ncount, int *stackindex, int *nBreaks,
                      int *dfnum, int *lowlink, int *stack, unsigned char *isnew, unsigned char *onstack,
                      int *tempList, int *zoneBreakList, scc_frame *frames) {
  /* sccsearch.F90, recursion unrolled onto an explicit frame stack; every test is
     evaluated at the same moment the recursive original evaluates it. */
  const orc_mesh *M = S->M;
  const int mf = M->maxFaces;
  int nfr = 0;
#define ENTER(z) do { (*ncount)++; dfnum[(z) - 1] = *ncount; lowlink[(z) - 1] = *ncount; isnew[(z) - 1] = 0; \
    stack[(*stackindex)++] = (z); onstack[(z) - 1] = 1; frames[nfr].zone = (z); frames[nfr].face = 1; frames[nfr].child = 0; nfr++; } while (0)
  ENTER(zone0);
  while (nfr > 0) {
    scc_frame *fr = &frames[nfr - 1];
    int zone = fr->zone, descended = 0;
    if (fr->child) { /* return from recursion: sccsearch.F90:103-105 */
      if (lowlink[fr->child - 1] < lowlink[zone - 1]) lowlink[zone - 1] = lowlink[fr->child - 1];
      fr->child = 0;
    }
    int nFaces = M->zoneFaces[zone - 1];
    while (fr->face <= nFaces) {
      int face = fr->face++;
      if (F2(S->exitFace, face, zone, mf)) {
        int zone2 = F2(M->zoneOpp, face, zone, mf);
        if (zone2 > 0) {
          if (isnew[zone2 - 1]) { fr->child = zone2; ENTER(zone2); descended = 1; break; }
          else if (dfnum[zone2 - 1] < dfnum[zone - 1] && onstack[zone2 - 1] && lowlink[zone2 - 1] < lowlink[zone - 1])
            lowlink[zone - 1] = lowlink[zone2 - 1];
        }
      }
    }
    if (descended) continue;
    if (lowlink[zone - 1] == dfnum[zone - 1]) { /* sccsearch.F90:122-184 */
      int zone2 = stack[--(*stackindex)];
      onstack[zone2 - 1] = 0;
      if (zone2 != zone) {
        int cyclesize = 0;
        while (zone2 != zone) { tempList[cyclesize++] = zone2; zone2 = stack[--(*stackindex)]; }
        tempList[cyclesize++] = zone2;
        onstack[tempList[0] - 1] = 1;
        int lowlinkZ = tempList[cyclesize - 1];
        for (int face = 1; face <= M->zoneFaces[lowlinkZ - 1]; face++) {
          int zoneBreak = F2(M->zoneOpp, face, lowlinkZ, mf), faceBreak = F2(M->faceOpp, face, lowlinkZ, mf);
          if (zoneBreak > 0 && onstack[zoneBreak - 1] && F2(S->exitFace, faceBreak, zoneBreak, mf)) {
            if (!S->onCycleList[zoneBreak - 1]) { add_zone_to_cycle_list(S, zoneBreak); S->onCycleList[zoneBreak - 1] = 1; }
            S->needZ[lowlinkZ - 1]--;
            F2(S->exitFace, faceBreak, zoneBreak, mf) = 0;
            if (S->needZ[lowlinkZ - 1] == 0) zoneBreakList[(*nBreaks)++] = lowlinkZ;
          }
        }
        for (int i = 0; i < cyclesize; i++) onstack[tempList[i] - 1] = 0;
      }
    }
    nfr--;
  }
#undef ENTER
}

static void cyclebreaker(sched_t *S, int ndoneZ, int *nextZone, int *addedZones) {
  const orc_mesh *M = S->M;
  int nzones = M->nzones, ngraph = nzones - ndoneZ;
  int *listZ = malloc(sizeof(int) * (ngraph + 1)), *zoneBreakList = malloc(sizeof(int) * (ngraph + 1));
  int *dfnum = calloc(nzones, sizeof(int)), *lowlink = calloc(nzones, sizeof(int));
  int *stack = calloc(ngraph + 1, sizeof(int)), *tempList = malloc(sizeof(int) * (ngraph + 1));
  unsigned char *isnew = malloc(nzones), *onstack = calloc(nzones, 1);
  scc_frame *frames = malloc(sizeof(scc_frame) * (ngraph + 1));
  memset(isnew, 1, nzones);
  int nBreaks = 0, ncount = 0, stackindex = 0, nleft = 0;
  for (int zone = 1; zone <= nzones; zone++) {
    if (S->needZ[zone - 1] == 0) isnew[zone - 1] = 0;
    else listZ[nleft++] = zone;
  }
  if (nleft != ngraph) die("Miscount of remaining zones in CYCLEBREAKER");
  for (int i = 0; i < ngraph; i++) {
    int zone = listZ[i];
    if (isnew[zone - 1])
      sccsearch(S, zone, &ncount, &stackindex, &nBreaks, dfnum, lowlink, stack, isnew, onstack, tempList, zoneBreakList, frames);
  }
  if (nBreaks == 0) die("CYCLEBREAKER: detection failed, no dependencies broken");
  *addedZones = 0;
  for (int i = 0; i < nBreaks; i++) {
    int zone = zoneBreakList[i];
    if (S->needZ[zone - 1] == 0) { S->listZone[(*nextZone)++] = zone; (*addedZones)++; }
    else if (S->needZ[zone - 1] < 0) die("CycleBreaker, needZ < 0");
  }
  if (*addedZones == 0) die("Cycles found, but not broken");
  free(listZ); free(zoneBreakList); free(dfnum); free(lowlink); free(stack); free(tempList);
  free(isnew); free(onstack); free(frames);
}

/* One angle.  Outputs: nextZ(nz) signed, nextC(nc), zonesInPlane(<=nz), cycleList(<=nc).
   Returns nHyperPlanes; *numCycles = meshCycles.  (snnext.F90:12-233) */
int orc_snnext(const orc_mesh *M, const double *A_fp, const double *A_ez, const double *omega,
               int *nextZ, int *nextC, int *zonesInPlane, int *numCycles, int *cycleList) {
  const int nzones = M->nzones, mf = M->maxFaces;
  sched_t S;
  S.M = M;
  S.needZ = malloc(sizeof(int) * nzones);
  S.listZone = malloc(sizeof(int) * nzones);
  S.cycleList = cycleList;
  S.exitFace = malloc((size_t)mf * nzones);
  S.onCycleList = calloc(nzones, 1);
  S.badZone = calloc(nzones, 1);
  S.doneZ = calloc(nzones, 1);
  S.meshCycles = 0;
  snneed(&S, omega, A_fp);
  int newZones = findseeds(&S);
  getDownStreamData(&S, omega, A_ez, nextC);
  int ndoneZ = 0, nextZone = 0, lastZone = 0, nHyperPlanes = 0;
  for (;;) {
    nHyperPlanes++;
    zonesInPlane[nHyperPlanes - 1] = newZones;
    nextZone = lastZone + newZones;
    int addedZones = 0;
    for (int zID = 1; zID <= newZones; zID++) {
      int zone = S.listZone[lastZone + zID - 1];
      ndoneZ++;
      S.doneZ[zone - 1] = 1;
      for (int face = 1; face <= M->zoneFaces[zone - 1]; face++) {
        if (F2(S.exitFace, face, zone, mf)) {
          int Zexit = F2(M->zoneOpp, face, zone, mf);
          if (Zexit > 0 && !S.doneZ[Zexit - 1]) {
            S.needZ[Zexit - 1]--;
            if (S.needZ[Zexit - 1] == 0) { S.listZone[nextZone++] = Zexit; addedZones++; }
            else if (S.needZ[Zexit - 1] < 0) die("needZ < 0 in SNNEXT!");
          }
        }
      }
      nextZ[ndoneZ - 1] = S.badZone[zone - 1] ? -zone : zone;
    }
    lastZone += newZones;
    if (lastZone == nzones) break;
    if (addedZones > 0) newZones = addedZones;
    else { cyclebreaker(&S, ndoneZ, &nextZone, &addedZones); newZones = addedZones; }
  }
  *numCycles = S.meshCycles;
  if (ndoneZ != nzones) die("Wrong number of zones in SNNEXT!");
  free(S.needZ); free(S.listZone); free(S.exitFace); free(S.onCycleList); free(S.badZone); free(S.doneZ);
  return nHyperPlanes;
}

/* findexit.F90:296-349, non-shared boundaries + shared send lists in boundary order.
   isExitShared[b] (may be NULL): for shared elements, 1 if the element is on this angle's ListSend. */
int orc_bdy_exit(int ndim, int nbelem, const double *A_bdy, const int *BdyToC, const unsigned char *isShared,
                 const unsigned char *isExitShared, const double *omega, int *bdyList /* (2,nxBdy) */) {
  int nx = 0;
  for (int pass = 0; pass < 2; pass++)
    for (int b = 1; b <= nbelem; b++) {
      int sh = isShared ? isShared[b - 1] : 0;
      if (sh != pass) continue;
      int ex = sh ? (isExitShared && isExitShared[b - 1]) : (dotn(omega, &F2(A_bdy, 1, b, ndim), ndim) > 0.0);
      if (ex) { bdyList[2 * nx] = b; bdyList[2 * nx + 1] = BdyToC[b - 1]; nx++; }
    }
  return nx;
}

/* ------------------------------------------------------------------ */
/* snac/SweepUCBxyz.F90:11-322                                         */
/* ------------------------------------------------------------------ */
void orc_sweep_xyz(const orc_mesh *M, int Groups, int nHyperPlanes, const int *zonesInPlane,
                   const int *nextZ, const int *nextC, const double *omega, double quadwt, double tau,
                   const double *STotal /* (G,nc) */, const double *Sigt /* (G,nz) */, const double *Volume,
                   const double *A_fp, const double *A_ez,
                   double *PsiA /* (G,nc) slice of Psi for this angle */, double *Psi1 /* (G,nc+nb) */,
                   double *PsiBA /* (G,nb) slice of PsiB for this angle */, double *Phi /* (G,nc) */, int savePsi) {
  const int G = Groups, nc = M->ncornr, nb = M->nbelem, mcf = M->maxcf, mC = M->maxCorner;
  const double fouralpha = 1.82;
  int *nxez = malloc(sizeof(int) * mC), *ez_exit = malloc(sizeof(int) * mcf * mC), *bdy_exit = malloc(sizeof(int) * 2 * mcf * mC);
  double *sumArea = malloc(sizeof(double) * mC), *psi_opp = calloc(G, sizeof(double));
  double *SigtVol = malloc(sizeof(double) * G * mC), *src = malloc(sizeof(double) * G * mC), *Q = malloc(sizeof(double) * G * mC);
  double *afp = malloc(sizeof(double) * mcf), *coefpsi = malloc(sizeof(double) * mcf * mC), *psifp = calloc((size_t)G * mcf, sizeof(double));

  for (int ib = 1; ib <= nb; ib++) /* :95-97 */
    memcpy(&F2(Psi1, 1, nc + ib, G), &F2(PsiBA, 1, ib, G), sizeof(double) * G);

  int ndoneZ = 0;
  for (int hyperPlane = 1; hyperPlane <= nHyperPlanes; hyperPlane++) {
    int nzones = zonesInPlane[hyperPlane - 1];
    for (int ii = 1; ii <= nzones; ii++) {
      int zone0 = nextZ[ndoneZ + ii - 1];
      int zone = abs(zone0), nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      int nxBdy = 0;
      for (int c = 1; c <= nCorner; c++)
        for (int g = 1; g <= G; g++) {
          double source = F2(STotal, g, c0 + c, G) + tau * F2(PsiA, g, c0 + c, G);
          F2(Q, g, c, G) = source;
          F2(src, g, c, G) = Volume[c0 + c - 1] * source;
          F2(SigtVol, g, c, G) = F2(Sigt, g, zone, G) * Volume[c0 + c - 1];
        }
      for (int c = 0; c < mC; c++) nxez[c] = 0;
      for (int c = 1; c <= nCorner; c++) {
        sumArea[c - 1] = 0.0;
        int nCFaces = M->nCFaces[c0 + c - 1];
        for (int cface = 1; cface <= nCFaces; cface++) {
          afp[cface - 1] = dotn(omega, &F3(A_fp, 1, cface, c0 + c, 3, mcf), 3);
          int cfp = F2(M->cFP, cface, c0 + c, mcf);
          if (afp[cface - 1] > 0.0) {
            sumArea[c - 1] += afp[cface - 1];
            if (cfp > nc) { bdy_exit[2 * nxBdy] = c; bdy_exit[2 * nxBdy + 1] = cfp - nc; nxBdy++; }
          } else if (afp[cface - 1] < 0.0) {
            for (int g = 1; g <= G; g++) {
              F2(psifp, g, cface, G) = F2(Psi1, g, cfp, G);
              F2(src, g, c, G) = F2(src, g, c, G) - afp[cface - 1] * F2(Psi1, g, cfp, G);
            }
          }
        }
        for (int cface = 1; cface <= nCFaces; cface++) {
          double aez = dotn(omega, &F3(A_ez, 1, cface, c0 + c, 3, mcf), 3);
          int cez = F2(M->cEZ, cface, c0 + c, mcf);
          if (cez > c) {
            if (aez > 0.0) {
              nxez[c - 1]++;
              F2(ez_exit, nxez[c - 1], c, mcf) = cez;
              F2(coefpsi, nxez[c - 1], c, mcf) = aez;
            } else if (aez < 0.0) {
              nxez[cez - 1]++;
              F2(ez_exit, nxez[cez - 1], cez, mcf) = c;
              F2(coefpsi, nxez[cez - 1], cez, mcf) = -aez;
            }
          }
          if (aez > 0.0) {
            sumArea[c - 1] += aez;
            double area_opp = 0.0;
            if (nCFaces == 3) {
              int ifp = cface % nCFaces + 1;
              if (afp[ifp - 1] < 0.0) {
                for (int g = 1; g <= G; g++) psi_opp[g - 1] = F2(psifp, g, ifp, G);
                area_opp = -afp[ifp - 1];
              }
            } else {
              int ifp = cface;
              area_opp = 0.0;
              for (int g = 0; g < G; g++) psi_opp[g] = 0.0;
              for (int k = 1; k <= nCFaces - 2; k++) {
                ifp = ifp % nCFaces + 1;
                if (afp[ifp - 1] < 0.0) {
                  area_opp = area_opp - afp[ifp - 1];
                  for (int g = 1; g <= G; g++) psi_opp[g - 1] = psi_opp[g - 1] - afp[ifp - 1] * F2(psifp, g, ifp, G);
                }
              }
              if (area_opp > 0.0) {
                double area_inv = 1.0 / area_opp;
                for (int g = 0; g < G; g++) psi_opp[g] = psi_opp[g] * area_inv;
              }
            }
            if (area_opp > 0.0) {
              double aez2 = aez * aez, vol = Volume[c0 + c - 1];
              for (int g = 1; g <= G; g++) {
                double sig = F2(Sigt, g, zone, G);
                double sigv = sig * vol, sigv2 = sigv * sigv;
                double gnum = aez2 * (fouralpha * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
                double gden = vol * (4.0 * sigv * sigv2 + aez * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
                double sez = (vol * gnum * (sig * psi_opp[g - 1] - F2(Q, g, c, G)) +
                              0.5 * aez * gden * (F2(Q, g, c, G) - F2(Q, g, cez, G))) / (gnum + gden * sig);
                F2(src, g, c, G) = F2(src, g, c, G) + sez;
                F2(src, g, cez, G) = F2(src, g, cez, G) - sez;
              }
            } else {
              for (int g = 1; g <= G; g++) {
                double sigInv = 1.0 / F2(Sigt, g, zone, G);
                double sez = 0.5 * aez * sigInv * (F2(Q, g, c, G) - F2(Q, g, cez, G));
                F2(src, g, c, G) = F2(src, g, c, G) + sez;
                F2(src, g, cez, G) = F2(src, g, cez, G) - sez;
              }
            }
          }
        }
      }
      if (zone0 > 0) { /* :261-281 */
        for (int i = 1; i <= nCorner; i++) {
          int c = nextC[c0 + i - 1];
          for (int g = 1; g <= G; g++) {
            F2(Psi1, g, c0 + c, G) = F2(src, g, c, G) / (sumArea[c - 1] + F2(SigtVol, g, c, G));
            F2(Phi, g, c0 + c, G) = F2(Phi, g, c0 + c, G) + quadwt * F2(Psi1, g, c0 + c, G);
          }
          for (int cface = 1; cface <= nxez[c - 1]; cface++) {
            int cez = F2(ez_exit, cface, c, mcf);
            double cf = F2(coefpsi, cface, c, mcf);
            for (int g = 1; g <= G; g++) F2(src, g, cez, G) = F2(src, g, cez, G) + cf * F2(Psi1, g, c0 + c, G);
          }
        }
      } else { /* :283-298 */
        for (int c = 1; c <= nCorner; c++)
          for (int cface = 1; cface <= nxez[c - 1]; cface++) {
            int cez = F2(ez_exit, cface, c, mcf);
            double cf = F2(coefpsi, cface, c, mcf);
            for (int g = 1; g <= G; g++) F2(src, g, cez, G) = F2(src, g, cez, G) + cf * F2(Psi1, g, c0 + c, G);
          }
        for (int c = 1; c <= nCorner; c++)
          for (int g = 1; g <= G; g++) {
            F2(Psi1, g, c0 + c, G) = F2(src, g, c, G) / (sumArea[c - 1] + F2(SigtVol, g, c, G));
            F2(Phi, g, c0 + c, G) = F2(Phi, g, c0 + c, G) + quadwt * F2(Psi1, g, c0 + c, G);
          }
      }
      for (int ib = 0; ib < nxBdy; ib++) { /* :302-306 */
        int c = bdy_exit[2 * ib], b = bdy_exit[2 * ib + 1];
        memcpy(&F2(PsiBA, 1, b, G), &F2(Psi1, 1, c0 + c, G), sizeof(double) * G);
      }
    }
    ndoneZ += nzones;
  }
  if (savePsi) memcpy(PsiA, Psi1, sizeof(double) * (size_t)G * nc); /* :314-318 */
  free(nxez); free(ez_exit); free(bdy_exit); free(sumArea); free(psi_opp); free(SigtVol); free(src); free(Q);
  free(afp); free(coefpsi); free(psifp);
}

/* ------------------------------------------------------------------ */
/* snac/SweepUCBrz.F90:11-282                                          */
/* ------------------------------------------------------------------ */
void orc_sweep_rz(const orc_mesh *M, int Groups, int nHyperPlanes, const int *zonesInPlane,
                  const int *nextZ, const int *nextC, const double *omega, double quadwt, double tau,
                  double fac, double quadTauW1, double quadTauW2, int StartingDirection, int setFinishingDirection,
                  const double *STotal, const double *Sigt, const double *Volume, const double *Area,
                  const double *A_fp, const double *A_ez, const double *RadiusFP, const double *RadiusEZ,
                  int nxBdy, const int *bdyList /* (2,nxBdy) */,
                  double *PsiA, double *PsiA1 /* Psi(:,:,Angle+1) or NULL */, double *Psi1, double *PsiM,
                  double *PsiBA, double *PsiBA1 /* PsiB(:,:,Angle+1) or NULL */, double *Phi, int savePsi) {
  const int G = Groups, nc = M->ncornr, nb = M->nbelem, mC = M->maxCorner;
  const double fouralpha = 1.82;
  int *nxez = malloc(sizeof(int) * mC), *ez_exit = malloc(sizeof(int) * 2 * mC);
  double *sumArea = malloc(sizeof(double) * mC), *coefpsi = malloc(sizeof(double) * 2 * mC);
  double *Q = malloc(sizeof(double) * G * mC), *src = malloc(sizeof(double) * G * mC);
  for (int ib = 1; ib <= nb; ib++) memcpy(&F2(Psi1, 1, nc + ib, G), &F2(PsiBA, 1, ib, G), sizeof(double) * G);
  int ndoneZ = 0, cfp = -1;
  for (int hyperPlane = 1; hyperPlane <= nHyperPlanes; hyperPlane++) {
    int nzones = zonesInPlane[hyperPlane - 1];
    for (int ii = 1; ii <= nzones; ii++) {
      int zone0 = nextZ[ndoneZ + ii - 1];
      int zone = abs(zone0), nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      for (int c = 0; c < mC; c++) nxez[c] = 0;
      for (int c = 1; c <= nCorner; c++) {
        for (int g = 1; g <= G; g++) {
          double source = F2(STotal, g, c0 + c, G) + tau * F2(PsiA, g, c0 + c, G);
          F2(Q, g, c, G) = source;
          F2(src, g, c, G) = Volume[c0 + c - 1] * source;
        }
        sumArea[c - 1] = fac * Area[c0 + c - 1];
      }
      for (int c = 1; c <= nCorner; c++) {
        for (int cface = 1; cface <= 2; cface++) {
          double afp = dotn(omega, &F3(A_fp, 1, cface, c0 + c, 2, 2), 2);
          double aez = dotn(omega, &F3(A_ez, 1, cface, c0 + c, 2, 2), 2);
          if (afp < 0.0) {
            cfp = F2(M->cFP, cface, c0 + c, 2);
            double R_afp = F2(RadiusFP, cface, c0 + c, 2) * afp;
            sumArea[c - 1] = sumArea[c - 1] - R_afp;
            for (int g = 1; g <= G; g++) F2(src, g, c, G) = F2(src, g, c, G) - R_afp * F2(Psi1, g, cfp, G);
          }
          if (aez > 0.0) {
            double R = F2(RadiusEZ, cface, c0 + c, 2);
            int cez = F2(M->cEZ, cface, c0 + c, 2);
            double area = Area[c0 + c - 1];
            nxez[c - 1]++;
            F2(ez_exit, nxez[c - 1], c, 2) = cez;
            F2(coefpsi, nxez[c - 1], c, 2) = R * aez;
            sumArea[cez - 1] = sumArea[cez - 1] + R * aez;
            if (afp < 0.0) {
              for (int g = 1; g <= G; g++) {
                double sig = F2(Sigt, g, zone, G);
                double sigA = sig * area, sigA2 = sigA * sigA;
                double gnum = aez * aez * (fouralpha * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
                double gden = area * (4.0 * sigA * sigA2 + aez * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
                double sez = R * (area * gnum * (sig * F2(Psi1, g, cfp, G) - F2(Q, g, c, G)) +
                                  0.5 * aez * gden * (F2(Q, g, c, G) - F2(Q, g, cez, G))) / (gnum + gden * sig);
                F2(src, g, c, G) = F2(src, g, c, G) + sez;
                F2(src, g, cez, G) = F2(src, g, cez, G) - sez;
              }
            } else {
              for (int g = 1; g <= G; g++) {
                double sez = 0.5 * R * aez * (F2(Q, g, c, G) - F2(Q, g, cez, G)) / F2(Sigt, g, zone, G);
                F2(src, g, c, G) = F2(src, g, c, G) + sez;
                F2(src, g, cez, G) = F2(src, g, cez, G) - sez;
              }
            }
          }
        }
      }
      for (int i = 1; i <= nCorner; i++) {
        int c = nextC[c0 + i - 1];
        for (int g = 1; g <= G; g++)
          F2(Psi1, g, c0 + c, G) = (F2(src, g, c, G) + Area[c0 + c - 1] * fac * F2(PsiM, g, c0 + c, G)) /
                                   (sumArea[c - 1] + F2(Sigt, g, zone, G) * Volume[c0 + c - 1]);
        for (int cface = 1; cface <= nxez[c - 1]; cface++) {
          int cez = F2(ez_exit, cface, c, 2);
          double cf = F2(coefpsi, cface, c, 2);
          for (int g = 1; g <= G; g++) F2(src, g, cez, G) = F2(src, g, cez, G) + cf * F2(Psi1, g, c0 + c, G);
        }
      }
      if (StartingDirection) {
        for (int c = 1; c <= nCorner; c++)
          memcpy(&F2(PsiM, 1, c0 + c, G), &F2(Psi1, 1, c0 + c, G), sizeof(double) * G);
      } else {
        for (int c = 1; c <= nCorner; c++)
          for (int g = 1; g <= G; g++) {
            F2(PsiM, g, c0 + c, G) = quadTauW1 * F2(Psi1, g, c0 + c, G) - quadTauW2 * F2(PsiM, g, c0 + c, G);
            F2(Phi, g, c0 + c, G) = F2(Phi, g, c0 + c, G) + quadwt * F2(Psi1, g, c0 + c, G);
          }
      }
    }
    ndoneZ += nzones;
  }
  for (int ib = 0; ib < nxBdy; ib++) {
    int b = bdyList[2 * ib], c = bdyList[2 * ib + 1];
    memcpy(&F2(PsiBA, 1, b, G), &F2(Psi1, 1, c, G), sizeof(double) * G);
  }
  if (setFinishingDirection)
    for (int ib = 0; ib < nxBdy; ib++) {
      int b = bdyList[2 * ib], c = bdyList[2 * ib + 1];
      memcpy(&F2(PsiBA1, 1, b, G), &F2(PsiM, 1, c, G), sizeof(double) * G);
    }
  if (savePsi) {
    memcpy(PsiA, Psi1, sizeof(double) * (size_t)G * nc);
    if (setFinishingDirection) memcpy(PsiA1, PsiM, sizeof(double) * (size_t)G * nc);
  }
  free(nxez); free(ez_exit); free(sumArea); free(coefpsi); free(Q); free(src);
}

/* ------------------------------------------------------------------ */
/* One flux pass of snac/SetSweep.F90:113-170 for a single-domain 3-D  */
/* problem with one angle per phase-space set, threaded over sets like */
/* the reference (omp parallel do schedule(dynamic)), followed by the  */
/* CPU branch of control/getPhiTotal_OMPOL.F90:134-160.                */
/* schedule arrays are per angle, concatenated with stride nz / nc.    */
/* cycle lists: initFromCycleList / updateCycleList                    */
/* (control/constructDynMemory.F90:111-213).                            */
/* ------------------------------------------------------------------ */
void orc_setsweep_xyz(const orc_mesh *M, int Groups, int NumAngles, const int *nHyperPlanes,
                      const int *zonesInPlane /* (nz,NA) */, const int *nextZ /* (nz,NA) */, const int *nextC /* (nc,NA) */,
                      const int *numCycles, const int *cycleOffSet, const int *cycleList, double *cyclePsi /* (G,totalCycles) */,
                      const double *omega /* (3,NA) */, const double *weight, double tau,
                      const double *STotal, const double *Sigt, const double *Volume,
                      const double *A_fp, const double *A_ez,
                      double *Psi /* (G,nc,NA) */, double *PsiB /* (G,nb,NA) */, double *PhiTotal /* (G,nc) */,
                      int savePsi, int nthreads) {
  const int G = Groups, nc = M->ncornr, nb = M->nbelem, nz = M->nzones;
  const size_t npsi1 = (size_t)G * (nc + nb), nphi = (size_t)G * nc;
  /* Set%Phi per set (= per angle) */
  double *PhiSet = calloc(nphi * NumAngles, sizeof(double));
  if (!PhiSet) die("out of memory for per-set Phi");
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
  {
    double *Psi1 = calloc(npsi1, sizeof(double));
#ifdef _OPENMP
#pragma omp for schedule(dynamic)
#endif
    for (int a = 0; a < NumAngles; a++) {
      for (int m = 0; m < numCycles[a]; m++) { /* initFromCycleList */
        int mC = cycleOffSet[a] + m, c = cycleList[mC];
        memcpy(&F2(Psi1, 1, c, G), &F2(cyclePsi, 1, mC + 1, G), sizeof(double) * G);
      }
      orc_sweep_xyz(M, G, nHyperPlanes[a], zonesInPlane + (size_t)nz * a, nextZ + (size_t)nz * a,
                    nextC + (size_t)nc * a, omega + 3 * a, weight[a], tau, STotal, Sigt, Volume, A_fp, A_ez,
                    Psi + nphi * a, Psi1, PsiB + (size_t)G * nb * a, PhiSet + nphi * a, savePsi);
      for (int m = 0; m < numCycles[a]; m++) { /* updateCycleList */
        int mC = cycleOffSet[a] + m, c = cycleList[mC];
        memcpy(&F2(cyclePsi, 1, mC + 1, G), &F2(Psi1, 1, c, G), sizeof(double) * G);
      }
    }
    free(Psi1);
  }
  /* getPhiTotal: PhiTotal = sum over sets in set order */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < (long)nphi; i++) {
    double s = 0.0;
    for (int a = 0; a < NumAngles; a++) s = s + PhiSet[nphi * a + i];
    PhiTotal[i] = s;
  }
  free(PhiSet);
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
