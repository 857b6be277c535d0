/* umt_oracle_gta.c — TEST INFRASTRUCTURE (see umt_oracle.c header; "parity unpinned": the reference
 * holds no golden vectors for GTA and, in the shipped mini-app build, never executes it,
 * rt/LinearSolver.F90:87-121).
 *
 * CPU restatement of the grey-transport-acceleration pieces of Teton (3-D, "new" GTA solver, the
 * variant the reference's GPU path uses):
 *   rt/setGTAOpacity.F90:10-113 (setGTAOpacityNEW)       -> orc_gta_set_opacity
 *   rt/getCollisionRate.F90:10-97                        -> orc_collision_rate
 *   snac/InitSweepGreyUCBxyz.F90:10-253                  -> orc_gta_init_tt
 *   snac/SweepGreyUCBxyz.F90:12-133,137-355 (KernelNew)  -> orc_gta_sweep_angle
 *   snac/GTASweep.F90:9-168 (GTA%ID == 1, single domain) + rt/GreySweep.F90:12-48 (GreySweepNEW)
 *   + snac/UpdateScalarIntensity.F90:11-170              -> orc_gta_grey_sweep
 *   rt/scat_prod.F90 / scat_prod1.F90, rt/GTASolver.F90:42-425 (BiCGSTAB) -> orc_gta_solver
 *   rt/addGreyCorrections.F90:70-91 (scalar part)        -> orc_add_grey_corrections
 * and their r-z counterparts:
 *   rt/quadrz.F90:82-160 (level-symmetric S2) + rtquad.F90:95-127 + AngleCoef2D.F90 + AngleSet_mod.F90:337-347 -> orc_gta_quad_rz
 *   snac/SweepGreyUCBrz.F90:12-133,137-330 (KernelNew)   -> orc_gta_sweep_angle_rz
 *   snac/InitSweepGreyUCBrz.F90:10-235                   -> orc_gta_init_tt_rz
 * Arrays are Fortran memory images with 1-based ids, as in umt_oracle.c.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define F2(a, i, j, n1) ((a)[((i) - 1) + (size_t)(n1) * ((j) - 1)])
#define F3(a, i, j, k, n1, n2) ((a)[((i) - 1) + (size_t)(n1) * (((j) - 1) + (size_t)(n2) * ((k) - 1))])

typedef struct {
  int ndim, nzones, ncornr, nbelem, maxcf, maxCorner, maxFaces;
  const int *numCorner, *cOffSet, *zoneFaces, *zoneOpp, *faceOpp, *nCFaces, *cFP, *cEZ, *CToFace;
  const unsigned char *BoundaryZone;
  const double *px;
} orc_mesh;

#define MAXC 8
#define MAXCF 3
static const double fouralpha = 1.82;

static double dot3(const double *a, const double *b) { /* DOT_PRODUCT order */
  double s = 0.0;
  for (int d = 0; d < 3; d++) s = s + a[d] * b[d];
  return s;
}

/* level-symmetric S2 set of the GTA sweeps: rt/quadxyz.F90 (order 2: one ordinate per octant,
   dircos 0.577350269189625, weight 1, QuadratureData_mod.F90:711-713,782-784) + rtquad.F90:95-105 */
int orc_gta_quad_xyz(double *omega /* (3,8) */, double *weight /* (8) */) {
  const double pi = 3.14159265358979323846, halfpi = 0.5 * pi;
  const double mu = 0.577350269189625;
  const int sx[8] = {1, -1, -1, 1, 1, -1, -1, 1}, sy[8] = {1, 1, -1, -1, 1, 1, -1, -1}, sz[8] = {1, 1, 1, 1, -1, -1, -1, -1};
  double sum = 0.0;
  for (int a = 0; a < 8; a++) {
    omega[3 * a] = sx[a] * mu; omega[3 * a + 1] = sy[a] * mu; omega[3 * a + 2] = sz[a] * mu;
    weight[a] = halfpi * 1.0;
    sum += weight[a];
  }
  const double wtiso = 1.0 / (4.0 * pi), fac = 1.0 / (wtiso * sum);
  for (int a = 0; a < 8; a++) weight[a] = fac * weight[a];
  return 8;
}

/* setGTAOpacityNEW for every zone; Chi(ngr,nc) is rescaled in place */
void orc_gta_set_opacity(const orc_mesh *M, int ngr, double tau, const double *Siga /* (ngr,nz) */, const double *Sigs,
                         const double *Eta /* (nc) */, double *Chi, const double *Volume, double *GreySigTotal,
                         double *GreySigScat, double *GreySigScatVol, double *GreySigtInv) {
  const double minRatio = 1.0e-10;
  for (int zone = 1; zone <= M->nzones; zone++) {
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    for (int c = 1; c <= nCorner; c++) {
      int cID = c0 + c;
      double SigtInvAve = 0.0, Sigt2InvAve = 0.0, SigaAve = 0.0;
      for (int g = 1; g <= ngr; g++) {
        double SigtInv = 1.0 / (F2(Siga, g, zone, ngr) + F2(Sigs, g, zone, ngr) + tau);
        SigtInvAve = SigtInvAve + F2(Chi, g, cID, ngr) * SigtInv;
        Sigt2InvAve = Sigt2InvAve + F2(Chi, g, cID, ngr) * SigtInv * SigtInv;
        SigaAve = SigaAve + F2(Chi, g, cID, ngr) * F2(Siga, g, zone, ngr) * SigtInv;
        F2(Chi, g, cID, ngr) = F2(Chi, g, cID, ngr) * SigtInv;
      }
      double greysigt, greysiga, greysigs;
      if (SigtInvAve > 0.0) {
        for (int g = 1; g <= ngr; g++) F2(Chi, g, cID, ngr) = F2(Chi, g, cID, ngr) / SigtInvAve;
        greysigt = SigtInvAve / Sigt2InvAve;
        greysiga = tau + (1.0 - Eta[cID - 1]) * SigaAve / SigtInvAve;
        greysigs = greysigt - greysiga;
      } else {
        greysigt = tau; greysiga = tau; greysigs = 0.0;
      }
      double scatRatio = greysigs / greysigt;
      if (scatRatio <= minRatio) { GreySigScat[cID - 1] = 0.0; GreySigTotal[cID - 1] = greysiga; }
      else { GreySigScat[cID - 1] = greysigs; GreySigTotal[cID - 1] = greysigt; }
      GreySigScatVol[cID - 1] = GreySigScat[cID - 1] * Volume[cID - 1];
      GreySigtInv[cID - 1] = 1.0 / GreySigTotal[cID - 1];
    }
  }
}

/* getCollisionRate: GreySource(c) = sum_g (Eta siga + sigs) PhiTotal (residualFlag 1: minus the previous value) */
void orc_collision_rate(const orc_mesh *M, int ngr, const double *Eta, const double *Siga, const double *Sigs,
                        const double *PhiTotal /* (ngr,nc) */, double *GreySource, int residualFlag) {
  for (int zone = 1; zone <= M->nzones; zone++)
    for (int c = M->cOffSet[zone - 1] + 1; c <= M->cOffSet[zone - 1] + M->numCorner[zone - 1]; c++) {
      double s = 0.0;
      for (int g = 1; g <= ngr; g++)
        s = s + (Eta[c - 1] * F2(Siga, g, zone, ngr) + F2(Sigs, g, zone, ngr)) * F2(PhiTotal, g, c, ngr);
      GreySource[c - 1] = residualFlag == 0 ? s : s - GreySource[c - 1];
    }
}

/* InitGreySweepUCBxyz for every zone: TT(maxCorner, nc) */
void orc_gta_init_tt(const orc_mesh *M, int nAng, const double *omegas, const double *weights, const double *Volume,
                     const double *A_fp, const double *A_ez, const double *GreySigTotal, double *TT) {
  const int mC = M->maxCorner;
  for (int zone = 1; zone <= M->nzones; zone++) {
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    for (int c = 1; c <= nCorner; c++)
      for (int k = 1; k <= mC; k++) F2(TT, k, c0 + c, mC) = 0.0;
    for (int Angle = 0; Angle < nAng; Angle++) {
      const double *omega = omegas + 3 * Angle;
      double quadwt = weights[Angle];
      int nxez[MAXC + 1] = {0}, need[MAXC + 1] = {0}, ez_exit[MAXCF + 1][MAXC + 1];
      double denom[MAXC + 1], afp[MAXCF + 1], coefpsi[MAXCF + 1][MAXC + 1], Sigt[MAXC + 1], Pvv[MAXC + 1][MAXC + 1];
      for (int i = 0; i <= MAXC; i++) for (int j = 0; j <= MAXC; j++) Pvv[i][j] = 0.0;
      for (int c = 1; c <= nCorner; c++) { Pvv[c][c] = Volume[c0 + c - 1]; Sigt[c] = GreySigTotal[c0 + c - 1]; }
      for (int c = 1; c <= nCorner; c++) {
        double sigv = Volume[c0 + c - 1] * Sigt[c];
        denom[c] = sigv;
        int nCFaces = M->nCFaces[c0 + c - 1];
        for (int cface = 1; cface <= nCFaces; cface++) {
          afp[cface] = dot3(omega, &F3(A_fp, 1, cface, c0 + c, 3, 3));
          if (afp[cface] > 0.0) denom[c] = denom[c] + afp[cface];
        }
        for (int cface = 1; cface <= nCFaces; cface++) {
          double aez = dot3(omega, &F3(A_ez, 1, cface, c0 + c, 3, 3));
          int cez = F2(M->cEZ, cface, c0 + c, 3);
          if (cez > c) {
            if (aez > 0.0) { need[cez]++; nxez[c]++; ez_exit[nxez[c]][c] = cez; coefpsi[nxez[c]][c] = aez; }
            else if (aez < 0.0) { need[c]++; nxez[cez]++; ez_exit[nxez[cez]][cez] = c; coefpsi[nxez[cez]][cez] = -aez; }
          }
          if (aez > 0.0) {
            denom[c] = denom[c] + aez;
            double area_opp = 0.0;
            int ifp;
            if (nCFaces == 3) {
              ifp = cface % nCFaces + 1;
              if (afp[ifp] < 0.0) area_opp = -afp[ifp];
            } else {
              ifp = cface;
              for (int k = 1; k <= nCFaces - 2; k++) {
                ifp = ifp % nCFaces + 1;
                if (afp[ifp] < 0.0) area_opp = area_opp - afp[ifp];
              }
            }
            double B1, B2;
            if (area_opp > 0.0) {
              double sigv2 = sigv * sigv;
              double gnum = aez * aez * (fouralpha * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
              double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + aez * sigv * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
              double B0 = 0.5 * aez * (1.0 - gtau);
              B1 = (B0 - gtau * sigv) / Sigt[c];
              B2 = B0 / Sigt[cez];
            } else {
              B1 = 0.5 * aez / Sigt[c];
              B2 = 0.5 * aez / Sigt[cez];
            }
            Pvv[c][c] = Pvv[c][c] + B1;       /* Pvv(row, col) = Pvv[row][col] */
            Pvv[cez][c] = Pvv[cez][c] - B2;
            Pvv[c][cez] = Pvv[c][cez] - B1;
            Pvv[cez][cez] = Pvv[cez][cez] + B2;
          }
        }
      }
      for (int i = 1; i <= nCorner; i++) {
        int c = 1;   /* minloc: first minimum */
        for (int k = 2; k <= nCorner; k++) if (need[k] < need[c]) c = k;
        double dInv = 1.0 / denom[c];
        for (int c1 = 1; c1 <= nCorner; c1++) Pvv[c1][c] = dInv * Pvv[c1][c];
        for (int cface = 1; cface <= nxez[c]; cface++) {
          int cez = ez_exit[cface][c];
          double coef = coefpsi[cface][c];
          need[cez]--;
          for (int c1 = 1; c1 <= nCorner; c1++) Pvv[c1][cez] = Pvv[c1][cez] + coef * Pvv[c1][c];
        }
        need[c] = 99;
      }
      for (int c1 = 1; c1 <= nCorner; c1++)
        for (int c = 1; c <= nCorner; c++) F2(TT, c, c0 + c1, mC) = F2(TT, c, c0 + c1, mC) + quadwt * Pvv[c][c1];
    }
  }
}

/* SweepGreyUCBxyz (useNewGTASolver) for one angle: tPsi(nc+nb), pInc(nc) scratch; PsiBa(nb) in/out; PhiInc += w pInc */
void orc_gta_sweep_angle(const orc_mesh *M, int nHyperPlanes, const int *zonesInPlane, const int *nextZ, const int *nextC,
                         const double *omega, double quadwt, const double *Volume, const double *A_fp, const double *A_ez,
                         const double *GreySigTotal, const double *GreySigtInv, const double *TsaSource, double *tPsi,
                         double *pInc, double *PsiBa, double *PhiInc) {
  const int nc = M->ncornr, nb = M->nbelem;
  for (int c = 0; c < nc; c++) { tPsi[c] = 0.0; pInc[c] = 0.0; }
  for (int ib = 0; ib < nb; ib++) tPsi[nc + ib] = PsiBa[ib];
  int ndoneZ = 0;
  for (int hp = 1; hp <= nHyperPlanes; hp++) {
    int nzones = zonesInPlane[hp - 1];
    for (int ii = 1; ii <= nzones; ii++) {
      int zone = abs(nextZ[ndoneZ + ii - 1]);
      int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      int nxBdy = 0, nxez[MAXC + 1] = {0}, ez_exit[MAXCF + 1][MAXC + 1], bdy_exit[2][MAXCF * MAXC + 1];
      double denom[MAXC + 1], afp[MAXCF + 1], coefpsi[MAXCF + 1][MAXC + 1], psifp[MAXCF + 1], src[MAXC + 1], Q[MAXC + 1];
      for (int c = 1; c <= nCorner; c++) {
        Q[c] = GreySigtInv[c0 + c - 1] * TsaSource[c0 + c - 1];
        src[c] = Volume[c0 + c - 1] * TsaSource[c0 + c - 1];
      }
      for (int c = 1; c <= nCorner; c++) {
        double sigv = Volume[c0 + c - 1] * GreySigTotal[c0 + c - 1];
        denom[c] = sigv;
        int nCFaces = M->nCFaces[c0 + c - 1];
        for (int cface = 1; cface <= nCFaces; cface++) {
          afp[cface] = dot3(omega, &F3(A_fp, 1, cface, c0 + c, 3, 3));
          int cfp = F2(M->cFP, cface, c0 + c, 3);
          if (afp[cface] > 0.0) {
            denom[c] = denom[c] + afp[cface];
            if (cfp > nc) { nxBdy++; bdy_exit[0][nxBdy] = c; bdy_exit[1][nxBdy] = cfp - nc; }
          } else if (afp[cface] < 0.0) {
            psifp[cface] = tPsi[cfp - 1];
            src[c] = src[c] - afp[cface] * psifp[cface];
            pInc[c0 + c - 1] = pInc[c0 + c - 1] - afp[cface] * psifp[cface];
          }
        }
        for (int cface = 1; cface <= nCFaces; cface++) {
          double aez = dot3(omega, &F3(A_ez, 1, cface, c0 + c, 3, 3));
          int cez = F2(M->cEZ, cface, c0 + c, 3);
          if (cez > c) {
            if (aez > 0.0) { nxez[c]++; ez_exit[nxez[c]][c] = cez; coefpsi[nxez[c]][c] = aez; }
            else if (aez < 0.0) { nxez[cez]++; ez_exit[nxez[cez]][cez] = c; coefpsi[nxez[cez]][cez] = -aez; }
          }
          if (aez > 0.0) {
            double psi_opp = 0.0, area_opp = 0.0;
            denom[c] = denom[c] + aez;
            int ifp = cface % nCFaces + 1;
            if (afp[ifp] < 0.0) { area_opp = -afp[ifp]; psi_opp = -afp[ifp] * psifp[ifp]; }
            for (int k = 2; k <= nCFaces - 2; k++) {
              ifp = ifp % nCFaces + 1;
              if (afp[ifp] < 0.0) { area_opp = area_opp - afp[ifp]; psi_opp = psi_opp - afp[ifp] * psifp[ifp]; }
            }
            double sez;
            if (area_opp > 0.0) {
              psi_opp = psi_opp / area_opp;
              double sigv2 = sigv * sigv;
              double gnum = aez * aez * (fouralpha * sigv2 + aez * (4.0 * sigv + 3.0 * aez));
              double gtau = gnum / (gnum + 4.0 * sigv2 * sigv2 + aez * sigv * (6.0 * sigv2 + 2.0 * aez * (2.0 * sigv + aez)));
              sez = gtau * sigv * (psi_opp - Q[c]) + 0.5 * aez * (1.0 - gtau) * (Q[c] - Q[cez]);
              src[c] = src[c] + sez;
              src[cez] = src[cez] - sez;
              pInc[c0 + c - 1] = pInc[c0 + c - 1] + gtau * sigv * psi_opp;
              pInc[c0 + cez - 1] = pInc[c0 + cez - 1] - gtau * sigv * psi_opp;
            } else {
              sez = 0.5 * aez * (Q[c] - Q[cez]);
              src[c] = src[c] + sez;
              src[cez] = src[cez] - sez;
            }
          }
        }
      }
      for (int i = 1; i <= nCorner; i++) {
        int c = nextC[c0 + i - 1];
        tPsi[c0 + c - 1] = src[c] / denom[c];
        pInc[c0 + c - 1] = pInc[c0 + c - 1] / denom[c];
        for (int cface = 1; cface <= nxez[c]; cface++) {
          int cez = ez_exit[cface][c];
          src[cez] = src[cez] + coefpsi[cface][c] * tPsi[c0 + c - 1];
          pInc[c0 + cez - 1] = pInc[c0 + cez - 1] + coefpsi[cface][c] * pInc[c0 + c - 1];
        }
      }
      for (int ib = 1; ib <= nxBdy; ib++) PsiBa[bdy_exit[1][ib] - 1] = tPsi[c0 + bdy_exit[0][ib] - 1];
    }
    ndoneZ += nzones;
  }
  for (int c = 0; c < nc; c++) PhiInc[c] = PhiInc[c] + quadwt * pInc[c];
}

typedef struct {
  const orc_mesh *M;
  int nAng;
  const int *nHyperPlanes, *zonesInPlane /* (nz,nAng) */, *nextZ /* (nz,nAng) */, *nextC /* (nc,nAng) */;
  const double *omega, *weight, *Volume, *A_fp, *A_ez;
  const double *GreySigTotal, *GreySigtInv, *GreySigScat, *GreySigScatVol;
  double *GreySource;   /* (nc) */
  double *TT;           /* (maxCorner,nc) — decomposed in place by the first (withSource) sweep */
  double wtiso;
  /* r-z only (M->ndim == 2): omega is (2,nAng), A_fp/A_ez are (2,2,nc) */
  const double *Area, *RadiusFP, *RadiusEZ, *angDerivFac, *quadTauW1, *quadTauW2;
  const unsigned char *start, *finish;
  /* reflecting boundaries (3-D): GTASweep.F90:151 calls snreflect before each angle; angles are taken stage by stage (mirror images
     first), the copies PsiB(first..first+n-1, minc) <- PsiB(.., mref) of a stage made before its sweeps */
  int nStagesR, nReflOps;
  const int *angleStage;   /* (nAng) */
  const int *reflOps;      /* (nReflOps, 5): stage, minc, mref (0-based angles), first (0-based element), n */
} orc_gta;

void orc_gta_sweep_angle_rz(const orc_mesh *M, int nHyperPlanes, const int *zonesInPlane, const int *nextZ, const int *nextC,
                            const double *omega, double quadwt, double fac, double quadTauW1, double quadTauW2, int StartingDirection,
                            const double *Volume, const double *Area, const double *A_fp, const double *A_ez, const double *RadiusFP,
                            const double *RadiusEZ, const double *GreySigTotal, const double *GreySigtInv, const double *TsaSource,
                            double *tPsi, double *pInc, double *tPsiM, double *tInc, double *PsiBa, double *PhiInc);
void orc_gta_init_tt_rz(const orc_mesh *M, int nAng, const int *nextC, const double *omegas, const double *weights, const unsigned char *start,
                        const unsigned char *finish, const double *angDerivFac, const double *quadTauW1, const double *quadTauW2,
                        const double *Volume, const double *Area, const double *A_fp, const double *A_ez, const double *RadiusFP,
                        const double *RadiusEZ, const double *GreySigTotal, double *TT);

/* ScalarIntensityDecompose + ScalarIntensitySolve for one zone */
static void scalar_intensity(const orc_gta *S, int zone, const double *PhiInc, double *P, int withSource) {
  const orc_mesh *M = S->M;
  const int mC = M->maxCorner;
  int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
  double Phi[MAXC + 1];
  double *TT = S->TT;
  if (withSource) {
    for (int c = 1; c <= nCorner; c++) {
      Phi[c] = PhiInc[c0 + c - 1];
      for (int cc = 1; cc <= nCorner; cc++) {
        Phi[c] = Phi[c] + F2(TT, cc, c0 + c, mC) * S->wtiso * S->GreySource[c0 + cc - 1];
        F2(TT, cc, c0 + c, mC) = -S->wtiso * S->GreySigScat[c0 + cc - 1] * F2(TT, cc, c0 + c, mC);
      }
      F2(TT, c, c0 + c, mC) = 1.0 + F2(TT, c, c0 + c, mC);
    }
    for (int i = 1; i <= nCorner; i++) {
      double t = 0.0;
      for (int k = 1; k <= i - 1; k++) t = t + F2(TT, k, c0 + i, mC) * F2(TT, i, c0 + k, mC);
      F2(TT, i, c0 + i, mC) = F2(TT, i, c0 + i, mC) - t;
      double diagInv = 1.0 / F2(TT, i, c0 + i, mC);
      for (int j = i + 1; j <= nCorner; j++) {
        t = 0.0;
        double v = 0.0;
        for (int k = 1; k <= i - 1; k++) {
          t = t + F2(TT, k, c0 + i, mC) * F2(TT, j, c0 + k, mC);
          v = v + F2(TT, k, c0 + j, mC) * F2(TT, i, c0 + k, mC);
        }
        F2(TT, j, c0 + i, mC) = F2(TT, j, c0 + i, mC) - t;
        F2(TT, i, c0 + j, mC) = diagInv * (F2(TT, i, c0 + j, mC) - v);
      }
    }
  } else {
    for (int c = 1; c <= nCorner; c++) Phi[c] = PhiInc[c0 + c - 1];
  }
  for (int j = 2; j <= nCorner; j++) {
    double t = 0.0;
    for (int i = 1; i <= j - 1; i++) t = t - F2(TT, i, c0 + j, mC) * Phi[i];
    Phi[j] = Phi[j] + t;
  }
  Phi[nCorner] = Phi[nCorner] / F2(TT, nCorner, c0 + nCorner, mC);
  for (int k = nCorner - 1; k >= 1; k--) {
    double t = 0.0;
    for (int i = k + 1; i <= nCorner; i++) t = t + Phi[i] * F2(TT, i, c0 + k, mC);
    Phi[k] = (Phi[k] - t) / F2(TT, k, c0 + k, mC);
  }
  for (int c = 1; c <= nCorner; c++) P[c0 + c - 1] = Phi[c];
}

/* GreySweepNEW: GTASweep(P, PsiB) with GTA%ID = 1, then the per-zone solves.  PsiB is (nb, nAng). */
void orc_gta_grey_sweep(const orc_gta *S, double *PsiB, double *P, int withSource) {
  const orc_mesh *M = S->M;
  const int nc = M->ncornr, nb = M->nbelem, nz = M->nzones;
  double *TsaSource = malloc(sizeof(double) * nc), *PhiInc = calloc(nc, sizeof(double));
  double *tPsi = malloc(sizeof(double) * (nc + nb)), *pInc = malloc(sizeof(double) * nc);
  for (int c = 0; c < nc; c++) TsaSource[c] = S->wtiso * (S->GreySigScat[c] * P[c] + S->GreySource[c]);
  if (M->ndim == 2) {
    /* GTASweep.F90:113-119: tPsiM = tInc = 0, then every non-finishing angle in turn (:149-159), snreflect before each (:151).
       With reflecting boundaries the angles are taken in stage order (levels advance together), so the half-angle arrays are
       kept per xi-level; PhiInc is summed in angle order afterwards */
    int nLev = 0;
    int levelOf[64];
    for (int a = 0; a < S->nAng; a++) { if (S->start[a]) nLev++; levelOf[a] = nLev - 1; }
    double *tPsiM = calloc((size_t)nc * nLev, sizeof(double)), *tInc = calloc((size_t)nc * nLev, sizeof(double));
    double *pAll = calloc((size_t)nc * S->nAng, sizeof(double)), *dummy = calloc(nc, sizeof(double));
    const int nSt = S->nReflOps > 0 ? S->nStagesR : 1;
    for (int st = 0; st < nSt; st++) {
      for (int a = 0; a < S->nAng; a++) {
        if (S->finish[a]) continue;
        if (S->nReflOps > 0 && S->angleStage[a] != st) continue;
        for (int o = 0; o < S->nReflOps; o++) {
          const int *op = S->reflOps + 5 * o;
          if (op[0] != st || op[1] != a) continue;
          for (int i = 0; i < op[4]; i++) PsiB[(size_t)nb * op[1] + op[3] + i] = PsiB[(size_t)nb * op[2] + op[3] + i];
        }
        orc_gta_sweep_angle_rz(M, S->nHyperPlanes[a], S->zonesInPlane + (size_t)nz * a, S->nextZ + (size_t)nz * a,
                               S->nextC + (size_t)nc * a, S->omega + 2 * a, S->weight[a], S->angDerivFac[a], S->quadTauW1[a],
                               S->quadTauW2[a], S->start[a], S->Volume, S->Area, S->A_fp, S->A_ez, S->RadiusFP, S->RadiusEZ,
                               S->GreySigTotal, S->GreySigtInv, TsaSource, tPsi, pInc, tPsiM + (size_t)nc * levelOf[a],
                               tInc + (size_t)nc * levelOf[a], PsiB + (size_t)nb * a, dummy);
        memcpy(pAll + (size_t)nc * a, pInc, sizeof(double) * nc);
      }
    }
    for (int a = 0; a < S->nAng; a++)
      for (int c = 0; c < nc; c++) PhiInc[c] = PhiInc[c] + S->weight[a] * pAll[(size_t)nc * a + c];
    free(tPsiM); free(tInc); free(pAll); free(dummy);
  } else {
    /* PhiInc is summed in angle order whatever the sweep order (as gta_phiinc_kernel does): each angle's w pInc is kept */
    const int nSt = S->nReflOps > 0 ? S->nStagesR : 1;
    double *pAll = calloc((size_t)nc * S->nAng, sizeof(double)), *dummy = calloc(nc, sizeof(double));
    for (int st = 0; st < nSt; st++) {
      for (int o = 0; o < S->nReflOps; o++) {
        const int *op = S->reflOps + 5 * o;
        if (op[0] != st) continue;
        for (int i = 0; i < op[4]; i++) PsiB[(size_t)nb * op[1] + op[3] + i] = PsiB[(size_t)nb * op[2] + op[3] + i];
      }
      for (int a = 0; a < S->nAng; a++) {
        if (S->nReflOps > 0 && S->angleStage[a] != st) continue;
        orc_gta_sweep_angle(M, S->nHyperPlanes[a], S->zonesInPlane + (size_t)nz * a, S->nextZ + (size_t)nz * a,
                            S->nextC + (size_t)nc * a, S->omega + 3 * a, S->weight[a], S->Volume, S->A_fp, S->A_ez,
                            S->GreySigTotal, S->GreySigtInv, TsaSource, tPsi, pInc, PsiB + (size_t)nb * a, dummy);
        memcpy(pAll + (size_t)nc * a, pInc, sizeof(double) * nc);
      }
    }
    for (int a = 0; a < S->nAng; a++)
      for (int c = 0; c < nc; c++) PhiInc[c] = PhiInc[c] + S->weight[a] * pAll[(size_t)nc * a + c];
    free(pAll); free(dummy);
  }
  for (int zone = 1; zone <= nz; zone++) scalar_intensity(S, zone, PhiInc, P, withSource);
  free(TsaSource); free(PhiInc); free(tPsi); free(pInc);
}

static double scat_prod1(const orc_gta *S, const double *x) {
  double s = 0.0;
  for (int i = 0; i < S->M->ncornr; i++) s = s + x[i] * S->GreySigScatVol[i];
  return s;
}
static double scat_prod(const orc_gta *S, const double *x, const double *y) {
  double s = 0.0;
  for (int i = 0; i < S->M->ncornr; i++) s = s + x[i] * y[i] * S->GreySigScatVol[i];
  return s;
}

/* GTASolver (single domain).  PhiTotal (ngr,nc) -> radEnergy per zone for the convergence test; GreySource in S is
   consumed (zeroed after the first sweep as in the reference).  Returns nGreyIter; GreyCorrection(nc) out. */
int orc_gta_solver(orc_gta *S, int ngr, const double *PhiTotal, const double *VolumeZone, double epsPoint, int maxIters,
                   double epsGrey, int enforceHardMax, double *GreyCorrection, double *maxRelErrOut) {
  const orc_mesh *M = S->M;
  const int nc = M->ncornr, nb = M->nbelem, nz = M->nzones, nA = S->nAng;
  const double adqtSmall = 1.e-150;
  double *radEnergy = calloc(nz, sizeof(double)), *pzOld = calloc(nz, sizeof(double));
  double *R = calloc(nc, sizeof(double)), *D = calloc(nc, sizeof(double)), *A = calloc(nc, sizeof(double)), *AS = calloc(nc, sizeof(double));
  size_t nB = (size_t)nb * nA > 0 ? (size_t)nb * nA : 1;
  double *RB = calloc(nB, sizeof(double)), *DB = calloc(nB, sizeof(double)), *AB = calloc(nB, sizeof(double)), *ASB = calloc(nB, sizeof(double));
  for (int zone = 1; zone <= nz; zone++) {
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    for (int c = 1; c <= nCorner; c++) {
      double sumRad = 0.0;
      for (int g = 1; g <= ngr; g++) sumRad = sumRad + F2(PhiTotal, g, c0 + c, ngr);
      radEnergy[zone - 1] = radEnergy[zone - 1] + S->Volume[c0 + c - 1] * sumRad;
    }
    radEnergy[zone - 1] = radEnergy[zone - 1] / VolumeZone[zone - 1];
  }
  if (M->ndim == 2)
    orc_gta_init_tt_rz(M, nA, S->nextC, S->omega, S->weight, S->start, S->finish, S->angDerivFac, S->quadTauW1, S->quadTauW2, S->Volume,
                       S->Area, S->A_fp, S->A_ez, S->RadiusFP, S->RadiusEZ, S->GreySigTotal, S->TT);
  else
    orc_gta_init_tt(M, nA, S->omega, S->weight, S->Volume, S->A_fp, S->A_ez, S->GreySigTotal, S->TT);
  for (int c = 0; c < nc; c++) GreyCorrection[c] = 0.0;
  int nGreyIter = 1;
  orc_gta_grey_sweep(S, RB, R, 1);
  memcpy(D, R, sizeof(double) * nc);
  memcpy(DB, RB, sizeof(double) * nB);
  double rrOld = scat_prod1(S, R), maxRelErrGrey = 0.0;
  for (int c = 0; c < nc; c++) S->GreySource[c] = 0.0;
  for (;;) {
    if (fabs(rrOld) < adqtSmall) {
      if (nGreyIter <= 2) memcpy(GreyCorrection, R, sizeof(double) * nc);
      break;
    }
    nGreyIter += 2;
    memcpy(A, D, sizeof(double) * nc);
    memcpy(AB, DB, sizeof(double) * nB);
    orc_gta_grey_sweep(S, AB, A, 0);
    for (int c = 0; c < nc; c++) A[c] = D[c] - A[c];
    for (size_t i = 0; i < nB; i++) AB[i] = DB[i] - AB[i];
    double dAd = scat_prod1(S, A);
    if (fabs(dAd) < adqtSmall) break;
    double alpha = rrOld / dAd;
    for (int c = 0; c < nc; c++) R[c] = R[c] - alpha * A[c];
    for (size_t i = 0; i < nB; i++) RB[i] = RB[i] - alpha * AB[i];
    memcpy(AS, R, sizeof(double) * nc);
    memcpy(ASB, RB, sizeof(double) * nB);
    orc_gta_grey_sweep(S, ASB, AS, 0);
    for (int c = 0; c < nc; c++) AS[c] = R[c] - AS[c];
    for (size_t i = 0; i < nB; i++) ASB[i] = RB[i] - ASB[i];
    double omegaNum = scat_prod(S, AS, R), omegaDen = scat_prod(S, AS, AS);
    if (fabs(omegaDen) < adqtSmall || fabs(omegaNum) < adqtSmall) {
      for (int c = 0; c < nc; c++) GreyCorrection[c] = GreyCorrection[c] + alpha * D[c];
      break;
    }
    double omegaCG = omegaNum / omegaDen;
    for (int c = 0; c < nc; c++) GreyCorrection[c] = GreyCorrection[c] + alpha * D[c] + omegaCG * R[c];
    for (int c = 0; c < nc; c++) R[c] = R[c] - omegaCG * AS[c];
    for (size_t i = 0; i < nB; i++) RB[i] = RB[i] - omegaCG * ASB[i];
    double rr = scat_prod1(S, R);
    double beta = (rr * alpha) / (rrOld * omegaCG);
    for (int c = 0; c < nc; c++) D[c] = R[c] + beta * (D[c] - omegaCG * A[c]);
    for (size_t i = 0; i < nB; i++) DB[i] = RB[i] + beta * (DB[i] - omegaCG * AB[i]);
    double errL2 = 0.0, phiL2 = 0.0, maxRelErrPoint = 0.0;
    for (int zone = 1; zone <= nz; zone++) {
      int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      double pz = 0.0;
      for (int c = 1; c <= nCorner; c++) pz = pz + S->Volume[c0 + c - 1] * GreyCorrection[c0 + c - 1];
      pz = pz / VolumeZone[zone - 1];
      double errZone = pz - pzOld[zone - 1];
      errL2 = errL2 + VolumeZone[zone - 1] * (errZone * errZone);
      double phiNew = radEnergy[zone - 1] + pz;
      phiL2 = phiL2 + VolumeZone[zone - 1] * (phiNew * phiNew);
      if (phiNew != 0.0) {
        double rel = fabs(errZone / phiNew);
        if (rel > maxRelErrPoint) maxRelErrPoint = rel;
      }
      pzOld[zone - 1] = pz;
    }
    double relErrL2 = phiL2 != 0.0 ? sqrt(fabs(errL2 / phiL2)) : 0.0;
    maxRelErrGrey = maxRelErrPoint > relErrL2 ? maxRelErrPoint : relErrL2;
    if (enforceHardMax && nGreyIter >= maxIters) break;
    else if ((maxRelErrGrey < epsPoint || nGreyIter >= maxIters) && maxRelErrGrey < epsGrey) break;
    else if (nGreyIter >= 100 * maxIters) { fprintf(stderr, "orc_gta_solver: not converging\n"); break; }
    rrOld = rr;
  }
  if (maxRelErrOut) *maxRelErrOut = maxRelErrGrey;
  free(radEnergy); free(pzOld); free(R); free(D); free(A); free(AS); free(RB); free(DB); free(AB); free(ASB);
  return nGreyIter;
}

/* addGreyCorrections.F90:70-91: PhiTotal(g,c) += GreyCorrection(c) Chi(g,c) */
void orc_add_grey_corrections(int ngr, int nc, const double *GreyCorrection, const double *Chi, double *PhiTotal) {
  for (int c = 1; c <= nc; c++)
    for (int g = 1; g <= ngr; g++)
      F2(PhiTotal, g, c, ngr) = F2(PhiTotal, g, c, ngr) + GreyCorrection[c - 1] * F2(Chi, g, c, ngr);
}

/* ------------------------------------------------------------------ */
/* r-z grey sweeps                                                     */
/* ------------------------------------------------------------------ */
int orc_rz_angle_coefs(int NA, const double *omega, double *weight, unsigned char *start, unsigned char *finish, int *angleToLevel,
                       double *alpha, double *tauc, double *angDerivFac, double *quadTauW1, double *quadTauW2);   /* umt_oracle.c */

/* the GTA angle set in r-z: level-symmetric S2 (quadrz.F90:82-160 with norder = 2: one pair of xi-levels, one ordinate per
   quadrant, dircos 0.577350269189625, weight 1), each level = starting direction, mu < 0, mu > 0, finishing direction */
int orc_gta_quad_rz(double *omega /* (2,8) */, double *weight, unsigned char *start, unsigned char *finish, int *angleToLevel,
                    double *angDerivFac, double *quadTauW1, double *quadTauW2) {
  const double pi = 3.14159265358979323846, halfpi = 0.5 * pi;
  const double xilev = 0.577350269189625, mu = 0.577350269189625, wgt = 1.0;
  int m = 0;
  for (int half = 0; half < 2; half++) {
    const double sxi = half == 0 ? -xilev : xilev;
    omega[2 * m] = -sqrt(1.0 - xilev * xilev); omega[2 * m + 1] = sxi; weight[m] = 0.0; m++;
    omega[2 * m] = -mu; omega[2 * m + 1] = sxi; weight[m] = halfpi * wgt; m++;
    omega[2 * m] = mu; omega[2 * m + 1] = sxi; weight[m] = halfpi * wgt; m++;
    omega[2 * m] = sqrt(1.0 - xilev * xilev); omega[2 * m + 1] = sxi; weight[m] = 0.0; m++;
  }
  double alpha[8], tauc[8];
  return orc_rz_angle_coefs(8, omega, weight, start, finish, angleToLevel, alpha, tauc, angDerivFac, quadTauW1, quadTauW2);
}

static double dot2(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1]; }

/* SweepGreyUCBrz + SweepGreyUCBrzKernelNew for one (non-finishing) angle */
void orc_gta_sweep_angle_rz(const orc_mesh *M, int nHyperPlanes, const int *zonesInPlane, const int *nextZ, const int *nextC,
                            const double *omega, double quadwt, double fac, double quadTauW1, double quadTauW2, int StartingDirection,
                            const double *Volume, const double *Area, const double *A_fp, const double *A_ez, const double *RadiusFP,
                            const double *RadiusEZ, const double *GreySigTotal, const double *GreySigtInv, const double *TsaSource,
                            double *tPsi, double *pInc, double *tPsiM, double *tInc, double *PsiBa, double *PhiInc) {
  const int nc = M->ncornr, nb = M->nbelem;
  for (int c = 0; c < nc; c++) { tPsi[c] = 0.0; pInc[c] = 0.0; }
  for (int ib = 0; ib < nb; ib++) tPsi[nc + ib] = PsiBa[ib];
  int ndoneZ = 0;
  for (int hp = 1; hp <= nHyperPlanes; hp++) {
    int nzones = zonesInPlane[hp - 1];
    for (int ii = 1; ii <= nzones; ii++) {
      int zone = abs(nextZ[ndoneZ + ii - 1]);
      int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
      int nxBdy = 0, nxez[MAXC + 1] = {0}, ez_exit[3][MAXC + 1], bdy_exit[2][2 * MAXC + 1];
      double denom[MAXC + 1], coefpsi[3][MAXC + 1], Sigt[MAXC + 1], Q[MAXC + 1], src[MAXC + 1];
      for (int c = 1; c <= nCorner; c++) {
        Q[c] = GreySigtInv[c0 + c - 1] * TsaSource[c0 + c - 1];
        src[c] = Volume[c0 + c - 1] * TsaSource[c0 + c - 1] + fac * Area[c0 + c - 1] * tPsiM[c0 + c - 1];
        Sigt[c] = GreySigTotal[c0 + c - 1];
        denom[c] = Sigt[c] * Volume[c0 + c - 1] + fac * Area[c0 + c - 1];
        pInc[c0 + c - 1] = fac * Area[c0 + c - 1] * tInc[c0 + c - 1];
      }
      for (int c = 1; c <= nCorner; c++) {
        for (int cface = 1; cface <= 2; cface++) {
          double afp = dot2(omega, &F3(A_fp, 1, cface, c0 + c, 2, 2));
          double aez = dot2(omega, &F3(A_ez, 1, cface, c0 + c, 2, 2));
          int cfp = F2(M->cFP, cface, c0 + c, 2);
          if (afp < 0.0) {
            double R_afp = F2(RadiusFP, cface, c0 + c, 2) * afp;
            denom[c] = denom[c] - R_afp;
            src[c] = src[c] - R_afp * tPsi[cfp - 1];
            pInc[c0 + c - 1] = pInc[c0 + c - 1] - R_afp * tPsi[cfp - 1];
          } else if (cfp > nc) {
            nxBdy++; bdy_exit[0][nxBdy] = c; bdy_exit[1][nxBdy] = cfp - nc;
          }
          if (aez > 0.0) {
            double R = F2(RadiusEZ, cface, c0 + c, 2);
            int cez = F2(M->cEZ, cface, c0 + c, 2);
            nxez[c]++; ez_exit[nxez[c]][c] = cez; coefpsi[nxez[c]][c] = R * aez;
            denom[cez] = denom[cez] + R * aez;
            double sez;
            if (afp < 0.0) {
              double sigA = Sigt[c] * Area[c0 + c - 1], sigA2 = sigA * sigA;
              double gnum = aez * aez * (fouralpha * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
              double gtau = gnum / (gnum + 4.0 * sigA2 * sigA2 + aez * sigA * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
              sez = R * (gtau * sigA * (tPsi[cfp - 1] - Q[c]) + 0.5 * aez * (1.0 - gtau) * (Q[c] - Q[cez]));
              src[c] = src[c] + sez;
              src[cez] = src[cez] - sez;
              pInc[c0 + c - 1] = pInc[c0 + c - 1] + R * gtau * sigA * tPsi[cfp - 1];
              pInc[c0 + cez - 1] = pInc[c0 + cez - 1] - R * gtau * sigA * tPsi[cfp - 1];
            } else {
              sez = 0.5 * R * aez * (Q[c] - Q[cez]);
              src[c] = src[c] + sez;
              src[cez] = src[cez] - sez;
            }
          }
        }
      }
      for (int i = 1; i <= nCorner; i++) {
        int c = nextC[c0 + i - 1];
        tPsi[c0 + c - 1] = src[c] / denom[c];
        pInc[c0 + c - 1] = pInc[c0 + c - 1] / denom[c];
        for (int cface = 1; cface <= nxez[c]; cface++) {
          int cez = ez_exit[cface][c];
          src[cez] = src[cez] + coefpsi[cface][c] * tPsi[c0 + c - 1];
          pInc[c0 + cez - 1] = pInc[c0 + cez - 1] + coefpsi[cface][c] * pInc[c0 + c - 1];
        }
      }
      for (int ib = 1; ib <= nxBdy; ib++) PsiBa[bdy_exit[1][ib] - 1] = tPsi[c0 + bdy_exit[0][ib] - 1];
      for (int c = 1; c <= nCorner; c++) {
        if (StartingDirection) tInc[c0 + c - 1] = pInc[c0 + c - 1];
        else tInc[c0 + c - 1] = quadTauW1 * pInc[c0 + c - 1] - quadTauW2 * tInc[c0 + c - 1];
      }
    }
    ndoneZ += nzones;
  }
  for (int c = 0; c < nc; c++) PhiInc[c] = PhiInc[c] + quadwt * pInc[c];
  for (int c = 0; c < nc; c++) {
    if (StartingDirection) tPsiM[c] = tPsi[c];
    else tPsiM[c] = quadTauW1 * tPsi[c] - quadTauW2 * tPsiM[c];
  }
}

/* InitGreySweepUCBrz for every zone: TT(c1, c0+c) = sum over the weighted angles of quadwt Pvv(c1,c), the starting direction's Pvv
   carried through Tvv like PsiM */
void orc_gta_init_tt_rz(const orc_mesh *M, int nAng, const int *nextC /* (nc,nAng) */, const double *omegas, const double *weights,
                        const unsigned char *start, const unsigned char *finish, const double *angDerivFac, const double *quadTauW1,
                        const double *quadTauW2, const double *Volume, const double *Area, const double *A_fp, const double *A_ez,
                        const double *RadiusFP, const double *RadiusEZ, const double *GreySigTotal, double *TT) {
  const int mC = M->maxCorner, nc = M->ncornr;
  for (int zone = 1; zone <= M->nzones; zone++) {
    int nCorner = M->numCorner[zone - 1], c0 = M->cOffSet[zone - 1];
    double Tvv[MAXC + 1][MAXC + 1], Pvv[MAXC + 1][MAXC + 1], Sigt[MAXC + 1], denom[MAXC + 1], coefpsi[3][MAXC + 1]; /* [column][row] */
    int nxez[MAXC + 1], ez_exit[3][MAXC + 1];
    for (int i = 0; i <= MAXC; i++) for (int j = 0; j <= MAXC; j++) Tvv[i][j] = 0.0;
    for (int c = 1; c <= nCorner; c++) {
      for (int c1 = 1; c1 <= mC; c1++) F2(TT, c1, c0 + c, mC) = 0.0;
      Sigt[c] = GreySigTotal[c0 + c - 1];
    }
    for (int a = 0; a < nAng; a++) {
      if (finish[a]) continue;
      const double *omega = omegas + 2 * a;
      double quadwt = weights[a], fac = angDerivFac[a];
      for (int i = 0; i <= MAXC; i++) { nxez[i] = 0; for (int j = 0; j <= MAXC; j++) Pvv[i][j] = 0.0; }
      for (int c = 1; c <= nCorner; c++) {
        Pvv[c][c] = Volume[c0 + c - 1];
        denom[c] = Sigt[c] * Volume[c0 + c - 1] + fac * Area[c0 + c - 1];
        for (int c1 = 1; c1 <= nCorner; c1++) Pvv[c1][c] = Pvv[c1][c] + fac * Area[c0 + c - 1] * Tvv[c1][c];
      }
      for (int c = 1; c <= nCorner; c++) {
        for (int cface = 1; cface <= 2; cface++) {
          double afp = dot2(omega, &F3(A_fp, 1, cface, c0 + c, 2, 2));
          double aez = dot2(omega, &F3(A_ez, 1, cface, c0 + c, 2, 2));
          if (afp < 0.0) denom[c] = denom[c] - F2(RadiusFP, cface, c0 + c, 2) * afp;
          if (aez > 0.0) {
            double R = F2(RadiusEZ, cface, c0 + c, 2);
            int cez = F2(M->cEZ, cface, c0 + c, 2);
            nxez[c]++; ez_exit[nxez[c]][c] = cez; coefpsi[nxez[c]][c] = R * aez;
            denom[cez] = denom[cez] + R * aez;
            double B1, B2;
            if (afp < 0.0) {
              double sigA = Sigt[c] * Area[c0 + c - 1], sigA2 = sigA * sigA;
              double gnum = aez * aez * (fouralpha * sigA2 + aez * (4.0 * sigA + 3.0 * aez));
              double gtau = gnum / (gnum + 4.0 * sigA2 * sigA2 + aez * sigA * (6.0 * sigA2 + 2.0 * aez * (2.0 * sigA + aez)));
              double B0 = 0.5 * aez * (1.0 - gtau) * R;
              B1 = (B0 - R * gtau * sigA) / Sigt[c];
              B2 = B0 / Sigt[cez];
            } else {
              B1 = 0.5 * R * aez / Sigt[c];
              B2 = 0.5 * R * aez / Sigt[cez];
            }
            Pvv[c][c] = Pvv[c][c] + B1;
            Pvv[cez][c] = Pvv[cez][c] - B2;
            Pvv[c][cez] = Pvv[c][cez] - B1;
            Pvv[cez][cez] = Pvv[cez][cez] + B2;
          }
        }
      }
      for (int i = 1; i <= nCorner; i++) {
        int c = nextC[(size_t)nc * a + c0 + i - 1];
        double dInv = 1.0 / denom[c];
        for (int c1 = 1; c1 <= nCorner; c1++) Pvv[c1][c] = dInv * Pvv[c1][c];
        for (int cface = 1; cface <= nxez[c]; cface++) {
          int cez = ez_exit[cface][c];
          double coef = coefpsi[cface][c];
          for (int c1 = 1; c1 <= nCorner; c1++) Pvv[c1][cez] = Pvv[c1][cez] + coef * Pvv[c1][c];
        }
      }
      if (start[a]) {
        for (int c = 1; c <= nCorner; c++) for (int c1 = 1; c1 <= nCorner; c1++) Tvv[c1][c] = Pvv[c1][c];
      } else {
        for (int c = 1; c <= nCorner; c++)
          for (int c1 = 1; c1 <= nCorner; c1++) {
            F2(TT, c1, c0 + c, mC) = F2(TT, c1, c0 + c, mC) + quadwt * Pvv[c1][c];
            Tvv[c1][c] = quadTauW1[a] * Pvv[c1][c] - quadTauW2[a] * Tvv[c1][c];
          }
      }
    }
  }
}
