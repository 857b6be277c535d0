/*
 * umt_sweep.h — C ABI of libumtsweep.so, the B200-native Sn sweep library.
 *
 * Drop-in seam: Teton's sweep dispatch rt/ControlSweep.F90:55-73
 * (`useCUDASweep .and. .not. useGPU` => SetSweep_CUDA + getPhiTotal).  The
 * reference's own C seam there is gpu_sweepucbxyz / gpu_streamsynchronize /
 * gpu_devicesynchronize (gpu/GPU_SweepUCBxyz.cu:520-572, Fortran interface
 * gpu/SweepUCBxyzToGPU.F90:30-91), a per-(set,angle) call that re-uploads every
 * array.  This header replaces it with a context API whose state lives on the
 * device; `umt_sweep` is one whole SetSweep + getPhiTotal.
 *
 * Conventions (same as Teton's Fortran callers, the BIND(C) routines in aux/):
 *   - arrays are flat, column-major images of the Fortran arrays named in the
 *     comments, group index fastest; integer ids inside arrays are 1-based;
 *   - scalar arguments are passed by value here (the Fortran glue in
 *     umt_b200/fortran/teton_b200_mod.F90 uses VALUE);
 *   - every function returns 0 on success, non-zero on failure and never aborts
 *     (the reference aborts through f90fatal/MPI_Abort, misc/f90errors.F90:40-68;
 *     the Fortran glue maps non-zero to f90fatal).  umt_last_error() gives text.
 *   - a context is used by one host thread at a time; contexts are independent.
 */
#ifndef UMT_SWEEP_H
#define UMT_SWEEP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct umt_ctx umt_ctx;

enum { UMT_OK = 0, UMT_ERR_ARG = 1, UMT_ERR_CUDA = 2, UMT_ERR_STATE = 3, UMT_ERR_NCCL = 4, UMT_ERR_SCHEDULE = 5 };

/* ---- life cycle ------------------------------------------------------- */
/* Sizes as in mods/Size_mod.F90:19-80 (ndim, nzones, ncornr, nbelem, maxcf,
   maxCorner, ngr).  device = CUDA ordinal; fails (UMT_ERR_CUDA) without a GPU.
   device = -1 makes a host-only context that can build quadratures and sweep
   schedules (pure host work) but refuses every call that needs a kernel. */
int umt_ctx_create(int device, int ndim, int nzones, int ncornr, int nbelem,
                   int maxcf, int maxCorner, int ngr, umt_ctx **ctx);
int umt_ctx_destroy(umt_ctx *ctx);
const char *umt_last_error(const umt_ctx *ctx); /* ctx may be NULL: last create error */
const char *umt_version(void);

/* Page-locked host staging memory on the NUMA node of the context's GPU, for the arrays that cross PCIe every sweep (GSet%Sigt,
   GSet%STotal, Rad%PhiTotal: rt/ControlSweep.F90 hands them over per call).  *numaNode (may be NULL) returns the node the pages were
   bound to, -1 if the platform gave no NUMA information (plain page-locked memory then).  Free with umt_host_free. */
int umt_host_alloc(umt_ctx *ctx, size_t bytes, void **ptr, int *numaNode);
int umt_host_free(umt_ctx *ctx, void *ptr);

/* ---- mesh connectivity: mods/Geometry_mod.F90:19-78, aux/setTetonZone.F90 ---- */
/* numCorner(nz), cOffSet(nz), nCFacesArray(nc), cFP(maxcf,nc), cEZ(maxcf,nc).
   The remaining arrays are only needed by umt_build_schedule / umt_compute_geometry
   and may be NULL otherwise: zoneFaces(nz), zoneOpp(maxFaces,nz), faceOpp(maxFaces,nz),
   CToFace(maxcf,nc), BoundaryZone(nz) (bytes), BdyToC(nbelem). */
int umt_set_connectivity(umt_ctx *ctx, const int *numCorner, const int *cOffSet,
                         const int *nCFacesArray, const int *cFP, const int *cEZ,
                         int maxFaces, const int *zoneFaces, const int *zoneOpp,
                         const int *faceOpp, const int *CToFace,
                         const unsigned char *BoundaryZone, const int *BdyToC);

/* ---- geometry: rt/geometryUCBxyz.F90, rt/geometryUCBrz.F90 ------------- */
/* Upload Teton's own arrays: Volume(nc), A_fp(ndim,maxcf,nc), A_ez(ndim,maxcf,nc);
   RZ only: Area(nc), RadiusFP(2,nc), RadiusEZ(2,nc); A_bdy(ndim,nbelem) optional
   (needed by umt_build_schedule's exit lists and the exchange tallies). */
int umt_set_geometry(umt_ctx *ctx, const double *Volume, const double *A_fp, const double *A_ez,
                     const double *Area, const double *RadiusFP, const double *RadiusEZ,
                     const double *A_bdy);
/* ...or compute them on the device from corner coordinates px(ndim,nc)
   (replaces getGeometry, control/initializeSets.F90:85). */
int umt_compute_geometry(umt_ctx *ctx, const double *px);
int umt_download_geometry(umt_ctx *ctx, double *Volume, double *A_fp, double *A_ez,
                          double *Area, double *RadiusFP, double *RadiusEZ, double *A_bdy,
                          double *VolumeZone);

/* ---- quadrature: mods/AngleSet_mod.F90:34-125, rt/rtquad.F90 ------------ */
/* omega(ndim,NA), weight(NA); RZ only: StartingDirection(NA), FinishingDirection(NA)
   (bytes), angDerivFac/quadTauW1/quadTauW2(NA). */
int umt_set_quadrature(umt_ctx *ctx, int nAngles, const double *omega, const double *weight,
                       const unsigned char *startingDirection, const unsigned char *finishingDirection,
                       const double *angDerivFac, const double *quadTauW1, const double *quadTauW2);
/* Build the product quadrature (rt/quadProduct.F90 | rt/quadrz.F90 product branch +
   rt/rtquad.F90 normalisation + rt/AngleCoef2D.F90) and install it.  Returns the
   number of angles through *nAngles (3-D: 8*P*A, RZ: 4*P*(A+1)). */
int umt_build_product_quadrature(umt_ctx *ctx, int npolar, int nazimuthal, int polaraxis, int *nAngles);
int umt_get_quadrature(umt_ctx *ctx, double *omega, double *weight);

/* ---- sweep schedule: snac/snnext.F90, mods/AngleSet_mod.F90 HypPlane/BdyExit ---- */
/* Install Teton's schedule for one angle (1-based `angle`): nHyperPlanes,
   zonesInPlane(nHyp), nextZ(nz) signed, nextC(nc), numCycles + cycleList(numCycles)
   (global corner ids), exit list bdyList(2,nxBdy) (RZ sweeps and radiation-field init). */
int umt_set_schedule(umt_ctx *ctx, int angle, int nHyperPlanes, const int *zonesInPlane,
                     const int *nextZ, const int *nextC, int numCycles, const int *cycleList,
                     int nxBdy, const int *bdyList);
/* ...or let the library do rtorder/snnext/findexit on the host for every angle. */
int umt_build_schedule(umt_ctx *ctx);
int umt_get_schedule_info(umt_ctx *ctx, int angle, int *nHyperPlanes, int *numCycles, int *nBadZones);
int umt_get_schedule(umt_ctx *ctx, int angle, int *zonesInPlane, int *nextZ, int *nextC, int *cycleList);

/* ---- state: mods/SetData_mod.F90:13-73, mods/GroupSet_mod.F90:16-27 ------ */
/* Whole-problem upload: Psi(ngr,nc,NA), PsiB(ngr,nb,NA), Sigt(ngr,nz), STotal(ngr,nc).
   Any pointer may be NULL to leave that array untouched. */
int umt_upload_state(umt_ctx *ctx, const double *Psi, const double *PsiB, const double *Sigt,
                     const double *STotal, double tau);
/* Per phase-space-set upload/download (Set%Psi(Groups,nc,NumAngles), Set%PsiB(Groups,nb,NumAngles)
   with Set%g0, Set%angle0 0-based offsets). */
int umt_upload_set(umt_ctx *ctx, int g0, int Groups, int angle0, int NumAngles,
                   const double *Psi, const double *PsiB);
int umt_download_set(umt_ctx *ctx, int g0, int Groups, int angle0, int NumAngles,
                     double *Psi, double *PsiB);
int umt_download_psi(umt_ctx *ctx, double *Psi);
int umt_download_psib(umt_ctx *ctx, double *PsiB);
int umt_download_phi(umt_ctx *ctx, double *PhiTotal); /* Rad%PhiTotal(ngr,nc) */

/* Planck group integrals B_g(T) (misc/NormalizedBlackBody.cc:151-186, Clark 1987); pure host. */
int umt_planck_groups(double T, double k, double Bnorm, int numGroups, const double *groupBounds, double *B);
/* aux/InitTeton.F90:82-118: Psi(g,c,a) = max(wtiso*B_g(Trz(zone)), efloor) for every angle,
   built on the device from the zone radiation temperatures Trz(nz) and group bounds gnu(ngr+1). */
int umt_init_teton(umt_ctx *ctx, const double *Trz, const double *groupBounds, double speedLight,
                   double radConstant, double wtiso, double efloor);

/* control/initPhiTotal_OMPOL.F90 (Psi *= VolumeOld/Volume when volRatio != NULL, PhiTotal = sum w Psi)
   and control/initializeRadiationField_OMPOL.F90:116-143 (exit PsiB <- Psi, cyclePsi <- Psi). */
int umt_set_boundary_sources(umt_ctx *ctx);   /* control/setBoundarySources.F90:42 without source profiles: PsiB = 0 */
int umt_init_phi_total(umt_ctx *ctx, const double *volRatio);
int umt_init_radiation_field(umt_ctx *ctx);
/* initCyclePsi alone (control/constructDynMemory.F90:56-109): Set%cyclePsi(:,m) <- Psi(:,c,angle) for the corners on the cycle
   lists.  A caller that uploads Psi and installs new schedules every cycle (umt_b200/fortran/SetSweep_B200.F90) calls it after the
   upload, as the reference's initializeRadiationField does.  (The library also re-seeds by itself when a cycle list changes.) */
int umt_init_cycle_psi(umt_ctx *ctx);

/* ---- device memory of the angular flux -------------------------------------- */
/* The reference holds Set%Psi once plus a one-angle scratch Set%Psi1 (mods/SetData_mod.F90:157,164).  So does this library on 3-D
   meshes without cycle lists, direct-solve zones, reflecting boundaries or staged comm sets ("single-psi" layout): a savePsi sweep
   writes Psi in place, the other sweeps keep Psi1 in a ring of angle batches whose PhiTotal contribution is tallied, in fixed angle
   order, as the batches retire.  umt_set_psi1_ring chooses the ring size in angle batches (0 = as many as the free device memory
   holds; a ring that holds every batch needs no in-sweep tally).  Every other problem keeps a full second buffer.
   umt_get_psi_layout: info6 = {1 if single-psi, Psi1 slabs allocated, angles per batch, ring size in batches, batches,
   angles tallied inside the sweep kernel}; bytes = device memory of Psi plus the Psi1 workspace. */
int umt_set_psi1_ring(umt_ctx *ctx, int nBatches);
int umt_get_psi_layout(umt_ctx *ctx, int *info6, double *bytes);

/* ---- the hot path: rt/ControlSweep.F90:15-81 = SetSweep + getPhiTotal ---- */
/* maxFluxIters / fluxTol: incidentFlux iteration control (SetSweep.F90:81-207,
   rt/testFluxConv.F90:55-59).  *itersDone returns the number of flux passes. */
int umt_sweep(umt_ctx *ctx, int savePsi, int maxFluxIters, double fluxTol, int *itersDone);
/* Device time of the kernels of the last umt_sweep, in ms: [0] sweep kernel(s),
   [1] psi->phi reduction, [2] exchange (pack/NCCL/unpack), [3] whole call. */
/* One rt/ControlSweep.F90 call with the host arrays the Fortran caller owns: GSet%Sigt (ngr,nzones) and GSet%STotal
   (ngr,ncornr) in (NULL: keep what is on the device), Rad%PhiTotal (ngr,ncornr) out; = umt_upload_state + umt_sweep +
   umt_download_phi with the phi reduction and its device-to-host copy overlapped.  Pinned host buffers recommended. */
int umt_control_sweep(umt_ctx *ctx, const double *Sigt, const double *STotal, double tau, int savePsi, int maxFluxIters,
                      double fluxTol, int *itersDone, double *PhiTotal);
/* One ControlSweep over the n group sets of a domain, one context per group set (the reference's phase-space sets split the groups
   the same way: every SetData has its g0 / Groups, mods/SetData_mod.F90:35-50, and GSet%Sigt / GSet%STotal are per group set,
   mods/GroupSet_mod.F90).  Sigt[k] (Groups_k, nzones), STotal[k] (Groups_k, ncornr) in, PhiTotal[k] (Groups_k, ncornr) out, all host
   arrays (page-locked recommended).  The sets are pipelined: the upload of set k+1 and the download of set k-1 run while set k is
   swept, so host<->device traffic in both directions hides behind the sweeps.  Same results as n umt_control_sweep calls (groups do
   not couple inside a sweep).  Contexts with shared boundaries are swept one after the other (each with its own exchange). */
int umt_control_sweep_sets(umt_ctx *const *ctxs, int n, const double *const *Sigt, const double *const *STotal, double tau, int savePsi,
                           int maxFluxIters, double fluxTol, int *itersDone, double *const *PhiTotal);
int umt_last_sweep_times(umt_ctx *ctx, double *ms4);
int umt_last_sweep_launches(umt_ctx *ctx, int *nLaunches);
int umt_synchronize(umt_ctx *ctx);

/* ---- reflecting boundaries: snac/snreflect.F90, rt/findReflectedAngles.F90 ---- */
/* One call per reflecting boundary (one plane each, as the reference requires): before an angle incident on it is swept,
   PsiB(:,b,Minc) <- PsiB(:,b,Mref) for its mirror angle Mref (snac/reflectAxis.F90, axis-aligned planes).  Angles are swept in
   stages so that Mref precedes Minc. */
int umt_add_reflecting_boundary(umt_ctx *ctx, int firstBdyElem, int nBdyElem);
int umt_get_reflected_angles(umt_ctx *ctx, int reflIndex, int *mref /* (NA) 1-based, -1 = not incident */);
int umt_get_reflect_stages(umt_ctx *ctx, int *stageOf /* (NA) */);

/* ---- domain decomposition: rt/findexit.F90:102-294, rt/SendFlux.F90, rt/RecvFlux.F90 ---- */
/* One call per shared boundary (neighbour): its boundary elements are
   firstBdyElem..firstBdyElem+nBdyElem-1 (1-based), matched element-by-element with
   the neighbour's list (aux/checkSharedBoundary.F90).  Shared boundaries are indexed
   0,1,... in the order they were added (sharedIndex below). */
int umt_add_shared_boundary(umt_ctx *ctx, int neighborRank, int firstBdyElem, int nBdyElem);
/* myRank/nRanks and a 128-byte ncclUniqueId (same on all ranks; rank 0 gets it from
   umt_nccl_unique_id).  NCCL is dlopen'ed; fails with UMT_ERR_NCCL if absent.  The psib
   rows then travel GPU-to-GPU with ncclSend/ncclRecv (replaces the persistent MPI
   requests of rt/initcomm.F90:88-101). */
/* (Between the domains of one process, and between ranks when UMT_EXCHANGE_PUT=1 is set, the rows travel over peer memory instead:
   umt_build_exchange opens the neighbours' receive buffers (CUDA IPC between ranks) and the pack kernel stores into them directly,
   over NVLink when the neighbour is another GPU; NCCL then only carries the small per-pass messages.) */
int umt_nccl_unique_id(unsigned char *id128);
int umt_set_comm(umt_ctx *ctx, int myRank, int nRanks, const unsigned char *id128);
/* In-process alternative: the n contexts become ranks 0..n-1 of one group (several domains
   per process, e.g. on one GPU); collective calls (umt_build_exchange, umt_sweep) must then be
   made from n host threads, one per context. */
int umt_connect_local(umt_ctx **ctxs, int n);
/* Rank only (host-only contexts that build exchange lists without a communicator). */
int umt_set_rank(umt_ctx *ctx, int myRank, int nRanks);
/* rt/findexit.F90:128-205: each side of a shared boundary classifies half of the angles of an
   angle set (lower rank the first half) as incident (-1) / exiting (+1) by the sign of
   omega . A_bdy and the two halves are exchanged.  incTest is (nBdyElem, NA) bytes, zero for
   the angles the other side decides.  umt_build_exchange trades them through the communicator;
   a caller with its own transport (MPI, gloo) passes the neighbour's array in instead. */
int umt_get_incident_test(umt_ctx *ctx, int sharedIndex, signed char *incTest);
int umt_set_incident_test(umt_ctx *ctx, int sharedIndex, const signed char *incTestNeighbor);
/* ListSend / ListRecv of every (shared boundary, angle) (findexit.F90:193-287) + device buffers. Collective. */
int umt_build_exchange(umt_ctx *ctx);
int umt_get_exchange_counts(umt_ctx *ctx, int sharedIndex, int *nSend /* (NA) */, int *nRecv /* (NA) */);
int umt_get_exchange_lists(umt_ctx *ctx, int sharedIndex, int angle, int *listSend, int *listRecv);
/* CSet%IncFlux / IncFluxOld per comm set after the last umt_sweep (rt/setIncidentFlux.F90:128-146):
   one bin per angle in 3-D, per xi-level in 2-D. */
/* rt/SweepScheduler.F90 + rt/setNetFlux.F90.  The scheduler orders angle BINS: angles in 3-D, xi-levels in r-z (SweepScheduler.F90:110-117).
   By default every bin is its own comm set: all are swept concurrently, the exchange is lagged one flux pass and the scheduler is the
   identity.  umt_set_comm_sets groups the bins into nCommSets consecutive comm sets (the reference's nSets < maxAngleSets,
   aux/ConstructPhaseSpaceSets.F90:247-296); umt_sweep_scheduler (collective over the domains, once per cycle like
   control/initializeSets.F90:515-523) then fixes CSet%AngleOrder / RecvOrder from the net flux on the shared boundaries
   (netFlux: rows of NA doubles per shared boundary, entry b = exiting - incident current of bin b -- the first nLevels entries of a row in
   r-z; NULL: tallied from the PsiB on the device) and the mirror dependencies, and umt_sweep runs the comm sets' steps one after the other
   with SendFlux/RecvFlux per step (snac/SetSweep.F90:113-170), so a neighbour that sweeps a bin later in the pass receives this pass's
   flux.  umt_get_angle_order returns the angles (a level's angles in level order in r-z).  r-z comm sets of several levels are not
   combined with reflecting boundaries. */
int umt_set_comm_sets(umt_ctx *ctx, int nCommSets);
int umt_sweep_scheduler(umt_ctx *ctx, const double *netFlux);
int umt_get_net_flux(umt_ctx *ctx, double *netFlux /* (nShared, NA): CSet%NetFlux of the last umt_sweep_scheduler */);
int umt_get_angle_order(umt_ctx *ctx, int *angleOrder /* (NA) 1-based, comm sets concatenated */, int *recvOrder /* (nShared, NA) or NULL */);
int umt_get_incident_flux(umt_ctx *ctx, double *incFlux, double *incFluxOld);
/* adqtEpsilon*speed_light*rad_constant*tr4floor of rt/testFluxConv.F90:73 (default 0). */
int umt_set_flux_floor(umt_ctx *ctx, double floorFlux);

/* ---- end-of-cycle edits on the device-resident fields: aux/rtedit.F90:142-232, control/BoundaryEdit.F90, setEnergyDensity.F90 ---- */
/* out5 = {EnergyRadiation, TrMax, PowerEscape, PowerIncident (0 without source boundaries), sum_zones sum_c V_c sum_g PhiTotal};
   optional (NULL to skip): Mat%trz(nzones), RadEdit%RadPowerEscape(ngr), Rad%RadEnergyDensity(nzones, ngr).  3-D and r-z (the r-z
   boundary edit carries geometryFactor = 2 pi and the radius of the boundary element, control/BoundaryEdit.F90:123).  Escape currents use Set%Psi at the boundary corners as the reference does (call after the savePsi sweep). */
int umt_cycle_edits(umt_ctx *ctx, double speedLight, double radConstant, double tr4floor, double *out5, double *trz,
                    double *RadPowerEscape, double *RadEnergyDensity);

/* ---- scattering + emission source build (extension; the mini-app reference never fills GSet%STotal, mods/GroupSet_mod.F90:74-77) ---- */
/* STotal(g,c) = wtiso [ sigs(g,z) PhiTotal(g,c) + Chi(g,c) Eta(c) sum_g' siga(g',z) PhiTotal(g',c) + EmissionRate(g,c) ] from the
   device-resident PhiTotal into the device-resident GSet%STotal; EmissionRate (Mat%EmissionRate(ngr,ncornr)) and STotalOut may be NULL.
   Consistent with rt/getCollisionRate.F90:60-75 and the Chi redistribution of rt/addGreyCorrections.F90:85-86.  Parity unpinned. */
int umt_build_source(umt_ctx *ctx, const double *Siga, const double *Sigs, const double *Eta, const double *Chi,
                     const double *EmissionRate, double *STotalOut);

/* ---- grey transport acceleration (3-D, "new" GTA solver): rt/GTASolver.F90, snac/GTASweep.F90 ---- */
/* GTA angle set (level-symmetric S2, 8 ordinates: rt/quadxyz.F90), its sweep order (rtorder/snnext) and device arrays.
   Needs full connectivity and geometry. */
int umt_gta_setup(umt_ctx *ctx);
int umt_gta_get_quadrature(umt_ctx *ctx, double *omega /* (ndim,8) */, double *weight /* (8) */);
/* GTA%GreySigTotal, GreySigScat, GreySigScatVol (ncornr) from the caller (GreySigtInv = 1/GreySigTotal) ... */
int umt_gta_set_opacity(umt_ctx *ctx, const double *GreySigTotal, const double *GreySigScat,
                        const double *GreySigScatVol);
/* ... or rt/setGTAOpacity.F90:10-113 (setGTAOpacityNEW) on the device from Mat%Siga, Mat%Sigs (ngr,nzones),
   Mat%Eta (ncornr) and GTA%Chi (ngr,ncornr); Chi is returned rescaled and kept on the device.  Uses tau of umt_upload_state. */
int umt_gta_compute_opacity(umt_ctx *ctx, const double *Siga, const double *Sigs, const double *Eta, double *Chi);
int umt_gta_get_opacity(umt_ctx *ctx, double *GreySigTotal, double *GreySigScat, double *GreySigScatVol, double *GreySigtInv);
/* rt/getCollisionRate.F90:10-97 on the device-resident PhiTotal: GTA%GreySource(c) = sum_g (Eta siga + sigs) PhiTotal
   (residualFlag 1: minus its previous value).  GreySource may be NULL (result stays on the device). */
int umt_collision_rate(umt_ctx *ctx, const double *Eta, const double *Siga, const double *Sigs, int residualFlag,
                       double *GreySource);
int umt_gta_set_source(umt_ctx *ctx, const double *GreySource);
/* snac/InitSweepGreyUCBxyz.F90: within-zone transfer matrices GTA%TT(maxCorner, ncornr); TT may be NULL. */
int umt_gta_init_tt(umt_ctx *ctx, double *TT);
/* snac/GTASweep.F90 (GTA%ID = 1) + snac/SweepGreyUCBxyz.F90 KernelNew for all 8 angles: TsaSource = wtiso (GreySigScat P +
   GreySource); GreySource NULL keeps the device copy, withSource 0 sweeps with GreySource = 0.
   PsiB_gta (nbelem, 8) in/out (may be NULL = 0), PhiInc (ncornr) out. */
int umt_gta_sweep(umt_ctx *ctx, const double *P, const double *GreySource, double *PsiB_gta,
                  double *PhiInc, int withSource);
/* rt/GreySweep.F90:12-48 (GreySweepNEW): sweep + snac/UpdateScalarIntensity.F90 per-zone LU/solve; P, PsiB_gta in/out.
   The withSource call decomposes TT in place (once per umt_gta_init_tt). */
int umt_gta_grey_sweep(umt_ctx *ctx, double *P, double *PsiB_gta, int withSource);
/* rt/GTASolver.F90:42-425: BiCGSTAB for the grey corrections from the device-resident PhiTotal and GreySource
   (epsPoint/maxIters = the "grey" iteration control, epsGrey = GTA%epsGrey, enforceHardMax = GTA%enforceHardGTAIterMax). */
int umt_gta_solve(umt_ctx *ctx, double epsPoint, int maxIters, double epsGrey, int enforceHardMax, int *nGreyIter,
                  double *maxRelErrGrey);
int umt_gta_get_correction(umt_ctx *ctx, double *GreyCorrection);
/* rt/addGreyCorrections.F90:70-91: PhiTotal(g,c) += GreyCorrection(c) Chi(g,c) on the device. */
int umt_add_grey_corrections(umt_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* UMT_SWEEP_H */
