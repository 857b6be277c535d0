/*
 * teton_gpu_compat.h — the reference's OWN C seam for the CUDA sweep, exported by libumtsweep.so with the
 * reference's exact signatures so that Teton's unchanged Fortran (gpu/SweepUCBxyzToGPU.F90:30-91 interface
 * blocks, call :198-240; gpu/SetSweep_CUDA.F90:125-189) links against this library instead of the object
 * built from gpu/GPU_SweepUCBxyz.cu.
 *
 *   gpu_sweepucbxyz        replaces gpu/GPU_SweepUCBxyz.cu:532-994
 *   gpu_streamsynchronize  replaces gpu/GPU_SweepUCBxyz.cu:520-523
 *   gpu_devicesynchronize  replaces gpu/GPU_SweepUCBxyz.cu:525-528
 *
 * Conventions are the reference's: every argument by reference (Fortran), host arrays in Teton's layout,
 * 1-based ids, void return, and any failure prints a message and ends the process (the reference's
 * CUDA_SAFE_CALL, GPU_SweepUCBxyz.cu:18-27).  One call sweeps ONE angle of one phase-space set:
 *
 *   in : schedule of the angle (nHyperPlanes, nZonesInPlane, nextZ, nextC), STotal(G,nc), tau, Psi(G,nc) [psi^n of
 *        the angle], Sigt(G,nz), geometry (Volume, A_fp, A_ez), connectivity (nCFacesArray, cFP, cEZ, Geom_numCorner,
 *        Geom_cOffSet), omega(3), quadwt, Psi1(G,nc+nb) [previous content: read by zones with an intra-zone cycle],
 *        Phi(G,nc), PsiB(G,nb) of the angle, cyclePsi/cycleList (whole arrays) with this angle's cycleOffSet and
 *        numCycles, and for one reflecting boundary b0 (0-based first element), nBdyElem, PsiBMref = PsiB(:,:,Mref);
 *   out: Psi1 (corner rows), PsiB (exiting rows; the reflected rows too), Phi += quadwt * Psi1, cyclePsi rows of the
 *        angle, and Psi <- Psi1 when *savePsi == 1.
 *
 * Unlike the reference the call is synchronous (results are in the host arrays on return), so the two
 * synchronize entry points only drain the device.  It forwards to the context API of umt_sweep.h with a
 * one-angle context per stream id; like the reference it moves every array across PCIe on every call, which is
 * why the fast path is umt_sweep / umt_control_sweep (state resident in HBM, all angles in one launch) and this
 * symbol exists for link compatibility.  mem0solve1, totalStreams, NumAngles, Angle and Mref carry no
 * One deliberate difference: for a reflecting boundary that does not start at element 1 the reference offsets its
 * device copy by b0 DOUBLES (`d_PsiB + *b0`, `PsiBMref + *b0`, GPU_SweepUCBxyz.cu:802-803) although b0 counts boundary
 * elements (SweepUCBxyzToGPU.F90:159); this library copies rows b0 .. b0+nBdyElem-1 as snac/snreflect.F90:62-70 does.
 * mem0solve1, totalStreams, NumAngles, Angle and Mref carry no
 * information the sweep needs and are ignored (the reference ignores them too, apart from a one-time memset).
 */
#ifndef TETON_GPU_COMPAT_H
#define TETON_GPU_COMPAT_H

#ifdef __cplusplus
extern "C" {
#endif

void gpu_sweepucbxyz(int *Angle, int *nHyperPlanes, int *nZonesInPlane, int *nextZ, int *nextC, double *STotal,
                     double *tau, double *Psi, int *Groups, double *Volume, double *Sigt, int *nCFacesArray, int *ndim,
                     int *maxcf, int *ncorner, double *A_fp, double *omega, int *cFP, double *Psi1, int *nbelem,
                     double *A_ez, int *cEZ, int *NumAngles, double *quadwt, double *Phi, double *PsiB, int *maxCorner,
                     int *mem0solve1, int *streamIdPtr, int *totalStreams, int *savePsi, int *numCycles, int *cycleOffSet,
                     double *cyclePsi, int *cycleList, int *b0, int *nBdyElem, double *PsiBMref, int *Mref,
                     int *Geom_numCorner, int *Geom_cOffSet);
void gpu_streamsynchronize(int *streamId);
void gpu_devicesynchronize(void);

#ifdef __cplusplus
}
#endif
#endif
