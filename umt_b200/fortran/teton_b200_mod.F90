!***********************************************************************
!  teton_b200_mod - ISO_C_BINDING interfaces to libumtsweep.so          *
!  (include/umt_sweep.h).  Add this file to teton/gpu/ and link         *
!  libumtsweep.so; see INTEGRATION.md.  Scalars are passed by VALUE,    *
!  arrays as assumed-size contiguous C_DOUBLE / C_INT (the Fortran      *
!  memory image is exactly what the library expects: group index        *
!  fastest, 1-based ids inside the arrays).                             *
!                                                                       *
!  NOT COMPILED in the umt_b200 repository's own CI: the build image    *
!  has no Fortran compiler (SURVEY.md section 0 fact 1).                *
!***********************************************************************
module teton_b200_mod

   use, intrinsic :: iso_c_binding
   implicit none

   type(C_PTR), save :: b200_ctx = C_NULL_PTR
   logical,     save :: b200_static_uploaded = .FALSE.     ! Psi/PsiB of this cycle are on the device
   ! geometry, quadrature and sweep schedules are on the device.  The host code clears it whenever it moves the mesh or
   ! changes the quadrature (the test driver runs with options/mesh_motion = 0, TetonConduitInterface.cc:603-612: it never does),
   ! so the schedules, the plan records built from them (0.16 s + 0.41 s at -d 20 -G 128) and the geometry are built once per run
   logical,     save :: b200_mesh_uploaded = .FALSE.

   interface

      integer(C_INT) function umt_ctx_create(device, ndim, nzones, ncornr, nbelem, maxcf, maxCorner, ngr, ctx) &
                              bind(C, name="umt_ctx_create")
         import :: C_INT, C_PTR
         integer(C_INT), value :: device, ndim, nzones, ncornr, nbelem, maxcf, maxCorner, ngr
         type(C_PTR)           :: ctx
      end function

      integer(C_INT) function umt_ctx_destroy(ctx) bind(C, name="umt_ctx_destroy")
         import :: C_INT, C_PTR
         type(C_PTR), value :: ctx
      end function

      function umt_last_error(ctx) result(msg) bind(C, name="umt_last_error")
         import :: C_PTR
         type(C_PTR), value :: ctx
         type(C_PTR)        :: msg
      end function

      integer(C_INT) function umt_set_connectivity(ctx, numCorner, cOffSet, nCFacesArray, cFP, cEZ, maxFaces, &
                              zoneFaces, zoneOpp, faceOpp, CToFace, BoundaryZone, BdyToC) &
                              bind(C, name="umt_set_connectivity")
         import :: C_INT, C_PTR, C_BOOL
         type(C_PTR),    value :: ctx
         integer(C_INT)        :: numCorner(*), cOffSet(*), nCFacesArray(*), cFP(*), cEZ(*)
         integer(C_INT), value :: maxFaces
         type(C_PTR),    value :: zoneFaces, zoneOpp, faceOpp, CToFace, BoundaryZone, BdyToC   ! C_NULL_PTR unless umt_build_schedule is used
      end function

      integer(C_INT) function umt_set_geometry(ctx, Volume, A_fp, A_ez, Area, RadiusFP, RadiusEZ, A_bdy) &
                              bind(C, name="umt_set_geometry")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR), value :: ctx
         real(C_DOUBLE)     :: Volume(*), A_fp(*), A_ez(*)
         type(C_PTR), value :: Area, RadiusFP, RadiusEZ, A_bdy      ! RZ only / optional
      end function

      integer(C_INT) function umt_set_quadrature(ctx, nAngles, omega, weight, StartingDirection, FinishingDirection, &
                              angDerivFac, quadTauW1, quadTauW2) bind(C, name="umt_set_quadrature")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: nAngles
         real(C_DOUBLE)        :: omega(*), weight(*)
         type(C_PTR),    value :: StartingDirection, FinishingDirection, angDerivFac, quadTauW1, quadTauW2
      end function

      integer(C_INT) function umt_set_schedule(ctx, angle, nHyperPlanes, zonesInPlane, nextZ, nextC, numCycles, &
                              cycleList, nxBdy, bdyList) bind(C, name="umt_set_schedule")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: angle, nHyperPlanes, numCycles, nxBdy
         integer(C_INT)        :: zonesInPlane(*), nextZ(*), nextC(*), cycleList(*), bdyList(*)
      end function

      integer(C_INT) function umt_upload_state(ctx, Psi, PsiB, Sigt, STotal, tau) bind(C, name="umt_upload_state")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         type(C_PTR),    value :: Psi, PsiB, Sigt, STotal     ! C_LOC of the arrays, C_NULL_PTR = leave untouched
         real(C_DOUBLE), value :: tau
      end function

      integer(C_INT) function umt_upload_set(ctx, g0, Groups, angle0, NumAngles, Psi, PsiB) bind(C, name="umt_upload_set")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: g0, Groups, angle0, NumAngles
         type(C_PTR),    value :: Psi, PsiB
      end function

      integer(C_INT) function umt_init_cycle_psi(ctx) bind(C, name="umt_init_cycle_psi")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
      end function

      integer(C_INT) function umt_set_psi1_ring(ctx, nBatches) bind(C, name="umt_set_psi1_ring")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: nBatches     ! angle batches the Psi1 ring holds (0: as the free device memory allows)
      end function

      integer(C_INT) function umt_download_set(ctx, g0, Groups, angle0, NumAngles, Psi, PsiB) bind(C, name="umt_download_set")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: g0, Groups, angle0, NumAngles
         type(C_PTR),    value :: Psi, PsiB
      end function

      integer(C_INT) function umt_sweep(ctx, savePsi, maxFluxIters, fluxTol, itersDone) bind(C, name="umt_sweep")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: savePsi, maxFluxIters
         real(C_DOUBLE), value :: fluxTol
         integer(C_INT)        :: itersDone
      end function

!     one whole ControlSweep with the caller's host arrays (upload, SetSweep, getPhiTotal, download overlapped)
      integer(C_INT) function umt_control_sweep(ctx, Sigt, STotal, tau, savePsi, maxFluxIters, fluxTol, itersDone, PhiTotal) &
                              bind(C, name="umt_control_sweep")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         real(C_DOUBLE)        :: Sigt(*), STotal(*), PhiTotal(*)
         real(C_DOUBLE), value :: tau, fluxTol
         integer(C_INT), value :: savePsi, maxFluxIters
         integer(C_INT)        :: itersDone
      end function

!     the same over the group sets of the domain, one context per group set: ctxs(n), Sigt(n), STotal(n), PhiTotal(n) are arrays of
!     C pointers (c_loc of GSet%Sigt, GSet%STotal of group set k and of the PhiTotal block of its groups); the sets are pipelined
      integer(C_INT) function umt_control_sweep_sets(ctxs, n, Sigt, STotal, tau, savePsi, maxFluxIters, fluxTol, itersDone, PhiTotal) &
                              bind(C, name="umt_control_sweep_sets")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR)           :: ctxs(*), Sigt(*), STotal(*), PhiTotal(*)
         integer(C_INT), value :: n
         real(C_DOUBLE), value :: tau, fluxTol
         integer(C_INT), value :: savePsi, maxFluxIters
         integer(C_INT)        :: itersDone
      end function

      integer(C_INT) function umt_download_phi(ctx, PhiTotal) bind(C, name="umt_download_phi")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR), value :: ctx
         real(C_DOUBLE)     :: PhiTotal(*)
      end function

      integer(C_INT) function umt_add_shared_boundary(ctx, neighborRank, firstBdyElem, nBdyElem) &
                              bind(C, name="umt_add_shared_boundary")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: neighborRank, firstBdyElem, nBdyElem
      end function

      integer(C_INT) function umt_nccl_unique_id(id128) bind(C, name="umt_nccl_unique_id")
         import :: C_INT, C_SIGNED_CHAR
         integer(C_SIGNED_CHAR) :: id128(128)
      end function

      integer(C_INT) function umt_set_comm(ctx, myRank, nRanks, id128) bind(C, name="umt_set_comm")
         import :: C_INT, C_PTR, C_SIGNED_CHAR
         type(C_PTR),    value  :: ctx
         integer(C_INT), value  :: myRank, nRanks
         integer(C_SIGNED_CHAR) :: id128(128)
      end function

!     SweepScheduler / setNetFlux for comm sets of several angle bins (rt/SweepScheduler.F90): optional, the default is one
!     comm set per angle set.  Called once per cycle where control/initializeSets.F90:515-523 calls SweepScheduler.
      integer(C_INT) function umt_set_comm_sets(ctx, nCommSets) bind(C, name="umt_set_comm_sets")
         import :: C_INT, C_PTR
         type(C_PTR),    value :: ctx
         integer(C_INT), value :: nCommSets
      end function

      integer(C_INT) function umt_sweep_scheduler(ctx, netFlux) bind(C, name="umt_sweep_scheduler")
         import :: C_INT, C_PTR
         type(C_PTR), value :: ctx
         type(C_PTR), value :: netFlux      ! C_NULL_PTR: tallied from the PsiB on the device
      end function

!     Grey transport acceleration (rt/GTASolver.F90 and friends): replaces the body of GTASolver + addGreyCorrections.
      integer(C_INT) function umt_gta_setup(ctx) bind(C, name="umt_gta_setup")
         import :: C_INT, C_PTR
         type(C_PTR), value :: ctx
      end function

      integer(C_INT) function umt_gta_compute_opacity(ctx, Siga, Sigs, Eta, Chi) bind(C, name="umt_gta_compute_opacity")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR), value :: ctx
         real(C_DOUBLE)     :: Siga(*), Sigs(*), Eta(*), Chi(*)   ! Mat%Siga, Mat%Sigs (ngr,nzones), Mat%Eta (ncornr), GTA%Chi (ngr,ncornr)
      end function

      integer(C_INT) function umt_collision_rate(ctx, Eta, Siga, Sigs, residualFlag, GreySource) bind(C, name="umt_collision_rate")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         real(C_DOUBLE)        :: Eta(*), Siga(*), Sigs(*)
         integer(C_INT), value :: residualFlag
         type(C_PTR),    value :: GreySource                       ! C_NULL_PTR: stays on the device
      end function

      integer(C_INT) function umt_gta_solve(ctx, epsPoint, maxIters, epsGrey, enforceHardMax, nGreyIter, maxRelErr) &
                              bind(C, name="umt_gta_solve")
         import :: C_INT, C_PTR, C_DOUBLE
         type(C_PTR),    value :: ctx
         real(C_DOUBLE), value :: epsPoint, epsGrey
         integer(C_INT), value :: maxIters, enforceHardMax
         integer(C_INT)        :: nGreyIter
         real(C_DOUBLE)        :: maxRelErr
      end function

      integer(C_INT) function umt_add_grey_corrections(ctx) bind(C, name="umt_add_grey_corrections")
         import :: C_INT, C_PTR
         type(C_PTR), value :: ctx
      end function

   end interface

contains

!  Map a non-zero library status to Teton's abort path (misc/f90errors.F90:40-68).
   subroutine b200_check(rc, where)
      integer(C_INT),   intent(in) :: rc
      character(len=*), intent(in) :: where
      character(kind=C_CHAR), pointer :: cmsg(:)
      character(len=512) :: msg
      integer :: i
      if (rc == 0) return
      msg = ' '
      call c_f_pointer(umt_last_error(b200_ctx), cmsg, [512])
      do i = 1, 512
         if (cmsg(i) == C_NULL_CHAR) exit
         msg(i:i) = cmsg(i)
      enddo
      call f90fatal(where // ": " // trim(msg))
   end subroutine b200_check

end module teton_b200_mod
