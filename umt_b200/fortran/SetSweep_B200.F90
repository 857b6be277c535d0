!***********************************************************************
!  SetSweep_B200 - drop-in for SetSweep_CUDA (gpu/SetSweep_CUDA.F90)    *
!  at the dispatch point rt/ControlSweep.F90:59-63: one call sweeps     *
!  every angle and group of the domain on the B200 and leaves           *
!  Rad%PhiTotal filled, so ControlSweep must NOT call getPhiTotal       *
!  after it (see INTEGRATION.md for the two-line change).               *
!                                                                       *
!  Once per mesh (b200_mesh_uploaded; cleared by the host when it moves *
!  the mesh): geometry, quadrature and sweep schedules go to the device.*
!  Per cycle (first call after initializeSets): Psi/PsiB go up once.    *
!  Per call: Sigt, STotal, tau up; PhiTotal down; on savePsi also       *
!  Psi/PsiB down, because finalizeSets/rtedit read them on the host.    *
!                                                                       *
!  NOT COMPILED here: no Fortran compiler in the build image.           *
!***********************************************************************
subroutine SetSweep_B200(savePsi)

   use, intrinsic :: iso_c_binding
   use kind_mod
   use constant_mod
   use Size_mod
   use Geometry_mod
   use QuadratureList_mod
   use SetData_mod
   use AngleSet_mod
   use GroupSet_mod
   use RadIntensity_mod
   use BoundaryList_mod
   use Boundary_mod
   use iter_control_list_mod
   use iter_control_mod
   use teton_b200_mod

   implicit none

   logical (kind=1), intent(in) :: savePsi

   type(SetData),     pointer :: Set
   type(AngleSet),    pointer :: ASet
   type(GroupSet),    pointer :: GSet
   type(HypPlane),    pointer :: HypPlanePtr
   type(BdyExit),     pointer :: BdyExitPtr
   type(Boundary),    pointer :: BdyT
   type(IterControl), pointer :: incidentFluxControl, temperatureControl

   integer(C_INT) :: rc, iters, device
   integer        :: setID, nSets, angle, a, sharedID, nShared, maxIters
   real(adqt)     :: fluxTol

   nSets = getNumberOfSets(Quad)

   if (.not. c_associated(b200_ctx)) then
      device = 0      ! one rank per GPU: the launcher sets CUDA_VISIBLE_DEVICES
      rc = umt_ctx_create(device, Size%ndim, Size%nzones, Size%ncornr, Size%nbelem, Size%maxcf, &
                          Size%maxCorner, Size%ngr, b200_ctx)
      call b200_check(rc, "umt_ctx_create")
      rc = umt_set_connectivity(b200_ctx, Geom%numCorner, Geom%cOffSet, Geom%nCFacesArray, Geom%cFP, Geom%cEZ, &
                                int(Size%maxFaces, C_INT), C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, &
                                C_NULL_PTR, C_NULL_PTR)
      call b200_check(rc, "umt_set_connectivity")
      nShared = getNumberOfShared(RadBoundary)
      do sharedID = 1, nShared
         BdyT => getShared(RadBoundary, sharedID)
         rc = umt_add_shared_boundary(b200_ctx, getNeighborID(BdyT), getFirstBdyElement(BdyT), &
                                      getNumberOfBdyElements(BdyT))
         call b200_check(rc, "umt_add_shared_boundary")
      enddo
      ! the NCCL id is created on rank 0 and broadcast with MPI_Bcast by the caller (128 bytes), then
      ! rc = umt_set_comm(b200_ctx, Size%myRankInGroup, Size%nprocs, id128)
   endif

!  Once per mesh: geometry, quadrature, schedules.  initializeSets rebuilds the host copies every cycle
!  (initializeSets.F90:85-105), identically while the mesh does not move; the context keeps what it built from them
   if (.not. b200_mesh_uploaded) then
      rc = umt_set_geometry(b200_ctx, Geom%Volume, Geom%A_fp, Geom%A_ez, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR)
      call b200_check(rc, "umt_set_geometry")
      ! no reflecting boundaries: set s holds angle s (decomposeAngleSets.F90:280-285); gather omega/weight
      ! of all sets into quadrature order and install them once
      call b200_install_quadrature_and_schedules()
      b200_mesh_uploaded = .TRUE.
   endif

!  Once per cycle: the angular flux
   if (.not. b200_static_uploaded) then
      do setID = 1, nSets
         Set => getSetData(Quad, setID)
         rc = umt_upload_set(b200_ctx, Set%g0, Set%Groups, Set%angle0, Set%NumAngles, c_loc(Set%Psi), c_loc(Set%PsiB))
         call b200_check(rc, "umt_upload_set")
      enddo
      ! Set%cyclePsi <- Psi on the cycle-list corners, as initializeRadiationField does every cycle
      ! (control/constructDynMemory.F90:56-109); the schedules installed above define the lists
      rc = umt_init_cycle_psi(b200_ctx)
      call b200_check(rc, "umt_init_cycle_psi")
      b200_static_uploaded = .TRUE.     ! finalizeSets resets it at the end of the cycle
   endif

   GSet => getGroupSetData(Quad, 1)     ! one group set spanning ngr, or loop g0 blocks with umt_upload_state per block
   rc = umt_upload_state(b200_ctx, C_NULL_PTR, C_NULL_PTR, c_loc(GSet%Sigt), c_loc(GSet%STotal), Size%tau)
   call b200_check(rc, "umt_upload_state")

!  incident-flux iteration controls exactly as SetSweep.F90:55-59 and testFluxConv.F90:55-59
   incidentFluxControl => getIterationControl(IterControls, "incidentFlux")
   temperatureControl  => getIterationControl(IterControls, "temperature")
   maxIters = getMaxNumberOfIterations(incidentFluxControl)
   fluxTol  = max(getEpsilonPoint(incidentFluxControl), getGlobalError(temperatureControl)/twenty)
   fluxTol  = min(fluxTol, 0.01_adqt)

   rc = umt_sweep(b200_ctx, merge(1_C_INT, 0_C_INT, savePsi), int(maxIters, C_INT), fluxTol, iters)
   call b200_check(rc, "umt_sweep")

   rc = umt_download_phi(b200_ctx, Rad%PhiTotal)
   call b200_check(rc, "umt_download_phi")

   if (savePsi) then
      do setID = 1, nSets
         Set => getSetData(Quad, setID)
         rc = umt_download_set(b200_ctx, Set%g0, Set%Groups, Set%angle0, Set%NumAngles, c_loc(Set%Psi), c_loc(Set%PsiB))
         call b200_check(rc, "umt_download_set")
      enddo
   endif

   return

contains

   subroutine b200_install_quadrature_and_schedules()
      real(C_DOUBLE), allocatable :: omega(:,:), weight(:)
      integer :: NA, aSetID, nAngleSets, a0
      NA = 0
      nAngleSets = getNumberOfAngleSets(Quad)
      do aSetID = 1, nAngleSets
         ASet => getAngleSetData(Quad, aSetID)
         NA = NA + ASet%NumAngles
      enddo
      allocate(omega(Size%ndim, NA), weight(NA))
      do aSetID = 1, nAngleSets
         ASet => getAngleSetData(Quad, aSetID)
         a0 = ASet%angle0
         omega(:, a0+1:a0+ASet%NumAngles) = ASet%omega(:, 1:ASet%NumAngles)
         weight(a0+1:a0+ASet%NumAngles)   = ASet%weight(1:ASet%NumAngles)
      enddo
      rc = umt_set_quadrature(b200_ctx, int(NA, C_INT), omega, weight, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR, C_NULL_PTR)
      call b200_check(rc, "umt_set_quadrature")
      do aSetID = 1, nAngleSets
         ASet => getAngleSetData(Quad, aSetID)
         do angle = 1, ASet%NumAngles
            a = ASet%angle0 + angle
            HypPlanePtr => ASet%HypPlanePtr(angle)
            BdyExitPtr  => ASet%BdyExitPtr(angle)
            rc = umt_set_schedule(b200_ctx, int(a, C_INT), int(ASet%nHyperPlanes(angle), C_INT), HypPlanePtr%zonesInPlane, &
                                  ASet%nextZ(:, angle), ASet%nextC(:, angle), int(ASet%numCycles(angle), C_INT), &
                                  ASet%cycleList(ASet%cycleOffSet(angle)+1:), int(BdyExitPtr%nxBdy, C_INT), BdyExitPtr%bdyList)
            call b200_check(rc, "umt_set_schedule")
         enddo
      enddo
      deallocate(omega, weight)
   end subroutine b200_install_quadrature_and_schedules

end subroutine SetSweep_B200
