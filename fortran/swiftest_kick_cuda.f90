!! swiftest_kick_cuda.f90 -- replacement bodies for the loops of src/swiftest/swiftest_kick.f90,
!! src/swiftest/swiftest_drift.f90 and src/encounter/encounter_check.f90 when built with -DUSE_CUDA.
!! The module-procedure INTERFACES (swiftest_module.f90:940-991, 513-522; encounter_module.f90:110-161) are unchanged,
!! so WHM, RMVS, HELIO and SyMBA call the new path without modification.  (Uncompiled here: no Fortran compiler.)

! ---- in submodule(swiftest) s_swiftest_kick ----------------------------------------------------------------------
   module subroutine swiftest_kick_getacch_int_all_tri_rad_pl(npl, nplm, r, Gmass, radius, acc)
      use swiftest_cuda
      implicit none
      integer(I4B),                 intent(in)    :: npl, nplm
      real(DP),     dimension(:,:), intent(in)    :: r
      real(DP),     dimension(:),   intent(in)    :: Gmass, radius
      real(DP),     dimension(:,:), intent(inout) :: acc
      real(DP), dimension(:), allocatable, target :: rad
      rad = radius(1:npl)
      call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, npl, nplm, r, Gmass, c_loc(rad), acc), "tri_rad_pl")
   end subroutine

   module subroutine swiftest_kick_getacch_int_all_tri_norad_pl(npl, nplm, r, Gmass, acc)
      use swiftest_cuda
      implicit none
      integer(I4B),                 intent(in)    :: npl, nplm
      real(DP),     dimension(:,:), intent(in)    :: r
      real(DP),     dimension(:),   intent(in)    :: Gmass
      real(DP),     dimension(:,:), intent(inout) :: acc
      call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, npl, nplm, r, Gmass, c_null_ptr, acc), "tri_norad_pl")
   end subroutine

   module subroutine swiftest_kick_getacch_int_all_flat_rad_pl(npl, nplpl, k_plpl, r, Gmass, radius, acc)
      use swiftest_cuda
      implicit none
      integer(I4B),                 intent(in)    :: npl
      integer(I8B),                 intent(in)    :: nplpl
      integer(I4B), dimension(:,:), intent(in), target :: k_plpl
      real(DP),     dimension(:,:), intent(in)    :: r
      real(DP),     dimension(:),   intent(in)    :: Gmass, radius
      real(DP),     dimension(:,:), intent(inout) :: acc
      real(DP), dimension(:), allocatable, target :: rad
      type(c_ptr) :: kp
      rad = radius(1:npl)
      ! pl%k_plpl built by swiftest_util_flatten_eucl_plpl is the canonical table: never shipped to the device.
      ! Only the SyMBA encounter list (symba_kick.f90:61-68) is an explicit table.
      if (size(k_plpl, 2) == int(npl, I8B) * (npl - 1) / 2) then
         kp = c_null_ptr
      else
         kp = c_loc(k_plpl)
      end if
      call swcu_check(swcu_kick_getacch_int_all_flat_pl(swcu_ctx, npl, nplpl, kp, r, Gmass, c_loc(rad), acc), "flat_rad_pl")
   end subroutine

   module subroutine swiftest_kick_getacch_int_all_tp(ntp, npl, rtp, rpl, GMpl, lmask, acc)
      use swiftest_cuda
      implicit none
      integer(I4B),                 intent(in)    :: ntp, npl
      real(DP),     dimension(:,:), intent(in)    :: rtp, rpl
      real(DP),     dimension(:),   intent(in)    :: GMpl
      logical,      dimension(:),   intent(in)    :: lmask
      real(DP),     dimension(:,:), intent(inout) :: acc
      integer(c_int), dimension(ntp) :: imask
      imask(:) = merge(1_c_int, 0_c_int, lmask(1:ntp))
      call swcu_check(swcu_kick_getacch_int_all_tp(swcu_ctx, ntp, npl, rtp, rpl, GMpl, imask, acc), "all_tp")
   end subroutine

! ---- in submodule(swiftest) s_swiftest_drift ---------------------------------------------------------------------
   module subroutine swiftest_drift_all(mu, x, v, n, param, dt, lmask, iflag)
      use swiftest_cuda
      implicit none
      real(DP), dimension(:),     intent(in)    :: mu
      real(DP), dimension(:,:),   intent(inout) :: x, v
      integer(I4B),               intent(in)    :: n
      class(swiftest_parameters), intent(in)    :: param
      real(DP),                   intent(in)    :: dt
      logical, dimension(:),      intent(in)    :: lmask
      integer(I4B), dimension(:), intent(out)   :: iflag
      integer(c_int), dimension(n) :: imask
      if (n == 0) return
      imask(:) = merge(1_c_int, 0_c_int, lmask(1:n))
      call swcu_check(swcu_drift_all(swcu_ctx, n, mu, x, v, dt, merge(1_c_int, 0_c_int, param%lgr), param%inv_c2, &
                                     imask, iflag), "drift_all")
   end subroutine

! ---- in submodule(encounter) s_encounter_check --------------------------------------------------------------------
   subroutine encounter_check_all_sort_and_sweep_plpl(npl, r, v, renc, dt, nenc, index1, index2, lvdotr)
      use swiftest_cuda
      implicit none
      integer(I4B),                            intent(in)  :: npl
      real(DP),     dimension(:,:),            intent(in)  :: r, v
      real(DP),     dimension(:),              intent(in)  :: renc
      real(DP),                                intent(in)  :: dt
      integer(I8B),                            intent(out) :: nenc
      integer(I4B), dimension(:), allocatable, intent(out) :: index1, index2
      logical,      dimension(:), allocatable, intent(out) :: lvdotr
      integer(c_int), dimension(:), allocatable :: ilv
      if (npl == 0) return
      call swcu_check(swcu_encounter_check_all_sort_and_sweep_plpl(swcu_ctx, npl, r, v, renc, dt, nenc), "sas_plpl")
      if (nenc == 0) return
      allocate(index1(nenc), index2(nenc), lvdotr(nenc), ilv(nenc))      ! two-phase: allocate after learning nenc
      call swcu_check(swcu_encounter_fetch(swcu_ctx, nenc, index1, index2, ilv), "encounter_fetch")
      lvdotr(:) = ilv(:) /= 0
   end subroutine

! ---- in submodule(symba) s_symba_kick: the encounter-pair removal of symba_kick_getacch_pl (symba_kick.f90:59-70) --
!            if (plpl_encounter%nenc > 0) then
!               call swcu_check(swcu_symba_kick_subtract_encounters(swcu_ctx, npl, int(plpl_encounter%nenc, c_int64_t), &
!                     plpl_encounter%index1, plpl_encounter%index2, pl%rh, pl%Gmass, pl%radius, pl%ah), "symba subtract")
!            end if

! ---- in submodule(helio) s_helio_step: the whole democratic-heliocentric step stays on the device (tier 2) ---------
!    helio_step_pl (helio_step.f90:37-78) and helio_step_tp (:81-123) become one call each.  The resident arrays were
!    uploaded by swcu_body_sync when swcu_generation_pl / _tp last changed; rh, vh (and vb for output) come back with
!    swcu_body_get / swcu_body_get_vb only when the driver writes a frame (INTEGRATION.md section 3, hook 3).
   module subroutine helio_step_pl(self, nbody_system, param, t, dt)
      use swiftest_cuda
      implicit none
      class(helio_pl),              intent(inout) :: self
      class(swiftest_nbody_system), intent(inout) :: nbody_system
      class(swiftest_parameters),   intent(inout) :: param
      real(DP),                     intent(in)    :: t, dt
      integer(c_int) :: nfail, variant
      if (self%nbody == 0) return
      variant = merge(SWCU_LOOP_FLAT, SWCU_LOOP_TRIANGULAR, param%lflatten_interactions)
      call swcu_check(swcu_helio_step_pl(swcu_ctx, nbody_system%cb%Gmass, dt, variant, &
                                         merge(1_c_int, 0_c_int, param%lclose), merge(1_c_int, 0_c_int, self%lfirst), nfail), &
                      "helio_step_pl")
      self%lfirst = .false.
      if (nfail > 0) call helio_step_fetch_drift_failures(self)   ! swcu_body_get(iflag) -> DISCARDED_DRIFTERR (helio_drift.f90:41-50)
   end subroutine

   module subroutine helio_step_tp(self, nbody_system, param, t, dt)
      use swiftest_cuda
      implicit none
      class(helio_tp),              intent(inout) :: self
      class(swiftest_nbody_system), intent(inout) :: nbody_system
      class(swiftest_parameters),   intent(inout) :: param
      real(DP),                     intent(in)    :: t, dt
      integer(c_int) :: nfail
      if (self%nbody == 0) return
      call swcu_check(swcu_helio_step_tp(swcu_ctx, nbody_system%cb%Gmass, dt, merge(1_c_int, 0_c_int, self%lfirst), nfail), &
                      "helio_step_tp")
      self%lfirst = .false.
      if (nfail > 0) call helio_step_fetch_drift_failures(self)
   end subroutine

! ---- in submodule(swiftest) s_swiftest_util: swiftest_util_get_potential_energy_flat / _triangular -----------------
   module subroutine swiftest_util_get_potential_energy_triangular(npl, lmask, GMcb, Gmass, mass, rb, pe)
      use swiftest_cuda
      implicit none
      integer(I4B),                 intent(in)  :: npl
      logical,      dimension(:),   intent(in)  :: lmask
      real(DP),                     intent(in)  :: GMcb
      real(DP),     dimension(:),   intent(in)  :: Gmass, mass
      real(DP),     dimension(:,:), intent(in)  :: rb
      real(DP),                     intent(out) :: pe
      integer(c_int), dimension(npl) :: imask
      imask(:) = merge(1_c_int, 0_c_int, lmask(1:npl))
      call swcu_check(swcu_util_get_potential_energy(swcu_ctx, npl, imask, GMcb, Gmass, mass, rb, pe), "potential energy")
   end subroutine

! ---- in submodule(swiftest) s_swiftest_discard: the double loop of swiftest_discard_pl_tp (:261-288) ---------------
!         call swcu_check(swcu_discard_pl_tp(swcu_ctx, ntp, npl, tp%rh, tp%vh, merge(1_c_int, 0_c_int, tp%status(1:ntp) == ACTIVE), &
!                                            pl%rh, pl%vh, pl%radius, param%dt, iplanet, ndiscard), "discard_pl_tp")
!         do i = 1, ntp            ! the bookkeeping (status, ldiscard, log line, info%set_value) stays as it is,
!            j = iplanet(i)        ! driven by the planet index the kernel found
!            if (j == 0) cycle
!            tp%status(i) = DISCARDED_PLR ; tp%lmask(i) = .false. ; pl%ldiscard(j) = .true. ; ...
!         end do
