!! swiftest_cuda.f90 -- iso_c_binding interface module over libswiftest_cuda.so (include/swiftest_cuda.h).
!!
!! This is the file a Swiftest maintainer adds as src/cuda/swiftest_cuda.f90 (INTEGRATION.md).  It could NOT be compiled
!! in the build container (no Fortran compiler in the image, SURVEY.md section 8c); every interface below is a literal
!! transcription of the C prototypes, which ARE exercised from Python/ctypes by tests/.
!!
!! Conventions: scalars by value, arrays by reference (assumed-size, contiguous), logical masks converted with
!! merge(1_c_int, 0_c_int, lmask), nplpl/nenc as integer(c_int64_t), status /= 0 -> base_util_exit(FAILURE).
module swiftest_cuda
   use, intrinsic :: iso_c_binding
   implicit none
   public

   integer(c_int), parameter :: SWCU_OK = 0
   integer(c_int), parameter :: SWCU_PL = 0, SWCU_TP = 1
   integer(c_int), parameter :: SWCU_LOOP_TRIANGULAR = 0, SWCU_LOOP_FLAT = 1, SWCU_LOOP_AUTO = 2

   type(c_ptr), save  :: swcu_ctx = c_null_ptr      !! one context per process / coarray image
   integer(c_int64_t), save :: swcu_generation_pl = 0_c_int64_t !! bumped in rearray_pl, pl%flatten (Fraggle), restart read-in
   integer(c_int64_t), save :: swcu_generation_tp = 0_c_int64_t !! bumped in tp%spill / rearray

   interface
      integer(c_int) function swcu_create(device, ctx) bind(C, name="swcu_create")
         import :: c_int, c_ptr
         integer(c_int), value :: device
         type(c_ptr), intent(out) :: ctx
      end function
      integer(c_int) function swcu_destroy(ctx) bind(C, name="swcu_destroy")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      type(c_ptr) function swcu_last_error(ctx) bind(C, name="swcu_last_error")
         import :: c_ptr
         type(c_ptr), value :: ctx
      end function

      ! ---- tier 1: array-level, host arrays ----
      integer(c_int) function swcu_kick_getacch_int_all_flat_pl(ctx, npl, nplpl, k_plpl, r, Gmass, radius, acc) &
            bind(C, name="swcu_kick_getacch_int_all_flat_pl")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl
         integer(c_int64_t), value :: nplpl
         type(c_ptr), value :: k_plpl           !! c_null_ptr: canonical flattened pairs; else c_loc(k_plpl_enc)
         real(c_double), intent(in) :: r(3,*), Gmass(*)
         type(c_ptr), value :: radius           !! c_null_ptr selects the norad variant
         real(c_double), intent(inout) :: acc(3,*)
      end function
      integer(c_int) function swcu_kick_getacch_int_all_tri_pl(ctx, npl, nplm, r, Gmass, radius, acc) &
            bind(C, name="swcu_kick_getacch_int_all_tri_pl")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl, nplm
         real(c_double), intent(in) :: r(3,*), Gmass(*)
         type(c_ptr), value :: radius
         real(c_double), intent(inout) :: acc(3,*)
      end function
      integer(c_int) function swcu_kick_getacch_int_all_tp(ctx, ntp, npl, rtp, rpl, GMpl, lmask, acc) &
            bind(C, name="swcu_kick_getacch_int_all_tp")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: ntp, npl
         real(c_double), intent(in) :: rtp(3,*), rpl(3,*), GMpl(*)
         integer(c_int), intent(in) :: lmask(*)
         real(c_double), intent(inout) :: acc(3,*)
      end function
      integer(c_int) function swcu_symba_kick_subtract_encounters(ctx, npl, nenc, index1, index2, rh, Gmass, radius, ah) &
            bind(C, name="swcu_symba_kick_subtract_encounters")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*)
         real(c_double), intent(in) :: rh(3,*), Gmass(*), radius(*)
         real(c_double), intent(inout) :: ah(3,*)
      end function
      integer(c_int) function swcu_drift_all(ctx, n, mu, x, v, dt, lgr, inv_c2, lmask, iflag) bind(C, name="swcu_drift_all")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: n, lgr
         real(c_double), value :: dt, inv_c2
         real(c_double), intent(in) :: mu(*)
         real(c_double), intent(inout) :: x(3,*), v(3,*)
         integer(c_int), intent(in) :: lmask(*)
         integer(c_int), intent(inout) :: iflag(*)
      end function
      integer(c_int) function swcu_encounter_check_all_sort_and_sweep_plpl(ctx, npl, r, v, renc, dt, nenc) &
            bind(C, name="swcu_encounter_check_all_sort_and_sweep_plpl")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl
         real(c_double), intent(in) :: r(3,*), v(3,*), renc(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_check_all_sort_and_sweep_pltp(ctx, npl, ntp, rpl, vpl, rtp, vtp, rencpl, dt, nenc) &
            bind(C, name="swcu_encounter_check_all_sort_and_sweep_pltp")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl, ntp
         real(c_double), intent(in) :: rpl(3,*), vpl(3,*), rtp(3,*), vtp(3,*), rencpl(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_check_all_plplm(ctx, nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt, nenc) &
            bind(C, name="swcu_encounter_check_all_plplm")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: nplm, nplt
         real(c_double), intent(in) :: rplm(3,*), vplm(3,*), rplt(3,*), vplt(3,*), rencm(*), renct(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_fetch(ctx, nenc, index1, index2, lvdotr) bind(C, name="swcu_encounter_fetch")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(out) :: index1(*), index2(*), lvdotr(*)
      end function

      ! ---- tier 2: device-resident populations (see include/swiftest_cuda.h for the full list) ----
      integer(c_int) function swcu_body_sync(ctx, kind, n, nplm, r, v, Gmass, radius, rhill, mu, lmask, generation) &
            bind(C, name="swcu_body_sync")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, n, nplm
         type(c_ptr), value :: r, v, Gmass, radius, rhill, mu, lmask   !! c_loc(array) or c_null_ptr
         integer(c_int64_t), value :: generation
      end function
      integer(c_int) function swcu_pl_accel_int(ctx, loop_variant, lclose) bind(C, name="swcu_pl_accel_int")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: loop_variant, lclose
      end function
      integer(c_int) function swcu_tp_accel_int(ctx) bind(C, name="swcu_tp_accel_int")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_body_zero_accel(ctx, kind) bind(C, name="swcu_body_zero_accel")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
      end function
      integer(c_int) function swcu_body_kick_velocity(ctx, kind, dt) bind(C, name="swcu_body_kick_velocity")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         real(c_double), value :: dt
      end function
      integer(c_int) function swcu_pl_set_renc(ctx, irec) bind(C, name="swcu_pl_set_renc")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: irec
      end function
      integer(c_int) function swcu_pl_encounter_check(ctx, dt, nenc) bind(C, name="swcu_pl_encounter_check")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_tp_encounter_check(ctx, dt, nenc) bind(C, name="swcu_tp_encounter_check")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_body_put(ctx, kind, r, v, a, lmask) bind(C, name="swcu_body_put")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         type(c_ptr), value :: r, v, a, lmask      !! c_loc(array) or c_null_ptr
      end function
      integer(c_int) function swcu_body_get(ctx, kind, r, v, a, iflag) bind(C, name="swcu_body_get")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         type(c_ptr), value :: r, v, a, iflag
      end function
      !! slice forms: bodies i0+1 .. i1 (0-based half-open [i0,i1) on the C side), arrays of that length
      integer(c_int) function swcu_body_put_range(ctx, kind, i0, i1, r, v) bind(C, name="swcu_body_put_range")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, i0, i1
         type(c_ptr), value :: r, v
      end function
      integer(c_int) function swcu_body_get_range(ctx, kind, i0, i1, r, v, a) bind(C, name="swcu_body_get_range")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, i0, i1
         type(c_ptr), value :: r, v, a
      end function
      !! whm_step_pl on the resident planets (whm_step.f90:37-69); the tp step that follows passes c_null_ptr as ah0
      integer(c_int) function swcu_whm_step_pl(ctx, GMcb, dt, loop_variant, lclose, lfirst, nfail) &
            bind(C, name="swcu_whm_step_pl")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: GMcb, dt
         integer(c_int), value :: loop_variant, lclose, lfirst
         integer(c_int), intent(out) :: nfail
      end function
      integer(c_int) function swcu_whm_tp_first_accel(ctx) bind(C, name="swcu_whm_tp_first_accel")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_whm_get_jacobi(ctx, xj, vj) bind(C, name="swcu_whm_get_jacobi")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         type(c_ptr), value :: xj, vj
      end function
      !! asynchronous slice forms (page-locked arrays, e.g. allocated with cudaHostAlloc through iso_c_binding or
      !! registered with cudaHostRegister); complete after swcu_io_wait
      integer(c_int) function swcu_body_put_range_async(ctx, kind, i0, i1, r, v) bind(C, name="swcu_body_put_range_async")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, i0, i1
         type(c_ptr), value :: r, v
      end function
      integer(c_int) function swcu_body_get_range_async(ctx, kind, i0, i1, r, v, a) bind(C, name="swcu_body_get_range_async")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, i0, i1
         type(c_ptr), value :: r, v, a
      end function
      integer(c_int) function swcu_io_wait(ctx) bind(C, name="swcu_io_wait")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      !! whm_step_tp in one kernel (whm/whm_step.f90:72-100); ah0 = whm_kick_getacch_ah0 at the end-of-step planets
      integer(c_int) function swcu_whm_tp_step(ctx, dt, ah0, nfail) bind(C, name="swcu_whm_tp_step")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         real(c_double), intent(in) :: ah0(3)
         integer(c_int), intent(out) :: nfail
      end function
      !! multi-GPU (one image / process per GPU): NCCL id exchange or CUDA-IPC handle exchange is done by the host
      !! (co_broadcast of the byte buffers in a Coarray build), see include/swiftest_cuda.h
      integer(c_int) function swcu_p2p_export(ctx, handles) bind(C, name="swcu_p2p_export")
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: ctx
         character(kind=c_char), intent(out) :: handles(512)
      end function
      integer(c_int) function swcu_p2p_import(ctx, nranks, rank, all_handles) bind(C, name="swcu_p2p_import")
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: ctx
         integer(c_int), value :: nranks, rank
         character(kind=c_char), intent(in) :: all_handles(*)
      end function
      integer(c_int) function swcu_pl_kick_drift_p2p(ctx, lclose, dt, nfail) bind(C, name="swcu_pl_kick_drift_p2p")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: lclose
         real(c_double), value :: dt
         integer(c_int), intent(out) :: nfail
      end function
      integer(c_int) function swcu_body_drift(ctx, kind, dt, lgr, inv_c2, nfail) bind(C, name="swcu_body_drift")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, lgr
         real(c_double), value :: dt, inv_c2
         integer(c_int), intent(out) :: nfail
      end function

      ! ---- tier 2: democratic-heliocentric glue on the resident populations (helio/helio_step.f90:37-123) ----
      !! out-vectors are passed as type(c_ptr): c_loc(vec) to receive the value, c_null_ptr to leave it on the device
      integer(c_int) function swcu_pl_vh2vb(ctx, GMcb, vbcb) bind(C, name="swcu_pl_vh2vb")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx, vbcb
         real(c_double), value :: GMcb
      end function
      integer(c_int) function swcu_pl_vb2vh(ctx, GMcb, vbcb) bind(C, name="swcu_pl_vb2vh")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx, vbcb
         real(c_double), value :: GMcb
      end function
      integer(c_int) function swcu_pl_lindrift(ctx, GMcb, dt, lbeg, pt) bind(C, name="swcu_pl_lindrift")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx, pt
         real(c_double), value :: GMcb, dt
         integer(c_int), value :: lbeg
      end function
      integer(c_int) function swcu_tp_lindrift(ctx, dt, lbeg) bind(C, name="swcu_tp_lindrift")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         integer(c_int), value :: lbeg
      end function
      integer(c_int) function swcu_cb_set_pt(ctx, ptbeg, ptend) bind(C, name="swcu_cb_set_pt")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, ptbeg, ptend
      end function
      integer(c_int) function swcu_cb_get_pt(ctx, ptbeg, ptend) bind(C, name="swcu_cb_get_pt")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, ptbeg, ptend
      end function
      integer(c_int) function swcu_tp_vh2vb(ctx, lbeg) bind(C, name="swcu_tp_vh2vb")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: lbeg
      end function
      integer(c_int) function swcu_tp_vb2vh(ctx, lbeg) bind(C, name="swcu_tp_vb2vh")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: lbeg
      end function
      integer(c_int) function swcu_body_kick_vb(ctx, kind, dt, lbeg) bind(C, name="swcu_body_kick_vb")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind, lbeg
         real(c_double), value :: dt
      end function
      integer(c_int) function swcu_body_drift_vb(ctx, kind, GMcb, dt, nfail) bind(C, name="swcu_body_drift_vb")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         real(c_double), value :: GMcb, dt
         integer(c_int), intent(out) :: nfail
      end function
      integer(c_int) function swcu_body_put_vb(ctx, kind, vb) bind(C, name="swcu_body_put_vb")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         real(c_double), intent(in) :: vb(3,*)
      end function
      integer(c_int) function swcu_body_set_active(ctx, kind, lactive) bind(C, name="swcu_body_set_active")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         type(c_ptr), value :: lactive     !! c_loc of merge(1, 0, body%status(1:n) /= INACTIVE), or c_null_ptr: all active
      end function
      integer(c_int) function swcu_body_get_vb(ctx, kind, vb, rbeg, rend) bind(C, name="swcu_body_get_vb")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, vb, rbeg, rend
         integer(c_int), value :: kind
      end function
      integer(c_int) function swcu_helio_step_pl(ctx, GMcb, dt, loop_variant, lclose, lfirst, nfail) &
            bind(C, name="swcu_helio_step_pl")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: GMcb, dt
         integer(c_int), value :: loop_variant, lclose, lfirst
         integer(c_int), intent(out) :: nfail
      end function
      integer(c_int) function swcu_helio_step_tp(ctx, GMcb, dt, lfirst, nfail) bind(C, name="swcu_helio_step_tp")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: GMcb, dt
         integer(c_int), value :: lfirst
         integer(c_int), intent(out) :: nfail
      end function

      ! ---- tier 1: energy sums, triangular encounter checks, pl-tp discard, SyMBA list check ----
      integer(c_int) function swcu_util_get_potential_energy(ctx, npl, lmask, GMcb, Gmass, mass, rb, pe) &
            bind(C, name="swcu_util_get_potential_energy")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl
         integer(c_int), intent(in) :: lmask(*)
         real(c_double), value :: GMcb
         real(c_double), intent(in) :: Gmass(*), mass(*), rb(3,*)
         real(c_double), intent(out) :: pe
      end function
      integer(c_int) function swcu_util_get_energy_and_momentum(ctx, npl, lmask, GMcb, mass_cb, rbcb, vbcb, Gmass, mass, &
            radius, rb, vb, lclose, out8) bind(C, name="swcu_util_get_energy_and_momentum")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl, lclose
         integer(c_int), intent(in) :: lmask(*)
         real(c_double), value :: GMcb, mass_cb
         real(c_double), intent(in) :: rbcb(3), vbcb(3), Gmass(*), mass(*), radius(*), rb(3,*), vb(3,*)
         real(c_double), intent(out) :: out8(8)   !! ke_orbit, pe, be, te, L_orbit(1:3), GMtot
      end function
      integer(c_int) function swcu_encounter_check_all_triangular_plpl(ctx, npl, r, v, renc, dt, nenc) &
            bind(C, name="swcu_encounter_check_all_triangular_plpl")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl
         real(c_double), intent(in) :: r(3,*), v(3,*), renc(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_check_all_triangular_pltp(ctx, npl, ntp, rpl, vpl, rtp, vtp, rencpl, dt, nenc) &
            bind(C, name="swcu_encounter_check_all_triangular_pltp")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: npl, ntp
         real(c_double), intent(in) :: rpl(3,*), vpl(3,*), rtp(3,*), vtp(3,*), rencpl(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_check_all_triangular_plplm(ctx, nplm, nplt, rplm, vplm, rplt, vplt, rencm, &
            renct, dt, nenc) bind(C, name="swcu_encounter_check_all_triangular_plplm")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: nplm, nplt
         real(c_double), intent(in) :: rplm(3,*), vplm(3,*), rplt(3,*), vplt(3,*), rencm(*), renct(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_discard_pl_tp(ctx, ntp, npl, rtp, vtp, lactive, rpl, vpl, radius, dt, iplanet, ndiscard) &
            bind(C, name="swcu_discard_pl_tp")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: ntp, npl
         real(c_double), intent(in) :: rtp(3,*), vtp(3,*), rpl(3,*), vpl(3,*), radius(*)
         integer(c_int), intent(in) :: lactive(*)   !! merge(1, 0, tp%status == ACTIVE)
         real(c_double), value :: dt
         integer(c_int), intent(out) :: iplanet(*), ndiscard
      end function
      integer(c_int) function swcu_pl_encounter_check_triangular(ctx, dt, nenc) bind(C, name="swcu_pl_encounter_check_triangular")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_tp_encounter_check_triangular(ctx, dt, nenc) bind(C, name="swcu_tp_encounter_check_triangular")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_tp_discard_pl(ctx, dt, iplanet, ndiscard) bind(C, name="swcu_tp_discard_pl")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), value :: dt
         type(c_ptr), value :: iplanet   !! c_loc of an integer(c_int) array of ntp elements, or c_null_ptr
         integer(c_int), intent(out) :: ndiscard
      end function
      integer(c_int) function swcu_symba_encounter_check_list(ctx, nenc, index1, index2, lencmask, n1, r1, v1, renc1, &
            radius1, n2, r2, v2, renc2, radius2, dt, lencounter, lvdotr, nfound) &
            bind(C, name="swcu_symba_encounter_check_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lencmask(*)
         integer(c_int), value :: n1, n2
         real(c_double), intent(in) :: r1(3,*), v1(3,*), renc1(*), radius1(*)
         type(c_ptr), value :: r2, v2, renc2, radius2   !! c_null_ptr for the pl-pl form / absent arrays
         real(c_double), value :: dt
         integer(c_int), intent(out) :: lencounter(*)
         integer(c_int), intent(inout) :: lvdotr(*)
         integer(c_int64_t), intent(out) :: nfound
      end function
      integer(c_int) function swcu_symba_kick_list_plpl(ctx, nenc, index1, index2, lactive, npl, levelg, rh, rhill, Gmass, &
            dt, irec, sgn, vb, lgood) bind(C, name="swcu_symba_kick_list_plpl")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lactive(*), levelg(*)
         integer(c_int), value :: npl, irec, sgn
         real(c_double), intent(in) :: rh(3,*), rhill(*), Gmass(*)
         real(c_double), value :: dt
         real(c_double), intent(inout) :: vb(3,*)
         integer(c_int), intent(out) :: lgood(*)
      end function
      integer(c_int) function swcu_symba_kick_list_pltp(ctx, nenc, index1, index2, lactive, npl, ntp, levelg_pl, levelg_tp, &
            rh_pl, rhill, Gmass, rh_tp, dt, irec, sgn, vb_tp, lgood) bind(C, name="swcu_symba_kick_list_pltp")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lactive(*), levelg_pl(*), levelg_tp(*)
         integer(c_int), value :: npl, ntp, irec, sgn
         real(c_double), intent(in) :: rh_pl(3,*), rhill(*), Gmass(*), rh_tp(3,*)
         real(c_double), value :: dt
         real(c_double), intent(inout) :: vb_tp(3,*)
         integer(c_int), intent(out) :: lgood(*)
      end function
      integer(c_int) function swcu_collision_check_list(ctx, nenc, index1, index2, lmask, lvdotr, n1, r1, v1, Gmass1, &
            radius1, n2, r2, v2, dt, lcollision, lclosest, ncollision) bind(C, name="swcu_collision_check_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lmask(*), lvdotr(*)
         integer(c_int), value :: n1, n2
         real(c_double), intent(in) :: r1(3,*), v1(3,*), Gmass1(*), radius1(*)
         type(c_ptr), value :: r2, v2            !! c_null_ptr (and n2 = 0) for the pl-pl form
         real(c_double), value :: dt
         integer(c_int), intent(out) :: lcollision(*), lclosest(*)
         integer(c_int64_t), intent(out) :: ncollision
      end function
      ! tier 2 of the list loops: resident populations (pl%rh, pl%vb, ... stay on the device)
      integer(c_int) function swcu_pl_symba_kick_list(ctx, nenc, index1, index2, lactive, levelg, dt, irec, sgn, lgood) &
            bind(C, name="swcu_pl_symba_kick_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lactive(*), levelg(*)
         real(c_double), value :: dt
         integer(c_int), value :: irec, sgn
         type(c_ptr), value :: lgood     !! c_loc of an integer(c_int) array of nenc elements, or c_null_ptr (no synchronisation)
      end function
      integer(c_int) function swcu_tp_symba_kick_list(ctx, nenc, index1, index2, lactive, levelg_pl, levelg_tp, dt, irec, &
            sgn, lgood) bind(C, name="swcu_tp_symba_kick_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lactive(*), levelg_pl(*), levelg_tp(*)
         real(c_double), value :: dt
         integer(c_int), value :: irec, sgn
         type(c_ptr), value :: lgood
      end function
      integer(c_int) function swcu_body_symba_encounter_check_list(ctx, kind, nenc, index1, index2, lencmask, dt, &
            lencounter, lvdotr, nfound) bind(C, name="swcu_body_symba_encounter_check_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind     !! SWCU_PL: pl-pl list, SWCU_TP: pl-tp list
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lencmask(*)
         real(c_double), value :: dt
         integer(c_int), intent(out) :: lencounter(*)
         integer(c_int), intent(inout) :: lvdotr(*)
         integer(c_int64_t), intent(out) :: nfound
      end function
      integer(c_int) function swcu_body_collision_check_list(ctx, kind, nenc, index1, index2, lmask, lvdotr, dt, &
            lcollision, lclosest, ncollision) bind(C, name="swcu_body_collision_check_list")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         integer(c_int64_t), value :: nenc
         integer(c_int), intent(in) :: index1(*), index2(*), lmask(*), lvdotr(*)
         real(c_double), value :: dt
         integer(c_int), intent(out) :: lcollision(*), lclosest(*)
         integer(c_int64_t), intent(out) :: ncollision
      end function

      ! ---- the rest of the ABI: context services, multi-GPU plumbing, statistics and measurement helpers ----
      integer(c_int) function swcu_version() bind(C, name="swcu_version")
         import :: c_int
      end function
      integer(c_int64_t) function swcu_launch_count(ctx) bind(C, name="swcu_launch_count")
         import :: c_int64_t, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_set_stream(ctx, cuda_stream) bind(C, name="swcu_set_stream")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, cuda_stream
      end function
      integer(c_int) function swcu_synchronize(ctx) bind(C, name="swcu_synchronize")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_device_info(ctx, sm_count, cc, mem_bytes) bind(C, name="swcu_device_info")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), intent(out) :: sm_count, cc
         integer(c_int64_t), intent(out) :: mem_bytes
      end function
      integer(c_int) function swcu_encounter_check_all_sort_and_sweep_plplm(ctx, nplm, nplt, rplm, vplm, rplt, vplt, rencm, &
            renct, dt, nenc) bind(C, name="swcu_encounter_check_all_sort_and_sweep_plplm")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: nplm, nplt
         real(c_double), intent(in) :: rplm(3,*), vplm(3,*), rplt(3,*), vplt(3,*), rencm(*), renct(*)
         real(c_double), value :: dt
         integer(c_int64_t), intent(out) :: nenc
      end function
      integer(c_int) function swcu_encounter_stats(ctx, nbox_total, ncandidates_emitted) bind(C, name="swcu_encounter_stats")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: nbox_total, ncandidates_emitted
      end function
      integer(c_int) function swcu_body_count(ctx, kind, n, nplm, generation) bind(C, name="swcu_body_count")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: kind
         integer(c_int), intent(out) :: n, nplm
         integer(c_int64_t), intent(out) :: generation
      end function
      integer(c_int) function swcu_comm_unique_id(ctx, id128) bind(C, name="swcu_comm_unique_id")
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: ctx
         character(kind=c_char), intent(out) :: id128(128)
      end function
      integer(c_int) function swcu_comm_init(ctx, nranks, rank, id128) bind(C, name="swcu_comm_init")
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: ctx
         integer(c_int), value :: nranks, rank
         character(kind=c_char), intent(in) :: id128(128)
      end function
      integer(c_int) function swcu_comm_finalize(ctx) bind(C, name="swcu_comm_finalize")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_pl_set_slice(ctx, i0, i1) bind(C, name="swcu_pl_set_slice")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: i0, i1
      end function
      integer(c_int) function swcu_pl_allgather(ctx, with_v) bind(C, name="swcu_pl_allgather")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: with_v
      end function
      integer(c_int) function swcu_partition(n, nranks, rank, i0, i1) bind(C, name="swcu_partition")
         import :: c_int
         integer(c_int), value :: n, nranks, rank
         integer(c_int), intent(out) :: i0, i1
      end function
      integer(c_int) function swcu_p2p_close(ctx) bind(C, name="swcu_p2p_close")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_timer_start(ctx) bind(C, name="swcu_timer_start")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_timer_stop(ctx, elapsed_ms) bind(C, name="swcu_timer_stop")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), intent(out) :: elapsed_ms
      end function
      integer(c_int) function swcu_probe_fp64_peak(ctx, tflops) bind(C, name="swcu_probe_fp64_peak")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), intent(out) :: tflops
      end function
      integer(c_int) function swcu_probe_hbm_copy(ctx, bytes, gbs) bind(C, name="swcu_probe_hbm_copy")
         import :: c_int, c_int64_t, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int64_t), value :: bytes
         real(c_double), intent(out) :: gbs
      end function
      integer(c_int) function swcu_flush_l2(ctx) bind(C, name="swcu_flush_l2")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_last_kernel_ms(ctx, family, ms) bind(C, name="swcu_last_kernel_ms")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: family
         real(c_double), intent(out) :: ms
      end function
      integer(c_int) function swcu_enable_kernel_timing(ctx, on) bind(C, name="swcu_enable_kernel_timing")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: on
      end function
      integer(c_int) function swcu_kernel_ms_accumulated(ctx, family, total_ms, count) &
            bind(C, name="swcu_kernel_ms_accumulated")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: family
         real(c_double), intent(out) :: total_ms
         integer(c_int), intent(out) :: count
      end function
      integer(c_int) function swcu_timer_lap_begin(ctx) bind(C, name="swcu_timer_lap_begin")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_timer_lap_end(ctx) bind(C, name="swcu_timer_lap_end")
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
      end function
      integer(c_int) function swcu_timer_laps(ctx, total_ms, count, each_ms, each_cap) bind(C, name="swcu_timer_laps")
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), intent(out) :: total_ms
         integer(c_int), intent(out) :: count
         type(c_ptr), value :: each_ms      !! c_null_ptr, or c_loc of a real(c_double) array of each_cap elements
         integer(c_int), value :: each_cap
      end function
      integer(c_int) function swcu_encounter_direct_count(ctx, direct, fallbacks) bind(C, name="swcu_encounter_direct_count")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: direct, fallbacks
      end function
      integer(c_int) function swcu_encounter_bucket_fallbacks(ctx, count) bind(C, name="swcu_encounter_bucket_fallbacks")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: count
      end function
      integer(c_int) function swcu_step_graph_replays(ctx, count) bind(C, name="swcu_step_graph_replays")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: count
      end function
      integer(c_int) function swcu_flat_redo_count(ctx, chunks) bind(C, name="swcu_flat_redo_count")
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int64_t), intent(out) :: chunks
      end function
   end interface

contains

   subroutine swcu_ensure_ctx()
      !! One context per process (per coarray image: device = this_image() - 1), created at the first call of any
      !! replacement body and kept for the run.  Without an sm_100 GPU swcu_create fails and the run stops: there is no
      !! CPU fallback behind USE_CUDA.
      integer(c_int) :: device
      if (c_associated(swcu_ctx)) return
      device = 0_c_int
#ifdef COARRAY
      device = int(this_image() - 1, c_int)
#endif
      call swcu_check(swcu_create(device, swcu_ctx), "swcu_create")
   end subroutine swcu_ensure_ctx

   subroutine swcu_finalize()
      !! Called once from the driver's shutdown path (optional: the process exit releases the device as well)
      integer(c_int) :: status
      if (.not. c_associated(swcu_ctx)) return
      status = swcu_destroy(swcu_ctx)
      swcu_ctx = c_null_ptr
   end subroutine swcu_finalize

   subroutine swcu_check(status, where)
      !! Maps a nonzero status to the reference's fatal-error convention (base_util_exit(FAILURE), base_module.f90:589)
      use base, only : base_util_exit, FAILURE
      integer(c_int),   intent(in) :: status
      character(len=*), intent(in) :: where
      if (status /= SWCU_OK) then
         write(*,*) "swiftest_cuda: ", where, " failed with status ", status
         call base_util_exit(FAILURE)
      end if
   end subroutine swcu_check

end module swiftest_cuda
