#!/usr/bin/env python
"""Generates patches/swiftest_use_cuda.diff: the edits a Swiftest maintainer applies to carlislewishard/swiftest
(2023.10.2) to route the force-and-drift hot path through libswiftest_cuda.so when configured with -DUSE_CUDA=ON.

    python patches/make_patch.py [/root/reference]          (writes patches/swiftest_use_cuda.diff)
    cd <swiftest checkout> && patch -p1 < swiftest_use_cuda.diff && cp <this repo>/fortran/swiftest_cuda.f90 src/cuda/

Every edit is an `#ifdef USE_CUDA ... #endif` block (the reference already runs its sources through the C preprocessor:
PROFILE, DOCONLOC, COARRAY) placed at the first executable statement of a procedure: a Fortran 2008 BLOCK that makes the
C call and RETURNs, so the original loop below it stays untouched and is what a build without USE_CUDA compiles.
Interfaces, argument lists and callers do not change.  The diff carries one line of context per hunk.

Which forms are used where
  * array-level module procedures (swiftest_kick_getacch_int_all_*, swiftest_drift_all, encounter_check_all_*):
    tier 1 of the C ABI -- host arrays in, host arrays out, per call.
  * the flat (k_plpl) forms: the ARRAY-LEVEL procedures always pass the table they are given (it is an explicit pair list:
    SyMBA's encounter list); the TYPE-BOUND procedures swiftest_kick_getacch_int_pl / symba_kick_getacch_int_pl, whose
    table is the canonical flattened triangle built by swiftest_util_flatten_eucl_plpl, say so explicitly by passing
    c_null_ptr with nplpl / nplplm -- nothing is inferred from the size of the table.
"""
import difflib
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# (file, anchor = first executable line of the procedure, unique text of the procedure header, block to insert BEFORE the anchor)
EDITS = []


def edit(path, header, anchor, block):
    EDITS.append((path, header, anchor, block))


KICK = "src/swiftest/swiftest_kick.f90"
edit(KICK, "module subroutine swiftest_kick_getacch_int_pl(self, param)", "      if (param%lflatten_interactions) then", """\
#ifdef USE_CUDA
      block   ! the table self%k_plpl is the canonical flattened triangle (swiftest_util_flatten_eucl_plpl): never shipped
         use swiftest_cuda
         real(DP), dimension(:), allocatable, target :: rad
         type(c_ptr) :: prad
         if (self%nbody == 0) return
         call swcu_ensure_ctx()
         prad = c_null_ptr
         if (param%lclose) then
            rad = self%radius(1:self%nbody)
            prad = c_loc(rad)
         end if
         if (param%lflatten_interactions) then
            call swcu_check(swcu_kick_getacch_int_all_flat_pl(swcu_ctx, self%nbody, int(self%nplpl, c_int64_t), c_null_ptr, &
                                                              self%rh, self%Gmass, prad, self%ah), "kick_getacch_int_pl (flat)")
         else
            call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, self%nbody, self%nbody, self%rh, self%Gmass, prad, &
                                                             self%ah), "kick_getacch_int_pl (triangular)")
         end if
         return
      end block
#endif
""")
edit(KICK, "module subroutine swiftest_kick_getacch_int_all_flat_rad_pl(", "      ahi(:,:) = 0.0_DP", """\
#ifdef USE_CUDA
      block   ! array-level form: k_plpl is an explicit pair list (SyMBA's encounter list) and is passed as it is
         use swiftest_cuda
         integer(c_int), dimension(:,:), allocatable, target :: kp
         real(DP), dimension(:), allocatable, target :: rad
         if (npl == 0 .or. nplpl == 0_I8B) return
         call swcu_ensure_ctx()
         kp = k_plpl(1:2, 1:nplpl)
         rad = radius(1:npl)
         call swcu_check(swcu_kick_getacch_int_all_flat_pl(swcu_ctx, npl, int(nplpl, c_int64_t), c_loc(kp), r, Gmass, &
                                                           c_loc(rad), acc), "kick_getacch_int_all_flat_rad_pl")
         return
      end block
#endif
""")
edit(KICK, "module subroutine swiftest_kick_getacch_int_all_flat_norad_pl(", "      ahi(:,:) = 0.0_DP", """\
#ifdef USE_CUDA
      block
         use swiftest_cuda
         integer(c_int), dimension(:,:), allocatable, target :: kp
         if (npl == 0 .or. nplpl == 0_I8B) return
         call swcu_ensure_ctx()
         kp = k_plpl(1:2, 1:nplpl)
         call swcu_check(swcu_kick_getacch_int_all_flat_pl(swcu_ctx, npl, int(nplpl, c_int64_t), c_loc(kp), r, Gmass, &
                                                           c_null_ptr, acc), "kick_getacch_int_all_flat_norad_pl")
         return
      end block
#endif
""")
edit(KICK, "module subroutine swiftest_kick_getacch_int_all_tri_rad_pl(", "      nplt = npl - nplm", """\
#ifdef USE_CUDA
      block   ! both branches of the loop below (upper triangle with reduction when nplt > nplm, full rows otherwise)
         use swiftest_cuda
         real(DP), dimension(:), allocatable, target :: rad
         if (npl == 0) return
         call swcu_ensure_ctx()
         rad = radius(1:npl)
         call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, npl, nplm, r, Gmass, c_loc(rad), acc), &
                         "kick_getacch_int_all_tri_rad_pl")
         return
      end block
#endif
""")
edit(KICK, "module subroutine swiftest_kick_getacch_int_all_tri_norad_pl(", "      nplt = npl - nplm", """\
#ifdef USE_CUDA
      block
         use swiftest_cuda
         if (npl == 0) return
         call swcu_ensure_ctx()
         call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, npl, nplm, r, Gmass, c_null_ptr, acc), &
                         "kick_getacch_int_all_tri_norad_pl")
         return
      end block
#endif
""")
edit(KICK, "module subroutine swiftest_kick_getacch_int_all_tp(", "      !$omp parallel do default(private) schedule(static)&", """\
#ifdef USE_CUDA
      block   ! logical masks cross the boundary as integers (gfortran and Intel disagree on the bits of .true.)
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: imask
         if (ntp == 0 .or. npl == 0) return
         call swcu_ensure_ctx()
         allocate(imask(ntp))
         imask(:) = merge(1_c_int, 0_c_int, lmask(1:ntp))
         call swcu_check(swcu_kick_getacch_int_all_tp(swcu_ctx, ntp, npl, rtp, rpl, GMpl, imask, acc), "kick_getacch_int_all_tp")
         return
      end block
#endif
""")

edit("src/swiftest/swiftest_drift.f90", "module subroutine swiftest_drift_all(", "      if (n == 0) return", """\
#ifdef USE_CUDA
      block   ! the GR step-size correction (dtp) is applied by the kernel; iflag keeps the reference's meaning
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: imask
         if (n == 0) return
         call swcu_ensure_ctx()
         allocate(imask(n))
         imask(:) = merge(1_c_int, 0_c_int, lmask(1:n))
         call swcu_check(swcu_drift_all(swcu_ctx, n, mu, x, v, dt, merge(1_c_int, 0_c_int, param%lgr), param%inv_c2, &
                                        imask, iflag), "drift_all")
         return
      end block
#endif
""")

ENC = "src/encounter/encounter_check.f90"
FETCH = """\
         if (nenc == 0_I8B) return
         allocate(index1(nenc), index2(nenc), lvdotr(nenc), ilv(nenc))     ! two-phase: sized after the count is known
         call swcu_check(swcu_encounter_fetch(swcu_ctx, int(nenc, c_int64_t), index1, index2, ilv), "encounter_fetch")
         lvdotr(:) = ilv(:) /= 0
         return
      end block
#endif
"""
edit(ENC, "   subroutine encounter_check_all_sort_and_sweep_plpl(", "      if (npl == 0) return", """\
#ifdef USE_CUDA
      block   ! the persistent bounding-box state of this call site (the save'd variables below) lives in the context
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: ilv
         integer(c_int64_t) :: nfound
         nenc = 0_I8B
         if (npl == 0) return
         call swcu_ensure_ctx()
         call swcu_check(swcu_encounter_check_all_sort_and_sweep_plpl(swcu_ctx, npl, r, v, renc, dt, nfound), "sort_and_sweep_plpl")
         nenc = int(nfound, I8B)
""" + FETCH)
edit(ENC, "   subroutine encounter_check_all_sort_and_sweep_plplm(", "      if ((nplm == 0) .or. (nplt == 0)) return", """\
#ifdef USE_CUDA
      block   ! index1 counts the plm list, index2 the plt list (1-based within each list), as below
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: ilv
         integer(c_int64_t) :: nfound
         nenc = 0_I8B
         if ((nplm == 0) .or. (nplt == 0)) return
         call swcu_ensure_ctx()
         call swcu_check(swcu_encounter_check_all_sort_and_sweep_plplm(swcu_ctx, nplm, nplt, rplm, vplm, rplt, vplt, rencm, &
                                                                       renct, dt, nfound), "sort_and_sweep_plplm")
         nenc = int(nfound, I8B)
""" + FETCH)
edit(ENC, "   subroutine encounter_check_all_sort_and_sweep_pltp(", "      if ((ntp == 0) .or. (npl == 0)) return", """\
#ifdef USE_CUDA
      block   ! test particles have renc = 0 (renctp below); the library knows
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: ilv
         integer(c_int64_t) :: nfound
         nenc = 0_I8B
         if ((ntp == 0) .or. (npl == 0)) return
         call swcu_ensure_ctx()
         call swcu_check(swcu_encounter_check_all_sort_and_sweep_pltp(swcu_ctx, npl, ntp, rpl, vpl, rtp, vtp, rencpl, dt, &
                                                                      nfound), "sort_and_sweep_pltp")
         nenc = int(nfound, I8B)
""" + FETCH)
edit(ENC, "   module subroutine encounter_check_all_plplm(", "      allocate(tmp_param, source=param)", """\
#ifdef USE_CUDA
      if (param%lencounter_sas_plpl) then   ! pl-pl sweep + plm-plt sweep + merge + index shift + sort in one call
      block
         use swiftest_cuda
         integer(c_int), dimension(:), allocatable :: ilv
         integer(c_int64_t) :: nfound
         nenc = 0_I8B
         if (nplm == 0) return
         call swcu_ensure_ctx()
         call swcu_check(swcu_encounter_check_all_plplm(swcu_ctx, nplm, nplt, rplm, vplm, rplt, vplt, rencm, renct, dt, &
                                                        nfound), "encounter_check_all_plplm")
         nenc = int(nfound, I8B)
""" + FETCH.replace("#endif\n", "      end if\n#endif\n"))

edit("src/symba/symba_kick.f90", "module subroutine symba_kick_getacch_int_pl(self, param)", "      if (param%lflatten_interactions) then", """\
#ifdef USE_CUDA
      block   ! canonical pairs with i <= nplm: the first nplplm entries of the flattened triangle (symba_util.f90:202)
         use swiftest_cuda
         real(DP), dimension(:), allocatable, target :: rad
         if (self%nbody == 0) return
         call swcu_ensure_ctx()
         rad = self%radius(1:self%nbody)
         if (param%lflatten_interactions) then
            call swcu_check(swcu_kick_getacch_int_all_flat_pl(swcu_ctx, self%nbody, int(self%nplplm, c_int64_t), c_null_ptr, &
                                                              self%rh, self%Gmass, c_loc(rad), self%ah), "symba_kick_getacch_int_pl (flat)")
         else
            call swcu_check(swcu_kick_getacch_int_all_tri_pl(swcu_ctx, self%nbody, self%nplm, self%rh, self%Gmass, c_loc(rad), &
                                                             self%ah), "symba_kick_getacch_int_pl (triangular)")
         end if
         return
      end block
#endif
""")
edit("src/symba/symba_kick.f90", "module subroutine symba_kick_getacch_pl(self, nbody_system, param, t, lbeg)",
     "               ah_enc(:,:) = 0.0_DP", """\
#ifdef USE_CUDA
               block   ! all pairs were kicked above; the listed pairs again into a zeroed ah_enc, then ah = ah - ah_enc
                  use swiftest_cuda
                  call swcu_ensure_ctx()
                  call swcu_check(swcu_symba_kick_subtract_encounters(swcu_ctx, npl, int(plpl_encounter%nenc, c_int64_t), &
                                  plpl_encounter%index1, plpl_encounter%index2, pl%rh, pl%Gmass, pl%radius, pl%ah), &
                                  "symba_kick_getacch_pl (encounter pairs)")
               end block
#else
""")
# ... and the matching #endif after the subtraction statement (second edit on the same procedure, anchored on the next line)
edit("src/symba/symba_kick.f90", "module subroutine symba_kick_getacch_pl(self, nbody_system, param, t, lbeg)",
     "            end if\n\n         end associate", "#endif\n")

edit("src/swiftest/swiftest_discard.f90", "   subroutine swiftest_discard_pl_tp(tp, nbody_system, param)",
     "         do i = 1, ntp\n            if (tp%status(i) == ACTIVE) then\n               do j = 1, npl", """\
#ifdef USE_CUDA
         ! the O(ntp*npl) search runs on the device and returns, per particle, the first planet that discards it (0: none);
         ! the loop below then visits only that planet and does the bookkeeping exactly as before
         if (ntp > 0 .and. npl > 0) then
            block
               use swiftest_cuda
               integer(c_int) :: ndiscard
               call swcu_ensure_ctx()
               if (allocated(swcu_iplanet)) deallocate(swcu_iplanet)
               allocate(swcu_iplanet(ntp))
               call swcu_check(swcu_discard_pl_tp(swcu_ctx, ntp, npl, tp%rh, tp%vh, merge(1_c_int, 0_c_int, tp%status(1:ntp) == ACTIVE), &
                                                  pl%rh, pl%vh, pl%radius, dt, swcu_iplanet, ndiscard), "discard_pl_tp")
            end block
         end if
#endif
""")
edit("src/swiftest/swiftest_discard.f90", "   subroutine swiftest_discard_pl_tp(tp, nbody_system, param)",
     "               do j = 1, npl\n                  dx(:) = tp%rh(:, i) - pl%rh(:, j)", """\
#ifdef USE_CUDA
               do j = max(swcu_iplanet(i), 1), swcu_iplanet(i)   ! the planet the kernel found (none: empty loop)
#else
""")
edit("src/swiftest/swiftest_discard.f90", "   subroutine swiftest_discard_pl_tp(tp, nbody_system, param)",
     "                  dx(:) = tp%rh(:, i) - pl%rh(:, j)\n                  dv(:) = tp%vh(:, i) - pl%vh(:, j)\n                  radius = pl%radius(j)",
     "#endif\n")

CMAKE_TOP = ("CMakeLists.txt", 'OPTION(USE_SIMD "Use SIMD vectorization" ON)\n',
             'OPTION(USE_CUDA "Run the force-and-drift hot path on an NVIDIA B200 through libswiftest_cuda.so" OFF)\n')
CMAKE_SRC_FILES = ("src/CMakeLists.txt", "SET(FAST_MATH_FILES\n", "            ${SRC}/cuda/swiftest_cuda.f90\n")


def apply_edits(text, path):
    """Insert every block of this file right BEFORE its anchor (searched from the procedure header on)."""
    for p, header, anchor, block in EDITS:
        if p != path:
            continue
        h = text.find(header)
        assert h >= 0, (path, header)
        a = text.find(anchor, h)
        assert a >= 0, (path, header, anchor)
        nxt = text.find("\n   end subroutine", h)
        assert a < nxt, (path, header, anchor, "anchor lies outside the procedure")
        text = text[:a] + block + text[a:]
    return text


def main():
    out = []
    files = sorted({p for p, *_ in EDITS})
    for path in files:
        old = open(os.path.join(REF, path)).read()
        new = apply_edits(old, path)
        if path.endswith("swiftest_discard.f90"):
            # the per-particle planet index lives next to the other module-level state of this submodule
            new = new.replace("contains\n", "   integer(I4B), dimension(:), allocatable, save :: swcu_iplanet !! USE_CUDA: first discarding planet per tp\ncontains\n", 1) \
                if "swcu_iplanet !!" not in new else new
        out += difflib.unified_diff(old.split("\n"), new.split("\n"), "a/" + path, "b/" + path, n=1, lineterm="")
    for path, anchor, add in (CMAKE_TOP, CMAKE_SRC_FILES):
        old = open(os.path.join(REF, path)).read()
        assert anchor in old, (path, anchor)
        new = old.replace(anchor, anchor + add, 1)
        if path == "src/CMakeLists.txt":
            new += ("\n# libswiftest_cuda.so (hand-written sm_100a kernels behind a C ABI; include/swiftest_cuda.h)\n"
                    "IF (USE_CUDA)\n"
                    "    FIND_LIBRARY(SWIFTEST_CUDA_LIB swiftest_cuda HINTS ${SWIFTEST_CUDA_ROOT}/lib $ENV{SWIFTEST_CUDA_ROOT}/lib REQUIRED)\n"
                    "    TARGET_COMPILE_DEFINITIONS(${SWIFTEST_LIBRARY} PUBLIC -DUSE_CUDA)\n"
                    "    TARGET_COMPILE_DEFINITIONS(${SWIFTEST_DRIVER} PUBLIC -DUSE_CUDA)\n"
                    "    TARGET_LINK_LIBRARIES(${SWIFTEST_LIBRARY} PUBLIC ${SWIFTEST_CUDA_LIB})\n"
                    "ENDIF ()\n")
        out += difflib.unified_diff(old.split("\n"), new.split("\n"), "a/" + path, "b/" + path, n=1, lineterm="")
    dst = os.environ.get("SWCU_PATCH_OUT") or os.path.join(HERE, "swiftest_use_cuda.diff")
    with open(dst, "w") as f:
        f.write("\n".join(out) + "\n")
    print(f"wrote {dst}: {sum(1 for l in out if l.startswith('+') and not l.startswith('+++'))} added lines, "
          f"{sum(1 for l in out if l.startswith('-') and not l.startswith('---'))} removed lines, {len(files) + 2} files")


if __name__ == "__main__":
    main()
