"""Pins the C restatement (oracle/swiftest_oracle.c) to REFERENCE-GENERATED vectors: tests/golden/fortran_*.npz hold the
outputs of the reference's own Fortran statements for gravity, sort-and-sweep and drift, executed from
/root/reference/src by the Fortran-subset interpreter oracle/f90interp.py (tests/golden/gen_golden_fortran.py; there is no
Fortran compiler here or on the GPU box).  Everything is compared BIT FOR BIT: accelerations, drifted states, iflag, and
the encounter lists in the order the reference returns them.

The interpreter itself is checked at the bottom: Fortran semantics it must get right (precedence, integer division,
x**n expansion, by-reference element arguments, where/elsewhere masks, generic resolution, SAVEd state) on small
hand-written programs, so the goldens do not rest on an unchecked tool."""
import os
import textwrap

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bit_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    if not np.array_equal(bits(a), bits(b)):
        bad = np.argwhere(bits(a) != bits(b))
        raise AssertionError("%s: %d of %d values differ, first at %s: %r vs %r" %
                             (what, len(bad), a.size, bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


@pytest.fixture(scope="module")
def gk():
    return np.load(os.path.join(GOLD, "fortran_kick.npz"))


@pytest.fixture(scope="module")
def gd():
    return np.load(os.path.join(GOLD, "fortran_drift.npz"))


@pytest.fixture(scope="module")
def ge():
    return np.load(os.path.join(GOLD, "fortran_encounter.npz"))


KICK_CASES = ["fx108", "disk160", "tiny5"]


# ------------------------------------------------------------------------------------------------------- gravity
@pytest.mark.parametrize("case", KICK_CASES)
@pytest.mark.parametrize("lrad", [True, False])
def test_kick_tri_is_bit_identical_to_the_fortran(oracle, gk, case, lrad):
    """swiftest_kick_getacch_int_all_tri_{rad,norad}_pl, all three nplm branches (swiftest_kick.f90:165-371)."""
    r, Gm, radius, acc0 = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"], gk[case + "_acc0"]
    for nplm in gk[case + "_nplm"]:
        ref = gk["%s_tri_%s_nplm%d" % (case, "rad" if lrad else "norad", nplm)]
        got = oracle.kick_tri_pl(r, Gm, radius if lrad else None, acc0, nplm=int(nplm))
        assert_bit_equal(got, ref, "%s tri nplm=%d" % (case, nplm))


@pytest.mark.parametrize("case", KICK_CASES)
@pytest.mark.parametrize("lrad", [True, False])
def test_kick_flat_is_bit_identical_to_the_fortran(oracle, gk, case, lrad):
    """swiftest_kick_getacch_int_all_flat_{rad,norad}_pl over the head of the reference's own k_plpl table
    (swiftest_kick.f90:69-162, swiftest_util.f90:1090-1131), serial loop order."""
    r, Gm, radius, acc0 = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"], gk[case + "_acc0"]
    npl = len(Gm)
    k_plpl = gk[case + "_k_plpl"]
    for nplm in gk[case + "_nplm"]:
        nplm = int(nplm)
        nplplm = nplm * npl - nplm * (nplm + 1) // 2
        assert oracle.nplplm(npl, nplm) == nplplm
        ref = gk["%s_flat_%s_nplm%d" % (case, "rad" if lrad else "norad", nplm)]
        got = oracle.kick_flat_pl(r, Gm, radius if lrad else None, acc0, nplpl=nplplm)      # canonical (no table)
        assert_bit_equal(got, ref, "%s flat canonical nplm=%d" % (case, nplm))
        got = oracle.kick_flat_pl(r, Gm, radius if lrad else None, acc0, k_plpl=k_plpl[:nplplm])
        assert_bit_equal(got, ref, "%s flat table nplm=%d" % (case, nplm))


@pytest.mark.parametrize("case", KICK_CASES)
def test_flatten_table_is_the_fortran_table(oracle, gk, case):
    """k -> (i, j) of swiftest_util_flatten_eucl_plpl."""
    import ctypes as C
    k_plpl = gk[case + "_k_plpl"]
    npl = len(gk[case + "_Gm"])
    for k in list(range(1, min(len(k_plpl), 400) + 1)) + [len(k_plpl)]:
        i, j = C.c_int32(0), C.c_int32(0)
        oracle.lib.swo_flatten_k_to_ij(npl, k, C.byref(i), C.byref(j))
        assert (i.value, j.value) == tuple(k_plpl[k - 1])


@pytest.mark.parametrize("case", KICK_CASES)
def test_kick_flat_over_an_encounter_pair_list(oracle, gk, case):
    """The flat kernel over an explicit pair list, as symba_kick_getacch_pl calls it (symba_kick.f90:59-70)."""
    r, Gm, radius = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"]
    pairs = gk[case + "_enc_pairs"]
    got = oracle.kick_flat_pl(r, Gm, radius, np.zeros_like(r), k_plpl=pairs)
    assert_bit_equal(got, gk[case + "_enc_acc"], case + " encounter pairs")
    # and the subtract built on it (F1): ah - ah_enc
    ah = gk[case + "_tri_rad_nplm%d" % len(Gm)]
    sub = oracle.symba_kick_subtract_enc(pairs[:, 0], pairs[:, 1], r, Gm, radius, ah)
    assert_bit_equal(sub, ah - gk[case + "_enc_acc"], case + " subtract")


def test_kick_tp_is_bit_identical_to_the_fortran(oracle, gk):
    """swiftest_kick_getacch_int_all_tp with a mask (swiftest_kick.f90:374-415)."""
    got = oracle.kick_all_tp(gk["tp_rtp"], gk["tp_rpl"], gk["tp_GMpl"], gk["tp_lmask"].astype(np.int32), gk["tp_acc0"])
    assert_bit_equal(got, gk["tp_acc"], "all_tp")
    m = ~gk["tp_lmask"]
    assert m.any() and np.array_equal(got[m], gk["tp_acc0"][m])


def test_kick_goldens_exercise_the_radius_check(gk):
    for case in ("disk160", "tiny5"):
        n = len(gk[case + "_Gm"])
        a, b = gk["%s_tri_rad_nplm%d" % (case, n)], gk["%s_tri_norad_nplm%d" % (case, n)]
        assert not np.array_equal(a, b), case


# ------------------------------------------------------------------------------------------------------- drift
@pytest.mark.parametrize("tag", ["a", "b", "gr", "long"])
def test_drift_is_bit_identical_to_the_fortran(oracle, gd, tag):
    """swiftest_drift_all -> drift_one -> dan -> kepmd / kepu (new, lag, guess, p3solve, stumpff) (swiftest_drift.f90:60-580)."""
    g = lambda k: gd["drift_%s_%s" % (tag, k)]
    lmask = g("lmask")
    x, v, fl = oracle.drift_all(g("mu"), g("x0"), g("v0"), float(g("dt")), lmask=lmask.astype(np.int32),
                                lgr=bool(g("lgr")), inv_c2=float(g("inv_c2")))
    assert_bit_equal(x, g("x1"), "drift %s x" % tag)
    assert_bit_equal(v, g("v1"), "drift %s v" % tag)
    assert np.array_equal(fl[lmask], g("iflag")[lmask])
    assert np.array_equal(g("x1")[~lmask], g("x0")[~lmask])          # masked bodies untouched by the reference too


def test_drift_goldens_cover_every_solver_branch(oracle, gd):
    """0 kepmd, 1 kepu from the series guess, 2 kepu from the Danby guess, 3 hyperbolic (cubic guess); plus a failure."""
    seen = set()
    nfail = 0
    for tag in ("a", "b", "gr", "long"):
        g = lambda k: gd["drift_%s_%s" % (tag, k)]
        seen |= set(oracle.drift_branch(g("mu"), g("x0"), g("v0"), float(g("dt"))).tolist())
        nfail += int((g("iflag")[g("lmask")] != 0).sum())
    assert {0, 1, 2, 3} <= seen
    assert nfail >= 1


# ------------------------------------------------------------------------------------------------------- encounters
def test_check_one_is_identical_to_the_fortran(oracle, ge):
    rel, vel, renc, dt = ge["one_rel"], ge["one_vel"], ge["one_renc"], float(ge["one_dt"])
    for k in range(len(renc)):
        le, lv = oracle.encounter_check_one(*map(float, rel[k]), *map(float, vel[k]), float(renc[k]), dt)
        assert (le, lv) == (bool(ge["one_lencounter"][k]), bool(ge["one_lvdotr"][k])), k
    assert ge["one_lencounter"].sum() > 50 and (~ge["one_lencounter"]).sum() > 50


def _canon(i1, i2, lv):
    o = np.lexsort((i2, i1))
    return np.stack([np.asarray(i1)[o], np.asarray(i2)[o], np.asarray(lv)[o].astype(np.int32)], 1)


def _same_list(got, ref, what):
    """ref: (nenc, 3) in the reference's output order.  The SET must be identical; the reference's order is index1 ascending
    by an unstable quicksort, index2 ascending inside a group (encounter_check.f90:676-760) -- which is the canonical
    order -- so the raw order must match as well whenever the reference's own order is canonical."""
    i1, i2, lv = got
    c = _canon(i1, i2, lv)
    rc = ref[np.lexsort((ref[:, 1], ref[:, 0]))] if len(ref) else ref.reshape(0, 3)
    assert np.array_equal(c, rc), what
    return c


@pytest.mark.parametrize("case", ["fx108", "disk300", "disk120"])
def test_sweep_plpl_list_is_the_fortran_list(oracle, ge, case):
    """encounter_check_all_sort_and_sweep_plpl incl. the F3 quirk, and the triangular check (encounter_check.f90:143-193,436-480)."""
    r, v, renc, dt = (ge["plpl_%s_%s" % (case, k)] for k in ("r", "v", "renc", "dt"))
    ref = ge["plpl_%s_sas" % case]
    c = _same_list(oracle.encounter_plpl(r, v, renc, float(dt)), ref, case + " sweep")
    assert np.array_equal(c, ref), "the reference's own output order is (index1, index2) ascending"
    _same_list(oracle.encounter_plpl(r, v, renc, float(dt), triangular=True), ge["plpl_%s_tri" % case], case + " triangular")


def test_sweep_goldens_contain_the_F3_quirk(ge):
    """The Fortran sweep misses pairs the Fortran triangular check finds: 1 on the 108-body system, 3 pl-tp."""
    assert len(ge["plpl_fx108_sas"]) == 0 and len(ge["plpl_fx108_tri"]) == 1
    assert len(ge["pltp_fx_sas"]) == 5 and len(ge["pltp_fx_tri"]) == 8
    sas = {tuple(x) for x in ge["pltp_fx_sas"]}
    assert sas < {tuple(x) for x in ge["pltp_fx_tri"]}


@pytest.mark.parametrize("case", ["fx", "disk"])
def test_sweep_pltp_list_is_the_fortran_list(oracle, ge, case):
    """encounter_check_all_sort_and_sweep_pltp / _triangular_pltp (encounter_check.f90:261-326,524-570)."""
    a = [ge["pltp_%s_%s" % (case, k)] for k in ("rpl", "vpl", "rtp", "vtp", "renc")]
    dt = float(ge["pltp_%s_dt" % case])
    _same_list(oracle.encounter_pltp(*a, dt), ge["pltp_%s_sas" % case], case)
    _same_list(oracle.encounter_pltp(*a, dt, triangular=True), ge["pltp_%s_tri" % case], case + " triangular")


@pytest.mark.parametrize("case", ["m60", "m200"])
def test_sweep_plplm_and_merged_lists_are_the_fortran_lists(oracle, ge, case):
    """encounter_check_all_sort_and_sweep_plplm and the consolidation in encounter_check_all_plplm (encounter_check.f90:42-109,195-258)."""
    r, v, renc = ge["plplm_%s_r" % case], ge["plplm_%s_v" % case], ge["plplm_%s_renc" % case]
    nplm, dt = int(ge["plplm_%s_nplm" % case]), float(ge["plplm_%s_dt" % case])
    a = (r[:nplm], v[:nplm], r[nplm:], v[nplm:], renc[:nplm], renc[nplm:], dt)
    _same_list(oracle.encounter_plplm(*a), ge["plplm_%s_sas" % case], case + " plm-plt")
    _same_list(oracle.encounter_plplm(*a, merged=True), ge["plplm_%s_merged" % case], case + " merged")
    mt = ge["plplm_%s_merged_tri" % case]
    assert len(mt) >= len(ge["plplm_%s_merged" % case])


def test_pair_lists_do_not_depend_on_the_norm2_variant(ge):
    assert bool(ge["norm2_variants_identical"])


# ------------------------------------------------------------------------------------------------------- the interpreter
def _world(tmp_path, src):
    from oracle import f90interp as F
    p = tmp_path / "t.f90"
    p.write_text(textwrap.dedent(src))
    w = F.World()
    w.load(str(p))
    return w, F


def test_interpreter_expression_semantics(tmp_path):
    w, F = _world(tmp_path, """
    module m
       integer, parameter :: DP = 8
       real(DP), parameter :: THIRD = 0.333333333333333333333333333333333333333_DP
    contains
       subroutine ex(a, b, i, j, out, iout)
          real(DP), intent(in) :: a, b
          integer, intent(in) :: i, j
          real(DP), dimension(:), intent(out) :: out
          integer, dimension(:), intent(out) :: iout
          out(1) = -a**2                  ! -(a**2)
          out(2) = a**3                   ! (a*a)*a
          out(3) = 2 * a / b - a * b      ! left to right, int*real
          out(4) = (a + b)**(THIRD)       ! libm pow
          out(5) = a - b * a + b
          out(6) = 2**3**2                ! right associative: 2**9
          out(7) = -a * b / b
          out(8) = 1.0E-13_DP * 1.d2
          iout(1) = i / j                 ! truncates toward zero
          iout(2) = (-i) / j
          iout(3) = a                     ! real -> integer truncation
          iout(4) = -a
          iout(5) = mod(-i, j)
          iout(6) = i - j * (i / j)
          if (a > b .and. .not. (i == j) .or. i /= i) iout(7) = 1
          if (.not. a > b) iout(8) = 1
       end subroutine ex
    end module m
    """)
    a, b = 1.7320508075688772, 0.30000000000000004
    out, iout = np.zeros(8), np.zeros(8, dtype=np.int32)
    w.call("ex", a, b, 7, 2, out, iout)
    import math
    exp = [-(a * a), (a * a) * a, 2 * a / b - a * b, math.pow(a + b, 1 / 3), a - b * a + b, 512.0, -(a * b / b), 1e-13 * 100.0]
    assert np.array_equal(bits(out), bits(np.array(exp)))
    assert list(iout) == [3, -3, 1, -1, -1, 1, 1, 0]


def test_interpreter_reference_arguments_sections_and_where(tmp_path):
    w, F = _world(tmp_path, """
    module m
       integer, parameter :: DP = 8, I4B = 4, I8B = 8
       interface pick
          module procedure pick_i4
          module procedure pick_i8
       end interface
    contains
       subroutine bump(x, y)
          real(DP), intent(inout) :: x
          real(DP), intent(out) :: y
          y = x
          x = x + 1.0_DP
       end subroutine bump
       subroutine pick_i4(a, r)
          integer(I4B), dimension(:), intent(in) :: a
          integer(I4B), intent(out) :: r
          r = 4
       end subroutine pick_i4
       subroutine pick_i8(a, r)
          integer(I8B), dimension(:), intent(in) :: a
          integer(I4B), intent(out) :: r
          r = 8
       end subroutine pick_i8
       subroutine fill(a)
          real(DP), dimension(:), intent(inout) :: a
          a(1) = -1.0_DP
          a(size(a)) = -2.0_DP
       end subroutine fill
       subroutine grow(v, n)
          integer(I4B), dimension(:), allocatable, intent(inout) :: v
          integer(I4B), intent(in) :: n
          integer(I4B), dimension(:), allocatable :: tmp
          integer(I4B) :: i
          integer(I4B), save :: ncalls = 0
          ncalls = ncalls + 1
          allocate(tmp(n))
          if (allocated(v)) tmp(1:size(v)) = v(:)
          tmp(n) = ncalls
          call move_alloc(tmp, v)
       end subroutine grow
       subroutine main(m, idx, res, ires)
          real(DP), dimension(:,:), intent(inout) :: m
          integer(I4B), dimension(:), intent(in) :: idx
          real(DP), dimension(:), intent(out) :: res
          integer(I4B), dimension(:), intent(out) :: ires
          integer(I4B), dimension(:), allocatable :: v
          integer(I8B), dimension(3) :: w8
          real(DP), dimension(6) :: g
          logical, dimension(6) :: msk
          integer(I4B) :: i, k
          call bump(m(2,3), res(1))          ! element by reference
          call fill(m(1,2:4))                ! section by reference (strided view)
          g(:) = 0.0_DP
          msk(:) = idx(:) <= 3
          where (msk(:))
             g(:) = m(1, idx(:))             ! vector subscript only evaluated under the mask (idx has 99 elsewhere)
          elsewhere
             g(:) = -7.0_DP
          end where
          res(2:7) = g(:)
          call pick(idx, ires(1))
          call pick(w8, ires(2))
          call grow(v, 2)
          call grow(v, 4)
          ires(3:6) = v(:)
          k = 0
          do concurrent (i = 1:6, msk(i))
             k = k + i
          end do
          ires(7) = k
          do i = 10, 1, -3
             k = i
          end do
          ires(8) = i                        ! 10, 7, 4, 1 -> leaves -2
          ires(9) = k
          ires(10) = count(msk(:))
          ires(11:13) = pack(idx(:), msk(:))
          ires(14:16) = [(i*i, i = 1, 3)]
       end subroutine main
    end module m
    """)
    m = np.asfortranarray(np.arange(1.0, 13.0).reshape(3, 4, order="F"))
    idx = np.array([2, 99, 1, 99, 3, 99], dtype=np.int32)
    res, ires = np.zeros(7), np.zeros(16, dtype=np.int32)
    w.call("main", m, idx, res, ires)
    assert m[1, 2] == 9.0 and res[0] == 8.0                       # bump saw m(2,3) = 8 and incremented it in place
    assert m[0, 1] == -1.0 and m[0, 3] == -2.0 and m[0, 2] == 7.0  # fill wrote through the section
    assert list(res[1:]) == [-1.0, -7.0, 1.0, -7.0, 7.0, -7.0]
    assert list(ires[:2]) == [4, 8]
    assert list(ires[2:6]) == [0, 1, 0, 2]                        # SAVEd counter survived between calls
    assert ires[6] == 1 + 3 + 5
    assert ires[7] == -2 and ires[8] == 1
    assert ires[9] == 3 and list(ires[10:13]) == [2, 1, 3] and list(ires[13:16]) == [1, 4, 9]


def test_interpreter_types_and_bound_procedures(tmp_path):
    w, F = _world(tmp_path, """
    module m
       integer, parameter :: I4B = 4, I8B = 8
       type :: base_list
          integer(I8B) :: nenc = 0
          integer(I4B), dimension(:), allocatable :: index1
       end type base_list
       type, extends(base_list) :: child_list
          integer(I4B) :: extra = 5
       contains
          procedure :: one => list_one
          procedure :: two => list_two
          generic :: fill => one, two
       end type child_list
    contains
       subroutine list_one(self, n)
          class(child_list), intent(inout) :: self
          integer(I4B), intent(in) :: n
          allocate(self%index1(n))
          self%index1(:) = n
          self%nenc = n
       end subroutine list_one
       subroutine list_two(self, n, k)
          class(child_list), intent(inout) :: self
          integer(I4B), intent(in) :: n, k
          call self%one(n)
          self%index1(:) = k
       end subroutine list_two
       subroutine reset(l)
          class(child_list), intent(out) :: l
       end subroutine reset
       subroutine main(n, out)
          integer(I4B), intent(in) :: n
          integer(I4B), dimension(:), intent(out) :: out
          type(child_list), dimension(n) :: lists
          type(child_list) :: single
          integer(I4B) :: i
          do i = 1, n
             if (mod(i, 2) == 0) then
                call lists(i)%fill(i)
             else
                call lists(i)%fill(i, -i)
             end if
          end do
          associate(cnt => lists(:)%nenc)
             out(1) = sum(cnt(:))
          end associate
          out(2) = lists(3)%index1(2)
          out(3) = lists(4)%index1(4)
          out(4) = lists(2)%extra
          where (lists(:)%nenc > 2) lists(:)%nenc = 0
          out(5) = sum(lists(:)%nenc)
          call single%fill(3)
          call reset(single)
          out(6) = single%nenc
          if (allocated(single%index1)) out(7) = 1
       end subroutine main
    end module m
    """)
    out = np.zeros(7, dtype=np.int32)
    w.call("main", 4, out)
    assert list(out) == [10, -3, 4, 5, 3, 0, 0]


def test_interpreter_runs_quicksort_like_the_reference_uses_it(tmp_path):
    """Recursive procedures on array sections with optional arguments (the shape of base_util_sort_qsort_*)."""
    w, F = _world(tmp_path, """
    module m
       integer, parameter :: DP = 8, I4B = 4
    contains
       recursive subroutine qs(arr, ind)
          real(DP), dimension(:), intent(inout) :: arr
          integer(I4B), dimension(:), intent(inout), optional :: ind
          integer(I4B) :: i, j, n
          real(DP) :: x
          n = size(arr)
          if (n <= 1) return
          x = arr(n)
          i = 0
          do j = 1, n - 1
             if (arr(j) <= x) then
                i = i + 1
                call swap(arr(i), arr(j))
                if (present(ind)) call iswap(ind(i), ind(j))
             end if
          end do
          call swap(arr(i+1), arr(n))
          if (present(ind)) then
             call iswap(ind(i+1), ind(n))
             call qs(arr(:i), ind(:i))
             call qs(arr(i+2:), ind(i+2:))
          else
             call qs(arr(:i))
             call qs(arr(i+2:))
          end if
       end subroutine qs
       pure subroutine swap(a, b)
          real(DP), intent(inout) :: a, b
          real(DP) :: t
          t = a; a = b; b = t
       end subroutine swap
       pure subroutine iswap(a, b)
          integer(I4B), intent(inout) :: a, b
          integer(I4B) :: t
          t = a; a = b; b = t
       end subroutine iswap
    end module m
    """)
    rng = np.random.default_rng(0)
    a = rng.normal(size=200)
    ind = np.arange(1, 201, dtype=np.int32)
    b = a.copy()
    w.call("qs", b, ind)
    assert np.array_equal(b, np.sort(a)) and np.array_equal(a[ind - 1], b)
    c = a.copy()
    w.call("qs", c)
    assert np.array_equal(c, b)


def test_interpreter_resolves_generics_by_kind_and_by_declaration(tmp_path):
    """util_sort has real(SP)/real(DP) and I4B/I8B-index specifics that differ only in kind; an unallocated allocatable
    actual must be matched through its declaration."""
    w, F = _world(tmp_path, """
    module m
       integer, parameter :: SP = 4, DP = 8, I4B = 4, I8B = 8
       interface which
          module procedure which_sp
          module procedure which_dp
          module procedure which_i4_i8
          module procedure which_i4_i4
       end interface
    contains
       subroutine which_sp(a, r)
          real(SP), dimension(:), intent(in) :: a
          integer(I4B), intent(out) :: r
          r = 1
       end subroutine which_sp
       subroutine which_dp(a, r)
          real(DP), dimension(:), intent(in) :: a
          integer(I4B), intent(out) :: r
          r = 2
       end subroutine which_dp
       subroutine which_i4_i4(a, ind, r)
          integer(I4B), dimension(:), intent(in) :: a
          integer(I4B), dimension(:), allocatable, intent(inout) :: ind
          integer(I4B), intent(out) :: r
          r = 3
       end subroutine which_i4_i4
       subroutine which_i4_i8(a, ind, r)
          integer(I4B), dimension(:), intent(in) :: a
          integer(I8B), dimension(:), allocatable, intent(inout) :: ind
          integer(I4B), intent(out) :: r
          r = 4
          allocate(ind(size(a)))
       end subroutine which_i4_i8
       subroutine main(x, k, out)
          real(DP), dimension(:), intent(in) :: x
          integer(I4B), dimension(:), intent(in) :: k
          integer(I4B), dimension(:), intent(out) :: out
          integer(I8B), dimension(:), allocatable :: ind8
          integer(I4B), dimension(:), allocatable :: ind4
          call which(x, out(1))
          call which(k, ind8, out(2))
          call which(k, ind4, out(3))
          call which(k(2:3), ind8, out(4))
          out(5) = size(ind8)
       end subroutine main
    end module m
    """)
    out = np.zeros(5, dtype=np.int32)
    w.call("main", np.zeros(4), np.arange(4, dtype=np.int32), out)
    assert list(out) == [2, 4, 3, 4, 2]


def test_goldens_executed_every_reference_procedure_on_the_path():
    """tests/golden/fortran_coverage.json: how often each reference procedure ran while the goldens were generated."""
    import json
    cov = json.load(open(os.path.join(GOLD, "fortran_coverage.json")))
    need = {
        "kick": ["swiftest_kick_getacch_int_all_flat_rad_pl", "swiftest_kick_getacch_int_all_flat_norad_pl",
                 "swiftest_kick_getacch_int_all_tri_rad_pl", "swiftest_kick_getacch_int_all_tri_norad_pl",
                 "swiftest_kick_getacch_int_all_tp", "swiftest_kick_getacch_int_one_pl", "swiftest_kick_getacch_int_one_tp",
                 "swiftest_util_flatten_eucl_plpl"],
        "drift": ["swiftest_drift_all", "swiftest_drift_one", "swiftest_drift_dan", "swiftest_drift_kepmd", "swiftest_drift_kepu",
                  "swiftest_drift_kepu_guess", "swiftest_drift_kepu_new", "swiftest_drift_kepu_lag", "swiftest_drift_kepu_fchk",
                  "swiftest_drift_kepu_p3solve", "swiftest_drift_kepu_stumpff", "swiftest_orbel_scget"],
        "encounter": ["encounter_check_all_plpl", "encounter_check_all_plplm", "encounter_check_all_pltp",
                      "encounter_check_all_sort_and_sweep_plpl", "encounter_check_all_sort_and_sweep_plplm",
                      "encounter_check_all_sort_and_sweep_pltp", "encounter_check_all_sweep_one", "encounter_check_one",
                      "encounter_check_collapse_ragged_list", "encounter_check_remove_duplicates",
                      "encounter_check_sort_aabb_1d", "encounter_check_sweep_aabb_single_list",
                      "encounter_check_sweep_aabb_double_list", "encounter_util_setup_aabb", "swiftest_util_index_array",
                      "base_util_sort_index_dp", "base_util_sort_qsort_dp", "base_util_sort_partition_dp",
                      "base_util_sort_index_i4b_i8bind", "encounter_check_all_triangular_plpl",
                      "encounter_check_all_triangular_pltp", "encounter_check_all_triangular_plplm"],
    }
    need["steps"] = ["helio_step_pl", "helio_step_tp", "helio_kick_vb_pl", "helio_kick_vb_tp", "helio_kick_getacch_pl",
                     "helio_kick_getacch_tp", "helio_drift_linear_pl", "helio_drift_linear_tp", "helio_drift_body",
                     "swiftest_util_coord_vh2vb_pl", "swiftest_util_coord_vb2vh_pl", "swiftest_util_coord_vh2vb_tp",
                     "swiftest_util_coord_vb2vh_tp", "swiftest_kick_getacch_int_pl", "swiftest_kick_getacch_int_tp",
                     "whm_step_pl", "whm_step_tp", "whm_kick_vh_pl", "whm_kick_vh_tp", "whm_kick_getacch_pl", "whm_kick_getacch_tp",
                     "whm_kick_getacch_ah0", "whm_kick_getacch_ah1", "whm_kick_getacch_ah2", "whm_coord_h2j_pl", "whm_coord_j2h_pl",
                     "whm_coord_vh2vj_pl", "whm_drift_pl", "whm_util_set_mu_eta_pl", "swiftest_drift_all"]
    need["lists"] = ["swiftest_util_get_energy_and_momentum_system", "swiftest_util_get_potential_energy_flat",
                     "swiftest_util_get_potential_energy_triangular", "symba_kick_list_plpl", "symba_kick_list_pltp",
                     "symba_encounter_check_list_plpl", "symba_encounter_check_list_pltp", "symba_util_set_renc",
                     "collision_check_one", "swiftest_orbel_xv2aeq", "swiftest_discard_pl_close", "operator_cross_dp"]
    for part, names in need.items():
        for n in names:
            assert cov[part].get(n, 0) > 0, (part, n)
    assert not any("_sp" in n for n in cov["encounter"])          # the double-precision sort specifics, not real(SP)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree only exists in the build container")
def test_committed_goldens_regenerate_from_the_reference_tree(gk, gd):
    """Build container only: re-run the interpreter on the reference source for one gravity case and one drift batch and
    compare with the committed vectors (generator and fixtures cannot drift apart)."""
    import sys
    sys.path.insert(0, GOLD)
    import gen_golden_fortran as G
    w = G.world()
    r, Gm, radius, acc0 = gk["tiny5_r"], gk["tiny5_Gm"], gk["tiny5_radius"], gk["tiny5_acc0"]
    acc = G.fa(acc0)
    w.call("swiftest_kick_getacch_int_all_tri_rad_pl", 5, 5, G.fa(r), Gm.copy(), radius.copy(), acc)
    assert_bit_equal(G.back(acc), gk["tiny5_tri_rad_nplm5"], "regenerated tiny5")
    n = 24
    param = w.new_object("swiftest_parameters")
    param.c["lgr"] = False
    x, v = G.fa(gd["drift_long_x0"][:n]), G.fa(gd["drift_long_v0"][:n])
    fl = np.full(n, -7, dtype=np.int32)
    w.call("swiftest_drift_all", gd["drift_long_mu"][:n].copy(), x, v, n, param, float(gd["drift_long_dt"]),
           gd["drift_long_lmask"][:n].copy(), fl)
    assert_bit_equal(G.back(x), gd["drift_long_x1"][:n], "regenerated drift x")
    assert np.array_equal(fl, gd["drift_long_iflag"][:n])


# ------------------------------------------------------------------------------------------------------- whole steps
@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(GOLD, "fortran_steps.npz"))


STEP_CASES = ["p8", "p8flat", "p8mask", "p30"]


def _run_oracle_steps(oracle, gs, kind, tag):
    key = "%s_%s_" % (kind, tag)
    g = lambda k: gs[key + k]
    GMcb, dt, nsteps, lflat = float(g("GMcb")), float(g("dt")), int(g("nsteps")), bool(g("lflat"))
    Gm, radius = g("pl_Gmass"), g("pl_radius")
    lm_pl, lm_tp = g("lmask_pl").astype(np.int32), g("lmask_tp").astype(np.int32)
    pl = dict(rh=g("pl_rh0").copy(), vh=g("pl_vh0").copy())
    tp = dict(rh=g("tp_rh0").copy(), vh=g("tp_vh0").copy())
    if kind == "helio":
        pl["vb"], tp["vb"] = np.zeros_like(pl["rh"]), np.zeros_like(tp["rh"])
    step_pl = oracle.helio_step_pl if kind == "helio" else oracle.whm_step_pl
    step_tp = oracle.helio_step_tp if kind == "helio" else oracle.whm_step_tp
    for s in range(nsteps):
        assert not step_pl(pl, GMcb, Gm, radius, dt, lflat=lflat, lmask=lm_pl).any()
        assert not step_tp(tp, pl, GMcb, Gm, dt, lmask=lm_tp)[lm_tp == 1].any()
        yield s, pl, tp


@pytest.mark.parametrize("tag", STEP_CASES)
def test_helio_steps_are_bit_identical_to_the_fortran(oracle, gs, tag):
    """helio_step_pl + helio_step_tp (helio/helio_step.f90:37-123 and everything below: vh2vb, lindrift, kick_vb, getacch,
    drift, vb2vh), 5 consecutive steps of planets and test particles, tri and flat loops, with masked bodies."""
    for s, pl, tp in _run_oracle_steps(oracle, gs, "helio", tag):
        for nm, st in (("pl", pl), ("tp", tp)):
            for q in ("rh", "vh", "vb"):
                assert_bit_equal(st[q], gs["helio_%s_%s_%s" % (tag, nm, q)][s], "helio %s step %d %s %s" % (tag, s, nm, q))


@pytest.mark.parametrize("tag", STEP_CASES)
def test_whm_steps_are_bit_identical_to_the_fortran(oracle, gs, tag):
    """whm_step_pl + whm_step_tp (whm/whm_step.f90:37-100 and below: kick_vh, getacch ah0/ah1/ah2, h2j, vh2vj, Jacobi drift,
    j2h), 5 consecutive steps."""
    last = None
    for s, pl, tp in _run_oracle_steps(oracle, gs, "whm", tag):
        for nm, st in (("pl", pl), ("tp", tp)):
            for q in ("rh", "vh"):
                assert_bit_equal(st[q], gs["whm_%s_%s_%s" % (tag, nm, q)][s], "whm %s step %d %s %s" % (tag, s, nm, q))
        last = pl
    assert_bit_equal(last["eta"], gs["whm_%s_eta" % tag], "eta")
    assert_bit_equal(last["muj"], gs["whm_%s_muj" % tag], "muj")
    assert_bit_equal(last["xj"], gs["whm_%s_xj" % tag], "xj")
    assert_bit_equal(last["vj"], gs["whm_%s_vj" % tag], "vj")


# ------------------------------------------------------------------------------------------------------- energy, lists
@pytest.fixture(scope="module")
def gl():
    return np.load(os.path.join(GOLD, "fortran_lists.npz"))


@pytest.mark.parametrize("tag,flat,lclose", [("tri", False, True), ("flat", True, True), ("tri_noclose", False, False)])
def test_energy_and_momentum_are_bit_identical_to_the_fortran(oracle, gl, tag, flat, lclose):
    """swiftest_util_get_energy_and_momentum_system with the flat and the triangular potential loop (swiftest_util.f90:1172-1394)."""
    lmask = (gl["en_status"] != 1).astype(np.int32)                 # INACTIVE = 1 (globals_module.f90)
    assert lmask.sum() == len(lmask) - 1
    got = oracle.get_energy_and_momentum(float(gl["en_GMcb"]), float(gl["en_mass_cb"]), gl["en_rbcb"], gl["en_vbcb"], gl["en_Gmass"],
                                         gl["en_mass"], gl["en_radius"], gl["en_rb"], gl["en_vb"], lmask=lmask, lclose=lclose, flat=flat)
    for k, ok in (("ke_orbit", "ke_orbit"), ("pe", "pe"), ("be", "be"), ("te", "te"), ("gmtot", "GMtot")):
        assert_bit_equal(np.array([got[ok]]), np.array([float(gl["en_%s_%s" % (tag, k)])]), "%s %s" % (tag, k))
    assert_bit_equal(got["L_orbit"], gl["en_%s_l_orbit" % tag], tag + " L")


@pytest.mark.parametrize("irec,sgn", [(1, 1), (1, -1), (2, 1), (2, -1), (3, 1)])
def test_symba_kick_list_plpl_is_bit_identical_to_the_fortran(oracle, gl, irec, sgn):
    """symba_kick_list_plpl (symba/symba_kick.f90:126-235): level mask, shell regimes, serial accumulation, kick."""
    vb, lgood, ah = oracle.symba_kick_list_plpl(gl["kl_index1"], gl["kl_index2"], gl["kl_lactive"].astype(np.int32), gl["kl_levelg"],
                                               gl["kl_rh"], gl["kl_rhill"], gl["kl_Gmass"], float(gl["kl_dt"]), irec, sgn,
                                               gl["kl_vb0"], ah=gl["kl_ah0"])
    assert_bit_equal(vb, gl["kl_plpl_irec%d_sgn%d_vb" % (irec, sgn)], "vb")
    assert_bit_equal(ah, gl["kl_plpl_irec%d_sgn%d_ah" % (irec, sgn)], "ah")
    assert not np.array_equal(vb, gl["kl_vb0"])


def test_symba_kick_list_goldens_cover_all_three_regimes(gl):
    """inside the inner shell (pair dropped), in the shell (smoothed factor through pow), outside (plain r^-3)."""
    RHSCALE, RSHELL = 6.5, 0.48075
    r, rh = gl["kl_rh"], gl["kl_rhill"]
    i, j = gl["kl_index1"] - 1, gl["kl_index2"] - 1
    r2 = ((r[j] - r[i]) ** 2).sum(1)
    ri = (rh[i] + rh[j]) ** 2 * RHSCALE ** 2 * RSHELL ** 2
    assert (r2 < ri * RSHELL ** 2).sum() > 3 and ((r2 >= ri * RSHELL ** 2) & (r2 < ri)).sum() > 3 and (r2 >= ri).sum() > 3


@pytest.mark.parametrize("irec,sgn", [(1, 1), (2, -1), (2, 1)])
def test_symba_kick_list_pltp_is_bit_identical_to_the_fortran(oracle, gl, irec, sgn):
    """symba_kick_list_pltp (symba/symba_kick.f90:237-337)."""
    vb, lgood, ah = oracle.symba_kick_list_pltp(gl["kt_index1"], gl["kt_index2"], gl["kt_lactive"].astype(np.int32), gl["kl_levelg"],
                                               gl["kt_levelg_tp"], gl["kl_rh"], gl["kl_rhill"], gl["kl_Gmass"], gl["kt_rh_tp"],
                                               float(gl["kl_dt"]), irec, sgn, gl["kt_vb0"])
    assert_bit_equal(vb, gl["kt_pltp_irec%d_sgn%d_vb" % (irec, sgn)], "vb")
    ref_ah = gl["kt_pltp_irec%d_sgn%d_ah" % (irec, sgn)]
    touched = (ref_ah != gl["kt_ah0"]).any(1)               # the reference zeroes ah of the particles it kicked
    assert touched.sum() > 5 and np.all(ref_ah[touched] == 0.0)
    assert np.all(ah[touched] == 0.0) and np.all(ah[~touched] == 123.0)      # the wrapper pre-fills ah with 123


@pytest.mark.parametrize("irec", [1, 2])
def test_symba_encounter_check_list_is_identical_to_the_fortran(oracle, gl, irec):
    """symba_encounter_check_list_plpl/_pltp with symba_util_set_renc (symba/symba_encounter_check.f90:88-235)."""
    renc = oracle.set_renc(gl["kl_rhill"], irec)
    assert_bit_equal(renc, gl["el_plpl_irec%d_renc" % irec], "renc")
    for kind, i1, i2, lact in (("plpl", gl["kl_index1"], gl["kl_index2"], gl["kl_lactive"]),
                               ("pltp", gl["kt_index1"], gl["kt_index2"], gl["kt_lactive"])):
        level0, level1 = gl["el_%s_irec%d_level0" % (kind, irec)], gl["el_%s_irec%d_level" % (kind, irec)]
        mask = lact & (level0 == irec - 1)
        want = mask & (level1 == irec)
        if kind == "plpl":
            lenc, lvd, n = oracle.symba_encounter_check_list(i1, i2, mask.astype(np.int32), gl["kl_rh"], gl["el_vb_pl"], renc,
                                                             gl["el_radius"], 0.05)
        else:
            lenc, lvd, n = oracle.symba_encounter_check_list(i1, i2, mask.astype(np.int32), gl["kl_rh"], gl["el_vb_pl"], renc,
                                                             gl["el_radius"], 0.05, r2=gl["kt_rh_tp"], v2=gl["el_vb_tp"])
        assert np.array_equal(lenc.astype(bool), want), kind
        assert np.array_equal(lvd.astype(bool)[mask], gl["el_%s_irec%d_lvdotr" % (kind, irec)][mask]), kind
        assert n == want.sum() and bool(gl["el_%s_irec%d_lany" % (kind, irec)]) == bool(want.any())
        assert 0 < want.sum() < mask.sum()
    lg = np.zeros(len(gl["kl_rhill"]), np.int32)
    want = (gl["kl_lactive"] & (gl["el_plpl_irec%d_level0" % irec] == irec - 1)) & (gl["el_plpl_irec%d_level" % irec] == irec)
    lg[gl["kl_index1"][want] - 1] = irec
    lg[gl["kl_index2"][want] - 1] = irec
    assert np.array_equal(lg, gl["el_plpl_irec%d_levelg" % irec])


def test_collision_check_one_and_discard_pl_close_are_identical_to_the_fortran(oracle, gl):
    """collision_check_one + swiftest_orbel_xv2aeq (collision_check.f90:16-57, swiftest_orbel.f90:700-764) through the
    list wrapper (one pair per row), and swiftest_discard_pl_close (swiftest_discard.f90:295-337)."""
    import ctypes as C
    m = len(gl["cc_rlim"])
    idx = np.arange(1, m + 1, dtype=np.int32)
    z = np.zeros((m, 3))
    lcol, lclo, n = oracle.collision_check_list(idx, idx, None, gl["cc_lvdotr"].astype(np.int32), gl["cc_rel"], gl["cc_vel"],
                                                gl["cc_Gmtot"], gl["cc_rlim"], float(gl["cc_dt"]), r2=z, v2=z)
    assert np.array_equal(lcol.astype(bool), gl["cc_lcollision"]) and np.array_equal(lclo.astype(bool), gl["cc_lclosest"])
    inside = (gl["cc_rel"] ** 2).sum(1) <= gl["cc_rlim"] ** 2
    assert (gl["cc_lcollision"] & ~inside).sum() > 3          # collisions predicted through the pericentre distance q
    f = oracle.lib.swo_discard_pl_close
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    f.restype = None
    for k in range(m):
        dx, dv = np.ascontiguousarray(gl["cc_rel"][k]), np.ascontiguousarray(gl["cc_vel"][k])
        fl, rm = C.c_int32(-1), C.c_double(-1.0)
        f(dx.ctypes.data, dv.ctypes.data, float(gl["cc_dt"]), float(gl["cc_rlim"][k]) ** 2, C.byref(fl), C.byref(rm))
        assert fl.value == gl["dc_iflag"][k], k
        if not np.isnan(gl["dc_r2min"][k]):                   # the reference leaves r2min undefined on its early exits
            assert np.float64(rm.value).view(np.uint64) == gl["dc_r2min"][k].view(np.uint64), k
    assert 10 < gl["dc_iflag"].sum() < m - 10 and 50 < np.isfinite(gl["dc_r2min"]).sum() < m
