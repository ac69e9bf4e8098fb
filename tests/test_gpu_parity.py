"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Bars (north star): encounter pair lists BIT-EXACT after canonical (index1,index2) ordering; accelerations within
1e-12 relative per component, relative to the per-component sum of |terms| (summation order and FMA contraction
differ; SURVEY.md section 7 "Tolerance definition"); drift BIT-EXACT on the kepmd path and the series-guess kepu
path, 1e-12 where CUDA's sin()/pow() stand in for the host libm (Danby and hyperbolic initial guesses).
"""
import os

import numpy as np
import pytest

from swiftest_b200 import workloads as W
from swiftest_b200 import PL, TP, LOOP_FLAT, LOOP_TRIANGULAR

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-12          # relative to sum_j |term_j| per component
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _scaled(a, ref, scale):
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - ref) / scale))


def _fixture108():
    f = W.fixture("108pl_50tp")
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    nplm = int((f["pl_Gmass"] >= float(f["GMTINY"])).sum())
    pl = {k: f["pl_" + k][order] for k in ("rh", "vh", "Gmass", "radius", "rhill")}
    return f, pl, nplm


# ---------------------------------------------------------------------------------------------- gravity
@pytest.mark.parametrize("lrad", [True, False])
@pytest.mark.parametrize("nplm_kind", ["all", "gmtiny", "lmtiny"])
def test_kick_tri_fixture_108pl(ctx, oracle, lrad, nplm_kind):
    f, pl, nplm57 = _fixture108()
    nplm = {"all": 108, "gmtiny": nplm57, "lmtiny": 20}[nplm_kind]
    rad = pl["radius"] if lrad else None
    acc0 = np.random.default_rng(1).normal(scale=1e-3, size=(108, 3))
    ref = oracle.kick_tri_pl(pl["rh"], pl["Gmass"], rad, acc0, nplm=nplm)
    got = acc0.copy()
    ctx.kick_getacch_int_all_tri_pl(108, nplm, pl["rh"], pl["Gmass"], rad, got)
    scale = oracle.kick_tri_abs_scale(pl["rh"], pl["Gmass"], rad, nplm=nplm) + np.abs(acc0)
    assert _scaled(got, ref, scale) < ACC_TOL


@pytest.mark.parametrize("n", [1, 2, 31, 257, 1000, 4099])
def test_kick_tri_disk_sizes(ctx, oracle, n):
    d = W.disk(n, seed=100 + n)
    ref = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)))
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_tri_pl(n, n, d["rh"], d["Gmass"], d["radius"], got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
    assert _scaled(got, ref, scale) < ACC_TOL


def test_kick_tri_1e4_disk_with_gmtiny_split(ctx, oracle):
    n, nplm = 10000, 3000
    d = W.disk(n, seed=3031179)
    ref = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)), nplm=nplm)
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_tri_pl(n, nplm, d["rh"], d["Gmass"], d["radius"], got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"], nplm=nplm)
    assert _scaled(got, ref, scale) < ACC_TOL


@pytest.mark.parametrize("lrad", [True, False])
def test_kick_flat_canonical_pairs(ctx, oracle, lrad):
    n = 1500
    d = W.disk(n, seed=77)
    rad = d["radius"] if lrad else None
    for nplm in (n, 400):
        nplpl = oracle.nplplm(n, nplm)
        ref = oracle.kick_flat_pl(d["rh"], d["Gmass"], rad, np.zeros((n, 3)), nplpl=nplpl)
        got = np.zeros((n, 3))
        ctx.kick_getacch_int_all_flat_pl(n, nplpl, None, d["rh"], d["Gmass"], rad, got)
        scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], rad, nplm=nplm)
        assert _scaled(got, ref, scale) < ACC_TOL


@pytest.mark.parametrize("n,nplm", [(2, 2), (127, 127), (128, 128), (129, 129), (256, 100), (1024, 1024), (1025, 512),
                                    (4099, 4099), (4224, 4224), (4224, 129), (5000, 1), (9000, 8999)])
def test_kick_flat_third_law_block_coverage(ctx, oracle, n, nplm):
    """Every block-pair shape of the third-law kernel: odd/even block counts, ragged last block, nplm on and off
    block boundaries, a single owner block."""
    d = W.disk(n, seed=n + nplm)
    nplpl = oracle.nplplm(n, nplm)
    acc0 = np.random.default_rng(n).normal(scale=1e-5, size=(n, 3))
    ref = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], acc0, nplm=nplm)
    got = acc0.copy()
    ctx.kick_getacch_int_all_flat_pl(n, nplpl, None, d["rh"], d["Gmass"], d["radius"], got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"], nplm=nplm) + np.abs(acc0)
    assert _scaled(got, ref, scale) < ACC_TOL
    # momentum conservation: the third-law kernel applies equal and opposite kicks
    if nplm == n:
        p = (d["Gmass"][:, None] * (got - acc0)).sum(0)
        assert np.max(np.abs(p)) < 1e-11 * np.abs(d["Gmass"][:, None] * (got - acc0)).sum()


def test_kick_flat_rejects_non_canonical_count(ctx):
    from swiftest_b200 import SwcuError
    d = W.disk(50, seed=1)
    with pytest.raises(SwcuError):
        ctx.kick_getacch_int_all_flat_pl(50, 7, None, d["rh"], d["Gmass"], d["radius"], np.zeros((50, 3)))


def test_kick_flat_explicit_pair_table(ctx, oracle):
    n = 300
    d = W.disk(n, seed=8)
    rng = np.random.default_rng(8)
    pairs = np.array([(i, j) for i in range(1, n + 1) for j in range(i + 1, n + 1)], np.int32)
    k = pairs[rng.choice(len(pairs), 5000, replace=False)]
    acc0 = rng.normal(scale=1e-4, size=(n, 3))
    ref = oracle.kick_flat_pl(d["rh"], d["Gmass"], d["radius"], acc0, k_plpl=k)
    got = acc0.copy()
    ctx.kick_getacch_int_all_flat_pl(n, len(k), k, d["rh"], d["Gmass"], d["radius"], got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"]) + np.abs(acc0)
    assert _scaled(got, ref, scale) < ACC_TOL


def test_kick_overlapping_radii_and_coincident_bodies(ctx, oracle):
    """radius check excludes touching pairs; exactly coincident bodies are skipped (r2 = 0 is not > rlim2)."""
    r = np.array([[0.0, 0, 0], [1e-3, 0, 0], [1.0, 0, 0], [1.0, 0, 0], [0, 2.0, 0]])
    Gm = np.array([1e-3, 2e-3, 1e-4, 3e-4, 5e-5])
    rad = np.array([1e-3, 1e-3, 1e-5, 1e-5, 1e-5])
    ref = oracle.kick_tri_pl(r, Gm, rad, np.zeros((5, 3)))
    got = np.zeros((5, 3))
    ctx.kick_getacch_int_all_tri_pl(5, 5, r, Gm, rad, got)
    assert np.all(np.isfinite(got))
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.abs(ref).max()


def test_kick_extreme_separations_take_ieee_path(ctx, oracle):
    """r^2 outside the normal FP32 range must not break the FP32-seeded inverse square root."""
    r = np.array([[0.0, 0, 0], [1e-25, 0, 0], [1e25, 0, 0], [0, 3e-21, 0]])
    Gm = np.array([1.0, 2.0, 3.0, 4.0])
    ref = oracle.kick_tri_pl(r, Gm, None, np.zeros((4, 3)))
    got = np.zeros((4, 3))
    ctx.kick_getacch_int_all_tri_pl(4, 4, r, Gm, None, got)
    assert np.all(np.isfinite(got))
    scale = oracle.kick_tri_abs_scale(r, Gm, None)
    assert _scaled(got, ref, scale) < ACC_TOL


@pytest.mark.parametrize("far", [1e15, 1e20, 1e30])
def test_kick_coordinate_guard_for_both_kernels(ctx, oracle, far):
    """The fast path drops the r^2 < 2^128 test only when every |coordinate| < 2^62; a body beyond that (or r^2 beyond
    the FP32 exponent range) must send the pairs through the checked / IEEE paths in both kernels."""
    n = 700
    d = W.disk(n, seed=3)
    r = d["rh"].copy()
    r[17] = [far, -0.5 * far, 0.25 * far]
    r[400] = [-far, far, far]
    for flat in (False, True):
        ref = oracle.kick_tri_pl(r, d["Gmass"], d["radius"], np.zeros((n, 3)))
        got = np.zeros((n, 3))
        if flat:
            ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, r, d["Gmass"], d["radius"], got)
        else:
            ctx.kick_getacch_int_all_tri_pl(n, n, r, d["Gmass"], d["radius"], got)
        assert np.all(np.isfinite(got))
        scale = oracle.kick_tri_abs_scale(r, d["Gmass"], d["radius"])
        assert _scaled(got, ref, scale) < ACC_TOL


def test_kick_flat_rollback_of_chunks_with_close_pairs(ctx, oracle):
    """The third-law fast path carries no per-pair test: a chunk that met a pair inside the block pair's radius bound,
    a coincident pair or a self pair outside a diagonal block is rolled back and redone with the IEEE expression.  Close
    pairs are planted inside one block, across two blocks and against the ragged last block."""
    n = 2100
    d = W.disk(n, seed=21)
    r, rad = d["rh"].copy(), d["radius"].copy()
    rmax = rad.max()
    plant = [(5, 40, 0.5), (10, 700, 1.5), (130, 2090, 0.0), (1500, 1501, 1.99), (64, 2050, 2.5)]  # (i, j, distance / rmax)
    for i, j, f in plant:
        r[j] = r[i] + np.array([f * rmax, 0.0, 0.0])
    n0 = ctx.flat_redo_count()
    ref = oracle.kick_tri_pl(r, d["Gmass"], rad, np.zeros((n, 3)))
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, r, d["Gmass"], rad, got)
    assert np.all(np.isfinite(got))
    scale = oracle.kick_tri_abs_scale(r, d["Gmass"], rad)
    assert _scaled(got, ref, scale) < ACC_TOL
    redone = ctx.flat_redo_count() - n0
    assert 4 <= redone <= 64, redone    # the planted pairs (the 2.5 rmax one is outside every bound), not the whole run
    # without the radius test only the coincident pair is special (r^2 = 0: the reference yields inf/NaN there, so
    # move it apart first)
    r[2090] = r[130] + np.array([1e-9, 0.0, 0.0])
    n0 = ctx.flat_redo_count()
    ref = oracle.kick_tri_pl(r, d["Gmass"], None, np.zeros((n, 3)))
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, r, d["Gmass"], None, got)
    scale = oracle.kick_tri_abs_scale(r, d["Gmass"], None)
    assert _scaled(got, ref, scale) < ACC_TOL
    assert ctx.flat_redo_count() == n0


def test_kick_flat_radius_bound_is_per_block(ctx, oracle):
    """One giant body must not push the whole population onto the exact path: the fast-path threshold of a block pair
    comes from the largest radii INSIDE the two blocks."""
    n = 3000
    d = W.disk(n, seed=22)
    rad = d["radius"].copy()
    rad[3] = 2e-2                      # a "planet" as large as the spacing of the disk bodies, in block 0
    n0 = ctx.flat_redo_count()
    ref = oracle.kick_tri_pl(d["rh"], d["Gmass"], rad, np.zeros((n, 3)))
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, d["rh"], d["Gmass"], rad, got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], rad)
    assert _scaled(got, ref, scale) < ACC_TOL
    redone = ctx.flat_redo_count() - n0
    nb = -(-n // 128)
    assert 0 < redone <= 4 * nb, (redone, nb)    # only chunks of block pairs that contain block 0


def test_kick_flat_plain_disk_never_leaves_the_fast_path(ctx):
    n = 4000
    d = W.disk(n, seed=23)
    n0 = ctx.flat_redo_count()
    ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)))
    assert ctx.flat_redo_count() == n0


def test_kick_flat_overflowing_r2_takes_exact_path(ctx, oracle):
    """|coordinate| beyond 2^500: r^2 overflows to inf, the reference's 1/(r2*sqrt(r2)) gives 0 for those pairs."""
    n = 300
    d = W.disk(n, seed=24)
    r = d["rh"].copy()
    r[7] = [1e160, 0.0, -1e158]
    ref = oracle.kick_tri_pl(r, d["Gmass"], d["radius"], np.zeros((n, 3)))
    got = np.zeros((n, 3))
    ctx.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, r, d["Gmass"], d["radius"], got)
    assert np.all(np.isfinite(got))
    scale = oracle.kick_tri_abs_scale(r, d["Gmass"], d["radius"])
    assert _scaled(got, ref, scale) < ACC_TOL


@pytest.mark.parametrize("ntp,npl", [(50, 108), (1000, 8), (100000, 8), (3000, 700)])
def test_kick_tp(ctx, oracle, ntp, npl):
    rng = np.random.default_rng(ntp + npl)
    if npl == 108:
        f, pl, _ = _fixture108()
        rtp, rpl, Gm = f["tp_rh"], pl["rh"], pl["Gmass"]
    elif npl == 8:
        p = W.planets8_year_units()
        rtp, rpl, Gm = W.tp_cloud(ntp, seed=ntp)["rh"], p["rh"], p["Gmass"]
    else:
        d = W.disk(npl, seed=4)
        rtp, rpl, Gm = W.tp_cloud(ntp, seed=ntp, a_lo=0.3, a_hi=2.0)["rh"], d["rh"], d["Gmass"]
    mask = (rng.uniform(size=ntp) > 0.1).astype(np.int32)
    acc0 = rng.normal(scale=1e-6, size=(ntp, 3))
    ref = oracle.kick_all_tp(rtp, rpl, Gm, mask, acc0)
    got = acc0.copy()
    ctx.kick_getacch_int_all_tp(ntp, npl, rtp, rpl, Gm, mask, got)
    dd = rtp[:, None, :] - rpl[None, :, :] if ntp * npl < 5e6 else None
    if dd is not None:
        scale = (Gm[None, :, None] * np.abs(dd) / (np.linalg.norm(dd, axis=2) ** 3)[:, :, None]).sum(1) + np.abs(acc0)
    else:
        scale = np.abs(ref) + np.abs(acc0) + 1e-300
    assert _scaled(got, ref, scale) < ACC_TOL
    off = mask == 0
    assert np.array_equal(got[off], acc0[off])  # masked-out particles are untouched


def test_symba_subtract_encounter_pairs(ctx, oracle):
    n = 2000
    d = W.disk(n, seed=5)
    renc = d["rhill"] * 6.5 * 4
    i1, i2, _ = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"])
    assert len(i1) > 10
    full = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)))
    ref = oracle.symba_kick_subtract_enc(i1, i2, d["rh"], d["Gmass"], d["radius"], full)
    got = full.copy()
    ctx.symba_kick_subtract_encounters(n, i1, i2, d["rh"], d["Gmass"], d["radius"], got)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
    assert _scaled(got, ref, scale) < ACC_TOL


# ---------------------------------------------------------------------------------------------- drift
def _gpu_drift(ctx, mu, x, v, dt, mask=None, lgr=False, inv_c2=0.0):
    n = len(x)
    mu = np.full(n, mu) if np.isscalar(mu) else mu
    mask = np.ones(n, np.int32) if mask is None else mask
    x, v, fl = x.copy(), v.copy(), np.zeros(n, np.int32)
    ctx.drift_all(mu, x, v, n, dt, mask, fl, lgr=lgr, inv_c2=inv_c2)
    return x, v, fl


@pytest.mark.parametrize("dt", [6.0875 / 365.25, 0.01, 0.25, 3.0])
def test_drift_bit_exact_where_no_libm_is_involved(ctx, oracle, dt):
    tp = W.tp_cloud(20000, seed=31, a_lo=0.3, a_hi=40.0)
    xr, vr, fr = oracle.drift_all(W.GMSUN, tp["rh"], tp["vh"], dt)
    xg, vg, fg = _gpu_drift(ctx, W.GMSUN, tp["rh"], tp["vh"], dt)
    br = oracle.drift_branch(W.GMSUN, tp["rh"], tp["vh"], dt)
    exact = (br == 0) | (br == 1)
    assert np.array_equal(fg, fr)
    assert np.array_equal(xg[exact], xr[exact]) and np.array_equal(vg[exact], vr[exact])
    rest = ~exact
    if rest.any():
        rs = np.linalg.norm(xr[rest], axis=1, keepdims=True)
        vs = np.linalg.norm(vr[rest], axis=1, keepdims=True)
        assert np.max(np.abs(xg[rest] - xr[rest]) / rs) < 1e-12
        assert np.max(np.abs(vg[rest] - vr[rest]) / vs) < 1e-12


def test_drift_reference_golden_vectors(ctx):
    """The CUDA drift against the states produced by the reference's own Python two-body code."""
    g = np.load(os.path.join(GOLD, "drift_kepler_ref.npz"))
    for dt in np.unique(g["dt"]):
        m = g["dt"] == dt
        x, v, fl = _gpu_drift(ctx, g["mu"][m], g["x0"][m], g["v0"][m], float(dt))
        assert not fl.any()
        assert np.max(np.abs(x - g["x1"][m]) / np.linalg.norm(g["x1"][m], axis=1, keepdims=True)) < 1e-11
        assert np.max(np.abs(v - g["v1"][m]) / np.linalg.norm(g["v1"][m], axis=1, keepdims=True)) < 1e-11


def test_drift_hyperbolic_mask_gr_and_failure_flags(ctx, oracle):
    rng = np.random.default_rng(3)
    n = 4000
    tp = W.tp_cloud(n, seed=9, a_lo=0.3, a_hi=5.0)
    v = tp["vh"] * rng.uniform(0.7, 1.8, size=(n, 1))  # a good fraction unbound
    mask = (rng.uniform(size=n) > 0.2).astype(np.int32)
    inv_c2 = 1.0 / 63241.077 ** 2
    for lgr in (False, True):
        xr, vr, fr = oracle.drift_all(W.GMSUN, tp["rh"], v, 0.05, lmask=mask, lgr=lgr, inv_c2=inv_c2)
        xg, vg, fg = _gpu_drift(ctx, W.GMSUN, tp["rh"], v, 0.05, mask=mask, lgr=lgr, inv_c2=inv_c2)
        assert (oracle.drift_branch(W.GMSUN, tp["rh"], v, 0.05) == 3).sum() > 100
        assert np.array_equal(fg, fr)
        assert np.max(np.abs(xg - xr) / np.linalg.norm(xr, axis=1, keepdims=True)) < 1e-12
        assert np.max(np.abs(vg - vr) / np.linalg.norm(vr, axis=1, keepdims=True)) < 1e-12
        off = mask == 0
        assert np.array_equal(xg[off], tp["rh"][off]) and np.array_equal(vg[off], v[off])
    # iflag of masked-out bodies keeps the caller's value
    fl = np.full(n, 7, np.int32)
    x2, v2 = tp["rh"].copy(), v.copy()
    ctx.drift_all(np.full(n, W.GMSUN), x2, v2, n, 0.05, mask, fl)
    assert np.all(fl[mask == 0] == 7) and np.all(fl[mask == 1] == 0)


def test_drift_round_trip_1e6_bodies(ctx):
    """Full-size property (BASELINE tp config): drift(+dt) then drift(-dt) returns to the start; energy and angular
    momentum of every two-body orbit are conserved."""
    n = 1_000_000
    tp = W.tp_cloud(n, seed=123)
    x1, v1, f1 = _gpu_drift(ctx, W.GMSUN, tp["rh"], tp["vh"], 0.01)
    x2, v2, f2 = _gpu_drift(ctx, W.GMSUN, x1, v1, -0.01)
    assert not f1.any() and not f2.any()
    assert np.max(np.abs(x2 - tp["rh"]) / np.linalg.norm(tp["rh"], axis=1, keepdims=True)) < 1e-11
    e0 = 0.5 * (tp["vh"] ** 2).sum(1) - W.GMSUN / np.linalg.norm(tp["rh"], axis=1)
    e1 = 0.5 * (v1 ** 2).sum(1) - W.GMSUN / np.linalg.norm(x1, axis=1)
    assert np.max(np.abs((e1 - e0) / e0)) < 1e-11
    assert np.max(np.abs(np.cross(x1, v1) - np.cross(tp["rh"], tp["vh"]))) < 1e-10


# ---------------------------------------------------------------------------------------------- encounters
def _same_pairs(got, ref):
    n, g1, g2, glv = got
    r1, r2, rlv = ref
    assert n == len(r1)
    assert np.array_equal(g1, r1) and np.array_equal(g2, r2)
    assert glv.all() and rlv.all()


@pytest.mark.parametrize("n,boost", [(2, 1.0), (3, 1.0), (108, 3.0), (600, 4.0), (5000, 2.0), (20000, 1.0)])
def test_sweep_plpl_bit_exact(ctx, oracle, n, boost):
    if n == 108:
        f, pl, _ = _fixture108()
        r, v, renc, dt = pl["rh"], pl["vh"], pl["rhill"] * 6.5 * boost, 0.05
    else:
        d = W.disk(n, seed=n)
        r, v, renc, dt = d["rh"], d["vh"], d["rhill"] * 6.5 * boost, d["dt"]
    ref = oracle.encounter_plpl(r, v, renc, dt)
    nbox = oracle.nbox_total()
    got = ctx.encounter_check_all_sort_and_sweep_plpl(n, r, v, renc, dt)
    _same_pairs(got, ref)
    assert ctx.encounter_stats()["nbox_total"] == nbox
    if n >= 600:
        assert got[0] > 0


def test_sweep_bucket_sort_equals_radix_sort_and_falls_back_on_clumps(ctx, oracle):
    """The extents are bucket sorted (equal slices of [min, max], bitonic sort on (key, id) per bucket = the stable sort);
    SWCU_SWEEP_BUCKET=0 takes the radix sort: same list, same nbox.  3000 bodies at bit-identical |r| overflow a bucket:
    the call must notice and repeat itself with the radix sort; so must a call with a non-finite extent."""
    d = W.disk(6000, seed=77)
    r, v, renc = d["rh"].copy(), d["vh"], d["rhill"] * 6.5 * 3
    ref = oracle.encounter_plpl(r, v, renc, d["dt"])
    nbox = oracle.nbox_total()
    f0 = ctx.encounter_bucket_fallbacks()
    got = ctx.encounter_check_all_sort_and_sweep_plpl(6000, r, v, renc, d["dt"])
    _same_pairs(got, ref)
    assert ctx.encounter_stats()["nbox_total"] == nbox and ctx.encounter_bucket_fallbacks() == f0 and got[0] > 0
    os.environ["SWCU_SWEEP_BUCKET"] = "0"
    try:
        _same_pairs(ctx.encounter_check_all_sort_and_sweep_plpl(6000, r, v, renc, d["dt"]), ref)
        assert ctx.encounter_stats()["nbox_total"] == nbox
    finally:
        del os.environ["SWCU_SWEEP_BUCKET"]
    # a clump: 3000 bodies on one sphere, |r| = 1 exactly (unit vectors along the axes scaled by exact powers of two)
    rng = np.random.default_rng(5)
    axis = rng.integers(0, 3, 3000)
    r[:3000] = 0.0
    r[np.arange(3000), axis] = np.where(rng.uniform(size=3000) < 0.5, 1.0, -1.0)
    renc2 = renc.copy()
    renc2[:3000] = 0.0  # all 3000 begin AND end extents equal 1.0: 6000 equal keys, in position order after the sort
    ref = oracle.encounter_plpl(r, v, renc2, d["dt"])
    nbox = oracle.nbox_total()
    got = ctx.encounter_check_all_sort_and_sweep_plpl(6000, r, v, renc2, d["dt"])
    _same_pairs(got, ref)
    assert ctx.encounter_stats()["nbox_total"] == nbox and ctx.encounter_bucket_fallbacks() == f0 + 1
    # the same clump below the capacity stays on the bucket sort: ties are ordered by position, as the stable sort does
    r[900:3000] = d["rh"][900:3000]
    renc2[900:3000] = renc[900:3000]
    ref = oracle.encounter_plpl(r, v, renc2, d["dt"])
    nbox = oracle.nbox_total()
    _same_pairs(ctx.encounter_check_all_sort_and_sweep_plpl(6000, r, v, renc2, d["dt"]), ref)
    assert ctx.encounter_stats()["nbox_total"] == nbox and ctx.encounter_bucket_fallbacks() == f0 + 1
    r[5] = np.inf
    ctx.encounter_check_all_sort_and_sweep_plpl(6000, r, v, renc2, d["dt"])  # garbage in; must not crash, must fall back
    assert ctx.encounter_bucket_fallbacks() == f0 + 2


def test_sweep_F3_quirk_is_reproduced(ctx, oracle):
    r = np.array([[1.0, 0, 0], [1.05, 0, 0]])
    v = np.array([[0.0, 6.0, 0], [0.0, -6.0, 0]])
    renc = np.array([0.1, 0.1])
    assert ctx.encounter_check_all_sort_and_sweep_plpl(2, r, v, renc, 0.01)[0] == 0
    r3 = np.vstack([r, [[0.0, 1.02, 0.0]]])
    v3 = np.vstack([v, [[0.0, 0.0, 0.0]]])
    renc3 = np.array([0.1, 0.1, 1e-4])
    _same_pairs(ctx.encounter_check_all_sort_and_sweep_plpl(3, r3, v3, renc3, 0.01),
                oracle.encounter_plpl(r3, v3, renc3, 0.01))


@pytest.mark.parametrize("ntp", [50, 20000, 300000])
def test_sweep_pltp_bit_exact(ctx, oracle, ntp):
    if ntp == 50:
        f, pl, _ = _fixture108()
        rpl, vpl, renc, rtp, vtp, dt = pl["rh"], pl["vh"], pl["rhill"] * 6.5 * 3, f["tp_rh"], f["tp_vh"], 0.05
    else:
        p = W.planets8_year_units()
        tp = W.tp_cloud(ntp, seed=ntp)
        rpl, vpl, renc, rtp, vtp, dt = p["rh"], p["vh"], p["rhill"] * 6.5, tp["rh"], tp["vh"], 0.05
    ref = oracle.encounter_pltp(rpl, vpl, rtp, vtp, renc, dt)
    got = ctx.encounter_check_all_sort_and_sweep_pltp(len(renc), ntp, rpl, vpl, rtp, vtp, renc, dt)
    _same_pairs(got, ref)
    if ntp >= 20000:
        assert got[0] > 0


def _pltp_both_paths(ctx, oracle, args, expect_fallback=False):
    """pl-tp sweep through the sort-free pass and through the sort path: both must return the oracle's list."""
    import os
    from tests import pltp_rule as R
    rpl, vpl, rtp, vtp, renc, dt = args
    ref = oracle.encounter_pltp(rpl, vpl, rtp, vtp, renc, dt)
    nbox = oracle.nbox_total()
    d0, f0 = ctx.encounter_direct_count()
    got = ctx.encounter_check_all_sort_and_sweep_pltp(len(renc), len(rtp), rpl, vpl, rtp, vtp, renc, dt)
    st = ctx.encounter_stats()
    d1, f1 = ctx.encounter_direct_count()
    _same_pairs(got, ref)
    assert d1 == d0 + 1 and f1 - f0 == (1 if expect_fallback else 0)
    if expect_fallback:
        assert st["nbox_total"] == nbox
    else:
        assert st["nbox_total"] + R.pairless_particle_boxes(rtp) == nbox
        assert st["emitted"] == 2 * got[0]
    os.environ["SWCU_PLTP_DIRECT_MAX"] = "0"
    try:
        got2 = ctx.encounter_check_all_sort_and_sweep_pltp(len(renc), len(rtp), rpl, vpl, rtp, vtp, renc, dt)
        assert ctx.encounter_stats()["nbox_total"] == nbox
    finally:
        del os.environ["SWCU_PLTP_DIRECT_MAX"]
    _same_pairs(got2, ref)
    assert ctx.encounter_direct_count() == (d1, f1)
    return got[0]


@pytest.mark.gpu
@pytest.mark.parametrize("ntp", [1, 31, 5000, 200000])
def test_sweep_pltp_sort_free_pass_equals_sort_path_and_oracle(ctx, oracle, ntp):
    p = W.planets8_year_units()
    tp = W.tp_cloud(ntp, seed=7 + ntp)
    n = _pltp_both_paths(ctx, oracle, (p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 6.5, 0.05))
    if ntp >= 5000:
        assert n > 0


@pytest.mark.gpu
def test_sweep_pltp_sort_free_pass_on_the_reference_fixture(ctx, oracle):
    f, pl, _ = _fixture108()
    for boost in (1.0, 3.0, 10.0):
        _pltp_both_paths(ctx, oracle, (pl["rh"], pl["vh"], f["tp_rh"], f["tp_vh"], pl["rhill"] * 6.5 * boost, 0.05))


@pytest.mark.gpu
def test_sweep_pltp_extent_ties(ctx, oracle):
    """Particles exactly on a planet's inner extent and particles sharing |r| stay on the sort-free pass; a particle
    exactly on an OUTER extent is the one case the sorted sequence decides: the call repeats itself on the sort path."""
    from tests import pltp_rule as R
    assert _pltp_both_paths(ctx, oracle, R.tie_case("rmin")) > 0
    assert _pltp_both_paths(ctx, oracle, R.tie_case("dup")) > 0
    assert _pltp_both_paths(ctx, oracle, R.tie_case("rmax"), expect_fallback=True) > 0
    # planets with identical extents, renc = 0 and renc < 0
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=8)
    rpl, renc = p["rh"].copy(), p["rhill"] * 6.5
    rpl[3] = rpl[2]
    renc[3], renc[6], renc[7] = renc[2], 0.0, -renc[7]
    _pltp_both_paths(ctx, oracle, (rpl, p["vh"], tp["rh"], tp["vh"], renc, 0.05))


@pytest.mark.gpu
def test_sweep_pltp_sort_free_pass_with_more_hits_than_the_first_candidate_buffer(ctx, oracle):
    """Every particle inside Jupiter's sphere: far more hits than the initial candidate capacity."""
    p = W.planets8_year_units()
    rng = np.random.default_rng(4)
    ntp = 300000
    rtp = p["rh"][4] + rng.normal(size=(ntp, 3)) * 0.02
    vtp = p["vh"][4] + rng.normal(size=(ntp, 3)) * 0.1
    n = _pltp_both_paths(ctx, oracle, (p["rh"], p["vh"], rtp, vtp, p["rhill"] * 6.5, 0.05))
    assert n > ntp // 4 + 65536


def test_sweep_plplm_and_merged_list_bit_exact(ctx, oracle):
    n, nplm = 3000, 700
    d = W.disk(n, seed=12)
    renc = d["rhill"] * 6.5 * 3
    a = (d["rh"][:nplm], d["vh"][:nplm], d["rh"][nplm:], d["vh"][nplm:], renc[:nplm], renc[nplm:])
    _same_pairs(ctx.encounter_check_all_sort_and_sweep_plplm(nplm, n - nplm, *a, d["dt"]),
                oracle.encounter_plplm(*a, d["dt"]))
    got = ctx.encounter_check_all_plplm(nplm, n - nplm, *a, d["dt"])
    _same_pairs(got, oracle.encounter_plplm(*a, d["dt"], merged=True))
    assert got[0] > 0 and got[2].max() > nplm


def test_sweep_1e5_disk_properties(ctx, oracle):
    """Full-size (BASELINE N=1e5) run: canonical order, no duplicates, index1 < index2, every reported pair
    satisfies the narrow-phase predicate, and a row sample agrees with the brute-force all-pairs check."""
    n = 100000
    d = W.disk(n, seed=3031179)
    renc = d["rhill"] * 6.5
    nenc, i1, i2, lv = ctx.encounter_check_all_sort_and_sweep_plpl(n, d["rh"], d["vh"], renc, d["dt"])
    assert nenc > 0 and lv.all()
    key = i1.astype(np.int64) * (1 << 32) + i2
    assert np.all(np.diff(key) > 0) and np.all(i1 < i2) and i1.min() >= 1 and i2.max() <= n
    pick = np.random.default_rng(0).choice(nenc, min(nenc, 2000), replace=False)
    for k in pick:
        a, b = i1[k] - 1, i2[k] - 1
        dr, dv = d["rh"][b] - d["rh"][a], d["vh"][b] - d["vh"][a]
        assert oracle.encounter_check_one(*dr, *dv, renc[a] + renc[b], d["dt"])[0]
    ref = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"])
    _same_pairs((nenc, i1, i2, lv), ref)


def test_empty_and_tiny_inputs(ctx):
    z3, z1 = np.zeros((0, 3)), np.zeros(0)
    assert ctx.encounter_check_all_sort_and_sweep_plpl(0, z3, z3, z1, 0.1)[0] == 0
    assert ctx.encounter_check_all_sort_and_sweep_pltp(0, 0, z3, z3, z3, z3, z1, 0.1)[0] == 0
    acc = np.zeros((0, 3))
    ctx.kick_getacch_int_all_tri_pl(0, 0, z3, z1, z1, acc)
    ctx.kick_getacch_int_all_tp(0, 0, z3, z3, z1, np.zeros(0, np.int32), acc)
    ctx.drift_all(z1, z3.copy(), z3.copy(), 0, 0.1, np.zeros(0, np.int32), np.zeros(0, np.int32))
    one = np.array([[1.0, 0.0, 0.0]])
    a1 = np.zeros((1, 3))
    ctx.kick_getacch_int_all_tri_pl(1, 1, one, np.array([1e-3]), np.array([1e-4]), a1)
    assert np.all(a1 == 0.0)
    assert ctx.encounter_check_all_sort_and_sweep_plpl(1, one, one, np.array([0.1]), 0.1)[0] == 0


# ---------------------------------------------------------------------------------------------- resident tier
def test_resident_step_matches_oracle_sequence(ctx, oracle):
    """Tier 2: sync once, then encounter check -> kick -> velocity update -> drift without host round trips."""
    n = 3000
    d = W.disk(n, seed=2)
    dt = d["dt"]
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=d["mu"], generation=1)
    ctx.pl_set_renc(0)
    got = ctx.pl_encounter_check(dt)
    renc = oracle.set_renc(d["rhill"], 0)
    _same_pairs(got, oracle.encounter_plpl(d["rh"], d["vh"], renc, dt))
    for variant in (LOOP_TRIANGULAR, LOOP_FLAT):
        ctx.body_put(PL, r=d["rh"], v=d["vh"])
        ctx.body_zero_accel(PL)
        ctx.pl_accel_int(variant, True)
        ctx.body_kick_velocity(PL, dt)
        assert ctx.body_drift(PL, dt) == 0
        out = ctx.body_get(PL, iflag=True)
        ah = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((n, 3)))
        scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
        assert _scaled(out["a"], ah, scale) < ACC_TOL
        vb = d["vh"] + ah * dt
        xr, vr, fr = oracle.drift_all(d["mu"], d["rh"], vb, dt)
        assert np.max(np.abs(out["r"] - xr) / np.linalg.norm(xr, axis=1, keepdims=True)) < 1e-12
        assert np.max(np.abs(out["v"] - vr) / np.linalg.norm(vr, axis=1, keepdims=True)) < 1e-12
        assert not out["iflag"].any()


def test_resident_generation_counter_and_resync(ctx, oracle):
    """Arrays are re-uploaded only when the generation changes (collision/discard => rearray_pl); after a body
    count change the device mirror must equal the host arrays in the new (mass-descending) order."""
    n = 500
    d = W.disk(n, seed=6)
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=d["mu"], generation=10)
    junk = np.zeros_like(d["rh"])
    ctx.body_sync(PL, n, nplm=n, r=junk, v=junk, Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                  generation=10)  # same generation: must be a no-op
    assert np.array_equal(ctx.body_get(PL)["r"], d["rh"])
    # merge/fragment event replay: delete 7 random bodies, append 3 fragments, re-sort by mass, bump generation
    rng = np.random.default_rng(6)
    keep = np.setdiff1d(np.arange(n), rng.choice(n, 7, replace=False))
    frag = W.disk(3, seed=99)
    arr = {k: np.concatenate([d[k][keep], frag[k]]) for k in ("rh", "vh", "Gmass", "radius", "rhill", "mu")}
    order = np.argsort(-arr["Gmass"], kind="stable")
    arr = {k: np.ascontiguousarray(q[order]) for k, q in arr.items()}
    m = len(order)
    ctx.body_sync(PL, m, nplm=m, r=arr["rh"], v=arr["vh"], Gmass=arr["Gmass"], radius=arr["radius"],
                  rhill=arr["rhill"], mu=arr["mu"], generation=11)
    assert ctx.body_count(PL) == (m, m, 11)
    out = ctx.body_get(PL)
    assert np.array_equal(out["r"], arr["rh"]) and np.array_equal(out["v"], arr["vh"])
    ctx.body_zero_accel(PL)
    ctx.pl_accel_int(LOOP_TRIANGULAR, True)
    ref = oracle.kick_tri_pl(arr["rh"], arr["Gmass"], arr["radius"], np.zeros((m, 3)))
    scale = oracle.kick_tri_abs_scale(arr["rh"], arr["Gmass"], arr["radius"])
    assert _scaled(ctx.body_get(PL)["a"], ref, scale) < ACC_TOL


def test_resident_tp_kick_drift_and_encounter(ctx, oracle):
    p = W.planets8_year_units()
    ntp = 50000
    tp = W.tp_cloud(ntp, seed=77)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=21)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, p["cb_Gmass"]), generation=22)
    ctx.tp_accel_int()
    ctx.body_kick_velocity(TP, 0.01)
    assert ctx.body_drift(TP, 0.01) == 0
    out = ctx.body_get(TP)
    acc = oracle.kick_all_tp(tp["rh"], p["rh"], p["Gmass"], np.ones(ntp, np.int32), np.zeros((ntp, 3)))
    assert np.max(np.abs(out["a"] - acc)) <= 1e-12 * np.abs(acc).max()
    xr, vr, _ = oracle.drift_all(p["cb_Gmass"], tp["rh"], tp["vh"] + acc * 0.01, 0.01)
    assert np.max(np.abs(out["r"] - xr) / np.linalg.norm(xr, axis=1, keepdims=True)) < 1e-12
    ctx.body_put(TP, r=tp["rh"], v=tp["vh"])
    ctx.pl_set_renc(0)
    got = ctx.tp_encounter_check(0.05)
    _same_pairs(got, oracle.encounter_pltp(p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 6.5, 0.05))


def test_fused_whm_tp_step_matches_unfused_oracle_sequence(ctx, oracle):
    """Next row: whm_step_tp as one kernel vs kick/drift/kick built from oracle calls
    (whm_step.f90:72-100, whm_kick.f90:70-149,265-314)."""
    p = W.planets8_year_units()
    ntp, dt = 30000, 0.01
    tp = W.tp_cloud(ntp, seed=5)
    rng = np.random.default_rng(5)
    mask = (rng.uniform(size=ntp) > 0.05).astype(np.int32)
    on = mask.astype(bool)

    def ah0_of(rpl):  # whm_kick_getacch_ah0
        a = np.zeros(3)
        for i in range(len(p["Gmass"])):
            r2 = float(rpl[i] @ rpl[i])
            a = a - p["Gmass"][i] * (1.0 / (r2 * np.sqrt(r2))) * rpl[i]
        return a

    rbeg = p["rh"]
    rend = p["rh"] + p["vh"] * dt  # any end-of-step planet positions will do for the comparison
    mu = np.full(ntp, p["cb_Gmass"])
    # reference sequence: lfirst accel at rbeg, kick, drift, accel at rend, kick
    ah = np.zeros((ntp, 3))
    ah[on] += ah0_of(rbeg)
    ah = oracle.kick_all_tp(tp["rh"], rbeg, p["Gmass"], mask, ah)
    v = tp["vh"].copy()
    v[on] = v[on] + ah[on] * (0.5 * dt)
    x, v, fl = oracle.drift_all(mu, tp["rh"], v, dt, lmask=mask)
    ah2 = np.zeros((ntp, 3))
    ah2[on] += ah0_of(rend)
    ah2 = oracle.kick_all_tp(x, rend, p["Gmass"], mask, ah2)
    v[on] = v[on] + ah2[on] * (0.5 * dt)
    # device: resident populations, first-step accelerations through the unfused calls, then the fused step
    ctx.body_sync(PL, 8, nplm=8, r=rbeg, v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=31)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=mu, lmask=mask, generation=32)
    a_first = np.zeros((ntp, 3))
    a_first[on] += ah0_of(rbeg)
    ctx.body_put(TP, a=a_first)
    ctx.tp_accel_int()
    ctx.body_put(PL, r=rend)
    assert ctx.whm_tp_step(dt, ah0_of(rend)) == 0
    out = ctx.body_get(TP, iflag=True)
    assert np.array_equal(out["iflag"][on], fl[on])
    assert np.max(np.abs(out["r"] - x) / np.linalg.norm(x, axis=1, keepdims=True)) < 1e-12
    assert np.max(np.abs(out["v"] - v) / np.linalg.norm(v, axis=1, keepdims=True)) < 1e-12
    assert np.max(np.abs(out["a"][on] - ah2[on])) <= 1e-12 * np.abs(ah2).max()
    off = ~on
    assert np.array_equal(out["r"][off], tp["rh"][off]) and np.array_equal(out["v"][off], tp["vh"][off])


# ---------------------------------------------------------------------------------------------- system level
def test_helio_integration_tracks_oracle_run(ctx, oracle):
    """Energy and angular-momentum error of a Sun + 8 planets run must track the CPU run step for step
    (north star); 2000 steps here, the 1e4-step run is in bench.py --conservation."""
    from tests.helio import GpuBackend, HelioSystem, OracleBackend
    p = W.planets8_year_units()
    a = HelioSystem(p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], OracleBackend(oracle))
    b = HelioSystem(p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], GpuBackend(ctx))
    E0, L0 = a.energy_and_momentum()
    for k in range(2000):
        a.step(0.01)
        b.step(0.01)
    Ea, La = a.energy_and_momentum()
    Eb, Lb = b.energy_and_momentum()
    assert abs((Ea - E0) / E0) < 1e-7 and abs((Eb - E0) / E0) < 1e-7
    assert abs((Eb - Ea) / E0) < 1e-11
    assert np.linalg.norm(Lb - La) / np.linalg.norm(L0) < 1e-12
    assert np.max(np.abs(a.rh - b.rh)) < 1e-9


# ---------------------------------------------------------------------------------------------- error behaviour
def test_slice_put_get_blocking_and_async(ctx):
    """swcu_body_put_range / _get_range and their asynchronous forms (copy streams, double-buffered staging): a slice
    round-trips bit for bit, bodies outside it are untouched, and back-to-back async calls keep their order."""
    import torch
    n = 5000
    d = W.disk(n, seed=71)
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                  generation=7101)
    i0, i1 = 1234, 4321
    rng = np.random.default_rng(1)
    r1, v1 = rng.normal(size=(i1 - i0, 3)), rng.normal(size=(i1 - i0, 3))
    ctx.body_put_range(PL, i0, i1, r=r1, v=v1)
    g = ctx.body_get(PL, a=False)
    assert np.array_equal(g["r"][i0:i1], r1) and np.array_equal(g["v"][i0:i1], v1)
    assert np.array_equal(g["r"][:i0], d["rh"][:i0]) and np.array_equal(g["v"][i1:], d["vh"][i1:])
    sl = ctx.body_get_range(PL, i0, i1)
    assert np.array_equal(sl["r"], r1) and np.array_equal(sl["v"], v1) and sl["a"].shape == (i1 - i0, 3)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    ins = [(pin(rng.normal(size=(i1 - i0, 3))), pin(rng.normal(size=(i1 - i0, 3)))) for _ in range(5)]
    outs = [{k: pin(np.zeros((i1 - i0, 3))) for k in ("r", "v", "a")} for _ in range(5)]
    for (r, v), o in zip(ins, outs):      # five put/get pairs in flight, no host synchronisation in between
        ctx.body_put_range_async(PL, i0, i1, r=r, v=v)
        ctx.body_get_range_async(PL, i0, i1, o)
    ctx.io_wait()
    for (r, v), o in zip(ins, outs):
        assert np.array_equal(o["r"], r) and np.array_equal(o["v"], v)
    g = ctx.body_get(PL, a=False)
    assert np.array_equal(g["r"][i0:i1], ins[-1][0]) and np.array_equal(g["r"][:i0], d["rh"][:i0])
    from swiftest_b200 import SwcuError
    with pytest.raises(SwcuError):
        ctx.body_put_range(PL, 10, n + 1, r=np.zeros((n - 9, 3)))


def test_error_paths_return_status_and_message(ctx):
    """The reference's conventions: early returns for empty populations, fatal status for inconsistent calls."""
    from swiftest_b200 import Context, SwcuError
    d = W.disk(64, seed=1)
    with Context(0) as c:  # a fresh context: nothing resident
        with pytest.raises(SwcuError, match="not resident"):
            c.pl_accel_int(LOOP_TRIANGULAR, True)
        with pytest.raises(SwcuError, match="not resident"):
            c.body_drift(PL, 0.01)
        with pytest.raises(SwcuError, match="not resident"):
            c.body_get(PL)
        with pytest.raises(SwcuError, match="not imported"):
            c.pl_kick_drift_p2p(0.01)
        with pytest.raises(SwcuError, match="bad npl"):
            c.kick_getacch_int_all_tri_pl(64, 65, d["rh"], d["Gmass"], d["radius"], np.zeros((64, 3)))
        with pytest.raises(SwcuError, match="bad nplm"):
            c.body_sync(PL, 64, nplm=70, r=d["rh"])
        # fetch must match the count of the last check
        n = c.encounter_check_all_sort_and_sweep_plpl(64, d["rh"], d["vh"], d["rhill"] * 100, d["dt"])[0]
        import ctypes as C
        i1 = np.zeros(n + 5, np.int32)
        rc = c._L.swcu_encounter_fetch(c._h, n + 1, i1.ctypes.data, i1.ctypes.data, None)
        assert rc == 2 and b"last check found" in c._L.swcu_last_error(c._h)
        assert C.sizeof(C.c_void_p) == 8
    # no-ops of the reference: npl == 0 or ntp == 0 return without touching acc (kick.f90:61, drift.f90:81)
    acc = np.full((5, 3), 7.0)
    ctx.kick_getacch_int_all_tp(5, 0, np.zeros((5, 3)), np.zeros((0, 3)), np.zeros(0), np.ones(5, np.int32), acc)
    assert np.all(acc == 7.0)
