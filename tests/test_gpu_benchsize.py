"""GPU parity at the sizes bench.py measures (BASELINE.json configs[1] and configs[3]), against the CPU oracle's
reference-shaped OpenMP loops on the same inputs -- the launches that produce the headline numbers (npl = 1e5: 782
blocks, 306 153 block pairs, graded claim schedule; ntp = 1e6) are compared element by element, not just timed.

Reference lines: swiftest_kick.f90:219-240 (full-row loop), :394-412 (pl -> tp), swiftest_drift.f90:60-108,
encounter_check.f90:261-326 (pl-tp sort and sweep).  The OpenMP oracle build (-O3, strict IEEE, no contraction: the same
source as the -O2 build the small tests use) keeps the CPU side of these tests to a few seconds each.
"""
import numpy as np
import pytest

from swiftest_b200 import workloads as W
from swiftest_b200 import PL, LOOP_FLAT, LOOP_TRIANGULAR

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-12          # relative to sum_j |term_j| per component (north star)
NPL = 100_000
NTP = 1_000_000


@pytest.fixture(scope="module")
def fast_oracle():
    from oracle import load
    return load(native=True)


@pytest.fixture(scope="module")
def disk1e5(fast_oracle):
    d = W.disk(NPL, seed=3031179)                    # bench.py's workload
    ref = np.zeros((NPL, 3))
    fast_oracle.omp_kick_tri_rad_pl_rows(d["rh"], d["Gmass"], d["radius"], ref, NPL, 0, NPL)
    scale = fast_oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
    return d, ref, scale


@pytest.fixture(scope="module")
def cloud1e6():
    return W.planets8_year_units(), W.tp_cloud(NTP, seed=123)      # bench.py's WHM workload


def _scaled(a, ref, scale):
    return float(np.max(np.abs(a - ref) / np.where(scale > 0, scale, 1.0)))


@pytest.mark.parametrize("variant", [LOOP_FLAT, LOOP_TRIANGULAR], ids=["flat", "tri"])
def test_plpl_accelerations_at_npl_1e5_all_rows(ctx, disk1e5, variant):
    """The resident launch bench.py times (pl%accel_int on 1e5 bodies, radius-checked): every row within 1e-12."""
    d, ref, scale = disk1e5
    ctx.body_sync(PL, NPL, nplm=NPL, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=d["mu"], generation=987001)
    n0 = ctx.flat_redo_count()
    ctx.body_zero_accel(PL)
    ctx.pl_accel_int(variant, True)
    got = ctx.body_get(PL, r=False, v=False)["a"]
    assert np.all(np.isfinite(got))
    assert _scaled(got, ref, scale) < ACC_TOL
    if variant == LOOP_FLAT:
        # Newton's third law at full size, and the benchmark disk never leaves the seeded fast path
        p = (d["Gmass"][:, None] * got).sum(0)
        assert np.max(np.abs(p)) < 1e-11 * np.abs(d["Gmass"][:, None] * got).sum()
        assert ctx.flat_redo_count() == n0


def test_plpl_accelerations_at_npl_1e5_tier1_host_pointers(ctx, disk1e5):
    """Same through the array-level entry point (upload, kernel, download), norad variant against its own oracle sum."""
    d, ref, scale = disk1e5
    got = np.zeros((NPL, 3))
    ctx.kick_getacch_int_all_flat_pl(NPL, NPL * (NPL - 1) // 2, None, d["rh"], d["Gmass"], d["radius"], got)
    assert _scaled(got, ref, scale) < ACC_TOL


def test_pltp_accelerations_at_ntp_1e6(ctx, fast_oracle, cloud1e6):
    p, tp = cloud1e6
    rng = np.random.default_rng(5)
    mask = (rng.uniform(size=NTP) > 0.02).astype(np.int32)
    acc0 = rng.normal(scale=1e-6, size=(NTP, 3))
    ref = acc0.copy()
    fast_oracle.omp_kick_all_tp(tp["rh"], p["rh"], p["Gmass"], mask, ref)
    got = acc0.copy()
    ctx.kick_getacch_int_all_tp(NTP, 8, tp["rh"], p["rh"], p["Gmass"], mask, got)
    scale = np.zeros((NTP, 3))
    for j in range(8):                               # sum of |terms| per component, planet by planet
        dd = tp["rh"] - p["rh"][j]
        scale += p["Gmass"][j] * np.abs(dd) / (np.linalg.norm(dd, axis=1) ** 3)[:, None]
    scale += np.abs(acc0)
    assert _scaled(got, ref, scale) < ACC_TOL
    off = mask == 0
    assert np.array_equal(got[off], acc0[off])


def test_drift_at_ntp_1e6_against_oracle(ctx, fast_oracle, cloud1e6):
    """Element by element against swiftest_drift_all restated on the CPU (not only the +dt/-dt round trip): iflag
    identical, bit-identical where no libm call is involved, 1e-12 elsewhere."""
    _, tp = cloud1e6
    dt = 0.01
    xr, vr, fr = fast_oracle.drift_all(W.GMSUN, tp["rh"], tp["vh"], dt, omp=True)
    x, v, fl = tp["rh"].copy(), tp["vh"].copy(), np.zeros(NTP, np.int32)
    ctx.drift_all(np.full(NTP, W.GMSUN), x, v, NTP, dt, np.ones(NTP, np.int32), fl)
    assert np.array_equal(fl, fr)
    assert np.max(np.abs(x - xr) / np.linalg.norm(xr, axis=1, keepdims=True)) < 1e-12
    assert np.max(np.abs(v - vr) / np.linalg.norm(vr, axis=1, keepdims=True)) < 1e-12
    br = fast_oracle.drift_branch(W.GMSUN, tp["rh"][:200000], tp["vh"][:200000], dt)
    exact = (br == 0) | (br == 1)
    assert exact.sum() > 1000
    assert np.array_equal(x[:200000][exact], xr[:200000][exact]) and np.array_equal(v[:200000][exact], vr[:200000][exact])


def test_sweep_pltp_at_ntp_1e6_bit_exact(ctx, fast_oracle, cloud1e6):
    """8 planets + 1e6 test particles (the WHM/RMVS configuration): pair list bit-exact in canonical order."""
    p, tp = cloud1e6
    renc = p["rhill"] * 6.5
    r1, r2, rlv = fast_oracle.encounter_pltp(p["rh"], p["vh"], tp["rh"], tp["vh"], renc, 0.05)
    n, g1, g2, glv = ctx.encounter_check_all_sort_and_sweep_pltp(8, NTP, p["rh"], p["vh"], tp["rh"], tp["vh"], renc, 0.05)
    assert n == len(r1) and n > 0
    assert np.array_equal(g1, r1) and np.array_equal(g2, r2)
    assert glv.all() and rlv.all()
