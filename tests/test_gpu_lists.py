"""GPU parity tests for SURVEY.md 8(f) ranks 3-4: triangular encounter checks, pl-tp discard, SyMBA list check.
All integer / index results: BIT-EXACT against the CPU restatement (the kernels are compiled without FMA contraction).
"""
import numpy as np
import pytest

from swiftest_b200 import workloads as W

pytestmark = pytest.mark.gpu


def _same(got, ref):
    n, g1, g2, glv = got
    r1, r2, rlv = ref
    assert n == len(r1)
    assert np.array_equal(g1, r1) and np.array_equal(g2, r2)
    assert glv.all()


@pytest.mark.parametrize("n,boost", [(1, 1.0), (2, 50.0), (129, 6.0), (1000, 4.0), (5000, 2.0)])
def test_triangular_plpl_matches_oracle(ctx, oracle, n, boost):
    d = W.disk(n, seed=500 + n)
    renc = oracle.set_renc(d["rhill"], 0) * boost
    ref = oracle.encounter_plpl(d["rh"], d["vh"], renc, d["dt"], triangular=True)
    got = ctx.encounter_check_all_triangular_plpl(n, d["rh"], d["vh"], renc, d["dt"])
    _same(got, ref)
    if n >= 1000:
        assert got[0] > 0
        # the sweep finds a subset of it
        sw = ctx.encounter_check_all_sort_and_sweep_plpl(n, d["rh"], d["vh"], renc, d["dt"])
        assert set(zip(sw[1].tolist(), sw[2].tolist())) <= set(zip(got[1].tolist(), got[2].tolist()))


def test_triangular_plpl_candidate_buffer_growth(ctx, oracle):
    """More hits than the initial candidate buffer: the call grows it and repeats."""
    n = 700
    rng = np.random.default_rng(3)
    r = rng.normal(size=(n, 3))
    v = np.zeros((n, 3))
    renc = np.full(n, 10.0)  # everybody is inside everybody's encounter radius: n(n-1)/2 = 244650 pairs
    got = ctx.encounter_check_all_triangular_plpl(n, r, v, renc, 0.1)
    assert got[0] == n * (n - 1) // 2
    iu = np.triu_indices(n, 1)
    assert np.array_equal(got[1], iu[0] + 1) and np.array_equal(got[2], iu[1] + 1)


def test_triangular_pltp_and_plplm_match_oracle(ctx, oracle):
    p = W.planets8_year_units()
    tp = W.tp_cloud(6000, seed=31)
    renc = p["rhill"] * 6.5 * 3
    ref = oracle.encounter_pltp(p["rh"], p["vh"], tp["rh"], tp["vh"], renc, 0.05, triangular=True)
    got = ctx.encounter_check_all_triangular_pltp(8, 6000, p["rh"], p["vh"], tp["rh"], tp["vh"], renc, 0.05)
    _same(got, ref)
    assert got[0] > 0
    f = W.fixture("108pl_50tp")
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    nplm = int((f["pl_Gmass"] >= float(f["GMTINY"])).sum())
    r, v = f["pl_rh"][order], f["pl_vh"][order]
    rc = oracle.set_renc(f["pl_rhill"][order], 0) * 20
    ref = oracle.encounter_plplm(r[:nplm], v[:nplm], r[nplm:], v[nplm:], rc[:nplm], rc[nplm:], 0.05, triangular=True)
    got = ctx.encounter_check_all_triangular_plplm(nplm, 108 - nplm, r[:nplm], v[:nplm], r[nplm:], v[nplm:], rc[:nplm],
                                                   rc[nplm:], 0.05)
    _same(got, ref)
    assert got[0] > 0
    assert ctx.encounter_check_all_triangular_pltp(8, 0, p["rh"], p["vh"], np.zeros((0, 3)), np.zeros((0, 3)), renc, 0.05)[0] == 0


@pytest.mark.parametrize("npl,ntp", [(8, 20000), (300, 5000), (1, 1)])
def test_discard_pl_tp_matches_oracle(ctx, oracle, npl, ntp):
    rng = np.random.default_rng(npl + ntp)
    if npl == 8:
        p = W.planets8_year_units()
        rpl, vpl = p["rh"], p["vh"]
    else:
        d = W.disk(npl, seed=npl)
        rpl, vpl = d["rh"], d["vh"]
    # particles scattered around the planets so that a good fraction is, or will be, inside the (inflated) radii
    host = rng.integers(0, npl, ntp)
    rtp = rpl[host] + rng.normal(scale=0.02, size=(ntp, 3))
    vtp = vpl[host] + rng.normal(scale=1.0, size=(ntp, 3))
    radius = np.full(npl, 0.01)
    act = (rng.uniform(size=ntp) > 0.1).astype(np.int32)
    ref, nref = oracle.discard_pl_tp(rtp, vtp, act, rpl, vpl, radius, 0.01)
    got, ngot = ctx.discard_pl_tp(rtp, vtp, act, rpl, vpl, radius, 0.01)
    assert np.array_equal(got, ref) and ngot == nref
    if ntp > 1:
        assert 0 < nref < ntp
        assert not got[act == 0].any()
    got2, _ = ctx.discard_pl_tp(rtp, vtp, None, rpl, vpl, radius, 0.01)
    ref2, _ = oracle.discard_pl_tp(rtp, vtp, None, rpl, vpl, radius, 0.01)
    assert np.array_equal(got2, ref2)


def test_symba_encounter_check_list_matches_oracle(ctx, oracle):
    n = 4000
    d = W.disk(n, seed=77)
    renc0 = oracle.set_renc(d["rhill"], 0) * 3
    _, i1, i2, _ = ctx.encounter_check_all_triangular_plpl(n, d["rh"], d["vh"], renc0, d["dt"])
    assert len(i1) > 50
    rng = np.random.default_rng(1)
    mask = (rng.uniform(size=len(i1)) > 0.3).astype(np.int32)
    renc1 = oracle.set_renc(d["rhill"], 1) * 3  # next recursion level: smaller shells
    radius = d["radius"] * 200                   # inflated so that some pairs count as overlapping
    lv0 = np.full(len(i1), 5, np.int32)
    ref = oracle.symba_encounter_check_list(i1, i2, mask, d["rh"], d["vh"], renc1, radius, d["dt"] / 3, lvdotr=lv0)
    got = ctx.symba_encounter_check_list(i1, i2, mask, d["rh"], d["vh"], renc1, radius, d["dt"] / 3, lvdotr=lv0)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert 0 < ref[2] < mask.sum()
    assert (got[1][mask == 0] == 5).all()
    # pl-tp form: second list without renc / radius
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=2)
    rc = p["rhill"] * 6.5 * 4
    _, j1, j2, _ = ctx.encounter_check_all_triangular_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], rc, 0.05)
    assert len(j1) > 5
    ref = oracle.symba_encounter_check_list(j1, j2, None, p["rh"], p["vh"], rc * 0.48075, p["radius"], 0.02,
                                            r2=tp["rh"], v2=tp["vh"])
    got = ctx.symba_encounter_check_list(j1, j2, None, p["rh"], p["vh"], rc * 0.48075, p["radius"], 0.02,
                                         r2=tp["rh"], v2=tp["vh"])
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    import swiftest_b200 as S
    with pytest.raises(S.SwcuError):  # an index outside the population is a checked error, not a wild read
        ctx.symba_encounter_check_list([1], [9999], None, p["rh"], p["vh"], rc, p["radius"], 0.02)
    assert ctx.symba_encounter_check_list([], [], None, p["rh"], p["vh"], rc, p["radius"], 0.02)[2] == 0


def _disk_list(ctx, oracle, n, seed, boost):
    d = W.disk(n, seed=seed)
    renc = oracle.set_renc(d["rhill"], 0) * boost
    _, i1, i2, _ = ctx.encounter_check_all_triangular_plpl(n, d["rh"], d["vh"], renc, d["dt"])
    return d, i1, i2


@pytest.mark.parametrize("irec,sgn", [(0, 1), (1, 1), (1, -1), (2, -1)])
def test_symba_kick_list_plpl_matches_serial_oracle(ctx, oracle, irec, sgn):
    """Bodies appear in many pairs; the device must add each body's contributions in list order.  Pairs outside every
    shell use IEEE arithmetic only -> bit-exact; pairs inside the shell call pow(r2,-1.5) -> 4 ulp on the factor."""
    n = 3000
    d, i1, i2 = _disk_list(ctx, oracle, n, 21, 6.0)
    assert len(i1) > 500
    rng = np.random.default_rng(irec * 7 + sgn)
    levelg = rng.integers(max(irec - 1, 0), irec + 2, n).astype(np.int32)
    active = (rng.uniform(size=len(i1)) > 0.1).astype(np.int32)
    rhill = d["rhill"] * (3.0 if irec == 0 else 6.0)   # inflate so that all three branches (inner, shell, outside) occur
    vb0 = d["vh"].copy()
    ref_vb, ref_good, _ = oracle.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], rhill, d["Gmass"], d["dt"], irec, sgn, vb0)
    got_vb, got_good = ctx.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], rhill, d["Gmass"], d["dt"], irec, sgn, vb0)
    assert np.array_equal(got_good, ref_good)
    assert 0 < ref_good.sum() < len(i1)
    dv_ref, dv_got = ref_vb - vb0, got_vb - vb0
    touched = np.abs(dv_ref).sum(1) > 0
    assert touched.sum() > 10
    assert np.array_equal(got_vb[~touched], vb0[~touched])
    scale = np.abs(dv_ref).max(1, keepdims=True)[touched]
    assert np.max(np.abs(dv_got[touched] - dv_ref[touched]) / scale) < 1e-14
    # with the shell pushed inside every pair (tiny Hill radii) nothing calls pow: identical bits
    ref_vb, ref_good, _ = oracle.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], rhill * 1e-6, d["Gmass"], d["dt"], irec,
                                                      sgn, vb0)
    got_vb, got_good = ctx.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], rhill * 1e-6, d["Gmass"], d["dt"], irec, sgn,
                                                vb0)
    assert np.array_equal(got_good, ref_good) and np.array_equal(got_vb, ref_vb)


def test_symba_kick_list_pltp_matches_serial_oracle(ctx, oracle):
    p = W.planets8_year_units()
    ntp = 4000
    tp = W.tp_cloud(ntp, seed=8)
    rc = p["rhill"] * 6.5 * 6
    _, i1, i2, _ = ctx.encounter_check_all_triangular_pltp(8, ntp, p["rh"], p["vh"], tp["rh"], tp["vh"], rc, 0.05)
    assert len(i1) > 20
    rng = np.random.default_rng(4)
    lev_pl = np.ones(8, np.int32)
    lev_tp = rng.integers(0, 2, ntp).astype(np.int32)
    vb0 = tp["vh"].copy()
    for rh_scale in (6.0, 1e-6):
        ref_vb, ref_good, _ = oracle.symba_kick_list_pltp(i1, i2, None, lev_pl, lev_tp, p["rh"], p["rhill"] * rh_scale,
                                                          p["Gmass"], tp["rh"], 0.01, 1, 1, vb0)
        got_vb, got_good = ctx.symba_kick_list_pltp(i1, i2, None, lev_pl, lev_tp, p["rh"], p["rhill"] * rh_scale,
                                                    p["Gmass"], tp["rh"], 0.01, 1, 1, vb0)
        assert np.array_equal(got_good, ref_good) and ref_good.sum() > 0
        if rh_scale < 1:
            assert np.array_equal(got_vb, ref_vb)
        else:
            dv = np.abs(ref_vb - vb0).max()
            assert np.max(np.abs(got_vb - ref_vb)) < 1e-14 * dv


def test_collision_check_list_matches_oracle(ctx, oracle):
    n = 3000
    d, i1, i2 = _disk_list(ctx, oracle, n, 33, 6.0)
    rng = np.random.default_rng(6)
    mask = (rng.uniform(size=len(i1)) > 0.2).astype(np.int32)
    lvdotr = (rng.uniform(size=len(i1)) > 0.3).astype(np.int32)
    radius = d["radius"] * 300   # inflated: some pairs overlap now, some have q < rlim
    for dt in (d["dt"], 50 * d["dt"]):
        ref = oracle.collision_check_list(i1, i2, mask, lvdotr, d["rh"], d["vh"], d["Gmass"], radius, dt)
        got = ctx.collision_check_list(i1, i2, mask, lvdotr, d["rh"], d["vh"], d["Gmass"], radius, dt)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert ref[0].sum() > 0 and ref[1].sum() > 0
    assert not got[0][mask == 0].any() and not got[1][mask == 0].any()
    # pl-tp form
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=12)
    rc = p["rhill"] * 6.5 * 6
    _, j1, j2, _ = ctx.encounter_check_all_triangular_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], rc, 0.05)
    lv = np.ones(len(j1), np.int32)
    ref = oracle.collision_check_list(j1, j2, None, lv, p["rh"], p["vh"], p["Gmass"], p["rhill"] * 2, 5.0, r2=tp["rh"],
                                      v2=tp["vh"])
    got = ctx.collision_check_list(j1, j2, None, lv, p["rh"], p["vh"], p["Gmass"], p["rhill"] * 2, 5.0, r2=tp["rh"],
                                   v2=tp["vh"])
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert ctx.collision_check_list([], [], None, [], p["rh"], p["vh"], p["Gmass"], p["radius"], 1.0)[2] == 0


# ------------------------------------------------------------------------------------------------- tier 2 (resident)
import itertools

_generation = itertools.count(1000)  # swcu_body_sync re-uploads only when the generation (or the count) changes

def _resident_disk(ctx, oracle, n, seed, boost, rhill_scale=1.0, radius_scale=1.0):
    """A disk resident on the device (rh, vb = vh, Gmass, radius, rhill) and its all-pairs encounter list."""
    from swiftest_b200 import PL
    d, i1, i2 = _disk_list(ctx, oracle, n, seed, boost)
    d = dict(d)
    d["rhill"] = d["rhill"] * rhill_scale
    d["radius"] = d["radius"] * radius_scale
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], generation=next(_generation))
    ctx.body_put_vb(PL, d["vh"])
    return d, i1, i2


@pytest.mark.parametrize("irec,sgn", [(0, 1), (1, 1), (1, -1), (2, -1)])
def test_resident_symba_kick_list_plpl_equals_host_pointer_form_and_oracle(ctx, oracle, irec, sgn):
    """The resident form kicks pl%vb on the device: same bits as the host-pointer form (same kernels, SoA instead of AoS),
    same bar against the serial oracle; only the pair list and the level array cross PCIe."""
    from swiftest_b200 import PL
    n = 3000
    d, i1, i2 = _resident_disk(ctx, oracle, n, 21, 6.0, rhill_scale=3.0 if irec == 0 else 6.0)
    rng = np.random.default_rng(irec * 7 + sgn)
    levelg = rng.integers(max(irec - 1, 0), irec + 2, n).astype(np.int32)
    active = (rng.uniform(size=len(i1)) > 0.1).astype(np.int32)
    vb0 = d["vh"].copy()
    ref_vb, ref_good, _ = oracle.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], d["rhill"], d["Gmass"], d["dt"], irec,
                                                      sgn, vb0)
    t1_vb, t1_good = ctx.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], d["rhill"], d["Gmass"], d["dt"], irec, sgn, vb0)
    good = ctx.pl_symba_kick_list(i1, i2, active, levelg, d["dt"], irec, sgn)
    vb = ctx.body_get_vb(PL)["vb"]
    assert np.array_equal(good, ref_good) and np.array_equal(good, t1_good) and 0 < good.sum() < len(i1)
    assert np.array_equal(vb, t1_vb)
    touched = np.abs(ref_vb - vb0).sum(1) > 0
    scale = np.abs(ref_vb - vb0).max(1, keepdims=True)[touched]
    assert np.max(np.abs(vb[touched] - ref_vb[touched]) / scale) < 1e-14
    assert np.array_equal(vb[~touched], vb0[~touched])
    # a second level on top of the first, without reading lgood back (asynchronous call): vb keeps accumulating on the device
    assert ctx.pl_symba_kick_list(i1, i2, active, levelg, d["dt"] / 3, irec, -sgn, want_lgood=False) is None
    t1_vb2, _ = ctx.symba_kick_list_plpl(i1, i2, active, levelg, d["rh"], d["rhill"], d["Gmass"], d["dt"] / 3, irec, -sgn, t1_vb)
    assert np.array_equal(ctx.body_get_vb(PL)["vb"], t1_vb2)


def test_resident_symba_kick_list_pltp_equals_host_pointer_form_and_oracle(ctx, oracle):
    from swiftest_b200 import PL, TP
    p = W.planets8_year_units()
    ntp = 4000
    tp = W.tp_cloud(ntp, seed=8)
    rc = p["rhill"] * 6.5 * 6
    _, i1, i2, _ = ctx.encounter_check_all_triangular_pltp(8, ntp, p["rh"], p["vh"], tp["rh"], tp["vh"], rc, 0.05)
    rng = np.random.default_rng(4)
    lev_pl, lev_tp = np.ones(8, np.int32), rng.integers(0, 2, ntp).astype(np.int32)
    vb0 = tp["vh"].copy()
    for rh_scale in (6.0, 1e-6):
        ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"] * rh_scale, generation=next(_generation))
        ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], generation=next(_generation))
        ctx.body_put_vb(TP, vb0)
        ref_vb, ref_good, _ = oracle.symba_kick_list_pltp(i1, i2, None, lev_pl, lev_tp, p["rh"], p["rhill"] * rh_scale,
                                                          p["Gmass"], tp["rh"], 0.01, 1, 1, vb0)
        t1_vb, t1_good = ctx.symba_kick_list_pltp(i1, i2, None, lev_pl, lev_tp, p["rh"], p["rhill"] * rh_scale, p["Gmass"],
                                                  tp["rh"], 0.01, 1, 1, vb0)
        good = ctx.tp_symba_kick_list(i1, i2, None, lev_pl, lev_tp, 0.01, 1, 1)
        vb = ctx.body_get_vb(TP)["vb"]
        assert np.array_equal(good, ref_good) and ref_good.sum() > 0
        assert np.array_equal(vb, t1_vb)
        if rh_scale < 1:
            assert np.array_equal(vb, ref_vb)
        else:
            assert np.max(np.abs(vb - ref_vb)) < 1e-14 * np.abs(ref_vb - vb0).max()


def test_resident_symba_encounter_check_list_matches_oracle(ctx, oracle):
    from swiftest_b200 import PL, TP
    import swiftest_b200 as S
    n = 4000
    d, i1, i2 = _resident_disk(ctx, oracle, n, 77, 3.0, rhill_scale=3.0, radius_scale=200.0)
    rng = np.random.default_rng(1)
    mask = (rng.uniform(size=len(i1)) > 0.3).astype(np.int32)
    lv0 = np.full(len(i1), 5, np.int32)
    ctx.pl_set_renc(1)
    renc1 = oracle.set_renc(d["rhill"], 1)
    ref = oracle.symba_encounter_check_list(i1, i2, mask, d["rh"], d["vh"], renc1, d["radius"], d["dt"] / 3, lvdotr=lv0)
    got = ctx.body_symba_encounter_check_list(PL, i1, i2, mask, d["dt"] / 3, lvdotr=lv0)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert 0 < ref[2] < mask.sum() and (got[1][mask == 0] == 5).all()
    # the check reads pl%vb, not pl%vh: change vb on the device and the answer follows
    vb2 = d["vh"] * (1.0 + 0.3 * rng.normal(size=(n, 1)))
    ctx.body_put_vb(PL, vb2)
    ref2 = oracle.symba_encounter_check_list(i1, i2, mask, d["rh"], vb2, renc1, d["radius"], d["dt"] / 3, lvdotr=lv0)
    got2 = ctx.body_symba_encounter_check_list(PL, i1, i2, mask, d["dt"] / 3, lvdotr=lv0)
    assert np.array_equal(got2[0], ref2[0]) and np.array_equal(got2[1], ref2[1]) and got2[2] == ref2[2]
    assert not np.array_equal(ref2[0], ref[0])
    # pl-tp list
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=2)
    rhill4 = p["rhill"] * 4
    _, j1, j2, _ = ctx.encounter_check_all_triangular_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], rhill4 * 6.5, 0.05)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=rhill4, generation=next(_generation))
    ctx.body_sync(TP, 3000, r=tp["rh"], v=tp["vh"], generation=next(_generation))
    ctx.pl_set_renc(1)
    ref = oracle.symba_encounter_check_list(j1, j2, None, p["rh"], p["vh"], oracle.set_renc(rhill4, 1), p["radius"], 0.02,
                                            r2=tp["rh"], v2=tp["vh"])
    got = ctx.body_symba_encounter_check_list(TP, j1, j2, None, 0.02)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2] and len(j1) > 5
    with pytest.raises(S.SwcuError):
        ctx.body_symba_encounter_check_list(PL, [1], [9999], None, 0.02)
    assert ctx.body_symba_encounter_check_list(PL, [], [], None, 0.02)[2] == 0


def test_resident_collision_check_list_matches_oracle(ctx, oracle):
    from swiftest_b200 import PL, TP
    n = 3000
    d, i1, i2 = _resident_disk(ctx, oracle, n, 33, 6.0, radius_scale=300.0)
    rng = np.random.default_rng(6)
    mask = (rng.uniform(size=len(i1)) > 0.2).astype(np.int32)
    lvdotr = (rng.uniform(size=len(i1)) > 0.3).astype(np.int32)
    for dt in (d["dt"], 50 * d["dt"]):
        ref = oracle.collision_check_list(i1, i2, mask, lvdotr, d["rh"], d["vh"], d["Gmass"], d["radius"], dt)
        got = ctx.body_collision_check_list(PL, i1, i2, mask, lvdotr, dt)
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]
    assert ref[0].sum() > 0 and ref[1].sum() > 0
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=12)
    _, j1, j2, _ = ctx.encounter_check_all_triangular_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 39, 0.05)
    lv = np.ones(len(j1), np.int32)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["rhill"] * 2, rhill=p["rhill"], generation=next(_generation))
    ctx.body_sync(TP, 3000, r=tp["rh"], v=tp["vh"], generation=next(_generation))
    ref = oracle.collision_check_list(j1, j2, None, lv, p["rh"], p["vh"], p["Gmass"], p["rhill"] * 2, 5.0, r2=tp["rh"],
                                      v2=tp["vh"])
    got = ctx.body_collision_check_list(TP, j1, j2, None, lv, 5.0)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and got[2] == ref[2]


@pytest.mark.parametrize("npl,ntp", [(8, 20000), (300, 5000)])
def test_resident_discard_pl_tp_matches_oracle(ctx, oracle, npl, ntp):
    """swcu_tp_discard_pl: the same double loop on the resident populations -- active flags from swcu_body_set_active, else
    lmask, else everybody; the indices come back only when somebody is discarded."""
    from swiftest_b200 import PL, TP
    rng = np.random.default_rng(npl + ntp)
    if npl == 8:
        p = W.planets8_year_units()
        rpl, vpl = p["rh"], p["vh"]
    else:
        d = W.disk(npl, seed=npl)
        rpl, vpl = d["rh"], d["vh"]
    host = rng.integers(0, npl, ntp)
    rtp = rpl[host] + rng.normal(scale=0.02, size=(ntp, 3))
    vtp = vpl[host] + rng.normal(scale=1.0, size=(ntp, 3))
    radius = np.full(npl, 0.01)
    act = (rng.uniform(size=ntp) > 0.1).astype(np.int32)
    ctx.body_sync(PL, npl, nplm=npl, r=rpl, v=vpl, Gmass=np.ones(npl), radius=radius, rhill=radius, generation=next(_generation))
    ctx.body_sync(TP, ntp, r=rtp, v=vtp, generation=next(_generation))
    ref_all, n_all = oracle.discard_pl_tp(rtp, vtp, None, rpl, vpl, radius, 0.01)
    got, n = ctx.tp_discard_pl(0.01)
    assert np.array_equal(got, ref_all) and n == n_all and 0 < n < ntp
    ctx.body_set_active(TP, act)
    ref, nref = oracle.discard_pl_tp(rtp, vtp, act, rpl, vpl, radius, 0.01)
    got, n = ctx.tp_discard_pl(0.01)
    assert np.array_equal(got, ref) and n == nref and not got[act == 0].any()
    assert ctx.tp_discard_pl(0.01, want_iplanet=False) == (None, nref)
    # nobody close: count 0, indices zero-filled without a copy
    ctx.body_put(TP, r=rtp + 50.0)
    got, n = ctx.tp_discard_pl(0.01)
    assert n == 0 and not got.any()
    ctx.body_set_active(TP, None)


def test_resident_triangular_encounter_checks_match_oracle(ctx, oracle):
    """swcu_pl/tp_encounter_check_triangular: ENCOUNTER_CHECK TRIANGULAR on the resident populations, incl. the GMTINY split
    (plm x plm + plm x plt merged like encounter_check_all_plplm)."""
    from swiftest_b200 import PL, TP
    f = W.fixture("108pl_50tp")
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    nplm = int((f["pl_Gmass"] >= float(f["GMTINY"])).sum())
    r, v, rhill = f["pl_rh"][order], f["pl_vh"][order], f["pl_rhill"][order] * 20
    ctx.body_sync(PL, 108, nplm=nplm, r=r, v=v, Gmass=f["pl_Gmass"][order], radius=f["pl_radius"][order], rhill=rhill,
                  generation=next(_generation))
    ctx.body_sync(TP, 50, r=f["tp_rh"], v=f["tp_vh"], generation=next(_generation))
    ctx.pl_set_renc(0)
    rc = oracle.set_renc(rhill, 0)
    a = oracle.encounter_plpl(r[:nplm], v[:nplm], rc[:nplm], 0.05, triangular=True)
    b = oracle.encounter_plplm(r[:nplm], v[:nplm], r[nplm:], v[nplm:], rc[:nplm], rc[nplm:], 0.05, triangular=True)
    ref = sorted(zip(a[0].tolist(), a[1].tolist())) + sorted((i, j + nplm) for i, j in zip(b[0].tolist(), b[1].tolist()))
    ref = sorted(ref)
    n, g1, g2, _ = ctx.pl_encounter_check_triangular(0.05)
    assert n == len(ref) and n > 0 and list(zip(g1.tolist(), g2.tolist())) == ref
    rt = oracle.encounter_pltp(r, v, f["tp_rh"], f["tp_vh"], rc, 0.05, triangular=True)
    _same(ctx.tp_encounter_check_triangular(0.05), rt)
    # fully interacting population: no split
    d = W.disk(900, seed=3)
    ctx.body_sync(PL, 900, nplm=900, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"] * 4,
                  generation=next(_generation))
    ctx.pl_set_renc(0)
    _same(ctx.pl_encounter_check_triangular(d["dt"]),
          oracle.encounter_plpl(d["rh"], d["vh"], oracle.set_renc(d["rhill"] * 4, 0), d["dt"], triangular=True))
