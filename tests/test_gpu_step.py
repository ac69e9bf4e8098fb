"""GPU parity tests for the rows SURVEY.md section 8(f) ranks next: the O(N) glue of the democratic-heliocentric step
on the device-resident populations (rank 1) and the energy / momentum sums (rank 2), against the CPU restatement.

Bars: the element-wise glue is BIT-EXACT whenever the sums are (n <= 1024: the device adds in the reference's serial
order); above that the fixed-tree sums are compared at 1e-14 relative to sum|terms|.  Whole steps inherit the 1e-12 bar
of the accelerations inside them.  Energy sums: 1e-13 relative to sum|terms|.
"""
import numpy as np
import pytest

from swiftest_b200 import workloads as W
from swiftest_b200 import PL, TP, LOOP_FLAT, LOOP_TRIANGULAR

pytestmark = pytest.mark.gpu


def _fixture108():
    f = W.fixture("108pl_50tp")
    return f, float(f["cb_Gmass"]), float(f["dt"])


def _sync_pl(ctx, gen, rh, vh, Gm, radius, GMcb, lmask=None):
    n = len(Gm)
    ctx.body_sync(PL, n, nplm=n, r=rh, v=vh, Gmass=Gm, radius=radius, mu=np.full(n, GMcb), lmask=lmask, generation=gen)


# ------------------------------------------------------------------------------------------------ glue, one call each
def test_glue_calls_bit_exact_on_the_108_body_fixture(ctx, oracle):
    f, GMcb, dt = _fixture108()
    Gm, rh, vh = f["pl_Gmass"], f["pl_rh"], f["pl_vh"]
    mask = np.ones(108, np.int32)
    mask[[3, 50, 107]] = 0
    _sync_pl(ctx, 9001, rh, vh, Gm, f["pl_radius"], GMcb, lmask=mask)
    # vh2vb: unmasked (swiftest_util.f90:440-455)
    vb_ref, vbcb_ref = oracle.coord_vh2vb_pl(GMcb, Gm, vh)
    vbcb = ctx.pl_vh2vb(GMcb)
    assert np.array_equal(vbcb, vbcb_ref)
    assert np.array_equal(ctx.body_get_vb(PL)["vb"], vb_ref)
    # lindrift: masked sum, masked update
    r_ref, pt_ref = oracle.helio_drift_linear_pl(GMcb, Gm, vb_ref, rh, 0.5 * dt, mask)
    pt = ctx.pl_lindrift(GMcb, 0.5 * dt, lbeg=True)
    assert np.array_equal(pt, pt_ref)
    assert np.array_equal(ctx.body_get(PL)["r"], r_ref)
    ptb, pte = ctx.cb_get_pt()
    assert np.array_equal(ptb, pt_ref) and pte.shape == (3,)
    # kick tail: vb += ah*dt under the mask, rbeg = rh for everyone
    ah = np.random.default_rng(2).normal(scale=1e-2, size=(108, 3))
    ctx.body_put(PL, a=ah)
    ctx.body_kick_vb(PL, 0.5 * dt, lbeg=True)
    got = ctx.body_get_vb(PL, vb=True, rbeg=True)
    assert np.array_equal(got["vb"], oracle.helio_kick_vb(ah, vb_ref, 0.5 * dt, mask))
    assert np.array_equal(got["rbeg"], r_ref)
    # vb2vh: reversed sum with a division per term, unmasked update.  The reference filters on status /= INACTIVE
    # (swiftest_util.f90:377), NOT on lmask: with no active flags loaded every body counts ...
    vb_now = got["vb"]
    vh_ref, vbcb2_ref = oracle.coord_vb2vh_pl(GMcb, Gm, vb_now, None)
    vbcb2 = ctx.pl_vb2vh(GMcb)
    assert np.array_equal(vbcb2, vbcb2_ref)
    assert np.array_equal(ctx.body_get(PL)["v"], vh_ref)
    # ... and with swcu_body_set_active the inactive ones drop out of the sum
    lactive = np.ones(108, np.int32)
    lactive[[3, 50, 107]] = 0
    ctx.body_set_active(PL, lactive)
    vh_ref, vbcb2_ref = oracle.coord_vb2vh_pl(GMcb, Gm, vb_now, lactive)
    assert np.array_equal(ctx.pl_vb2vh(GMcb), vbcb2_ref)
    assert np.array_equal(ctx.body_get(PL)["v"], vh_ref)
    ctx.body_set_active(PL, None)
    assert np.array_equal(ctx.pl_vb2vh(GMcb), oracle.coord_vb2vh_pl(GMcb, Gm, vb_now, None)[1])


@pytest.mark.parametrize("n", [1025, 5000, 70001])
def test_glue_tree_sums_are_reproducible_and_within_tolerance(ctx, oracle, n):
    d = W.disk(n, seed=n)
    GMcb = 39.476926408897626
    vbcbs = []
    for rep in range(2):
        _sync_pl(ctx, 9100 + 2 * n + rep, d["rh"], d["vh"], d["Gmass"], d["radius"], GMcb)
        vbcbs.append(ctx.pl_vh2vb(GMcb))
    assert np.array_equal(vbcbs[0], vbcbs[1])  # same bits run to run
    vb_ref, vbcb_ref = oracle.coord_vh2vb_pl(GMcb, d["Gmass"], d["vh"])
    scale = (d["Gmass"][:, None] * np.abs(d["vh"])).sum(0) / (GMcb + d["Gmass"].sum())
    assert np.all(np.abs(vbcbs[0] - vbcb_ref) <= 1e-14 * scale)
    vb = ctx.body_get_vb(PL)["vb"]
    assert np.max(np.abs(vb - vb_ref)) <= 1e-14 * scale.max() + 2e-16 * np.abs(vb_ref).max()
    rh_ref, pt_ref = oracle.helio_drift_linear_pl(GMcb, d["Gmass"], vb, d["rh"], 0.01)
    pt = ctx.pl_lindrift(GMcb, 0.01, lbeg=False)
    scale_pt = (d["Gmass"][:, None] * np.abs(vb)).sum(0) / GMcb
    assert np.all(np.abs(pt - pt_ref) <= 1e-14 * scale_pt)
    vh_ref, vbcb2_ref = oracle.coord_vb2vh_pl(GMcb, d["Gmass"], vb)
    assert np.all(np.abs(ctx.pl_vb2vh(GMcb) - vbcb2_ref) <= 1e-14 * scale_pt)


def test_glue_state_errors_and_empty(ctx):
    import swiftest_b200 as S
    with S.Context() as c:
        with pytest.raises(S.SwcuError):
            c.pl_vh2vb(1.0)
        with pytest.raises(S.SwcuError):
            c.helio_step_tp(1.0, 0.01)
        c.body_sync(PL, 0, generation=1)
        c.body_sync(TP, 0, generation=2)
        assert c.helio_step_pl(1.0, 0.01, lfirst=True) == 0  # helio_step.f90:54: nothing to do
        assert c.helio_step_tp(1.0, 0.01, lfirst=True) == 0
        with pytest.raises(S.SwcuError):
            c.helio_step_pl(0.0, 0.01)


# ------------------------------------------------------------------------------------------------ whole steps
@pytest.mark.parametrize("variant", [LOOP_TRIANGULAR, LOOP_FLAT])
def test_helio_step_pl_tracks_oracle_on_fixture(ctx, oracle, variant):
    f, GMcb, dt = _fixture108()
    Gm, rad = f["pl_Gmass"], f["pl_radius"]
    _sync_pl(ctx, 9200 + variant, f["pl_rh"], f["pl_vh"], Gm, rad, GMcb)
    st = dict(rh=f["pl_rh"].copy(), vh=f["pl_vh"].copy(), vb=np.zeros((108, 3)), lfirst=True)
    nsteps = 40
    for k in range(nsteps):
        assert not oracle.helio_step_pl(st, GMcb, Gm, rad, dt, lflat=(variant == LOOP_FLAT)).any()
        assert ctx.helio_step_pl(GMcb, dt, loop_variant=variant, lclose=True, lfirst=(k == 0)) == 0
    out = ctx.body_get(PL)
    hv = ctx.body_get_vb(PL, vb=True, rbeg=True, rend=True)
    rs, vs = np.abs(st["rh"]).max(), np.abs(st["vh"]).max()
    # 40 steps of 1e-12-accurate kicks; errors grow linearly at worst
    assert np.max(np.abs(out["r"] - st["rh"])) < 1e-11 * rs
    assert np.max(np.abs(out["v"] - st["vh"])) < 1e-11 * vs
    assert np.max(np.abs(hv["vb"] - st["vb"])) < 1e-11 * vs
    assert np.max(np.abs(hv["rbeg"] - st["rbeg"])) < 1e-11 * rs
    assert np.max(np.abs(hv["rend"] - st["rend"])) < 1e-11 * rs
    ptb, pte = ctx.cb_get_pt()
    assert np.max(np.abs(ptb - st["ptbeg"])) < 1e-11 * np.abs(st["ptbeg"]).max()
    assert np.max(np.abs(pte - st["ptend"])) < 1e-11 * np.abs(st["ptend"]).max()


@pytest.mark.parametrize("variant", [LOOP_TRIANGULAR, LOOP_FLAT])
def test_helio_step_pl_graph_replay_equals_stream_ordered_launches(ctx, variant):
    """npl > 128: the step is ~21 launches; from the third step on it is replayed as one CUDA graph.  Same kernels, same
    arguments: the full-row variant must give identical bits with SWCU_STEP_GRAPH=0, the third-law variant (FP64 atomics)
    agrees to rounding; the launch count per step is the same; the drift-failure count still comes back."""
    import os
    n, nsteps = 700, 9
    d = W.disk(n, seed=41)
    GMcb = W.GMSUN

    def run(graph, gen):
        if not graph:
            os.environ["SWCU_STEP_GRAPH"] = "0"
        try:
            ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                          mu=np.full(n, GMcb), generation=gen)
            r0, l0 = ctx.step_graph_replays(), ctx.launch_count()
            for k in range(nsteps):
                assert ctx.helio_step_pl(GMcb, d["dt"], loop_variant=variant, lclose=True, lfirst=(k == 0),
                                         want_nfail=(k % 2 == 0)) == 0
            out = ctx.body_get(PL)
            return out["r"], out["v"], ctx.body_get_vb(PL)["vb"], ctx.step_graph_replays() - r0, ctx.launch_count() - l0
        finally:
            os.environ.pop("SWCU_STEP_GRAPH", None)

    rg, vg, bg, nrep, nl = run(True, 9400 + variant)
    rs, vs, bs, nrep0, nl0 = run(False, 9410 + variant)
    assert nrep == nsteps - 3 and nrep0 == 0      # step 0 is the first step, 1 warms up, 2 is captured (and run), 3.. replay
    assert nl == nl0
    if variant == LOOP_TRIANGULAR:
        assert np.array_equal(rg, rs) and np.array_equal(vg, vs) and np.array_equal(bg, bs)
    else:
        assert np.max(np.abs(rg - rs)) < 1e-13 * np.abs(rs).max() and np.max(np.abs(vg - vs)) < 1e-13 * np.abs(vs).max()


def test_helio_step_pl_first_step_matches_unfused_device_calls(ctx):
    """The one-call step and the same step issued call by call through the ABI give identical bits."""
    f, GMcb, dt = _fixture108()
    Gm, rad = f["pl_Gmass"], f["pl_radius"]
    _sync_pl(ctx, 9301, f["pl_rh"], f["pl_vh"], Gm, rad, GMcb)
    ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_TRIANGULAR, lclose=True, lfirst=True)
    a = ctx.body_get(PL)
    avb = ctx.body_get_vb(PL)["vb"]
    _sync_pl(ctx, 9302, f["pl_rh"], f["pl_vh"], Gm, rad, GMcb)
    dth = 0.5 * dt
    ctx.pl_vh2vb(GMcb, want=False)
    ctx.pl_lindrift(GMcb, dth, True, want=False)
    for lbeg in (True, False):
        ctx.body_zero_accel(PL)
        ctx.pl_accel_int(LOOP_TRIANGULAR, True)
        ctx.body_kick_vb(PL, dth, lbeg)
        if lbeg:
            assert ctx.body_drift_vb(PL, GMcb, dt) == 0
    ctx.pl_lindrift(GMcb, dth, False, want=False)
    ctx.pl_vb2vh(GMcb, want=False)
    b = ctx.body_get(PL)
    assert np.array_equal(a["r"], b["r"]) and np.array_equal(a["v"], b["v"])
    assert np.array_equal(avb, ctx.body_get_vb(PL)["vb"])


def test_helio_step_tp_fused_kernel_tracks_oracle(ctx, oracle):
    p = W.planets8_year_units()
    GMcb, dt = float(p["cb_Gmass"]), 0.01
    ntp = 20000
    tp = W.tp_cloud(ntp, seed=9)
    mask = (np.random.default_rng(9).uniform(size=ntp) > 0.03).astype(np.int32)
    on = mask.astype(bool)
    _sync_pl(ctx, 9401, p["rh"], p["vh"], p["Gmass"], p["radius"], GMcb)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, GMcb), lmask=mask, generation=9402)
    pl = dict(rh=p["rh"].copy(), vh=p["vh"].copy(), vb=np.zeros((8, 3)), lfirst=True)
    st = dict(rh=tp["rh"].copy(), vh=tp["vh"].copy(), vb=np.zeros((ntp, 3)), lfirst=True)
    for k in range(10):
        oracle.helio_step_pl(pl, GMcb, p["Gmass"], p["radius"], dt)
        fl = oracle.helio_step_tp(st, pl, GMcb, p["Gmass"], dt, lmask=mask)
        assert ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_TRIANGULAR, lclose=True, lfirst=(k == 0)) == 0
        assert ctx.helio_step_tp(GMcb, dt, lfirst=(k == 0)) == int(np.count_nonzero(fl[on]))
    out = ctx.body_get(TP, iflag=True)
    vb = ctx.body_get_vb(TP)["vb"]
    rn = np.linalg.norm(st["rh"], axis=1, keepdims=True)
    vn = np.linalg.norm(st["vh"], axis=1, keepdims=True)
    assert np.max(np.abs(out["r"][on] - st["rh"][on]) / rn[on]) < 1e-11
    assert np.max(np.abs(out["v"][on] - st["vh"][on]) / vn[on]) < 1e-11
    assert np.max(np.abs(vb[on] - st["vb"][on]) / vn[on]) < 1e-11
    assert np.max(np.abs(out["a"][on] - st["ah"][on])) <= 1e-11 * np.abs(st["ah"]).max()
    off = ~on  # masked particles are not touched (helio_step.f90 loops are all under lmask)
    assert np.array_equal(out["r"][off], tp["rh"][off]) and np.array_equal(out["v"][off], tp["vh"][off])


def test_helio_step_tp_with_planets_stepped_elsewhere(ctx, oracle):
    """tp shards on other GPUs receive ptbeg / ptend from the planets' owner (swcu_cb_set_pt) and rbeg / rend as
    resident planet positions: same result as the local sequence."""
    p = W.planets8_year_units()
    GMcb, dt = float(p["cb_Gmass"]), 0.01
    ntp = 3000
    tp = W.tp_cloud(ntp, seed=19)
    _sync_pl(ctx, 9501, p["rh"], p["vh"], p["Gmass"], p["radius"], GMcb)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, GMcb), generation=9502)
    ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_TRIANGULAR, lfirst=True)
    ctx.helio_step_tp(GMcb, dt, lfirst=True)
    a = ctx.body_get(TP)
    ptb, pte = ctx.cb_get_pt()
    # unfused tp sequence with ptbeg/ptend loaded explicitly
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, GMcb), generation=9503)
    ctx.cb_set_pt(ptb, pte)
    ctx.tp_vh2vb(lbeg=True)
    ctx.tp_lindrift(0.5 * dt, lbeg=True)
    b = ctx.body_get(TP)
    st = dict(rh=tp["rh"] + ptb * (0.5 * dt))
    assert np.array_equal(b["r"], st["rh"])
    assert np.array_equal(ctx.body_get_vb(TP)["vb"], tp["vh"] + (-ptb))
    assert np.all(np.isfinite(a["r"]))


# ------------------------------------------------------------------------------------------------ energy
def _energy_inputs(n, seed, nmask=0):
    d = W.disk(n, seed=seed)
    GMcb = 39.476926408897626
    GU = GMcb
    mass = d["Gmass"] / GU
    mask = np.ones(n, np.int32)
    if nmask:
        mask[np.random.default_rng(seed).choice(n, nmask, replace=False)] = 0
    return d, GMcb, mass, mask


@pytest.mark.parametrize("n,nmask", [(1, 0), (2, 0), (108, 5), (257, 0), (1500, 40), (4100, 0)])
def test_potential_energy_matches_oracle(ctx, oracle, n, nmask):
    d, GMcb, mass, mask = _energy_inputs(n, 600 + n, nmask)
    lm = mask if nmask else None
    ref = oracle.get_potential_energy(GMcb, d["Gmass"], mass, d["rh"], lm)
    got = ctx.util_get_potential_energy(n, lm, GMcb, d["Gmass"], mass, d["rh"])
    assert abs(got - ref) <= 1e-13 * abs(ref)
    # flat and triangular reference variants are the same number up to summation order
    flat = oracle.get_potential_energy(GMcb, d["Gmass"], mass, d["rh"], lm, flat=True)
    assert abs(got - flat) <= 1e-13 * abs(ref)


def test_potential_energy_edge_cases(ctx, oracle):
    d, GMcb, mass, mask = _energy_inputs(300, 77)
    rb = d["rh"].copy()
    # a masked-out body sitting exactly on an active one and one at the origin: no contribution, no NaN
    rb[10] = rb[200]
    rb[11] = 0.0
    mask[[10, 11]] = 0
    ref = oracle.get_potential_energy(GMcb, d["Gmass"], mass, rb, mask)
    got = ctx.util_get_potential_energy(300, mask, GMcb, d["Gmass"], mass, rb)
    assert np.isfinite(got) and abs(got - ref) <= 1e-13 * abs(ref)
    # two ACTIVE bodies at the same place: the reference divides by zero -> -inf; so does the device (IEEE redo path)
    mask[10] = 1
    ref = oracle.get_potential_energy(GMcb, d["Gmass"], mass, rb, mask)
    got = ctx.util_get_potential_energy(300, mask, GMcb, d["Gmass"], mass, rb)
    assert ref == -np.inf and got == -np.inf
    # coordinates far outside the FP32 exponent range take the IEEE path too
    big = d["rh"] * 1e60
    ref = oracle.get_potential_energy(GMcb, d["Gmass"], mass, big)
    got = ctx.util_get_potential_energy(300, None, GMcb, d["Gmass"], mass, big)
    assert abs(got - ref) <= 1e-13 * abs(ref)
    # beyond 2^500 the FP64 seed is not used either (r2 may overflow): the whole tile takes the IEEE path
    huge = d["rh"] * 1e160
    ref = oracle.get_potential_energy(GMcb, d["Gmass"], mass, huge)
    got = ctx.util_get_potential_energy(300, None, GMcb, d["Gmass"], mass, huge)
    assert abs(got - ref) <= 1e-13 * abs(ref)
    assert ctx.util_get_potential_energy(0, None, GMcb, np.zeros(0), np.zeros(0), np.zeros((0, 3))) == 0.0


def test_energy_and_momentum_matches_oracle(ctx, oracle):
    f, GMcb, _ = _fixture108()
    Gm = f["pl_Gmass"]
    mass = Gm / GMcb
    mask = np.ones(108, np.int32)
    mask[[7, 90]] = 0
    rb, vb, rbcb, vbcb = oracle.coord_h2b_pl(GMcb, Gm, f["pl_rh"], f["pl_vh"])
    for lm, lclose in ((None, True), (mask, True), (mask, False)):
        ref = oracle.get_energy_and_momentum(GMcb, 1.0, rbcb, vbcb, Gm, mass, f["pl_radius"], rb, vb, lm, lclose)
        got = ctx.util_get_energy_and_momentum(108, lm, GMcb, 1.0, rbcb, vbcb, Gm, mass, f["pl_radius"], rb, vb, lclose)
        for k in ("ke_orbit", "pe", "be", "te", "GMtot"):
            assert abs(got[k] - ref[k]) <= 1e-13 * max(abs(ref[k]), abs(ref["pe"]) if k == "te" else 0.0), k
        assert np.max(np.abs(got["L_orbit"] - ref["L_orbit"])) <= 1e-13 * np.abs(ref["L_orbit"]).max()


def test_potential_energy_large_disk_against_blocked_numpy(ctx):
    """Full-size property check (no oracle pass at this size): 2e4 bodies against a blocked float64 numpy sum."""
    n = 20000
    d, GMcb, mass, _ = _energy_inputs(n, 4242)
    rb = d["rh"]
    tot = 0.0
    for i0 in range(0, n, 2000):
        blk = rb[i0:i0 + 2000]
        dist = np.sqrt(((blk[:, None, :] - rb[None, :, :]) ** 2).sum(2))
        w = d["Gmass"][i0:i0 + 2000, None] * mass[None, :]
        jj = np.arange(n)[None, :]
        ii = np.arange(i0, min(i0 + 2000, n))[:, None]
        sel = jj > ii
        tot += (w[sel] / dist[sel]).sum()
    ref = -tot - (GMcb * mass / np.linalg.norm(rb, axis=1)).sum()
    got = ctx.util_get_potential_energy(n, None, GMcb, d["Gmass"], mass, rb)
    assert abs(got - ref) <= 1e-12 * abs(ref)
    assert got == ctx.util_get_potential_energy(n, None, GMcb, d["Gmass"], mass, rb)  # reproducible bits


# ---------------------------------------------------------------------------------------------- WHM planet step
def test_whm_step_pl_resident_matches_oracle(ctx, oracle):
    """whm_step_pl on the device-resident planets (Jacobi chains, ah0 + ah1 + ah2 + accel_int, kick, drift, kick) against
    the C restatement of whm/whm_step.f90:37-69, step after step: the serial chains keep the reference's order, so the
    Jacobi coordinates and the state agree to rounding of the O(N^2) term only."""
    from swiftest_b200 import LOOP_TRIANGULAR, LOOP_FLAT
    for name, n_steps in (("8pl", 40), ("108pl", 10)):
        if name == "8pl":
            p = W.planets8_year_units()
            GMcb, Gm, rh, vh, rad, rhill, dt = p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], p["rhill"], 0.01
        else:
            f = W.fixture("108pl_50tp")
            order = np.argsort(-f["pl_Gmass"], kind="stable")
            GMcb, dt = float(f["cb_Gmass"]), float(f["dt"])
            Gm, rh, vh, rad, rhill = (f["pl_" + k][order] for k in ("Gmass", "rh", "vh", "radius", "rhill"))
        n = len(Gm)
        mask = np.ones(n, np.int32)
        if name == "108pl":
            mask[[5, 77]] = 0
        for variant, lflat in ((LOOP_TRIANGULAR, False), (LOOP_FLAT, True)):
            st = {"rh": rh.copy(), "vh": vh.copy(), "lfirst": True}
            ctx.body_sync(PL, n, nplm=n, r=rh, v=vh, Gmass=Gm, radius=rad, rhill=rhill, mu=GMcb + Gm, lmask=mask,
                          generation=8800 + 10 * n + int(lflat))
            for k in range(n_steps):
                fl = oracle.whm_step_pl(st, GMcb, Gm, rad, dt, lflat=lflat, lmask=mask)
                assert not fl.any()
                assert ctx.whm_step_pl(GMcb, dt, variant, True, lfirst=(k == 0)) == 0
            got = ctx.body_get(PL)
            xj, vj = ctx.whm_get_jacobi()
            scale_r, scale_v = np.linalg.norm(st["rh"], axis=1, keepdims=True), np.linalg.norm(st["vh"], axis=1, keepdims=True)
            assert np.max(np.abs(got["r"] - st["rh"]) / scale_r) < 1e-11, (name, lflat)
            assert np.max(np.abs(got["v"] - st["vh"]) / scale_v) < 1e-11
            assert np.max(np.abs(xj - st["xj"]) / scale_r) < 1e-11 and np.max(np.abs(vj - st["vj"]) / scale_v) < 1e-11
            on = mask.astype(bool)
            assert np.max(np.abs(got["a"][on] - st["ah"][on])) <= 1e-11 * np.abs(st["ah"]).max()
            vb = ctx.body_get_vb(PL, vb=False, rbeg=True, rend=True)
            assert np.max(np.abs(vb["rbeg"] - st["rbeg"])) <= 1e-11 * np.abs(st["rbeg"]).max()
            assert np.max(np.abs(vb["rend"] - st["rend"])) <= 1e-11 * np.abs(st["rend"]).max()


@pytest.mark.parametrize("n", [130, 300, 1500])
def test_whm_multi_launch_step_on_larger_systems_matches_oracle(ctx, oracle, n):
    """npl > 128: the multi-launch form.  Its Jacobi chains run tile by tile (256 bodies: terms in parallel, additions in the
    reference's order from shared memory): after every step the heliocentric positions must be EXACTLY j2h of the device's
    own Jacobi coordinates (the chain, bit for bit, across tile boundaries), and the step tracks the C restatement of
    whm_step_pl to the rounding of the O(N^2) term; masked bodies included."""
    from swiftest_b200 import LOOP_TRIANGULAR
    d = W.disk(n, seed=900 + n)
    GMcb, Gm, dt = W.GMSUN, d["Gmass"] * 300.0, d["dt"]   # heavier bodies: the Jacobi corrections are far above rounding
    mask = np.ones(n, np.int32)
    mask[[7, n // 2, n - 2]] = 0
    st = {"rh": d["rh"].copy(), "vh": d["vh"].copy(), "lfirst": True}
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=Gm, radius=d["radius"], rhill=d["rhill"], mu=GMcb + Gm,
                  lmask=mask, generation=8700 + n)
    _, eta, _ = oracle.whm_set_mu_eta(GMcb, Gm)
    for k in range(4):
        assert not oracle.whm_step_pl(st, GMcb, Gm, d["radius"], dt, lflat=False, lmask=mask).any()
        assert ctx.whm_step_pl(GMcb, dt, LOOP_TRIANGULAR, True, lfirst=(k == 0)) == 0
        got = ctx.body_get(PL)
        xj, vj = ctx.whm_get_jacobi()
        rh_from_xj, _ = oracle.whm_coord_j2h(Gm, eta, xj, vj)
        assert np.array_equal(got["r"], rh_from_xj)
    sr, sv = np.linalg.norm(st["rh"], axis=1, keepdims=True), np.linalg.norm(st["vh"], axis=1, keepdims=True)
    assert np.max(np.abs(got["r"] - st["rh"]) / sr) < 1e-11 and np.max(np.abs(got["v"] - st["vh"]) / sv) < 1e-11
    assert np.max(np.abs(xj - st["xj"]) / sr) < 1e-11 and np.max(np.abs(vj - st["vj"]) / sv) < 1e-11
    on = mask.astype(bool)
    assert np.max(np.abs(got["a"][on] - st["ah"][on])) <= 1e-11 * np.abs(st["ah"]).max()


def test_whm_first_step_is_bit_identical_where_the_chains_decide(ctx, oracle):
    """One first step of the Sun + 8 planets system: eta/muj, h2j and the ah0/ah1/ah2 chains run in the reference's serial
    order with no FMA contraction; only pl%accel_int (28 pairs) and libm-free drift arithmetic follow, so positions agree
    to a few ulp."""
    from swiftest_b200 import LOOP_TRIANGULAR
    p = W.planets8_year_units()
    st = {"rh": p["rh"].copy(), "vh": p["vh"].copy(), "lfirst": True}
    oracle.whm_step_pl(st, p["cb_Gmass"], p["Gmass"], p["radius"], 0.01)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=8899)
    assert ctx.whm_step_pl(p["cb_Gmass"], 0.01, LOOP_TRIANGULAR, True, lfirst=True) == 0
    got = ctx.body_get(PL)
    assert np.max(np.abs(got["r"] - st["rh"]) / np.linalg.norm(st["rh"], axis=1, keepdims=True)) < 2e-15
    assert np.max(np.abs(got["v"] - st["vh"]) / np.linalg.norm(st["vh"], axis=1, keepdims=True)) < 2e-15


def test_whm_resident_planets_and_test_particles_never_leave_the_device(ctx, oracle):
    """BASELINE configs[1] flow: swcu_whm_tp_first_accel, then per step swcu_whm_step_pl + swcu_whm_tp_step(ah0 = NULL: the
    value the planet step left on the device) against swo_whm_step_pl + swo_whm_step_tp."""
    from swiftest_b200 import LOOP_TRIANGULAR
    p = W.planets8_year_units()
    ntp, dt, nsteps = 20000, 0.01, 5
    tp = W.tp_cloud(ntp, seed=17)
    GMcb = p["cb_Gmass"]
    spl = {"rh": p["rh"].copy(), "vh": p["vh"].copy(), "lfirst": True}
    stp = {"rh": tp["rh"].copy(), "vh": tp["vh"].copy(), "lfirst": True}
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=GMcb + p["Gmass"], generation=8901)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, GMcb), generation=8902)
    ctx.whm_tp_first_accel()
    for k in range(nsteps):
        oracle.whm_step_pl(spl, GMcb, p["Gmass"], p["radius"], dt)
        fl = oracle.whm_step_tp(stp, spl, GMcb, p["Gmass"], dt)
        assert not fl.any()
        assert ctx.whm_step_pl(GMcb, dt, LOOP_TRIANGULAR, True, lfirst=(k == 0)) == 0
        assert ctx.whm_tp_step(dt, None) == 0
    got = ctx.body_get(TP)
    assert np.max(np.abs(got["r"] - stp["rh"]) / np.linalg.norm(stp["rh"], axis=1, keepdims=True)) < 1e-11
    assert np.max(np.abs(got["v"] - stp["vh"]) / np.linalg.norm(stp["vh"], axis=1, keepdims=True)) < 1e-11
    assert np.max(np.abs(got["a"] - stp["ah"])) <= 1e-11 * np.abs(stp["ah"]).max()


@pytest.mark.parametrize("lflat", [False, True])
def test_whm_small_system_step_is_one_launch_and_bit_identical_to_the_reference_order(ctx, oracle, lflat):
    """npl <= 128: swcu_whm_step_pl is ONE kernel (drift_kernels.cu::whm_step_pl_small_kernel).  Its pair sums run in the
    reference's own order with the reference's expression, so 40 steps of Sun + 8 planets reproduce the CPU restatement
    (itself bit-identical to the interpreted Fortran, tests/test_oracle_fortran_goldens.py) BIT FOR BIT, for the full-row
    and for the flat loop: r, v, Jacobi coordinates, accelerations, rbeg / rend."""
    from swiftest_b200 import LOOP_TRIANGULAR, LOOP_FLAT
    p = W.planets8_year_units()
    GMcb, dt = p["cb_Gmass"], 0.01
    st = {"rh": p["rh"].copy(), "vh": p["vh"].copy(), "lfirst": True}
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=GMcb + p["Gmass"], generation=8950 + int(lflat))
    n0 = ctx.launch_count()
    for k in range(40):
        assert not oracle.whm_step_pl(st, GMcb, p["Gmass"], p["radius"], dt, lflat=lflat).any()
        assert ctx.whm_step_pl(GMcb, dt, LOOP_FLAT if lflat else LOOP_TRIANGULAR, True, lfirst=(k == 0)) == 0
    assert ctx.launch_count() - n0 <= 41          # one launch per step (+ the one-off eta / muj chain)
    got = ctx.body_get(PL)
    xj, vj = ctx.whm_get_jacobi()
    assert np.array_equal(got["r"], st["rh"]) and np.array_equal(got["v"], st["vh"])
    assert np.array_equal(xj, st["xj"]) and np.array_equal(vj, st["vj"])
    assert np.array_equal(got["a"], st["ah"])
    vb = ctx.body_get_vb(PL, vb=False, rbeg=True, rend=True)
    assert np.array_equal(vb["rbeg"], st["rbeg"]) and np.array_equal(vb["rend"], st["rend"])


def test_whm_fused_planet_step_with_masked_bodies_is_bit_identical_on_the_108_body_system(ctx, oracle):
    """The 108-body fixture through the one-launch step (npl <= 128) against the oracle with masked bodies; the multi-launch
    form (what larger systems use, SWCU_WHM_FUSED=0 forces it) is covered by test_whm_step_pl_resident_matches_oracle."""
    from swiftest_b200 import LOOP_TRIANGULAR
    f = W.fixture("108pl_50tp")
    order = np.argsort(-f["pl_Gmass"], kind="stable")
    GMcb, dt = float(f["cb_Gmass"]), float(f["dt"])
    Gm, rh, vh, rad, rhill = (f["pl_" + k][order] for k in ("Gmass", "rh", "vh", "radius", "rhill"))
    mask = np.ones(108, np.int32)
    mask[[5, 77]] = 0
    st = {"rh": rh.copy(), "vh": vh.copy(), "lfirst": True}
    ctx.body_sync(PL, 108, nplm=108, r=rh, v=vh, Gmass=Gm, radius=rad, rhill=rhill, mu=GMcb + Gm, lmask=mask, generation=8960)
    for k in range(10):
        assert not oracle.whm_step_pl(st, GMcb, Gm, rad, dt, lmask=mask).any()
        assert ctx.whm_step_pl(GMcb, dt, LOOP_TRIANGULAR, True, lfirst=(k == 0)) == 0
    got = ctx.body_get(PL)
    assert np.array_equal(got["r"], st["rh"]) and np.array_equal(got["v"], st["vh"])


@pytest.mark.parametrize("lflat", [False, True])
def test_helio_small_system_step_is_one_launch_and_bit_identical_to_the_reference_order(ctx, oracle, lflat):
    """npl <= 128: swcu_helio_step_pl is ONE kernel (drift_kernels.cu::helio_step_pl_small_kernel); 40 steps of the 108-body
    system with masked bodies reproduce the CPU restatement (bit-identical to the interpreted Fortran) BIT FOR BIT."""
    f, GMcb, dt = _fixture108()
    Gm, rad = f["pl_Gmass"], f["pl_radius"]
    mask = np.ones(108, np.int32)
    mask[[4, 90]] = 0
    rhill = np.linalg.norm(f["pl_rh"], axis=1) * (Gm / (3 * GMcb)) ** (1.0 / 3.0)
    ctx.body_sync(PL, 108, nplm=108, r=f["pl_rh"], v=f["pl_vh"], Gmass=Gm, radius=rad, rhill=rhill, mu=GMcb + Gm, lmask=mask,
                  generation=9300 + int(lflat))
    st = dict(rh=f["pl_rh"].copy(), vh=f["pl_vh"].copy(), vb=np.zeros((108, 3)), lfirst=True)
    n0 = ctx.launch_count()
    for k in range(40):
        assert not oracle.helio_step_pl(st, GMcb, Gm, rad, dt, lflat=lflat, lmask=mask).any()
        assert ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_FLAT if lflat else LOOP_TRIANGULAR, lclose=True, lfirst=(k == 0)) == 0
    assert ctx.launch_count() - n0 <= 43
    out = ctx.body_get(PL)
    hv = ctx.body_get_vb(PL, vb=True, rbeg=True, rend=True)
    assert np.array_equal(out["r"], st["rh"]) and np.array_equal(out["v"], st["vh"]) and np.array_equal(hv["vb"], st["vb"])
    assert np.array_equal(out["a"], st["ah"])
    assert np.array_equal(hv["rbeg"], st["rbeg"]) and np.array_equal(hv["rend"], st["rend"])
    ptb, pte = ctx.cb_get_pt()
    assert np.array_equal(ptb, st["ptbeg"]) and np.array_equal(pte, st["ptend"])


def test_reference_conservation_test_on_the_device(ctx):
    """The reference's only quantitative test (tests/test_swiftest.py:112-169: Sun + 8 planets, dt = 0.01 y, slopes of the
    energy and angular-momentum errors) over 1e5 of its 1e6 steps, planets resident, one launch per step, against the
    oracle's C stepper on the same steps; bench.py runs the full 1e6 (extra.conservation_reference_test)."""
    import bench
    r = bench.conservation_reference_test(ctx, 100000, nout=100, budget_s=120.0)
    assert r["steps_done_gpu"] == 100000 and r["steps_done_cpu"] == 100000 and r["encounters_seen"] == 0
    assert abs(r["E_slope_per_year_gpu"]) < 1e-8 and abs(r["L_slope_per_year_gpu"]) < 1e-10 and abs(r["GM_error_final"]) < 1e-14
    assert r["gpu_minus_cpu_E_error_max"] < 1e-12 and r["gpu_minus_cpu_L_error_max"] < 1e-12
    assert r["max_abs_true_energy_error_gpu"] < 1e-7
    assert r["kernel_launches_per_step_gpu"] < 1.01
