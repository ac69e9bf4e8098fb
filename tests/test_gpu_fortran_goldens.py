"""The CUDA path, through the C ABI, against REFERENCE-GENERATED vectors: the outputs of the reference's own Fortran
statements (tests/golden/fortran_*.npz, made by tests/golden/gen_golden_fortran.py with the interpreter
oracle/f90interp.py from /root/reference/src -- nothing of that is read here).

Bars (north star): accelerations within 1e-12 of the per-component sum of |terms| (summation order and the seeded
r^-3 differ from the Fortran's 1/(r2*sqrt(r2))); encounter pair lists bit-exact after canonical (index1, index2) order;
drift bit-exact where no libm call is involved (kepmd and series-guess kepu), 1e-12 elsewhere, iflag identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ACC_TOL = 1e-12
KICK_CASES = ["fx108", "disk160", "tiny5"]


@pytest.fixture(scope="module")
def gk():
    return np.load(os.path.join(GOLD, "fortran_kick.npz"))


@pytest.fixture(scope="module")
def gd():
    return np.load(os.path.join(GOLD, "fortran_drift.npz"))


@pytest.fixture(scope="module")
def ge():
    return np.load(os.path.join(GOLD, "fortran_encounter.npz"))


def _scaled(a, ref, scale):
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a - ref) / scale))


def _pair_scale(r, Gm, radius, nplm):
    """sum over the pairs the Fortran loops visit of |G m_j (r_j - r_i)| / r^3, per body and component (numpy, n <= 200)."""
    n = len(Gm)
    d = r[None, :, :] - r[:, None, :]
    r2 = (d ** 2).sum(2)
    np.fill_diagonal(r2, np.inf)
    ok = np.ones((n, n), bool) if radius is None else r2 > (radius[:, None] + radius[None, :]) ** 2
    visit = (np.arange(n)[:, None] < nplm) | (np.arange(n)[None, :] < nplm)
    w = np.where(ok & visit, Gm[None, :] / (r2 * np.sqrt(r2)), 0.0)
    return (w[:, :, None] * np.abs(d)).sum(1)


@pytest.mark.parametrize("case", KICK_CASES)
@pytest.mark.parametrize("lrad", [True, False])
def test_cuda_tri_kick_against_the_fortran(ctx, gk, case, lrad):
    r, Gm, radius, acc0 = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"], gk[case + "_acc0"]
    npl = len(Gm)
    for nplm in gk[case + "_nplm"]:
        ref = gk["%s_tri_%s_nplm%d" % (case, "rad" if lrad else "norad", nplm)]
        got = acc0.copy()
        ctx.kick_getacch_int_all_tri_pl(npl, int(nplm), r, Gm, radius if lrad else None, got)
        scale = _pair_scale(r, Gm, radius if lrad else None, int(nplm)) + np.abs(acc0)
        assert _scaled(got, ref, scale) < ACC_TOL, (case, nplm)


@pytest.mark.parametrize("case", KICK_CASES)
@pytest.mark.parametrize("lrad", [True, False])
def test_cuda_flat_kick_against_the_fortran(ctx, gk, case, lrad):
    """Canonical pairs (third-law kernel, no table) and the reference's explicit k_plpl table (pair-list kernel)."""
    r, Gm, radius, acc0 = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"], gk[case + "_acc0"]
    npl = len(Gm)
    k_plpl = np.ascontiguousarray(gk[case + "_k_plpl"], dtype=np.int32)
    for nplm in gk[case + "_nplm"]:
        nplm = int(nplm)
        nplplm = nplm * npl - nplm * (nplm + 1) // 2
        ref = gk["%s_flat_%s_nplm%d" % (case, "rad" if lrad else "norad", nplm)]
        scale = _pair_scale(r, Gm, radius if lrad else None, nplm) + np.abs(acc0)
        got = acc0.copy()
        ctx.kick_getacch_int_all_flat_pl(npl, nplplm, None, r, Gm, radius if lrad else None, got)
        assert _scaled(got, ref, scale) < ACC_TOL, (case, nplm, "canonical")
        if nplplm:
            got = acc0.copy()
            ctx.kick_getacch_int_all_flat_pl(npl, nplplm, k_plpl[:nplplm], r, Gm, radius if lrad else None, got)
            assert _scaled(got, ref, scale) < ACC_TOL, (case, nplm, "table")


@pytest.mark.parametrize("case", KICK_CASES)
def test_cuda_encounter_pair_subtract_against_the_fortran(ctx, gk, case):
    """symba_kick_getacch_pl: all pairs, then the encounter pairs again through the flat kernel, subtracted (F1)."""
    r, Gm, radius = gk[case + "_r"], gk[case + "_Gm"], gk[case + "_radius"]
    npl = len(Gm)
    pairs = np.ascontiguousarray(gk[case + "_enc_pairs"], dtype=np.int32)
    full = gk["%s_tri_rad_nplm%d" % (case, npl)] - gk[case + "_acc0"]
    ref = full - gk[case + "_enc_acc"]
    got = full.copy()
    ctx.symba_kick_subtract_encounters(npl, pairs[:, 0].copy(), pairs[:, 1].copy(), r, Gm, radius, got)
    scale = _pair_scale(r, Gm, radius, npl)
    assert _scaled(got, ref, scale) < ACC_TOL


def test_cuda_tp_kick_against_the_fortran(ctx, gk):
    rtp, rpl, Gm, mask, acc0 = gk["tp_rtp"], gk["tp_rpl"], gk["tp_GMpl"], gk["tp_lmask"].astype(np.int32), gk["tp_acc0"]
    got = acc0.copy()
    ctx.kick_getacch_int_all_tp(len(rtp), len(Gm), rtp, rpl, Gm, mask, got)
    dd = rtp[:, None, :] - rpl[None, :, :]
    scale = (Gm[None, :, None] * np.abs(dd) / (np.linalg.norm(dd, axis=2) ** 3)[:, :, None]).sum(1) + np.abs(acc0)
    assert _scaled(got, gk["tp_acc"], scale) < ACC_TOL
    assert np.array_equal(got[mask == 0], acc0[mask == 0])


@pytest.mark.parametrize("tag", ["a", "b", "gr", "long"])
def test_cuda_drift_against_the_fortran(ctx, oracle, gd, tag):
    g = lambda k: gd["drift_%s_%s" % (tag, k)]
    n = len(g("mu"))
    mask = g("lmask").astype(np.int32)
    x, v, fl = g("x0").copy(), g("v0").copy(), np.full(n, -7, np.int32)
    ctx.drift_all(g("mu").copy(), x, v, n, float(g("dt")), mask, fl, lgr=bool(g("lgr")), inv_c2=float(g("inv_c2")))
    assert np.array_equal(fl, g("iflag"))                       # masked bodies keep the caller's value, like the Fortran
    ok = (g("iflag") == 0) & g("lmask")
    br = oracle.drift_branch(g("mu"), g("x0"), g("v0"), float(g("dt")))     # which solver path (classification only)
    exact = ok & ((br == 0) | (br == 1)) & (not bool(g("lgr")))
    assert np.array_equal(x[exact], g("x1")[exact]) and np.array_equal(v[exact], g("v1")[exact])
    rs = np.linalg.norm(g("x1"), axis=1, keepdims=True)
    vs = np.linalg.norm(g("v1"), axis=1, keepdims=True)
    assert np.max((np.abs(x - g("x1")) / rs)[ok]) < 1e-12
    assert np.max((np.abs(v - g("v1")) / vs)[ok]) < 1e-12
    off = ~g("lmask")
    assert np.array_equal(x[off], g("x0")[off]) and np.array_equal(v[off], g("v0")[off])


def _canon_ref(ref):
    return ref[np.lexsort((ref[:, 1], ref[:, 0]))] if len(ref) else ref.reshape(0, 3)


def _same(got, ref):
    n, i1, i2, lv = got
    rc = _canon_ref(ref)
    assert n == len(rc)
    assert np.array_equal(i1, rc[:, 0]) and np.array_equal(i2, rc[:, 1])
    assert np.array_equal(np.asarray(lv).astype(bool), rc[:, 2].astype(bool))


@pytest.mark.parametrize("case", ["fx108", "disk300", "disk120"])
def test_cuda_plpl_lists_are_the_fortran_lists(ctx, ge, case):
    r, v, renc, dt = (ge["plpl_%s_%s" % (case, k)] for k in ("r", "v", "renc", "dt"))
    _same(ctx.encounter_check_all_sort_and_sweep_plpl(len(renc), r, v, renc, float(dt)), ge["plpl_%s_sas" % case])
    _same(ctx.encounter_check_all_triangular_plpl(len(renc), r, v, renc, float(dt)), ge["plpl_%s_tri" % case])


@pytest.mark.parametrize("case", ["fx", "disk"])
def test_cuda_pltp_lists_are_the_fortran_lists(ctx, ge, case):
    a = [ge["pltp_%s_%s" % (case, k)] for k in ("rpl", "vpl", "rtp", "vtp", "renc")]
    dt = float(ge["pltp_%s_dt" % case])
    npl, ntp = len(a[4]), len(a[2])
    _same(ctx.encounter_check_all_sort_and_sweep_pltp(npl, ntp, *a, dt), ge["pltp_%s_sas" % case])
    _same(ctx.encounter_check_all_triangular_pltp(npl, ntp, *a, dt), ge["pltp_%s_tri" % case])


@pytest.mark.parametrize("case", ["m60", "m200"])
def test_cuda_plplm_and_merged_lists_are_the_fortran_lists(ctx, ge, case):
    r, v, renc = ge["plplm_%s_r" % case], ge["plplm_%s_v" % case], ge["plplm_%s_renc" % case]
    nplm, dt = int(ge["plplm_%s_nplm" % case]), float(ge["plplm_%s_dt" % case])
    nplt = len(renc) - nplm
    a = (r[:nplm], v[:nplm], r[nplm:], v[nplm:], renc[:nplm], renc[nplm:], dt)
    _same(ctx.encounter_check_all_sort_and_sweep_plplm(nplm, nplt, *a), ge["plplm_%s_sas" % case])
    _same(ctx.encounter_check_all_plplm(nplm, nplt, *a), ge["plplm_%s_merged" % case])


# ------------------------------------------------------------------------------------------------------- whole steps
@pytest.fixture(scope="module")
def gs():
    return np.load(os.path.join(GOLD, "fortran_steps.npz"))


def _sync_system(ctx, gs, kind, tag, gen):
    from swiftest_b200 import PL, TP
    key = "%s_%s_" % (kind, tag)
    g = lambda k: gs[key + k]
    GMcb = float(g("GMcb"))
    Gm, radius = g("pl_Gmass"), g("pl_radius")
    n, ntp = len(Gm), len(g("tp_rh0"))
    rhill = np.linalg.norm(g("pl_rh0"), axis=1) * (Gm / (3 * GMcb)) ** (1.0 / 3.0)
    ctx.body_sync(PL, n, nplm=n, r=g("pl_rh0"), v=g("pl_vh0"), Gmass=Gm, radius=radius, rhill=rhill, mu=GMcb + Gm,
                  lmask=g("lmask_pl").astype(np.int32), generation=gen)
    ctx.body_sync(TP, ntp, r=g("tp_rh0"), v=g("tp_vh0"), mu=np.full(ntp, GMcb), lmask=g("lmask_tp").astype(np.int32),
                  generation=gen + 1)
    return g, GMcb, float(g("dt")), int(g("nsteps")), bool(g("lflat"))


def _close(a, b, tol):
    s = np.linalg.norm(b, axis=-1, keepdims=True)
    return float(np.max(np.abs(a - b) / np.where(s > 0, s, 1.0))) < tol


@pytest.mark.parametrize("tag", ["p8", "p8flat", "p8mask", "p30"])
def test_cuda_helio_steps_against_the_fortran(ctx, gs, tag):
    """swcu_helio_step_pl + swcu_helio_step_tp (device resident, nothing crosses PCIe between steps) against the states the
    reference's helio_step_pl / helio_step_tp produce, step after step."""
    from swiftest_b200 import PL, TP, LOOP_TRIANGULAR, LOOP_FLAT
    g, GMcb, dt, nsteps, lflat = _sync_system(ctx, gs, "helio", tag, 7100 + 10 * ["p8", "p8flat", "p8mask", "p30"].index(tag))
    on_pl, on_tp = g("lmask_pl"), g("lmask_tp")
    for s in range(nsteps):
        assert ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_FLAT if lflat else LOOP_TRIANGULAR, lclose=True, lfirst=(s == 0)) == 0
        assert ctx.helio_step_tp(GMcb, dt, lfirst=(s == 0)) == 0
        pl, tp = ctx.body_get(PL), ctx.body_get(TP)
        assert _close(pl["r"], g("pl_rh")[s], 1e-12) and _close(pl["v"], g("pl_vh")[s], 1e-12), (tag, s)
        assert _close(tp["r"][on_tp], g("tp_rh")[s][on_tp], 1e-12) and _close(tp["v"][on_tp], g("tp_vh")[s][on_tp], 1e-12), (tag, s)
        assert np.array_equal(tp["r"][~on_tp], g("tp_rh0")[~on_tp])          # masked particles never move
        assert _close(ctx.body_get_vb(PL)["vb"][on_pl], g("pl_vb")[s][on_pl], 1e-12)


@pytest.mark.parametrize("tag", ["p8", "p8flat", "p8mask", "p30"])
def test_cuda_whm_steps_against_the_fortran(ctx, gs, tag):
    """swcu_whm_step_pl + swcu_whm_tp_step(ah0 from the device) against the reference's whm_step_pl / whm_step_tp."""
    from swiftest_b200 import PL, TP, LOOP_TRIANGULAR, LOOP_FLAT
    g, GMcb, dt, nsteps, lflat = _sync_system(ctx, gs, "whm", tag, 7200 + 10 * ["p8", "p8flat", "p8mask", "p30"].index(tag))
    on_tp = g("lmask_tp")
    ctx.whm_tp_first_accel()
    for s in range(nsteps):
        assert ctx.whm_step_pl(GMcb, dt, LOOP_FLAT if lflat else LOOP_TRIANGULAR, True, lfirst=(s == 0)) == 0
        assert ctx.whm_tp_step(dt, None) == 0
        pl, tp = ctx.body_get(PL), ctx.body_get(TP)
        assert _close(pl["r"], g("pl_rh")[s], 1e-12) and _close(pl["v"], g("pl_vh")[s], 1e-12), (tag, s)
        assert _close(tp["r"][on_tp], g("tp_rh")[s][on_tp], 1e-12) and _close(tp["v"][on_tp], g("tp_vh")[s][on_tp], 1e-12), (tag, s)
    xj, vj = ctx.whm_get_jacobi()
    assert _close(xj, g("xj"), 1e-12) and _close(vj, g("vj"), 1e-12)
