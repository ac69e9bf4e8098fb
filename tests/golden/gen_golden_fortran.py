"""REFERENCE-GENERATED golden vectors for gravity, sort-and-sweep and drift, produced by EXECUTING THE REFERENCE'S FORTRAN
SOURCE (read from /root/reference/src when this script runs) with the Fortran-subset interpreter oracle/f90interp.py.

  python tests/golden/gen_golden_fortran.py          # ~ a few minutes; writes tests/golden/fortran_*.npz

Run in the build container only: /root/reference does not exist on the GPU box and nothing reads it at test time; the
committed .npz files are what the tests load.  No Fortran compiler exists here or on the GPU box
(profiles/r02_fortran_probe.txt), so the reference cannot be compiled into oracle/_ref; instead its own statements are
interpreted one IEEE operation at a time (serial loop order, no FMA contraction), which is exactly what the C restatement
oracle/swiftest_oracle.c claims to reproduce.  tests/test_oracle_fortran_goldens.py holds the restatement to these vectors
BIT FOR BIT (accelerations, drifted states, iflag, pair lists in the reference's own output order), and the GPU tests
compare the CUDA path with the same vectors at the tolerance of the north star.

Routines executed (file:line in /root/reference/src):
  swiftest/swiftest_kick.f90:69-470     flat_rad/flat_norad/tri_rad/tri_norad_pl (all nplm branches), all_tp, one_pl, one_tp
  swiftest/swiftest_util.f90:1031-1131  flatten_eucl_plpl (the k_plpl table the flat kernels index)
  swiftest/swiftest_drift.f90:60-580    drift_all (with and without GR), drift_one, dan, kepmd, kepu family, stumpff
  swiftest/swiftest_orbel.f90:147-172   orbel_scget
  encounter/encounter_check.f90:14-990  all_plpl/_plplm/_pltp dispatch, sort_and_sweep_*, triangular_*, sweep_one, check_one,
                                        collapse_ragged_list, remove_duplicates, sort_aabb_1D, sweep_aabb_single/double_list
  encounter/encounter_util.f90:332-365  setup_aabb
  base/base_module.f90:1346-2060        util_sort (index quicksorts), util_sort_rearrange
  swiftest/swiftest_util.f90:1494-1526  index_array
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import f90interp as F  # noqa: E402

SRC = "/root/reference/src"
FILES = ["globals/globals_module.f90", "base/base_module.f90", "swiftest/swiftest_module.f90",
         "swiftest/swiftest_kick.f90", "swiftest/swiftest_drift.f90", "swiftest/swiftest_orbel.f90",
         "swiftest/swiftest_util.f90", "encounter/encounter_module.f90", "encounter/encounter_check.f90",
         "encounter/encounter_util.f90", "collision/collision_module.f90",
         "helio/helio_module.f90", "helio/helio_step.f90", "helio/helio_kick.f90", "helio/helio_drift.f90",
         "helio/helio_util.f90", "whm/whm_module.f90", "whm/whm_step.f90", "whm/whm_kick.f90", "whm/whm_drift.f90",
         "whm/whm_coord.f90", "whm/whm_util.f90",
         "operator/operator_module.f90", "operator/operator_cross.f90", "symba/symba_module.f90", "symba/symba_kick.f90",
         "symba/symba_encounter_check.f90", "symba/symba_util.f90", "collision/collision_check.f90",
         "swiftest/swiftest_discard.f90"]


def world():
    return F.load_world(SRC, FILES)


def fa(a):
    """(n,3) C-order (the repo's convention) -> Fortran r(3,n)."""
    return np.array(np.asarray(a, dtype=np.float64).T, order="F", copy=True)


def back(a):
    return np.array(a.T, order="C", copy=True)


def disk(n, seed, a0=1.0, a1=1.3, mscale=1e-7, hot=0.02):
    """A thin planetesimal disk around a unit-GM star: dense enough for many encounters at a few Hill radii."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(a0, a1, n)
    th = rng.uniform(0, 2 * np.pi, n)
    z = rng.normal(0, 0.002, n) * a
    r = np.stack([a * np.cos(th), a * np.sin(th), z], 1)
    vc = 1.0 / np.sqrt(a)
    v = np.stack([-vc * np.sin(th), vc * np.cos(th), np.zeros(n)], 1) * (1 + rng.normal(0, hot, (n, 1)))
    v += rng.normal(0, hot * 0.3, (n, 3))
    Gm = mscale * rng.lognormal(0, 1.0, n)
    rhill = a * (Gm / 3.0) ** (1.0 / 3.0)
    radius = 0.02 * rhill
    return r, v, Gm, rhill, radius


# ---------------------------------------------------------------------------------------------------------- gravity
def gen_kick(w, out):
    fx = np.load(os.path.join(HERE, "fixture_108pl_50tp.npz"))
    cases = {}
    # (1) the 108 massive bodies of the reference's own example system (heliocentric; the Sun is the central body)
    cases["fx108"] = dict(r=fx["pl_rh"][:], Gm=fx["pl_Gmass"][:], radius=fx["pl_radius"][:])
    # (2) a disk with physical overlaps (radius check rejects pairs) and a few coincident-distance cases
    r, v, Gm, rhill, radius = disk(160, 11)
    radius = radius * 400.0          # 63 of the 12720 pairs overlap physically
    cases["disk160"] = dict(r=r, Gm=Gm, radius=radius)
    # (3) small hand-made system: touching pair exactly at the limit (rji2 == rlim2 is rejected by '>')
    r3 = np.array([[1.0, 0, 0], [1.5, 0, 0], [0, 2.0, 0], [0, 0, -3.0], [1.0, 1.0, 1.0]])
    rad3 = np.array([0.25, 0.25, 0.1, 0.1, 0.05])
    cases["tiny5"] = dict(r=r3, Gm=np.array([1e-3, 2e-3, 3e-4, 0.0, 5e-5]), radius=rad3)
    pl = w.new_object("swiftest_pl")
    param = w.new_object("swiftest_parameters")
    param.c["lflatten_interactions"] = True
    for name, c in cases.items():
        r, Gm, radius = fa(c["r"]), np.array(c["Gm"], dtype=np.float64), np.array(c["radius"], dtype=np.float64)
        npl = len(Gm)
        rng = np.random.default_rng(5)
        acc0 = np.asfortranarray(rng.normal(0, 1e-6, (3, npl)))      # accumulates onto a non-zero acc (intent inout)
        out[name + "_r"], out[name + "_Gm"], out[name + "_radius"], out[name + "_acc0"] = c["r"], Gm, radius, back(acc0)
        # k_plpl by the reference's flatten
        pl.c["nbody"] = npl
        w.call("swiftest_util_flatten_eucl_plpl", pl, param)
        k_plpl = pl.c["k_plpl"]
        nplpl = int(pl.c["nplpl"])
        assert k_plpl.shape == (2, nplpl)
        out[name + "_k_plpl"] = np.ascontiguousarray(k_plpl.T)
        nplms = sorted({npl, (2 * npl) // 3, npl // 3, 1})           # full, nplt<=nplm, lmtiny, single massive
        out[name + "_nplm"] = np.array(nplms)
        for nplm in nplms:
            for rad in (True, False):
                acc = acc0.copy(order="F")
                if rad:
                    w.call("swiftest_kick_getacch_int_all_tri_rad_pl", npl, nplm, r, Gm, radius, acc)
                else:
                    w.call("swiftest_kick_getacch_int_all_tri_norad_pl", npl, nplm, r, Gm, acc)
                out["%s_tri_%s_nplm%d" % (name, "rad" if rad else "norad", nplm)] = back(acc)
                # flat: nplplm = nplm*npl - nplm*(nplm+1)/2 pairs (symba_kick.f90:23-29), the head of the k_plpl table
                nplplm = nplm * npl - nplm * (nplm + 1) // 2
                acc = acc0.copy(order="F")
                if rad:
                    w.call("swiftest_kick_getacch_int_all_flat_rad_pl", npl, nplplm, k_plpl, r, Gm, radius, acc)
                else:
                    w.call("swiftest_kick_getacch_int_all_flat_norad_pl", npl, nplplm, k_plpl, r, Gm, acc)
                out["%s_flat_%s_nplm%d" % (name, "rad" if rad else "norad", nplm)] = back(acc)
        # flat kernel over an explicit (encounter) pair list, as symba_kick_getacch_pl uses it (symba_kick.f90:59-70)
        rng = np.random.default_rng(9)
        npair = min(40, nplpl)
        sel = np.sort(rng.choice(nplpl, npair, replace=False))
        k_enc = np.asfortranarray(k_plpl[:, sel])
        acc = np.zeros((3, npl), order="F")
        w.call("swiftest_kick_getacch_int_all_flat_rad_pl", npl, npair, k_enc, r, Gm, radius, acc)
        out[name + "_enc_pairs"] = np.ascontiguousarray(k_enc.T)
        out[name + "_enc_acc"] = back(acc)
    # pl -> tp
    rpl = fa(fx["pl_rh"][:8]); GMpl = np.array(fx["pl_Gmass"][:8])
    rtp = fa(fx["tp_rh"])
    ntp = rtp.shape[1]
    lmask = np.ones(ntp, dtype=bool); lmask[::7] = False
    acc0 = np.asfortranarray(np.random.default_rng(3).normal(0, 1e-6, (3, ntp)))
    acc = acc0.copy(order="F")
    w.call("swiftest_kick_getacch_int_all_tp", ntp, 8, rtp, rpl, GMpl, lmask, acc)
    out["tp_rtp"], out["tp_rpl"], out["tp_GMpl"] = back(rtp), back(rpl), GMpl
    out["tp_lmask"], out["tp_acc0"], out["tp_acc"] = lmask, back(acc0), back(acc)


# ---------------------------------------------------------------------------------------------------------- drift
def drift_sample(seed, n):
    """States covering every branch: kepmd (small dM, e), universal-variable Newton, Laguerre fallback, hyperbolic with the
    cubic guess, near-parabolic, a step that fails (iflag /= 0) and masked bodies."""
    rng = np.random.default_rng(seed)
    mu = np.full(n, 4 * np.pi ** 2) * rng.uniform(0.5, 1.5, n)
    x = np.zeros((n, 3)); v = np.zeros((n, 3))
    for i in range(n):
        kind = i % 8
        rmag = 10 ** rng.uniform(-1.0, 1.5)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        t = np.cross(d, rng.normal(size=3)); t /= np.linalg.norm(t)
        vc = np.sqrt(mu[i] / rmag)
        if kind in (0, 1, 2):                 # moderate ellipse
            f, rad = rng.uniform(0.7, 1.15), rng.uniform(-0.2, 0.2)
        elif kind == 3:                       # very eccentric ellipse
            f, rad = rng.uniform(0.05, 0.4), rng.uniform(-0.5, 0.5)
        elif kind == 4:                       # near parabolic
            f, rad = np.sqrt(2.0) * (1 + rng.uniform(-1e-6, 1e-6)), rng.uniform(-0.1, 0.1)
        elif kind == 5:                       # hyperbolic
            f, rad = rng.uniform(1.5, 4.0), rng.uniform(-1.0, 1.0)
        elif kind == 6:                       # plunging, nearly radial
            f, rad = rng.uniform(0.0, 0.05), -rng.uniform(0.5, 1.3)
        else:                                 # strongly hyperbolic, deep
            f, rad = rng.uniform(5.0, 30.0), rng.uniform(-3.0, 3.0)
        x[i] = rmag * d
        v[i] = vc * (f * t + rad * d)
    return mu, x, v


def gen_drift(w, out):
    param = w.new_object("swiftest_parameters")
    for tag, (seed, n, dt, lgr) in {"a": (21, 240, 0.05, False), "b": (22, 160, 2.5, False),
                                    "gr": (23, 120, 0.3, True), "long": (24, 80, 400.0, False)}.items():
        mu, x, v = drift_sample(seed, n)
        lmask = np.ones(n, dtype=bool); lmask[5::11] = False
        param.c["lgr"] = lgr
        param.c["inv_c2"] = 1.0 / 63241.077 ** 2 if lgr else 0.0      # (AU/yr)^-2
        xf, vf = fa(x), fa(v)
        iflag = np.full(n, -7, dtype=np.int32)      # intent(out), only written where lmask: the rest must stay untouched
        w.call("swiftest_drift_all", mu, xf, vf, n, param, float(dt), lmask, iflag)
        out["drift_%s_mu" % tag], out["drift_%s_x0" % tag], out["drift_%s_v0" % tag] = mu, x, v
        out["drift_%s_dt" % tag], out["drift_%s_lgr" % tag] = float(dt), lgr
        out["drift_%s_inv_c2" % tag] = param.c["inv_c2"]
        out["drift_%s_lmask" % tag] = lmask
        out["drift_%s_x1" % tag], out["drift_%s_v1" % tag], out["drift_%s_iflag" % tag] = back(xf), back(vf), iflag
        print("  drift", tag, "iflag counts", dict(zip(*np.unique(iflag, return_counts=True))), flush=True)


# ---------------------------------------------------------------------------------------------------------- encounters
def reset_encounter_state(w):
    for name in ("encounter_check_all_sort_and_sweep_plpl", "encounter_check_all_sort_and_sweep_pltp",
                 "encounter_check_all_sort_and_sweep_plplm", "encounter_check_sweep_aabb_single_list",
                 "encounter_check_sweep_aabb_double_list", "encounter_check_all_triangular_plpl",
                 "encounter_check_all_triangular_pltp", "encounter_check_all_triangular_plplm"):
        w.procs[name].static.clear()


def lists(res, k):
    nenc, i1, i2, lv = res[k], res[k + 1], res[k + 2], res[k + 3]
    nenc = int(nenc)
    if nenc == 0 or i1 is None:
        z = np.zeros(0, dtype=np.int32)
        return z, z.copy(), np.zeros(0, dtype=bool)
    return np.array(i1[:nenc], dtype=np.int32), np.array(i2[:nenc], dtype=np.int32), np.array(lv[:nenc], dtype=bool)


def gen_encounter(w, out):
    fx = np.load(os.path.join(HERE, "fixture_108pl_50tp.npz"))
    param_sas = w.new_object("base_parameters")
    param_sas.c["lencounter_sas_plpl"] = True
    param_sas.c["lencounter_sas_pltp"] = True
    param_tri = w.new_object("base_parameters")
    param_tri.c["lencounter_sas_plpl"] = False
    param_tri.c["lencounter_sas_pltp"] = False

    # ---- check_one on a spread of relative states (every branch: r2 <= r2crit, receding, v2 tiny, tmin < dt, tmin >= dt)
    rng = np.random.default_rng(31)
    n1 = 400
    rel = rng.normal(0, 1.0, (n1, 3)) * 10 ** rng.uniform(-2, 0.5, (n1, 1))
    vel = rng.normal(0, 1.0, (n1, 3)) * 10 ** rng.uniform(-3, 1, (n1, 1))
    vel[::37] = 0.0
    vel[5::41] *= 1e-160
    renc = 10 ** rng.uniform(-2, 0.3, n1)
    dt = 0.3
    lenc = np.zeros(n1, dtype=bool); lvd = np.zeros(n1, dtype=bool)
    for k in range(n1):
        a, b = F.Cell(False), F.Cell(False)
        w.call("encounter_check_one", *map(float, rel[k]), *map(float, vel[k]), float(renc[k]), dt, a, b)
        lenc[k], lvd[k] = a.v, b.v
    out["one_rel"], out["one_vel"], out["one_renc"], out["one_dt"] = rel, vel, renc, dt
    out["one_lencounter"], out["one_lvdotr"] = lenc, lvd
    print("  check_one: %d encounters, %d approaching of %d" % (lenc.sum(), lvd.sum(), n1), flush=True)

    # ---- pl-pl: the 108-body fixture, then a dense disk, then a smaller population (the saved bounding box shrinks)
    plpl_cases = []
    r, v, rh = fx["pl_rh"][:], fx["pl_vh"][:], fx["pl_rhill"][:]
    plpl_cases.append(("fx108", r, v, 3.0 * rh, 0.05))
    r, v, Gm, rhill, radius = disk(300, 41)
    plpl_cases.append(("disk300", r, v, 4.0 * rhill, 0.02))
    r, v, Gm, rhill, radius = disk(120, 42, hot=0.05)
    plpl_cases.append(("disk120", r, v, 6.0 * rhill, 0.05))
    reset_encounter_state(w)
    for name, r, v, renc, dt in plpl_cases:       # consecutive calls share the reference's SAVEd bounding box, as in a run
        t0 = time.time()
        npl = len(renc)
        res = w.call("encounter_check_all_plpl", param_sas, npl, fa(r), fa(v), np.array(renc), float(dt), 0, None, None, None)
        i1, i2, lv = lists(res, 6)
        res = w.call("encounter_check_all_plpl", param_tri, npl, fa(r), fa(v), np.array(renc), float(dt), 0, None, None, None)
        t1, t2, tl = lists(res, 6)
        out["plpl_%s_r" % name], out["plpl_%s_v" % name], out["plpl_%s_renc" % name] = r, v, np.array(renc)
        out["plpl_%s_dt" % name] = float(dt)
        out["plpl_%s_sas" % name] = np.stack([i1, i2, lv.astype(np.int32)], 1)
        out["plpl_%s_tri" % name] = np.stack([t1, t2, tl.astype(np.int32)], 1)
        print("  plpl %-8s npl=%d  sweep nenc=%d  triangular nenc=%d  (%.1fs)" % (name, npl, len(i1), len(t1), time.time() - t0),
              flush=True)

    # ---- pl-tp: fixture planets and particles; disk planets against a particle swarm
    pltp_cases = []
    pltp_cases.append(("fx", fx["pl_rh"][:8], fx["pl_vh"][:8], fx["tp_rh"], fx["tp_vh"], 8.0 * fx["pl_rhill"][:8], 0.5))
    r, v, Gm, rhill, radius = disk(40, 51, mscale=3e-6)
    rt, vt, _, _, _ = disk(500, 52)
    pltp_cases.append(("disk", r, v, rt, vt, 3.5 * rhill, 0.02))
    reset_encounter_state(w)
    for name, rpl, vpl, rtp, vtp, renc, dt in pltp_cases:
        t0 = time.time()
        npl, ntp = len(renc), len(rtp)
        res = w.call("encounter_check_all_pltp", param_sas, npl, ntp, fa(rpl), fa(vpl), fa(rtp), fa(vtp), np.array(renc),
                     float(dt), 0, None, None, None)
        i1, i2, lv = lists(res, 9)
        res = w.call("encounter_check_all_pltp", param_tri, npl, ntp, fa(rpl), fa(vpl), fa(rtp), fa(vtp), np.array(renc),
                     float(dt), 0, None, None, None)
        t1, t2, tl = lists(res, 9)
        for key, val in dict(rpl=rpl, vpl=vpl, rtp=rtp, vtp=vtp, renc=np.array(renc), dt=float(dt)).items():
            out["pltp_%s_%s" % (name, key)] = val
        out["pltp_%s_sas" % name] = np.stack([i1, i2, lv.astype(np.int32)], 1)
        out["pltp_%s_tri" % name] = np.stack([t1, t2, tl.astype(np.int32)], 1)
        print("  pltp %-8s npl=%d ntp=%d  sweep nenc=%d  triangular nenc=%d  (%.1fs)" %
              (name, npl, ntp, len(i1), len(t1), time.time() - t0), flush=True)

    # ---- plm-plt and the merged list (SyMBA with GMTINY): bodies ordered massive first
    r, v, Gm, rhill, radius = disk(260, 61)
    order = np.argsort(-Gm, kind="stable")
    r, v, rhill = r[order], v[order], rhill[order]
    for name, nplm in (("m60", 60), ("m200", 200)):
        t0 = time.time()
        nplt = len(rhill) - nplm
        renc = 4.0 * rhill
        args = (nplm, nplt, fa(r[:nplm]), fa(v[:nplm]), fa(r[nplm:]), fa(v[nplm:]), np.array(renc[:nplm]),
                np.array(renc[nplm:]), 0.03)
        reset_encounter_state(w)
        res = w.call("encounter_check_all_sort_and_sweep_plplm", *args, 0, None, None, None)
        i1, i2, lv = lists(res, 9)
        reset_encounter_state(w)
        res = w.call("encounter_check_all_plplm", param_sas, *args, 0, None, None, None)
        m1, m2, ml = lists(res, 10)
        res = w.call("encounter_check_all_plplm", param_tri, *args, 0, None, None, None)
        t1, t2, tl = lists(res, 10)
        out["plplm_%s_r" % name], out["plplm_%s_v" % name], out["plplm_%s_renc" % name] = r, v, renc
        out["plplm_%s_nplm" % name], out["plplm_%s_dt" % name] = nplm, 0.03
        out["plplm_%s_sas" % name] = np.stack([i1, i2, lv.astype(np.int32)], 1)
        out["plplm_%s_merged" % name] = np.stack([m1, m2, ml.astype(np.int32)], 1)
        out["plplm_%s_merged_tri" % name] = np.stack([t1, t2, tl.astype(np.int32)], 1)
        print("  plplm %-5s nplm=%d nplt=%d  plm-plt nenc=%d  merged nenc=%d  merged triangular nenc=%d  (%.1fs)" %
              (name, nplm, nplt, len(i1), len(m1), len(t1), time.time() - t0), flush=True)


# ---------------------------------------------------------------------------------------------------------- whole steps
def setc(obj, **kw):
    for k, v in kw.items():
        obj.c[k.lower()] = v


def make_system(w, kind, fx, npl, lflat, masked):
    """A helio or whm nbody_system object as the reference's setup leaves it before the first step."""
    GMcb = float(fx["cb_Gmass"])
    ntp = len(fx["tp_rh"])
    z = lambda m: np.zeros((3, m), order="F")
    pl, tp = w.new_object(kind + "_pl"), w.new_object(kind + "_tp")
    cb = w.new_object("helio_cb" if kind == "helio" else "swiftest_cb")
    system, param = w.new_object(kind + "_nbody_system"), w.new_object("swiftest_parameters")
    lm_pl, lm_tp = np.ones(npl, dtype=bool), np.ones(ntp, dtype=bool)
    if masked:
        lm_pl[npl // 2] = False
        lm_tp[::9] = False
    setc(pl, nbody=npl, rh=fa(fx["pl_rh"][:npl]), vh=fa(fx["pl_vh"][:npl]), vb=z(npl), ah=z(npl),
         Gmass=np.array(fx["pl_Gmass"][:npl]), radius=np.array(fx["pl_radius"][:npl]), mu=GMcb + np.array(fx["pl_Gmass"][:npl]),
         lmask=lm_pl, lfirst=True, status=np.zeros(npl, dtype=np.int32))
    setc(tp, nbody=ntp, rh=fa(fx["tp_rh"]), vh=fa(fx["tp_vh"]), vb=z(ntp), ah=z(ntp), mu=np.full(ntp, GMcb), lmask=lm_tp,
         lfirst=True, status=np.zeros(ntp, dtype=np.int32))
    setc(cb, Gmass=GMcb)
    setc(system, cb=cb, pl=pl, tp=tp)
    setc(param, lflatten_interactions=lflat, lclose=True, lgr=False, loblatecb=False, lextra_force=False)
    if lflat:
        w.call("swiftest_util_flatten_eucl_plpl", pl, param)
    if kind == "whm":
        setc(pl, xj=z(npl), vj=z(npl), eta=np.zeros(npl), muj=np.zeros(npl), ir3j=np.zeros(npl), ir3h=np.zeros(npl))
        setc(tp, ir3h=np.zeros(ntp))
        w.call("whm_util_set_mu_eta_pl", pl, cb)
    return pl, tp, cb, system, param


def gen_steps(w, out):
    """helio_step_pl/_tp (helio/helio_step.f90:37-123) and whm_step_pl/_tp (whm/whm_step.f90:37-100) with everything they
    call (coordinate changes, linear drift, kicks, accelerations, Jacobi chains, Kepler drift), several consecutive steps."""
    fx = np.load(os.path.join(HERE, "fixture_108pl_50tp.npz"))
    nsteps, dt = 5, 0.02
    for kind in ("helio", "whm"):
        for tag, npl, lflat, masked in (("p8", 8, False, False), ("p8flat", 8, True, False), ("p8mask", 8, False, True),
                                        ("p30", 30, False, False)):
            t0 = time.time()
            pl, tp, cb, system, param = make_system(w, kind, fx, npl, lflat, masked)
            key = "%s_%s_" % (kind, tag)
            out[key + "npl"], out[key + "lflat"], out[key + "dt"], out[key + "nsteps"] = npl, lflat, dt, nsteps
            out[key + "GMcb"] = float(fx["cb_Gmass"])
            out[key + "lmask_pl"], out[key + "lmask_tp"] = pl.c["lmask"].copy(), tp.c["lmask"].copy()
            for nm in ("rh", "vh"):
                out[key + "pl_%s0" % nm], out[key + "tp_%s0" % nm] = back(pl.c[nm]), back(tp.c[nm])
            out[key + "pl_Gmass"], out[key + "pl_radius"] = pl.c["gmass"].copy(), pl.c["radius"].copy()
            hist = {k: [] for k in ("pl_rh", "pl_vh", "tp_rh", "tp_vh", "pl_vb", "tp_vb")}
            for s in range(nsteps):
                w.call(kind + "_step_pl", pl, system, param, s * dt, dt)
                w.call(kind + "_step_tp", tp, system, param, s * dt, dt)
                for o, nm in ((pl, "pl"), (tp, "tp")):
                    hist[nm + "_rh"].append(back(o.c["rh"])); hist[nm + "_vh"].append(back(o.c["vh"]))
                    hist[nm + "_vb"].append(back(o.c["vb"]))
            for k2, v in hist.items():
                out[key + k2] = np.array(v)
            if kind == "whm":
                out[key + "eta"], out[key + "muj"] = pl.c["eta"].copy(), pl.c["muj"].copy()
                out[key + "xj"], out[key + "vj"] = back(pl.c["xj"]), back(pl.c["vj"])
            print("  %s %-7s npl=%d ntp=%d  %d steps (%.1fs)" % (kind, tag, npl, tp.c["nbody"], nsteps, time.time() - t0), flush=True)


# ---------------------------------------------------------------------------------------------------------- energy, lists
def gen_lists(w, out):
    """swiftest_util_get_energy_and_momentum_system (swiftest_util.f90:1172-1394), symba_kick_list_plpl/_pltp
    (symba/symba_kick.f90:126-337), symba_encounter_check_list_plpl/_pltp (symba/symba_encounter_check.f90:88-235) with
    symba_util_set_renc, collision_check_one + swiftest_orbel_xv2aeq (collision/collision_check.f90:16-57,
    swiftest_orbel.f90:700-764), swiftest_discard_pl_close (swiftest_discard.f90:295-337)."""
    ACTIVE, INACTIVE = int(w.const("active")), int(w.const("inactive"))
    n, ntp = 70, 90
    r, v, Gm, rhill, radius = disk(n, 71, mscale=3e-6)
    rt, vt, _, _, _ = disk(ntp, 72)
    rng = np.random.default_rng(73)

    # ---- energy and momentum: barycentric state, one inactive body, flat and triangular potential loops, lclose on/off
    GMcb, mcb = 39.476926408897626, 1.0
    status = np.full(n, ACTIVE, dtype=np.int32); status[7] = INACTIVE
    mass = Gm / GMcb
    rbcb, vbcb = np.array([1e-5, -2e-5, 3e-7]), np.array([2e-6, 1e-6, -4e-8])
    out["en_GMcb"], out["en_mass_cb"], out["en_radius_cb"], out["en_rbcb"], out["en_vbcb"] = GMcb, mcb, 0.00465, rbcb, vbcb
    out["en_Gmass"], out["en_mass"], out["en_radius"], out["en_rb"], out["en_vb"], out["en_status"] = Gm, mass, radius, r, v, status
    for tag, lflat, lclose in (("tri", False, True), ("flat", True, True), ("tri_noclose", False, False)):
        pl, cb = w.new_object("symba_pl"), w.new_object("symba_cb")
        system, param = w.new_object("symba_nbody_system"), w.new_object("swiftest_parameters")
        setc(pl, nbody=n, rb=fa(r), vb=fa(v), Gmass=Gm.copy(), mass=mass.copy(), radius=radius.copy(), lmask=np.ones(n, dtype=bool),
             status=status.copy(), ip=np.zeros((3, n), order="F"), rot=np.zeros((3, n), order="F"))
        setc(cb, Gmass=GMcb, mass=mcb, radius=0.00465, rb=rbcb.copy(), vb=vbcb.copy(), ip=np.zeros(3), rot=np.zeros(3))
        setc(system, pl=pl, cb=cb)
        setc(param, lrotation=False, lflatten_interactions=lflat, loblatecb=False, lclose=lclose)
        if lflat:
            w.call("swiftest_util_flatten_eucl_plpl", pl, param)
        w.call("swiftest_util_get_energy_and_momentum_system", system, param)
        for k in ("ke_orbit", "pe", "be", "te", "gmtot"):
            out["en_%s_%s" % (tag, k)] = float(system.c[k])
        out["en_%s_l_orbit" % tag] = np.array(system.c["l_orbit"])

    # ---- SyMBA recursion kicks over an encounter list: close pairs so that all three regimes (inside the inner shell,
    #      in the shell, outside) occur; mixed levels, some inactive pairs, both signs, several recursion levels
    rhill_big = np.full(n, 0.012) * rng.uniform(0.7, 1.3, n)
    d = np.linalg.norm(r[:, None] - r[None], axis=2)
    iu = np.triu_indices(n, 1)
    order = np.argsort(d[iu])
    pick = np.concatenate([order[:110], rng.choice(order[110:], 40, replace=False)])
    rng.shuffle(pick)
    i1, i2 = (iu[0][pick] + 1).astype(np.int32), (iu[1][pick] + 1).astype(np.int32)
    nenc = len(i1)
    lstatus = np.full(nenc, ACTIVE, dtype=np.int32); lstatus[::13] = INACTIVE
    levelg = rng.integers(0, 3, n).astype(np.int32)
    out["kl_rh"], out["kl_rhill"], out["kl_Gmass"], out["kl_levelg"] = r, rhill_big, Gm, levelg
    out["kl_index1"], out["kl_index2"], out["kl_lactive"] = i1, i2, lstatus == ACTIVE
    vb0 = v.copy()
    ah0 = rng.normal(0, 1e-3, (n, 3))
    out["kl_vb0"], out["kl_ah0"] = vb0, ah0
    dtk = 0.004
    out["kl_dt"] = dtk
    for irec, sgn in ((1, 1), (1, -1), (2, 1), (2, -1), (3, 1)):
        pl, system, lst = w.new_object("symba_pl"), w.new_object("symba_nbody_system"), w.new_object("symba_list_plpl")
        setc(pl, nbody=n, rh=fa(r), vb=fa(vb0), ah=fa(ah0), Gmass=Gm.copy(), rhill=rhill_big.copy(), lmask=np.ones(n, dtype=bool),
             status=np.full(n, ACTIVE, dtype=np.int32), levelg=levelg.copy())
        setc(system, pl=pl)
        setc(lst, nenc=nenc, index1=i1.copy(), index2=i2.copy(), status=lstatus.copy())
        w.call("symba_kick_list_plpl", lst, system, dtk, irec, sgn)
        out["kl_plpl_irec%d_sgn%d_vb" % (irec, sgn)] = back(pl.c["vb"])
        out["kl_plpl_irec%d_sgn%d_ah" % (irec, sgn)] = back(pl.c["ah"])
    # pl-tp list
    dd = np.linalg.norm(r[:, None] - rt[None], axis=2)
    flat_order = np.argsort(dd.ravel())
    pick = np.concatenate([flat_order[:100], rng.choice(flat_order[100:], 30, replace=False)])
    rng.shuffle(pick)
    p1, p2 = (pick // ntp + 1).astype(np.int32), (pick % ntp + 1).astype(np.int32)
    ne2 = len(p1)
    st2 = np.full(ne2, ACTIVE, dtype=np.int32); st2[5::11] = INACTIVE
    levelg_tp = rng.integers(0, 3, ntp).astype(np.int32)
    vbt0, aht0 = vt.copy(), rng.normal(0, 1e-3, (ntp, 3))
    out["kt_rh_tp"], out["kt_levelg_tp"], out["kt_index1"], out["kt_index2"], out["kt_lactive"] = rt, levelg_tp, p1, p2, st2 == ACTIVE
    out["kt_vb0"], out["kt_ah0"] = vbt0, aht0
    for irec, sgn in ((1, 1), (2, -1), (2, 1)):
        pl, tp = w.new_object("symba_pl"), w.new_object("symba_tp")
        system, lst = w.new_object("symba_nbody_system"), w.new_object("symba_list_pltp")
        setc(pl, nbody=n, rh=fa(r), Gmass=Gm.copy(), rhill=rhill_big.copy(), lmask=np.ones(n, dtype=bool),
             status=np.full(n, ACTIVE, dtype=np.int32), levelg=levelg.copy())
        setc(tp, nbody=ntp, rh=fa(rt), vb=fa(vbt0), ah=fa(aht0), lmask=np.ones(ntp, dtype=bool),
             status=np.full(ntp, ACTIVE, dtype=np.int32), levelg=levelg_tp.copy())
        setc(system, pl=pl, tp=tp)
        setc(lst, nenc=ne2, index1=p1.copy(), index2=p2.copy(), status=st2.copy())
        w.call("symba_kick_list_pltp", lst, system, dtk, irec, sgn)
        out["kt_pltp_irec%d_sgn%d_vb" % (irec, sgn)] = back(tp.c["vb"])
        out["kt_pltp_irec%d_sgn%d_ah" % (irec, sgn)] = back(tp.c["ah"])

    # ---- the recursion's encounter re-check over the list (with symba_util_set_renc)
    param = w.new_object("swiftest_parameters")
    radius_big = rhill_big * 0.35                       # some listed pairs overlap physically and must be dropped
    out["el_radius"], out["el_vb_pl"], out["el_vb_tp"] = radius_big, v, vt
    for irec in (1, 2):
        level = np.where(np.arange(nenc) % 5 == 0, irec, irec - 1).astype(np.int32)      # only level == irec-1 is examined
        pl, system, lst = w.new_object("symba_pl"), w.new_object("symba_nbody_system"), w.new_object("symba_list_plpl")
        setc(pl, nbody=n, rh=fa(r), vb=fa(v), rhill=rhill_big.copy(), radius=radius_big.copy(), renc=np.zeros(n),
             levelg=np.zeros(n, dtype=np.int32), levelm=np.zeros(n, dtype=np.int32))
        setc(system, pl=pl)
        setc(lst, nenc=nenc, index1=i1.copy(), index2=i2.copy(), status=lstatus.copy(), level=level.copy(),
             lvdotr=np.zeros(nenc, dtype=bool))
        lany = w.call("symba_encounter_check_list_plpl", lst, param, system, 0.05, irec)
        out["el_plpl_irec%d_level0" % irec] = level
        out["el_plpl_irec%d_level" % irec], out["el_plpl_irec%d_lvdotr" % irec] = lst.c["level"].copy(), lst.c["lvdotr"].copy()
        out["el_plpl_irec%d_renc" % irec], out["el_plpl_irec%d_lany" % irec] = pl.c["renc"].copy(), bool(lany)
        out["el_plpl_irec%d_levelg" % irec] = pl.c["levelg"].copy()
        level2 = np.where(np.arange(ne2) % 4 == 0, irec, irec - 1).astype(np.int32)
        pl, tp = w.new_object("symba_pl"), w.new_object("symba_tp")
        system, lst = w.new_object("symba_nbody_system"), w.new_object("symba_list_pltp")
        setc(pl, nbody=n, rh=fa(r), vb=fa(v), rhill=rhill_big.copy(), radius=radius_big.copy(), renc=np.zeros(n),
             levelg=np.zeros(n, dtype=np.int32), levelm=np.zeros(n, dtype=np.int32))
        setc(tp, nbody=ntp, rh=fa(rt), vb=fa(vt), levelg=np.zeros(ntp, dtype=np.int32), levelm=np.zeros(ntp, dtype=np.int32))
        setc(system, pl=pl, tp=tp)
        setc(lst, nenc=ne2, index1=p1.copy(), index2=p2.copy(), status=st2.copy(), level=level2.copy(),
             lvdotr=np.zeros(ne2, dtype=bool))
        lany = w.call("symba_encounter_check_list_pltp", lst, param, system, 0.05, irec)
        out["el_pltp_irec%d_level0" % irec] = level2
        out["el_pltp_irec%d_level" % irec], out["el_pltp_irec%d_lvdotr" % irec] = lst.c["level"].copy(), lst.c["lvdotr"].copy()
        out["el_pltp_irec%d_lany" % irec] = bool(lany)

    # ---- collision_check_one (with xv2aeq) and discard_pl_close on a spread of relative states
    m = 400
    rel = rng.normal(0, 1.0, (m, 3)) * 10 ** rng.uniform(-3, 0, (m, 1))
    vel = rng.normal(0, 1.0, (m, 3)) * 10 ** rng.uniform(-1, 1.5, (m, 1))
    vel[::3] = rel[::3] * rng.uniform(0.5, 50, (len(rel[::3]), 1)) + vel[::3] * 0.05     # mostly receding: the xv2aeq branch
    Gmtot = 10 ** rng.uniform(-7, -3, m)
    rlim = 10 ** rng.uniform(-4, -1.5, m)
    lvd = rng.uniform(size=m) > 0.3
    dtc = 0.02
    lcol, lclo = np.zeros(m, dtype=bool), np.zeros(m, dtype=bool)
    iflag, r2min = np.zeros(m, dtype=np.int32), np.zeros(m)
    for k in range(m):
        a, b = F.Cell(False), F.Cell(False)
        w.call("collision_check_one", *map(float, rel[k]), *map(float, vel[k]), float(Gmtot[k]), float(rlim[k]), dtc, bool(lvd[k]), a, b)
        lcol[k], lclo[k] = a.v, b.v
        fl, rm = F.Cell(0), F.Cell(float("nan"))     # r2min stays undefined on the early exits: NaN marks them
        w.call("swiftest_discard_pl_close", rel[k].copy(), vel[k].copy(), dtc, float(rlim[k]) ** 2, fl, rm)
        iflag[k], r2min[k] = fl.v, rm.v
    out["cc_rel"], out["cc_vel"], out["cc_Gmtot"], out["cc_rlim"], out["cc_lvdotr"], out["cc_dt"] = rel, vel, Gmtot, rlim, lvd, dtc
    out["cc_lcollision"], out["cc_lclosest"], out["dc_iflag"], out["dc_r2min"] = lcol, lclo, iflag, r2min
    print("  collision_check_one: %d collisions, %d closest; discard_pl_close: %d flagged" % (lcol.sum(), lclo.sum(), iflag.sum()),
          flush=True)


def main():
    import collections
    import json
    which = sys.argv[1:] or ["kick", "drift", "encounter", "steps", "lists"]
    w = world()
    cov_path = os.path.join(HERE, "fortran_coverage.json")
    coverage = json.load(open(cov_path)) if os.path.exists(cov_path) else {}
    for part, fn in (("kick", gen_kick), ("drift", gen_drift), ("encounter", gen_encounter), ("steps", gen_steps),
                     ("lists", gen_lists)):
        if part not in which:
            continue
        t0 = time.time()
        out = {}
        w.trace = []
        fn(w, out)
        # which reference procedures were executed, and how often (tests assert the list)
        coverage[part] = dict(sorted(collections.Counter(t.strip() for t in w.trace).items()))
        w.trace = None
        if part == "encounter":
            # norm2 is processor dependent: the pair lists must not depend on the variant
            F.NORM2_MODE = "libgfortran"
            w2 = world()
            out2 = {}
            gen_encounter(w2, out2)
            F.NORM2_MODE = "plain"
            same = all(np.array_equal(out[k], out2[k]) for k in out)
            print("  pair lists identical under libgfortran's scaled norm2:", same)
            assert same
            out["norm2_variants_identical"] = np.array(same)
        path = os.path.join(HERE, "fortran_%s.npz" % part)
        np.savez_compressed(path, **out)
        print("%s: %d arrays -> %s (%.0f s)" % (part, len(out), path, time.time() - t0), flush=True)
    with open(cov_path, "w") as f:
        json.dump(coverage, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
