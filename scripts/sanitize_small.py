"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, TP, LOOP_FLAT, LOOP_TRIANGULAR, workloads as W  # noqa: E402

with Context(0) as c:
    for n, nplm in ((700, 700), (515, 130)):
        d = W.disk(n, seed=n)
        acc = np.zeros((n, 3))
        c.kick_getacch_int_all_tri_pl(n, nplm, d["rh"], d["Gmass"], d["radius"], acc)
        c.kick_getacch_int_all_flat_pl(n, nplm * n - nplm * (nplm + 1) // 2, None, d["rh"], d["Gmass"], d["radius"], acc)
        k = np.array([[1, 2], [3, 7], [5, 100]], np.int32)
        c.kick_getacch_int_all_flat_pl(n, 3, k, d["rh"], d["Gmass"], d["radius"], acc)
        c.symba_kick_subtract_encounters(n, k[:, 0], k[:, 1], d["rh"], d["Gmass"], d["radius"], acc)
        renc = d["rhill"] * 6.5 * 4
        print("plpl", c.encounter_check_all_sort_and_sweep_plpl(n, d["rh"], d["vh"], renc, d["dt"])[0])
        if nplm < n:
            print("plplm", c.encounter_check_all_plplm(nplm, n - nplm, d["rh"][:nplm], d["vh"][:nplm], d["rh"][nplm:],
                                                      d["vh"][nplm:], renc[:nplm], renc[nplm:], d["dt"])[0])
        x, v, fl = d["rh"].copy(), d["vh"].copy(), np.zeros(n, np.int32)
        c.drift_all(d["mu"], x, v, n, d["dt"], np.ones(n, np.int32), fl)
    p = W.planets8_year_units()
    tp = W.tp_cloud(3000, seed=1)
    at = np.zeros((3000, 3))
    c.kick_getacch_int_all_tp(3000, 8, tp["rh"], p["rh"], p["Gmass"], np.ones(3000, np.int32), at)
    print("pltp", c.encounter_check_all_sort_and_sweep_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 6.5, 0.05)[0])
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                mu=p["cb_Gmass"] + p["Gmass"], generation=1)
    c.body_sync(TP, 3000, r=tp["rh"], v=tp["vh"], mu=np.full(3000, p["cb_Gmass"]), generation=2)
    c.body_zero_accel(TP)
    c.tp_accel_int()
    print("fused nfail", c.whm_tp_step(0.01, np.zeros(3)))
    # round 2: third-law rollback path (planted overlapping / coincident pairs), WHM planet step with the device ah0,
    # helio steps, slice I/O (blocking and asynchronous), energy sums
    from swiftest_b200 import LOOP_AUTO
    n = 900
    d = W.disk(n, seed=5)
    r = d["rh"].copy()
    r[700] = r[3] + np.array([0.5 * d["radius"].max(), 0.0, 0.0])
    r[650] = r[140]
    acc = np.zeros((n, 3))
    c.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, r, d["Gmass"], d["radius"], acc)
    print("flat redo chunks", c.flat_redo_count())
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                mu=p["cb_Gmass"] + p["Gmass"], generation=3)
    c.body_put(TP, r=tp["rh"], v=tp["vh"])
    c.whm_tp_first_accel()
    for k in range(2):
        c.whm_step_pl(p["cb_Gmass"], 0.01, LOOP_TRIANGULAR, True, lfirst=(k == 0))
        c.whm_tp_step(0.01, None)
    c.whm_get_jacobi()
    for k in range(2):
        c.helio_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=(k == 0))
        c.helio_step_tp(p["cb_Gmass"], 0.01, lfirst=(k == 0))
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                generation=4)
    import torch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    hr, hv = pin(d["rh"][100:400]), pin(d["vh"][100:400])
    out = {k: pin(np.zeros((300, 3))) for k in ("r", "v", "a")}
    for _ in range(3):
        c.body_put_range_async(PL, 100, 400, r=hr, v=hv)
        c.body_zero_accel(PL)
        c.pl_accel_int(LOOP_FLAT, True)
        c.body_kick_velocity(PL, d["dt"])
        c.body_drift(PL, d["dt"])
        c.body_get_range_async(PL, 100, 400, out)
    c.io_wait()
    c.body_put_range(PL, 0, 10, r=d["rh"][:10])
    c.body_get_range(PL, 5, 50)
    mass = d["Gmass"] / W.GMSUN
    print("pe", c.util_get_potential_energy(n, None, W.GMSUN, d["Gmass"], mass, d["rh"]))
    # round 2, second half: sort-free pl-tp sweep (incl. a particle on an outer extent -> sort path), resident list kernels
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import pltp_rule as R
    for kind in ("rmin", "dup", "rmax"):
        a = R.tie_case(kind, ntp=1500)
        print("pltp", kind, c.encounter_check_all_sort_and_sweep_pltp(8, 1500, *a[:4], a[4], a[5])[0], c.encounter_direct_count())
    os.environ["SWCU_PLTP_DIRECT_MAX"] = "0"
    print("pltp sort path", c.encounter_check_all_sort_and_sweep_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 6.5, 0.05)[0])
    del os.environ["SWCU_PLTP_DIRECT_MAX"]
    n = 600
    d = W.disk(n, seed=9)
    _, i1, i2, _ = c.encounter_check_all_triangular_plpl(n, d["rh"], d["vh"], d["rhill"] * 6.5 * 6, d["dt"])
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"] * 200, rhill=d["rhill"] * 6, generation=5)
    c.body_put_vb(PL, d["vh"])
    lev = np.ones(n, np.int32)
    c.pl_set_renc(1)
    print("resident lists", len(i1), c.pl_symba_kick_list(i1, i2, None, lev, d["dt"], 1, 1).sum(),
          c.body_symba_encounter_check_list(PL, i1, i2, None, d["dt"])[2],
          c.body_collision_check_list(PL, i1, i2, None, np.ones(len(i1), np.int32), d["dt"])[2])
    _, j1, j2, _ = c.encounter_check_all_triangular_pltp(8, 3000, p["rh"], p["vh"], tp["rh"], tp["vh"], p["rhill"] * 39, 0.05)
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"], generation=6)
    c.body_sync(TP, 3000, r=tp["rh"], v=tp["vh"], generation=7)
    c.pl_set_renc(0)
    print("resident pl-tp lists", len(j1), c.tp_symba_kick_list(j1, j2, None, np.ones(8, np.int32), np.ones(3000, np.int32), 0.01, 1, 1).sum(),
          c.body_symba_encounter_check_list(TP, j1, j2, None, 0.02)[2],
          c.body_collision_check_list(TP, j1, j2, None, np.ones(len(j1), np.int32), 5.0)[2])
    # second half of round 2, later additions: multi-launch WHM step (tiled chains), helio step replayed as a CUDA graph,
    # bucket sort with a clump (radix fallback) and device-side finalisation, serial-order sums, RSQ64H potential energy
    n = 300
    d = W.disk(n, seed=11)
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"] * 100, radius=d["radius"], rhill=d["rhill"],
                mu=W.GMSUN + d["Gmass"] * 100, generation=8)
    for k in range(3):
        c.whm_step_pl(W.GMSUN, d["dt"], LOOP_TRIANGULAR, True, lfirst=(k == 0))
    n = 700
    d = W.disk(n, seed=12)
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                mu=np.full(n, W.GMSUN), generation=9)
    for k in range(6):
        c.helio_step_pl(W.GMSUN, d["dt"], LOOP_AUTO, True, lfirst=(k == 0), want_nfail=(k == 5))
    print("graph replays", c.step_graph_replays())
    r = d["rh"].copy()
    r[:3000 if n > 3000 else 0] = 0.0
    big = W.disk(6000, seed=13)
    rb = big["rh"].copy()
    rb[:3000] = 0.0
    rb[np.arange(3000), np.arange(3000) % 3] = 1.0
    rencb = big["rhill"] * 6.5 * 3
    rencb[:3000] = 0.0
    print("clump sweep", c.encounter_check_all_sort_and_sweep_plpl(6000, rb, big["vh"], rencb, big["dt"])[0],
          "bucket fallbacks", c.encounter_bucket_fallbacks())
    print("pe", c.util_get_potential_energy(n, None, W.GMSUN, d["Gmass"], d["Gmass"] / W.GMSUN, d["rh"]))
print("sanitize pass done")
