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
print("sanitize pass done")
