"""Timing of the test-particle kernels of the WHM / HELIO configurations (8 planets + ntp test particles): pl->tp kick,
Kepler drift, fused WHM tp step, fused HELIO tp step (development aid; also the ncu target for these kernels).
usage: python scripts/tp_bench.py [ntp] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, TP, LOOP_AUTO, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLTP, FAM_DRIFT  # noqa: E402

ntp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
HBM = 6556.8
p = W.planets8_year_units()
tp = W.tp_cloud(min(ntp, 1000000), seed=123)
rh, vh = tp["rh"], tp["vh"]
if ntp > len(rh):
    k = -(-ntp // len(rh))
    rh, vh = np.tile(rh, (k, 1))[:ntp], np.tile(vh, (k, 1))[:ntp]
with Context(0) as c:
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                mu=p["cb_Gmass"] + p["Gmass"], generation=2)
    c.body_sync(TP, ntp, r=rh, v=vh, mu=np.full(ntp, p["cb_Gmass"]), generation=3)
    c.enable_kernel_timing(True)
    ah0 = np.zeros(3)
    for i in range(8):
        r2 = float(p["rh"][i] @ p["rh"][i])
        ah0 -= p["Gmass"][i] / (r2 * np.sqrt(r2)) * p["rh"][i]

    def leg(name, fn, fam, bytes_per_tp):
        ms = []
        for it in range(reps + 2):
            c.flush_l2()
            fn()
            if it >= 2:
                ms.append(c.last_kernel_ms(fam))
        t = float(np.mean(ms))
        gbs = bytes_per_tp * ntp / (t * 1e-3) / 1e9
        print(f"{name:28s} {t * 1e3:8.1f} us  {gbs:7.0f} GB/s algorithmic = {100 * gbs / HBM:5.1f} % of {HBM:.0f}", flush=True)

    def kick():
        c.body_zero_accel(TP)
        c.tp_accel_int()
    leg("pl->tp kick (76 B/tp)", kick, FAM_PLTP, 76.0)
    leg("Kepler drift (112 B/tp)", lambda: c.body_drift(TP, 0.01, want_nfail=False), FAM_DRIFT, 112.0)
    c.body_put(TP, r=rh, v=vh)
    kick()
    leg("fused WHM tp step (152 B/tp)", lambda: c.whm_tp_step(0.01, ah0, want_nfail=False), FAM_DRIFT, 152.0)
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                mu=np.full(8, p["cb_Gmass"]), generation=12)
    c.body_put(TP, r=rh, v=vh)
    first = [True]

    def helio():
        c.helio_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=first[0], want_nfail=False)
        c.helio_step_tp(p["cb_Gmass"], 0.01, lfirst=first[0], want_nfail=False)
        first[0] = False
    leg("fused HELIO tp step (152 B/tp)", helio, FAM_DRIFT, 152.0)
