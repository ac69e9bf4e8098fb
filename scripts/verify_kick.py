"""One-shot check of the gravity kernels after a change to the pair arithmetic: parity against the oracle at small
sizes, then timing at npl = 1e5 (development aid).  usage: python scripts/verify_kick.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import load  # noqa: E402
from swiftest_b200 import Context, PL, LOOP_FLAT, LOOP_TRIANGULAR, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLPL  # noqa: E402

t00 = time.time()
o = load()
worst = 0.0
with Context(0) as c:
    for n, rads in ((1000, (True, False)), (4099, (True,))):
        d = W.disk(n, seed=100 + n)
        for use_rad in rads:
            rad = d["radius"] if use_rad else None
            ref = o.kick_tri_pl(d["rh"], d["Gmass"], rad, np.zeros((n, 3)))
            scale = o.kick_tri_abs_scale(d["rh"], d["Gmass"], rad)
            a = np.zeros((n, 3))
            c.kick_getacch_int_all_tri_pl(n, n, d["rh"], d["Gmass"], rad, a)
            b = np.zeros((n, 3))
            c.kick_getacch_int_all_flat_pl(n, n * (n - 1) // 2, None, d["rh"], d["Gmass"], rad, b)
            e1, e2 = float(np.max(np.abs(a - ref) / scale)), float(np.max(np.abs(b - ref) / scale))
            worst = max(worst, e1, e2)
            print(f"n={n} rad={use_rad}: tri {e1:.2e} flat {e2:.2e}", flush=True)
    p = W.planets8_year_units()
    tp = W.tp_cloud(20000, seed=3)
    ones = np.ones(20000, np.int32)
    ref = o.kick_all_tp(tp["rh"], p["rh"], p["Gmass"], ones, np.zeros((20000, 3)))
    got = np.zeros((20000, 3))
    c.kick_getacch_int_all_tp(20000, 8, tp["rh"], p["rh"], p["Gmass"], ones, got)
    e = float(np.max(np.abs(got - ref)) / np.abs(ref).max())
    worst = max(worst, e)
    d = W.disk(300, seed=9)
    ref = o.kick_all_tp(tp["rh"][:5000], d["rh"], d["Gmass"], ones[:5000], np.zeros((5000, 3)))
    got = np.zeros((5000, 3))
    c.kick_getacch_int_all_tp(5000, 300, tp["rh"][:5000], d["rh"], d["Gmass"], ones[:5000], got)
    e2 = float(np.max(np.abs(got - ref)) / np.abs(ref).max())
    worst = max(worst, e2)
    print(f"tp small {e:.2e} rows {e2:.2e}  WORST {worst:.2e} {'PARITY-OK' if worst < 1e-12 else 'PARITY-FAIL'}", flush=True)
    n = 100000
    d = W.disk(n, seed=3031179)
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                generation=1)
    c.enable_kernel_timing(True)
    for name, var, reps in (("flat", LOOP_FLAT, 3), ("tri", LOOP_TRIANGULAR, 2)):
        ms = []
        for _ in range(reps):
            c.body_zero_accel(PL)
            c.pl_accel_int(var, True)
            ms.append(c.last_kernel_ms(FAM_PLPL))
        print(f"{name} npl=1e5: {min(ms):.4f} ms  {n * (n - 1) / 2 / min(ms) / 1e6:.1f} Gpairs/s", flush=True)
print(f"total {time.time() - t00:.1f} s")
