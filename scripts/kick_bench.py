"""Quick device-resident timing of the pl-pl gravity variants (development aid, not the headline bench).
usage: python scripts/kick_bench.py [npl] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, LOOP_FLAT, LOOP_TRIANGULAR, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLPL  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
d = W.disk(n, seed=3031179)
pairs = n * (n - 1) / 2
with Context(0) as c:
    peak = c.probe_fp64_peak()
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                generation=1)
    c.enable_kernel_timing(True)
    res = {}
    for name, var in (("tri", LOOP_TRIANGULAR), ("flat", LOOP_FLAT)):
        for lclose in (True, False):
            ms = []
            for it in range(reps + 2):
                c.body_zero_accel(PL)
                c.pl_accel_int(var, lclose)
                t = c.last_kernel_ms(FAM_PLPL)
                if it >= 2:
                    ms.append(t)
            a = c.body_get(PL, r=False, v=False)["a"]
            res[(name, lclose)] = a
            t = float(np.min(ms))
            flop = (28.0 if lclose else 25.0) * pairs
            print(f"{name:4s} lclose={int(lclose)}  {t:8.3f} ms  {pairs / t / 1e6:9.1f} Gpairs/s  "
                  f"{flop / t / 1e9:6.2f} TFLOP/s algorithmic = {100 * flop / t / 1e9 / peak:5.1f}% of {peak:.1f} TF measured")
    for lclose in (True, False):
        a, b = res[("tri", lclose)], res[("flat", lclose)]
        print(f"flat vs tri lclose={int(lclose)}: max rel diff {np.max(np.abs(a - b)) / np.max(np.abs(a)):.2e}")
