"""Fixed cost vs work of the third-law gravity launch group: time it at several npl and fit t = F + c * pairs.
usage: python scripts/flat_scan.py [reps]   (development aid)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, LOOP_FLAT, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLPL  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
noflush = len(sys.argv) > 2 and sys.argv[2] == "noflush"
rows = []
with Context(0) as c:
    c.enable_kernel_timing(True)
    for gen, n in enumerate((12800, 25600, 36096, 51200, 72448, 100000)):
        d = W.disk(n, seed=7)
        c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                    mu=d["mu"], generation=100 + gen)
        ms = []
        for it in range(reps + 2):
            if not noflush:
                c.flush_l2()
            c.body_zero_accel(PL)
            c.pl_accel_int(LOOP_FLAT, True)
            t = c.last_kernel_ms(FAM_PLPL)
            if it >= 2:
                ms.append(t)
        pairs = n * (n - 1) / 2
        rows.append((n, pairs, float(np.min(ms)), float(np.median(ms))))
        print(f"npl={n:7d} pairs={pairs:.4e} min {rows[-1][2]:8.4f} ms  median {rows[-1][3]:8.4f} ms  "
              f"{pairs / rows[-1][2] / 1e6:8.1f} Gpairs/s")
p = np.array([r[1] for r in rows])
t = np.array([r[2] for r in rows])
A = np.vstack([np.ones_like(p), p]).T
F, cst = np.linalg.lstsq(A, t, rcond=None)[0]
print(f"fit: t = {F * 1e3:.1f} us + pairs / {1 / cst / 1e6:.1f} Gpairs/s")
