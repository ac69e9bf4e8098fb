"""Third-law (flat) vs full-row (triangular) pl-pl gravity kernel as a function of npl: kernel launch-group time of both
through the resident entry point, best of several launches (development aid; the table goes to profiles/r02_crossover.md
and the SWCU_LOOP_AUTO threshold in swcu_api.cu follows it).  usage: python scripts/crossover_scan.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, LOOP_FLAT, LOOP_TRIANGULAR, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLPL  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 8192, 10000, 16384, 32768]
print("| npl | flat (third-law) us | tri (full-row) us | flat/tri | faster |\n|---|---|---|---|---|")
with Context(0) as c:
    c.enable_kernel_timing(True)
    for n in sizes:
        d = W.disk(n, seed=3031179)
        c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                    generation=1000 + n)
        t = {}
        for name, var in (("flat", LOOP_FLAT), ("tri", LOOP_TRIANGULAR)):
            ms = []
            for it in range(12):
                c.flush_l2()
                c.body_zero_accel(PL)
                c.pl_accel_int(var, True)
                if it >= 4:
                    ms.append(c.last_kernel_ms(FAM_PLPL))
            t[name] = float(np.median(ms)) * 1e3
        print(f"| {n} | {t['flat']:.1f} | {t['tri']:.1f} | {t['flat'] / t['tri']:.2f} | {'flat' if t['flat'] < t['tri'] else 'tri'} |", flush=True)
