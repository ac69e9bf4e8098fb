"""swiftest_discard_pl_tp on resident populations: 8 planets x ntp particles.  python scripts/discard_bench.py [ntp]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, TP, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLTP  # noqa: E402

ntp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
p = W.planets8_year_units()
tp = W.tp_cloud(ntp, seed=3)
with Context(0) as c:
    c.enable_kernel_timing(True)
    c.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"])
    c.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"])
    ms, wall = [], []
    for it in range(8):
        c.flush_l2()
        c.synchronize()
        t0 = time.perf_counter()
        n = c.tp_discard_pl(0.01, want_iplanet=False)[1]
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(c.last_kernel_ms(FAM_PLTP))
    print("resident discard 8 pl x %d tp: kernel ms %.4f wall ms %.4f discarded %d" % (ntp, np.mean(ms[3:]), np.mean(wall[3:]), n))
