"""pl-pl sort-and-sweep of the SyMBA disk (BASELINE configs[2]/[3]): time per call and launches per call.
Run on the GPU box:  python scripts/plpl_sweep_bench.py [npl]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_SWEEP  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10000
d = W.disk(n, seed=3031179)
with Context(0) as ctx:
    ctx.enable_kernel_timing(True)
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"])
    ctx.pl_set_renc(0)
    ms, wall = [], []
    for it in range(10):
        ctx.flush_l2()
        ctx.synchronize()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        nenc = ctx.pl_encounter_check(d["dt"], fetch=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        ms.append(ctx.last_kernel_ms(FAM_SWEEP))
        nl = ctx.launch_count() - n0
    print("npl", n, "nenc", nenc, ctx.encounter_stats(), "launches", nl, "event ms", np.round(np.mean(ms[3:]), 4),
          "wall ms", np.round(np.mean(wall[3:]), 4))
