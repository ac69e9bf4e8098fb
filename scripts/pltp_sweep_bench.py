"""pl-tp encounter sweep of BASELINE configs[1] (8 planets + 1e6 test particles): the sort-free pass over the particles
against the sort path (SWCU_PLTP_DIRECT_MAX=0).  Run on the GPU box:  python scripts/pltp_sweep_bench.py [ntp]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, TP, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_SWEEP  # noqa: E402

ntp = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
p = W.planets8_year_units()
tp = W.tp_cloud(ntp, seed=123)
with Context(0) as ctx:
    ctx.enable_kernel_timing(True)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"])
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"])
    ctx.pl_set_renc(0)
    for mode in ("direct", "sort"):
        if mode == "sort":
            os.environ["SWCU_PLTP_DIRECT_MAX"] = "0"
        ms, wall = [], []
        for it in range(10):
            ctx.flush_l2()
            ctx.synchronize()
            n0 = ctx.launch_count()
            t0 = time.perf_counter()
            nenc = ctx.tp_encounter_check(0.01, fetch=False)
            wall.append((time.perf_counter() - t0) * 1e3)
            ms.append(ctx.last_kernel_ms(FAM_SWEEP))
            nl = ctx.launch_count() - n0
        print(mode, "nenc", nenc, ctx.encounter_stats(), "launches", nl, "event ms", np.round(np.mean(ms[3:]), 4),
              "wall ms", np.round(np.mean(wall[3:]), 4), "direct/fallback", ctx.encounter_direct_count())
