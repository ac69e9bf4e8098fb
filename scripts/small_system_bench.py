"""Times the planet step of a small system (Sun + 8 planets) and the whole WHM step of BASELINE configs[1]
(8 planets + 1e6 test particles, everything device resident) -- fused one-launch planet step vs the multi-launch form
(SWCU_WHM_FUSED=0 in the environment selects the latter).  python scripts/small_system_bench.py [nsteps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, TP, LOOP_TRIANGULAR  # noqa: E402
from swiftest_b200 import workloads as W  # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ctx = Context(0)
p = W.planets8_year_units()
GMcb, dt = p["cb_Gmass"], 0.01
ntp = 1000000
tp = W.tp_cloud(ntp, seed=17)
ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"], mu=GMcb + p["Gmass"],
              generation=1)
ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, GMcb), generation=2)
ctx.whm_tp_first_accel()
res = {}
for what in ("whm_step_pl", "whm_step_pl+tp", "helio_step_pl"):
    for rep in range(2):
        ctx.synchronize()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        for k in range(nsteps):
            if what == "helio_step_pl":
                ctx.helio_step_pl(GMcb, dt, loop_variant=LOOP_TRIANGULAR, lclose=True, lfirst=(k == 0 and rep == 0), want_nfail=False)
            else:
                ctx.whm_step_pl(GMcb, dt, LOOP_TRIANGULAR, True, lfirst=(k == 0 and rep == 0 and what == "whm_step_pl"),
                                want_nfail=False)
                if what.endswith("+tp"):
                    ctx.whm_tp_step(dt, None, want_nfail=False)
        ctx.synchronize()
        el = time.perf_counter() - t0
    res[what] = dict(us_per_step=1e6 * el / nsteps, launches_per_step=(ctx.launch_count() - n0) / nsteps)
print("SWCU_WHM_FUSED=%s" % os.environ.get("SWCU_WHM_FUSED", "1"), res)
