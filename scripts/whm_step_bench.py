"""whm_step_pl on a resident system of npl massive bodies (multi-launch form above 128): wall time and launches per step.
python scripts/whm_step_bench.py [npl] [nsteps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, LOOP_AUTO, workloads as W  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
d = W.disk(n, seed=3031179)
with Context(0) as ctx:
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=W.GMSUN + d["Gmass"])
    for rep in range(2):
        ctx.synchronize()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        for k in range(nsteps):
            ctx.whm_step_pl(W.GMSUN, d["dt"], LOOP_AUTO, True, lfirst=(k == 0 and rep == 0), want_nfail=False)
        ctx.synchronize()
        el = time.perf_counter() - t0
    print("npl", n, "whm_step_pl us/step", round(1e6 * el / nsteps, 1), "launches/step", (ctx.launch_count() - n0) / nsteps)
