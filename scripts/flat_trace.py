"""Per-warp timeline of one third-law gravity launch (development aid; needs SWCU_FLAT_TRACE=<file> in the environment).
usage: SWCU_FLAT_TRACE=/tmp/t.bin python scripts/flat_trace.py [npl]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from swiftest_b200 import Context, PL, LOOP_FLAT, workloads as W  # noqa: E402
from swiftest_b200.context import FAM_PLPL  # noqa: E402

path = os.environ["SWCU_FLAT_TRACE"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
d = W.disk(n, seed=7)
with Context(0) as c:
    c.enable_kernel_timing(True)
    c.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"], mu=d["mu"],
                generation=1)
    for it in range(4):
        c.flush_l2()
        c.body_zero_accel(PL)
        c.pl_accel_int(LOOP_FLAT, True)
        ms = c.last_kernel_ms(FAM_PLPL)
    t = np.fromfile(path, dtype=np.uint64).reshape(-1, 4).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    st, en, fl = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    print(f"npl={n} group {ms * 1e3:.1f} us; warps {len(t)}; kernel span {fl.max():.1f} us")
    print("start skew  us: p50 %.1f p99 %.1f max %.1f" % tuple(np.percentile(st, [50, 99, 100])))
    print("loop end    us: min %.1f p1 %.1f p50 %.1f p99 %.1f max %.1f" % tuple(np.percentile(en, [0, 1, 50, 99, 100])))
    print("after flush us: max %.1f" % fl.max())
    busy = (en - st).sum() / (len(t) * fl.max())
    print(f"warp-busy fraction of the span: {busy:.4f}; chunks/warp min {t[:, 2].min()} median {np.median(t[:, 2])} max {t[:, 2].max()}")
    print(f"idle at the end (sum over warps of span-end)/warps: {(fl.max() - en).mean():.1f} us; at start: {st.mean():.1f} us")
