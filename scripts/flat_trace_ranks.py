"""Summarise the per-rank per-warp timelines of one multi-GPU third-law launch (SWCU_FLAT_TRACE=<prefix> writes
<prefix>.rank<r>; %globaltimer is a per-GPU clock, so only the spans within a rank are compared).
usage: python scripts/flat_trace_ranks.py <prefix> <nranks>"""
import sys

import numpy as np

prefix, nr = sys.argv[1], int(sys.argv[2])
print(f"{'rank':>4} {'warps':>6} {'span us':>9} {'loop end p1':>11} {'p50':>8} {'p99':>8} {'max':>8} {'idle at end us':>14} {'chunks/warp min/med/max':>24}")
spans = []
for r in range(nr):
    t = np.fromfile(f"{prefix}.rank{r}" if nr > 1 else prefix, dtype=np.uint64).reshape(-1, 4).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    st, en, fl = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 3] - t0) / 1e3
    spans.append(fl.max())
    p = np.percentile(en, [1, 50, 99, 100])
    print(f"{r:>4} {len(t):>6} {fl.max():>9.1f} {p[0]:>11.1f} {p[1]:>8.1f} {p[2]:>8.1f} {p[3]:>8.1f} {(fl.max() - en).mean():>14.1f} "
          f"{t[:, 2].min():>8}/{int(np.median(t[:, 2]))}/{t[:, 2].max()}")
print(f"kernel span over ranks: min {min(spans):.1f} us, max {max(spans):.1f} us, spread {100 * (max(spans) - min(spans)) / max(spans):.1f} %")
