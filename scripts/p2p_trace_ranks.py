"""Phases of the fused reduce/kick/drift/allgather kernel per rank (SWCU_P2P_TRACE=<prefix> writes <prefix>.rank<r>:
per CTA the %globaltimer at entry / flags seen / partial sums loaded / drift done / stores issued / system fence passed).
usage: python scripts/p2p_trace_ranks.py <prefix> <nranks>"""
import sys

import numpy as np

prefix, nr = sys.argv[1], int(sys.argv[2])
names = ["entry->flags", "flags->loaded", "loaded->drifted", "drifted->stored", "stored->fenced"]
print(f"{'rank':>4} {'CTAs':>5} {'kernel span us':>14}   " + "  ".join(f"{n:>16}" for n in names) + "   (median / max over CTAs, us)")
for r in range(nr):
    t = np.fromfile(f"{prefix}.rank{r}", dtype=np.uint64).reshape(-1, 8).astype(np.int64)
    t = t[(t[:, 0] > 0) & (t[:, 5] > 0)]
    span = (t[:, 5].max() - t[:, 0].min()) / 1e3
    cols = []
    for k in range(5):
        dtk = (t[:, k + 1] - t[:, k]) / 1e3
        dtk = dtk[t[:, k + 1] > 0]
        cols.append(f"{np.median(dtk):7.1f}/{dtk.max():7.1f}")
    first_flags = (t[:, 1].min() - t[:, 0].min()) / 1e3
    print(f"{r:>4} {len(t):>5} {span:>14.1f}   " + "  ".join(f"{c:>16}" for c in cols) + f"   first CTA past the flags after {first_flags:.1f} us")
