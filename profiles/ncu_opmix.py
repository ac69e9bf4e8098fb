"""Per-opcode executed warp-instruction mix from `ncu -i X.ncu-rep --page source --csv` (needs -lineinfo builds).
usage: python profiles/ncu_opmix.py source.csv <units> [label]   -- units = warp-level work units (e.g. warp-pairs)"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hdr = rows[1]
ci, ie = hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
tot, stalls = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        n = int(r[ie])
    except Exception:
        continue
    t = r[ci].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") else t[0]
    tot[op.split(".")[0]] += n
    for h, i in stall_cols.items():
        try:
            stalls[h] += int(r[i])
        except Exception:
            pass
s = sum(tot.values())
print(f"total warp instructions {s}  ({s / units:.2f} per unit)")
for k, v in tot.most_common(30):
    print(f"  {k:10s} {v:14d} {v / units:8.2f}")
fp64 = sum(v for k, v in tot.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"FP64-pipe (DFMA+DMUL+DADD+DSETP): {fp64 / units:.2f} per unit")
ss = sum(stalls.values())
print("stall samples:", ", ".join(f"{k[6:]}={100 * v / ss:.1f}%" for k, v in stalls.most_common(8)))
