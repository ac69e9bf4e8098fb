"""Instruction mix of the loops of a kernel in an object file (development aid).
usage: python scripts/sass_loops.py <obj> <kernel-substring> [min_mufu]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
min_mufu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); funcs[name] = []; continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name: funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
def op(t): return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
for fn, ins in funcs.items():
    if pat not in fn: continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    print(fn, len(ins), "instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:U?P\d,\s*)?0x([0-9a-f]+)", t)
        if not m: continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = ins[addr[tgt]:i + 1]
            nm = sum("MUFU" in x for _, x in body)
            if nm < min_mufu: continue
            ops = collections.Counter(op(x).split(".")[0] for _, x in body)
            fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
            print(f"  loop {tgt:#x}..{a:#x}: {len(body)} instr, MUFU {nm}, FP64 {fp64} ({fp64/nm:.2f}/pair), other {len(body)-fp64} ({(len(body)-fp64)/nm:.2f}/pair)")
            print("   ", dict(sorted(ops.items(), key=lambda kv: -kv[1])))
