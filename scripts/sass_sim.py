"""Crude issue simulation of a kernel's hot loop from its SASS (development aid): NW warps per SMSP run the same loop
body round-robin (oldest ready first); an FP64 instruction holds the issue port 2 cycles, everything else 1; register
dependencies use fixed latencies (FP64 8, MUFU 24, LDS 30, other 5).  Prints cycles per loop iteration per warp and the
issue-bound value, i.e. how much of the loss is the instruction ORDER (lack of independent work between dependent
instructions) rather than the instruction count.
usage: python scripts/sass_sim.py <obj> <kernel-substring> <mufu-per-iteration> [warps]"""
import collections, re, subprocess, sys

obj, pat, nmufu = sys.argv[1], sys.argv[2], int(sys.argv[3])
NW = int(sys.argv[4]) if len(sys.argv) > 4 else 3
import os
LAT_FP64 = int(os.environ.get("LAT_FP64", 8)); LAT_MUFU = int(os.environ.get("LAT_MUFU", 24))
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
funcs, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1); funcs[name] = []; continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name: funcs[name].append((int(m.group(1), 16), m.group(2).strip()))

def regs(tok, wide):
    out = []
    for m in re.finditer(r"(?<![U\w])R(\d+)", tok):
        r = int(m.group(1)); out.append(r)
        if wide: out.append(r + 1)
    return out

def parse(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op, _, rest = t.partition(" ")
    base = op.split(".")[0]
    ops = [o.strip() for o in rest.split(",")] if rest else []
    fp64 = base in ("DFMA", "DMUL", "DADD")
    wide_dst = fp64 or ".64" in op or ".128" in op
    dst, src = [], []
    if base in ("STS", "ST", "STG", "BRA", "ISETP", "RED", "REDG", "YIELD", "NOP", "BSYNC", "BSSY"):
        for o in ops: src += regs(o, ".64" in op or ".128" in op or fp64)
        if ".128" in op:
            for o in ops[1:]:
                for r in regs(o, False): src += [r, r + 1, r + 2, r + 3]
    else:
        if ops:
            d = regs(ops[0], False)
            for r in d:
                n = 4 if ".128" in op else (2 if wide_dst else 1)
                dst += list(range(r, r + n))
        for o in ops[1:]: src += regs(o, fp64)
    lat = LAT_FP64 if fp64 else LAT_MUFU if base == "MUFU" else 30 if base in ("LDS", "LD", "LDG") else 5
    return base, fp64, dst, src, lat

for fn, ins in funcs.items():
    if pat not in fn: continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:U?P\d,\s*)?0x([0-9a-f]+)", t)
        if not m: continue
        tgt = int(m.group(1), 16)
        if not (tgt < a and tgt in addr): continue
        body = [parse(x) for _, x in ins[addr[tgt]:i + 1]]
        if sum(b[0] == "MUFU" for b in body) != nmufu: continue
        nfp = sum(b[1] for b in body)
        bound = 2 * nfp + (len(body) - nfp)
        ITER = 6
        pc = [0] * NW; it = [0] * NW; ready = [collections.defaultdict(int) for _ in range(NW)]
        cyc = 0; port_free = 0; done = [None] * NW; start = [None] * NW
        last = 0
        while any(d is None for d in done):
            issued = False
            if cyc >= port_free:
                for k in range(NW):
                    w = (last + 1 + k) % NW
                    if done[w] is not None: continue
                    b = body[pc[w]]
                    if all(ready[w][r] <= cyc for r in b[3]) and all(ready[w][r] <= cyc for r in b[2]):
                        for r in b[2]: ready[w][r] = cyc + b[4]
                        port_free = cyc + (2 if b[1] else 1)
                        pc[w] += 1
                        if pc[w] == len(body):
                            pc[w] = 0; it[w] += 1
                            if it[w] == 1: start[w] = cyc
                            if it[w] == ITER: done[w] = cyc
                        last = w; issued = True
                        break
            cyc += 1
        per = sum((done[w] - start[w]) / (ITER - 1) for w in range(NW)) / NW
        print(f"{fn[-40:]} loop {tgt:#x}: {len(body)} instr ({nfp} FP64), issue bound {bound} cycles/iter/warp x{NW} = {bound*NW}; "
              f"simulated {per:.0f} cycles/iter for {NW} warps -> efficiency {bound*NW/per:.3f}; per pair {per/NW/nmufu:.2f} (bound {bound/nmufu:.2f})")
