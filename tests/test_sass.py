"""Static checks of the compiled sm_100a code (no GPU needed): the instruction mix of the third-law hot loop that the
issue model of profiles/r01_fp64_pipe.md is built on, the TMA bulk copy in the full-row kernel, and no register spills
in the hot kernels.  Skipped when the object files or cuobjdump are not there (e.g. on a box that only received the .so).
"""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "swiftest_b200", "csrc", "build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(os.path.join(BUILD, "kick_flat_kernels.o")) and os.path.exists(CUOBJDUMP)),
                                reason="object files or cuobjdump not available")


def _sass(obj):
    out = subprocess.run([CUOBJDUMP, "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
    funcs, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and name:
            funcs[name].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def _loops(ins):
    """(start, end, body) of every backward branch."""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA\s+(?:U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                yield ins[addr[tgt]:i + 1]


def _op(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]


FLAT_KERNEL = "kick_flat_kernelILb1ELi3ELi2ELb1"   # radius-checked, 3 CTAs/SM, 2 steps per iteration, column prefetch


def _flat_hot_loop():
    funcs = _sass("kick_flat_kernels.o")
    name = next(n for n in funcs if FLAT_KERNEL in n)
    best = None
    for body in _loops(funcs[name]):
        if sum("MUFU.RSQ64H" in t for _, t in body) != 8:               # 2 steps x 4 row bodies per iteration
            continue
        if best is None or len(body) < len(best):
            best = body                                                 # the unchecked variant is the shortest
    assert best is not None, "hot loop not found"
    return best


def test_third_law_hot_loop_instruction_mix():
    body = _flat_hot_loop()
    ops = collections.Counter(_op(t) for _, t in body)
    fp64 = ops["DFMA"] + ops["DMUL"] + ops["DADD"]
    other = len(body) - fp64
    # 20 FP64 per pair: 3 differences, 3 for r^2, 6 for r^-3, 2 mass factors, 6 accumulations
    assert fp64 == 160, (fp64, dict(ops))
    # everything else: one MUFU.RSQ64H per pair, half a 3-input min, LDS/STS of the column and its accumulators, the
    # per-step __syncwarp (BRA.DIV + NOP) and the loop -- at most 4.5 per pair (profiles/r02_kick_flat.md); no seed
    # conversion, compare or select on the fast path
    assert other <= 36, (other, dict(ops))
    assert ops["MUFU"] == 8 and ops["VIMNMX3"] == 4, dict(ops)
    assert not any(o in ops for o in ("F2F", "LDL", "STL", "DSETP", "SEL", "FSEL", "LEA", "SHFL", "BAR")), dict(ops)
    assert ops["ISETP"] <= 1, dict(ops)                                 # the loop test only


def test_third_law_accumulator_accesses_stay_in_program_order():
    """The j-side accumulators are read by one lane one step after the neighbouring lane wrote them (a __syncwarp ends
    every step); the volatile shared-memory accesses must also keep their program order in the SASS: in every step the
    accumulator loads come before its stores, and the next step's loads after them."""
    body = _flat_hot_loop()
    seq = []
    for _, t in body:
        m = re.match(r"(?:@!?P\d+\s+)?(LDS|STS)\.(128|64)\s+(.*)", t)
        if not m:
            continue
        kind, width, rest = m.groups()
        off = re.search(r"\[R\d+(?:\+(-?0x[0-9a-f]+))?\]", rest)
        o = int(off.group(1), 16) if off and off.group(1) else 0
        if kind == "STS" or width == "64" or (o & 0xf00) == 0x800:     # column bodies are LDS.128 below +0x800
            seq.append((kind, width, o))
    # per step: LDS.128 axy, LDS.64 az (either order), then STS.128 axy, STS.64 az; steps in ascending slot order
    assert [k for k, _, _ in seq] == ["LDS", "LDS", "STS", "STS"] * 2, seq
    slots = [o & 0xff for _, _, o in seq]
    assert slots[:4] == [slots[0]] * 4 and slots[4:] == [slots[0] + 0x10] * 4, seq


def test_full_row_kernel_uses_tma_bulk_copies_and_16_fp64_per_evaluation():
    funcs = _sass("kick_kernels.o")
    rows = [n for n in funcs if "kick_rows_kernelILi4" in n]
    assert rows
    text = [t for _, t in funcs[rows[0]]]
    assert any("UBLKCP" in t for t in text), "cp.async.bulk (TMA) not in the full-row kernel"
    assert any("SYNCS" in t for t in text), "mbarrier handshake missing"
    counts = []
    for body in _loops(funcs[rows[0]]):
        nm = sum("MUFU.RSQ" in t for _, t in body)
        if nm and nm % 4 == 0:
            ops = collections.Counter(_op(t) for _, t in body)
            counts.append((ops["DFMA"] + ops["DMUL"] + ops["DADD"]) / nm)
    assert counts and min(counts) == 16.0, counts


def test_third_law_kernel_keeps_its_bookkeeping_spills_small():
    """The third-law kernel runs at the 168-register cap of 3 CTAs/SM; a few words of claim bookkeeping may spill
    OUTSIDE the step loop (the hot-loop test above forbids LDL/STL inside it), the parameter block must not be copied
    to local memory (that was a 576-byte frame in round 1)."""
    log = open(os.path.join(BUILD, "kick_flat_kernels.ptxas.log")).read()
    for k in ("kick_flat_kernelILb1ELi3ELi2ELb1", "kick_flat_kernelILb0ELi3ELi2ELb1"):
        m = re.search(r"Compiling entry function '[^']*" + re.escape(k) + r"[^']*' for 'sm_100a'.*?\n\s*(\d+) bytes stack frame, "
                      r"(\d+) bytes spill stores, (\d+) bytes spill loads", log, flags=re.S)
        assert m, k
        assert int(m.group(1)) <= 64 and int(m.group(2)) <= 64, (k, m.groups())


def test_hot_kernels_do_not_spill():
    logs = {f: open(os.path.join(BUILD, f)).read() for f in os.listdir(BUILD) if f.endswith(".ptxas.log")}
    want = {"kick_kernels.ptxas.log": ["kick_rows_kernelILi4", "kick_tp_small_kernel"],
            "energy_kernels.ptxas.log": ["pe_pairs_kernel"],
            "encounter_kernels.ptxas.log": ["sweep_kernel", "tri_check_kernel"]}
    for log, kernels in want.items():
        assert log in logs, log
        for k in kernels:
            m = re.search(r"Compiling entry function '[^']*" + re.escape(k) + r"[^']*' for 'sm_100a'.*?\n(.*?spill loads)",
                          logs[log], flags=re.S)
            assert m, (log, k)
            assert "0 bytes spill stores, 0 bytes spill loads" in m.group(1), (k, m.group(1))
