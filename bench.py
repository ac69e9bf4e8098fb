#!/usr/bin/env python
"""bench.py -- headline benchmark of the force-and-drift hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU path on the host cores
  torchrun ... bench.py --gpus N ...                        N>1: one rank per GPU

Metric: FP64 pair-interactions/s of the SyMBA planetesimal disk, npl = 1e5 fully interacting massive bodies
(BASELINE.json configs[3]; fits one GPU).  One STEP = one pass of the hot path over the resident system:
   ah = 0 ; pl%accel_int (pl-pl gravity, radius-checked, all N(N-1)/2 pairs) ; vb += ah*dt ; pl%drift (Kepler drift)
   ; [N>1: third-law kernel on this rank's run of block pairs, then ONE kernel that reduce-scatters the partial
   accelerations out of every rank's memory, kicks and drifts this rank's slice and allgathers it into every rank's
   arrays over NVLink peer memory (--collective nccl: ncclAllReduce instead; --variant tri: row slices + allgather)].
`value` = N(N-1)/2 pairs per step / step time with everything resident in HBM (strong scaling: the system is fixed).
Timing: K laps, one CUDA-event pair per step on the library's stream, L2 flushed (160 MiB written, then another 160 MiB read so that no dirty lines remain) BETWEEN the laps,
barrier + synchronize on both sides, max over ranks.  `e2e` = the same step with that step's positions and velocities
copied from pinned host memory and accelerations / positions / velocities read back, every step (N>1: every rank moves
its own slice of the bodies and the slices are allgathered on the device), host wall clock, max over ranks.
After the timed region `parity_check` compares the result of the timed code path with the CPU oracle (N=1) or with a
one-GPU run of the same steps (N>1; all ranks must hold identical bits).
The sort-and-sweep encounter check, the WHM test-particle configuration (8 planets + 1e6..1e8 tp), the 1e4-body SyMBA
disk (flat vs full-row) and the 1e4-step conservation runs are side legs reported under "extra".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PAIR_RAD = 28.0   # algorithmic flop per unordered pair with the radius check
FLOP_PER_TPEVAL = 17.0
DRIFT_BYTES_PER_BODY = 112.0
# potential energy, per unordered pair (swiftest_util.f90:1325-1329): 3 sub, 5 for r^2, sqrt, divide, Gm*m, add
FLOP_PER_PE_PAIR = 12.0
# fused helio tp step, per tp: r, vb, lmask read (52 B); r, vb, vh, ah, iflag written (100 B)
HELIO_TP_BYTES = 152.0
SWEEP_BYTES = dict(body=56.0, sort=2 * 24.0 * 2, gather=2 * 56.0, cand=56.0, out=9.0)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npl", type=int, default=100000)
    ap.add_argument("--ntp", type=int, default=1000000)
    ap.add_argument("--variant", default="auto", choices=["auto", "tri", "flat"])
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"],
                    help="N>1, third-law kernel: fused reduce+update+allgather over NVLink peer memory (p2p) or "
                         "ncclAllReduce of the partial accelerations (nccl)")
    ap.add_argument("--no-extra", action="store_true", help="skip the sweep / tp side legs and the CPU baseline")
    ap.add_argument("--cpu-evals", type=int, default=3, help="full CPU force evaluations timed for cpu_baseline")
    ap.add_argument("--conservation-long", type=int, default=1000000,
                    help="steps of the reference's own conservation test (Sun + 8 planets, dt = 0.01 y, 1e4 y = 1e6 steps, 1000 "
                         "outputs; extra.conservation_reference_test; 0 = skip; stops early after ~100 s)")
    ap.add_argument("--conservation", type=int, default=10000,
                    help="steps of the Sun + 8 planets energy/L tracking run (extra.conservation; 0 = skip)")
    ap.add_argument("--disk-steps", type=int, default=1000,
                    help="steps of the 1e4-body disk energy/L tracking run (extra.conservation_disk; 0 = skip)")
    ap.add_argument("--tp-sizes", default="1e6,1e7,1e8", help="total test particles of the sharded WHM tp legs")
    return ap.parse_args()


def common_config(n):
    """The workload both arms run; identical in the two JSON lines."""
    return {"workload": f"symba_disk_npl{n}_fully_interacting", "npl": n, "nplm": n, "lclose": True,
            "step": "zero_accel + pl-pl accel_int (all N(N-1)/2 pairs, radius-checked) + kick_velocity + Kepler drift",
            "generator": "Chambers-style disk, swiftest_b200/workloads.py::disk", "seed": 3031179,
            "l2": "GPU arm: L2 flushed between timed steps (160 MiB = 1.33 x L2 written, then 160 MiB read: clean lines only; outside the per-step event pairs); "
                  "CPU arm: not applicable"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  NVML is polled from a thread every
    few milliseconds (an nvidia-smi child process takes longer to start than a short timed region lasts); nvidia-smi
    -lms is the fallback when pynvml is missing."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index, self.rows, self.stop_flag, self.thread, self.mode = index, [], False, None, None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.mode = "smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:  # wait for the first sample before timing starts
                time.sleep(0.02)
        except Exception:
            self.mode = None

    def _poll_nvml(self):
        nv = self.nv
        k, watts = 0, 0.0
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                if k % 8 == 0:  # NVML queries take milliseconds under load: power only now and then
                    try:
                        watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                    except Exception:
                        pass
                k += 1
                self.rows.append((mhz, self.max_mhz, watts, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [q.strip() for q in line.split(",")]
            try:
                mask = 0
                for bit, val in zip((0x8, 0x40, 0x20, 0x4), r[3:7]):
                    if val.lower().startswith("active"):
                        mask |= bit
                self.rows.append((float(r[0]), float(r[1]), float(r[2]), mask))
            except Exception:
                pass

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        if self.mode == "nvml":
            time.sleep(0.01)
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[3]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "power_w_max": float(max(r[2] for r in self.rows)), "samples": len(sm), "source": self.mode,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b)}


# =====================================================================================================================
# reference arm / CPU baseline: the oracle's reference-shaped OpenMP loops on the host cores
# =====================================================================================================================
def cpu_full_evaluation(o, d):
    """One full pl-pl force evaluation: ALL rows of the full-row loop (swiftest_kick.f90:219-240 shape, OpenMP
    schedule(static), every thread this process may use).  Returns seconds."""
    n = d["n"]
    acc = np.zeros((n, 3))
    t0 = time.perf_counter()
    o.omp_kick_tri_rad_pl_rows(d["rh"], d["Gmass"], d["radius"], acc, n, 0, n)
    return time.perf_counter() - t0


def cpu_baseline(d, evals):
    from oracle import load
    o = load(native=True)
    threads = o.omp_set_threads()
    n = d["n"]
    cpu_full_evaluation(o, d)  # warm-up
    ts = [cpu_full_evaluation(o, d) for _ in range(max(1, evals))]
    t = float(np.mean(ts))
    pairs = n * (n - 1) / 2.0
    return {"value": pairs / t, "unit": "pair-interactions/s", "cores": threads, "kind": "port",
            "sample": f"{len(ts)} full force evaluations (all {n} rows of the full-row pl-pl loop, kick.f90:219-240 shape, "
                      f"OpenMP schedule(static), {threads} threads, gcc -O3 -march=x86-64-v3 strict IEEE), "
                      f"{t:.2f} s each; nothing extrapolated",
            "seconds_full_evaluation": t}


def run_reference(args):
    """--impl reference: the CPU path (oracle port; the Fortran reference cannot be built in this image or on the GPU
    box: profiles/r02_fortran_probe.txt) on the host cores.  A step = the full pl-pl kick (all rows, all threads this
    process may use -- set explicitly, torchrun exports OMP_NUM_THREADS=1) + the serial Kepler drift of all bodies."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import load
    from swiftest_b200 import workloads as W
    o = load(native=True)
    threads = o.omp_set_threads()
    d = W.disk(args.npl, seed=3031179)
    n = d["n"]
    pairs = n * (n - 1) / 2.0
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        acc = np.zeros((n, 3))
        o.omp_kick_tri_rad_pl_rows(d["rh"], d["Gmass"], d["radius"], acc, n, 0, n)
        v = d["vh"] + acc * d["dt"]
        x, v, fl = o.drift_all(d["mu"], d["rh"], v, d["dt"])  # serial, as in the reference (drift.f90:99-103)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    step = float(np.mean(times))
    val = pairs / step
    sample = (f"per step: all {n} rows of the full-row pl-pl loop on {threads} OpenMP threads + velocity kick + serial "
              f"Kepler drift of all {n} bodies; nothing extrapolated")
    print(json.dumps({
        "impl": "reference", "metric": "FP64 pair-interactions/s (pl-pl N=1e5)", "value": val,
        "unit": "pair-interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": common_config(n),
        "impl_detail": {"loop": "triangular full-row, radius-checked", "host": "cpu", "threads": threads},
        "cpu_baseline": {"value": val, "unit": "pair-interactions/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# =====================================================================================================================
# our arm
# =====================================================================================================================
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    from swiftest_b200 import Context, PL, TP, LOOP_FLAT, LOOP_TRIANGULAR, shard, workloads as W
    from swiftest_b200.context import FAM_PLPL, FAM_PLTP, FAM_DRIFT, FAM_SWEEP, FAM_ALLGATHER

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def max_over_ranks(x):
        if not dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(ident, 0)
        ctx.comm_init(world, rank, bytes(ident.cpu().numpy().tobytes()))

    # ---------------- workload: SyMBA disk, npl fully interacting bodies ----------------
    d = W.disk(args.npl, seed=3031179)
    n, dt = d["n"], d["dt"]
    pairs = n * (n - 1) / 2.0
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=d["mu"], generation=1)
    # auto = the third-law (flat) kernel: measured 1.43x faster than the full-row kernel at npl = 1e5
    variant = {"auto": LOOP_FLAT, "tri": LOOP_TRIANGULAR, "flat": LOOP_FLAT}[args.variant]
    i0, i1 = shard.partition(n, world, rank)
    if variant == LOOP_TRIANGULAR:
        ctx.pl_set_slice(i0, i1)   # full-row kernel: balanced i-slices, allgather of the drifted slices
    # flat kernel: balanced runs of block pairs per rank, allreduce of the partial accelerations inside
    # swcu_pl_accel_int; every rank then kicks and drifts all bodies (O(N), identical results), no allgather

    use_p2p = world > 1 and variant == LOOP_FLAT and args.collective == "p2p"
    if use_p2p:
        # CUDA-IPC handles of every rank's exchange buffers, gathered with torch.distributed (host plumbing only)
        ok = torch.ones(1, device="cuda")
        try:
            mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        except Exception:
            mine, ok = torch.zeros(512, dtype=torch.uint8, device="cuda"), torch.zeros(1, device="cuda")
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        if ok.item():
            try:
                ctx.p2p_import(world, rank, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
            except Exception:
                ok = torch.zeros(1, device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not ok.item():  # peer mapping unavailable on this box: every rank falls back to the NCCL collective
            use_p2p = False
        dist.barrier()

    def step():  # one pass of the hot path over the resident system (no flush in here)
        if use_p2p:
            ctx.pl_kick_drift_p2p(dt, True, want_nfail=False)
            return
        ctx.body_zero_accel(PL)
        ctx.pl_accel_int(variant, True)
        ctx.body_kick_velocity(PL, dt)
        ctx.body_drift(PL, dt, want_nfail=False)
        if world > 1 and variant == LOOP_TRIANGULAR:
            ctx.pl_allgather(with_v=True)

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        ctx.flush_l2()
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    ctx.enable_kernel_timing(2)  # an event pair per launch group, read after the loop: nothing stalls the stream
    for _ in range(args.steps):
        # the flush goes first and is NOT inside the lap: while it runs (~26 us) the host queues the step's first
        # launches, so the lap starts on a busy stream exactly when the flush ends
        ctx.flush_l2()
        ctx.timer_lap_begin()
        step()
        ctx.timer_lap_end()
    ms_total, nlaps, laps = ctx.timer_laps(each=args.steps)
    assert nlaps == args.steps
    barrier()
    fam_ms = {name: ctx.kernel_ms_accumulated(f) for name, f in
              (("gravity", FAM_PLPL), ("drift", FAM_DRIFT), ("collective", FAM_ALLGATHER))}
    ctx.enable_kernel_timing(0)
    launches = ctx.launch_count() - launches0 - args.steps  # the flush kernels are not part of the step
    ms_step = max_over_ranks(ms_total / args.steps)
    value = pairs / (ms_step * 1e-3)
    kick_ms_avg = max_over_ranks(fam_ms["gravity"][0] / max(1, fam_ms["gravity"][1]))
    breakdown = {k: (max_over_ranks(v[0] / args.steps) if v[1] else 0.0) for k, v in fam_ms.items()}
    breakdown["step_min"], breakdown["step_max"] = float(min(laps)), float(max(laps))
    breakdown["step_median"] = float(np.median(laps))  # `value` is the MEAN lap (the contract's K steps / time); one lap
    # disturbed by the box (seen once: 9.8 ms among nine of 6.89) moves the mean by 4 % and shows here as max >> median
    fp64_peak = ctx.probe_fp64_peak()

    # ---------------- end-to-end: host buffers in, results out, every step ----------------
    # Every rank owns a slice of the bodies on the host side too (the shape of swiftest_coarray_distribute_system; the
    # whole population at N = 1): each step it uploads the slice's r, v from pinned memory, [N>1: the slices are
    # allgathered on the device], runs the step and reads the slice's r, v, a back.  Headline form: the asynchronous
    # calls (copies on the library's copy streams, double-buffered staging) -- the upload of step k+1 and the read-back
    # of step k overlap the kernels of the neighbouring steps; every step still moves its own inputs and results.
    # `e2e_blocking` is the same with the blocking calls (each returns when its copy is complete).
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    m = i1 - i0
    h_r, h_v = pin(d["rh"][i0:i1]), pin(d["vh"][i0:i1])
    out = {"r": pin(np.zeros((m, 3))), "v": pin(np.zeros((m, 3))), "a": pin(np.zeros((m, 3)))}

    def gather_slices():
        if world > 1:
            ctx.pl_set_slice(i0, i1)
            ctx.pl_allgather(with_v=True)
            if variant != LOOP_TRIANGULAR:
                ctx.pl_set_slice(0, n)   # the third-law step kicks/drifts through its own partition

    def step_e2e_async():
        ctx.body_put_range_async(PL, i0, i1, r=h_r, v=h_v)
        gather_slices()
        step()
        ctx.body_get_range_async(PL, i0, i1, out)

    def step_e2e_blocking():
        ctx.body_put_range(PL, i0, i1, r=h_r, v=h_v)
        gather_slices()
        step()
        ctx.body_get_range(PL, i0, i1, out=out)

    def time_e2e(fn):
        for _ in range(max(1, args.warmup)):
            fn()
        ctx.io_wait()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        ctx.io_wait()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / args.steps)

    h2d, d2h = 2 * 3 * n * 8, 3 * 3 * n * 8   # summed over the ranks: every body crosses PCIe once each way per step
    e2e_s = time_e2e(step_e2e_async)
    # the last step's results arrived where the caller asked for them: r, v of the slice moved, a is finite
    e2e_ok = bool(np.all(np.isfinite(out["a"])) and np.any(out["a"] != 0.0) and not np.array_equal(out["r"], h_r))
    e2e_blk_s = time_e2e(step_e2e_blocking)
    sl = "slice " if world > 1 else ""
    e2e = {"value": pairs / e2e_s, "unit": "pair-interactions/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "results_delivered": e2e_ok,
           "path": f"per rank and step: swcu_body_put_range_async({sl}r,v; pinned) -> "
                   + ("swcu_pl_allgather (device) -> " if world > 1 else "")
                   + f"step -> swcu_body_get_range_async({sl}r,v,a; pinned); swcu_io_wait after the K steps; host wall "
                     "clock between barriers, max over ranks"}
    e2e_blocking = {"value": pairs / e2e_blk_s, "unit": "pair-interactions/s", "ms_per_step": e2e_blk_s * 1e3,
                    "path": "same with swcu_body_put_range / swcu_body_get_range (each call returns when its copy is complete)"}

    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (device-resident and e2e)

    # ---------------- parity of the timed code path (outside the timed regions) ----------------
    try:
        parity = parity_check(ctx, args, d, variant, use_p2p, rank, world, local, dist, step, i0, i1)
    except Exception as e:
        parity = {"ok": False, "error": str(e)}

    if use_p2p:  # the side legs re-sync other populations: the peer mappings of the 1e5-body arrays go first
        barrier()
        ctx.p2p_close()
        barrier()
    hbm_peak, peak_src = peaks()
    flops_kernel = FLOP_PER_PAIR_RAD * pairs / world  # algorithmic flop of one rank's launch (balanced shares)
    achieved = flops_kernel / (kick_ms_avg * 1e-3) / 1e12
    kname = "kick_flat_kernel (third-law pl-pl gravity)" if variant == LOOP_FLAT else "kick_rows_kernel (full-row pl-pl gravity)"
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if variant == LOOP_FLAT and n == 100000 and world == 1:
            traffic = tj["kick_flat_kernel"]["dram_bytes_read"] + tj["kick_flat_kernel"]["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"bound": "fp64", "kernel": kname, "achieved": achieved, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": achieved / fp64_peak, "traffic": traffic,
                "peak_source": "DFMA microbenchmark run in this process (swcu_probe_fp64_peak); MEASURED_PEAKS.json holds "
                               "no FP64 figure",
                "algorithmic_flop_per_pair": FLOP_PER_PAIR_RAD, "kernel_ms": kick_ms_avg,
                "note": "kernel_ms brackets the gravity launch group (memset, radius bounds, kernel, +ah update); an FP64 "
                        "instruction holds the B200 issue port 2 cycles and nothing co-issues, and three-register FP64 "
                        "streams sustain ~0.455 of the 0.5 inst/cycle/SMSP the peak probe reaches "
                        "(profiles/r02_kick_flat.md)"}

    extra = {}
    if not args.no_extra and rank == 0 and world == 1:
        try:
            extra = side_legs(ctx, args, d, hbm_peak, peak_src)
            extra["next_rows"] = next_row_legs(ctx, args, d, hbm_peak, fp64_peak)
        except Exception as e:  # side legs are reported figures, never a dependency of the headline number
            extra["error"] = str(e)
        try:
            extra["symba_1e4"] = symba_1e4_leg(ctx, hbm_peak, fp64_peak)
        except Exception as e:
            extra["symba_1e4"] = {"error": str(e)}
    cpu = None
    if not args.no_extra and rank == 0 and world == 1:
        try:
            cpu = cpu_baseline(d, args.cpu_evals)
        except Exception as e:  # the CPU baseline is a reported figure, never a dependency of the GPU number
            cpu = {"value": None, "unit": "pair-interactions/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if not args.no_extra and rank == 0 and world == 1:
        if args.conservation:
            try:
                extra["conservation"] = conservation_run(ctx, args.conservation)
            except Exception as e:
                extra["conservation"] = {"error": str(e)}
        if args.disk_steps:
            try:
                extra["conservation_disk"] = conservation_disk_run(ctx, args.disk_steps)
            except Exception as e:
                extra["conservation_disk"] = {"error": str(e)}
        if args.conservation_long:
            try:
                extra["conservation_reference_test"] = conservation_reference_test(ctx, args.conservation_long)
            except Exception as e:
                extra["conservation_reference_test"] = {"error": str(e)}

    if not args.no_extra:
        legs = {}
        for tok in args.tp_sizes.split(","):
            ntp_total = int(float(tok))
            try:  # a side leg must never cost the headline line
                legs[f"ntp_{tok.strip()}"] = tp_sharded_leg(ctx, ntp_total, rank, world, barrier, max_over_ranks, hbm_peak, dist)
            except Exception as e:
                legs[f"ntp_{tok.strip()}"] = {"error": str(e)}
        extra["whm_tp_sharded"] = legs

    if rank == 0:
        line = {
            "metric": "FP64 pair-interactions/s (pl-pl N=1e5)", "value": value, "unit": "pair-interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": common_config(n),
            "impl_detail": {"loop": "triangular full-row" if variant == LOOP_TRIANGULAR else "flat third-law",
                            "multi_gpu_step": ((("reduce-scatter + kick + drift + allgather fused in one kernel over NVLink "
                                                 "peer memory" if use_p2p else "ncclAllReduce(ah)") if variant == LOOP_FLAT
                                                else "allgather(r,v)") if world > 1 else "none"),
                            "sharding": (f"balanced runs of block pairs over {world} rank(s), per-rank work counters"
                                         if variant == LOOP_FLAT else f"i-slices over {world} rank(s)"),
                            "collective": ("p2p-fused" if use_p2p else ("nccl" if world > 1 else "none")),
                            "timing": "one CUDA-event pair per step on the library's stream (laps), L2 flush between laps, "
                                      "max over ranks of the mean lap"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_blocking": e2e_blocking,
            "gpu_launches": int(launches), "clocks": clocks,
            "parity_check": parity, "breakdown_ms_per_step": breakdown,
            "peaks": {"fp64_tflops_measured": fp64_peak, "hbm_gbs": hbm_peak, "hbm_source": peak_src},
            "extra": extra}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        ctx.comm_finalize()
        dist.destroy_process_group()
    ctx.close()


def parity_check(ctx, args, d, variant, use_p2p, rank, world, local, dist, step, i0, i1):
    """Result of the timed code path against an independent computation, outside the timed regions.
    N = 1: the accelerations of one pass against the CPU oracle's full-row loop on 1e4 sampled rows (1e-12 of the
    per-component sum of |terms|), and r, v after the pass against oracle kick + drift on those rows.
    N > 1: M passes of the multi-GPU step from the initial state; every rank must end with bit-identical r, v, and
    they must agree with the same M passes of the ONE-GPU path (a second context on rank 0's GPU) to 1e-12."""
    import torch
    from swiftest_b200 import Context, PL, LOOP_FLAT
    n, dt = d["n"], d["dt"]
    tol = 1e-12
    if world == 1:
        from oracle import load
        o = load(native=True)
        o.omp_set_threads()
        ctx.body_put(PL, r=d["rh"], v=d["vh"])
        step()
        got = ctx.body_get(PL)
        blocks = [(b0, min(n, b0 + 2500)) for b0 in np.linspace(0, max(0, n - 2500), 4).astype(int)] if n > 10000 else [(0, n)]
        rows = np.unique(np.concatenate([np.arange(a, b) for a, b in blocks]))
        ref = np.zeros((n, 3))
        for a, b in blocks:
            ref[a:b] = 0.0
            o.omp_kick_tri_rad_pl_rows(d["rh"], d["Gmass"], d["radius"], ref, n, int(a), int(b))
        scale = o.kick_tri_abs_scale(d["rh"], d["Gmass"], d["radius"])
        acc_err = float(np.max(np.abs(got["a"][rows] - ref[rows]) / np.where(scale[rows] > 0, scale[rows], 1.0)))
        v1 = d["vh"][rows] + ref[rows] * dt
        xr, vr, fr = o.drift_all(d["mu"][rows], d["rh"][rows], v1, dt)
        r_err = float(np.max(np.abs(got["r"][rows] - xr) / np.linalg.norm(xr, axis=1, keepdims=True)))
        v_err = float(np.max(np.abs(got["v"][rows] - vr) / np.linalg.norm(vr, axis=1, keepdims=True)))
        return {"against": "CPU oracle (full-row OpenMP loop + Kepler drift)", "rows_compared": int(len(rows)),
                "acc_max_err_over_sum_abs_terms": acc_err, "r_max_rel_err": r_err, "v_max_rel_err": v_err, "tol": tol,
                "redo_chunks": ctx.flat_redo_count(), "ok": bool(acc_err < tol and r_err < tol and v_err < tol)}
    M = 3
    ctx.body_put(PL, r=d["rh"], v=d["vh"])
    ctx.synchronize()
    dist.barrier()
    for _ in range(M):
        step()
    got = ctx.body_get(PL, a=False)
    # identical bits on every rank: compare 64-bit digests of r and v
    words = np.concatenate([got["r"].ravel(), got["v"].ravel()]).view(np.uint64)
    digest = np.array([np.bitwise_xor.reduce(words), np.add.reduce(words, dtype=np.uint64)], dtype=np.uint64)
    t = torch.from_numpy(digest.view(np.int64)).cuda()
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    identical = all(bool(torch.equal(allt[0], x)) for x in allt)
    res = {"against": f"one-GPU run of the same {M} steps (second context on rank 0)", "steps": M, "ranks": world,
           "ranks_bit_identical": identical, "rows_compared": n, "tol": tol}
    if rank == 0:
        with Context(local) as c1:
            c1.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                         mu=d["mu"], generation=1)
            for _ in range(M):
                c1.body_zero_accel(PL)
                c1.pl_accel_int(variant, True)
                c1.body_kick_velocity(PL, dt)
                c1.body_drift(PL, dt, want_nfail=False)
            one = c1.body_get(PL, a=False)
        res["r_max_rel_err"] = float(np.max(np.abs(got["r"] - one["r"]) / np.linalg.norm(one["r"], axis=1, keepdims=True)))
        res["v_max_rel_err"] = float(np.max(np.abs(got["v"] - one["v"]) / np.linalg.norm(one["v"], axis=1, keepdims=True)))
        res["ok"] = bool(identical and res["r_max_rel_err"] < tol and res["v_max_rel_err"] < tol)
    dist.barrier()
    return res


def tp_shard(ntp_total, t0, t1):
    """Bodies [t0, t1) of a synthetic cloud of ntp_total test particles: the 1e6-particle cloud of the WHM configuration
    (workloads.tp_cloud, seed 123) tiled, replica k scaled by (1 + 1e-3 k / replicas) in r and by the matching
    Keplerian factor in v, so every particle stays on a bound, distinct orbit.  Generated shard by shard: a rank never
    materialises the whole cloud."""
    from swiftest_b200 import workloads as W
    base_n = min(ntp_total, 1_000_000)
    base = _TP_BASE.get(base_n)
    if base is None:
        base = _TP_BASE[base_n] = W.tp_cloud(base_n, seed=123)
    reps = max(1, -(-ntp_total // base_n))
    idx = np.arange(t0, t1, dtype=np.int64)
    k = (idx // base_n).astype(np.float64) / reps
    src = idx % base_n
    f = (1.0 + 1e-3 * k)[:, None]
    return base["rh"][src] * f, base["vh"][src] / np.sqrt(f)


_TP_BASE = {}


def tp_sharded_leg(ctx, ntp_total, rank, world, barrier, max_over_ranks, hbm_peak, dist):
    """WHM test particles over the GPUs of the run: block partition like swiftest_coarray_distribute_system
    (swiftest_coarray.f90:705-711), planets replicated, NO per-step communication.  The fused tp step (152 B per tp) is
    timed lap by lap (one event pair per launch); the L2 flush between laps is outside them and is skipped when the
    shard is larger than L2 anyway."""
    import psutil
    import torch
    from swiftest_b200 import PL, TP, shard, workloads as W
    p = W.planets8_year_units()
    t0, t1 = shard.tp_block_partition(ntp_total, world, rank)
    m = t1 - t0
    need_host = m * 8 * (3 + 3 + 1) * 2.5   # r, v, mu + temporaries of the generator
    need_dev = m * 130 + 2 * m * 24         # resident arrays + staging of the upload
    ok = psutil.virtual_memory().available > need_host + (8 << 30) and torch.cuda.mem_get_info()[0] > need_dev + (4 << 30)
    if dist is not None:
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item() > 0)
    if not ok:
        return {"skipped": f"not enough free host or device memory for {m} particles per rank"}
    r, v = tp_shard(ntp_total, t0, t1)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=902 + ntp_total % 1000)
    ctx.body_sync(TP, m, r=r, v=v, mu=np.full(m, p["cb_Gmass"]), generation=903 + ntp_total)
    del r, v
    ah0 = np.zeros(3)
    for i in range(8):
        r2 = float(p["rh"][i] @ p["rh"][i])
        ah0 -= p["Gmass"][i] / (r2 * np.sqrt(r2)) * p["rh"][i]
    ctx.body_zero_accel(TP)
    ctx.tp_accel_int()
    flush = m * 152 < 2 * 126e6
    for _ in range(3):
        ctx.whm_tp_step(0.01, ah0, want_nfail=False)
    nsteps = 20 if ntp_total <= 10_000_000 else 8
    barrier()
    for _ in range(nsteps):
        if flush:
            ctx.flush_l2()
        ctx.timer_lap_begin()
        ctx.whm_tp_step(0.01, ah0, want_nfail=False)
        ctx.timer_lap_end()
    tot, cnt = ctx.timer_laps()
    ms = max_over_ranks(tot / cnt)
    nfail = ctx.body_get(TP, r=False, v=False, a=False, iflag=True)["iflag"]
    bad = max_over_ranks(float(np.count_nonzero(nfail)))
    barrier()
    gbs = 152.0 * ntp_total / (ms * 1e-3) / 1e9
    return {"ntp_total": ntp_total, "ranks": world, "ntp_per_rank": m, "ms_per_step": ms,
            "tp_steps_per_s": ntp_total / (ms * 1e-3), "drift_failures": int(bad),
            "partition": "block (coarray_distribute shape), planets replicated, no per-step communication",
            "timing": "mean of per-launch event pairs, max over ranks; L2 flush between launches"
                      + ("" if flush else " skipped (shard > 2 x L2)"),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak * world, "unit": "GB/s",
                         "frac": gbs / (hbm_peak * world), "bytes_per_tp": 152,
                         "note": "aggregate over the ranks against ranks x measured HBM copy bandwidth"}}


def symba_1e4_leg(ctx, hbm_peak, fp64_peak):
    """BASELINE.json configs[2]: SyMBA planetesimal disk, 1e4 fully interacting bodies -- flat (third-law) vs
    triangular (full-row) pl-pl kernel, and the pl-pl sort-and-sweep."""
    from swiftest_b200 import PL, LOOP_FLAT, LOOP_TRIANGULAR, workloads as W
    from swiftest_b200.context import FAM_PLPL, FAM_SWEEP
    n = 10000
    d = W.disk(n, seed=3031179)
    pairs = n * (n - 1) / 2.0
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=d["mu"], generation=41)
    ctx.enable_kernel_timing(True)
    res = {"npl": n, "pairs": pairs}
    acc = {}
    for name, var in (("flat", LOOP_FLAT), ("tri", LOOP_TRIANGULAR)):
        ms = []
        for it in range(8):
            ctx.flush_l2()
            ctx.body_zero_accel(PL)
            ctx.pl_accel_int(var, True)
            if it >= 3:
                ms.append(ctx.last_kernel_ms(FAM_PLPL))
        acc[name] = ctx.body_get(PL, r=False, v=False)["a"]
        t = float(np.mean(ms)) * 1e-3
        tf = FLOP_PER_PAIR_RAD * pairs / t / 1e12
        res[name] = {"ms": t * 1e3, "pairs_per_s": pairs / t,
                     "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak}}
    res["flat_vs_tri_max_rel_diff"] = float(np.max(np.abs(acc["flat"] - acc["tri"])) / np.max(np.abs(acc["tri"])))
    res["auto_picks"] = "flat" if res["flat"]["ms"] < res["tri"]["ms"] else "tri"
    ctx.pl_set_renc(0)
    ms = []
    for it in range(6):
        ctx.flush_l2()
        nenc = ctx.pl_encounter_check(d["dt"], fetch=False)
        if it >= 2:
            ms.append(ctx.last_kernel_ms(FAM_SWEEP))
    st = ctx.encounter_stats()
    b = SWEEP_BYTES
    bytes_alg = n * b["body"] + 2 * n * b["sort"] / 2 + 2 * n * 56.0 + st["nbox_total"] * b["cand"] + nenc * b["out"]
    t = float(np.mean(ms)) * 1e-3
    res["sweep_plpl"] = {"nenc": int(nenc), "nbox_total": int(st["nbox_total"]), "ms": t * 1e3, "algorithmic_bytes": bytes_alg,
                         "roofline": {"bound": "hbm", "achieved": bytes_alg / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                      "frac": bytes_alg / t / 1e9 / hbm_peak}}
    ctx.enable_kernel_timing(False)
    # the whole device-resident step of this configuration: helio_step_pl (two force evaluations, glue, drift; replayed as one
    # CUDA graph from the third step) + the pl-pl encounter sweep, host wall clock over 200 steps, nothing crossing PCIe but
    # the pair count
    from swiftest_b200 import LOOP_AUTO
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=np.full(n, W.GMSUN), generation=42)
    r0 = ctx.step_graph_replays()
    for rep in range(2):
        ctx.synchronize()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        for k in range(200):
            ctx.helio_step_pl(W.GMSUN, d["dt"], LOOP_AUTO, True, lfirst=(rep == 0 and k == 0), want_nfail=False)
        ctx.synchronize()
        t_step = (time.perf_counter() - t0) / 200
        lps = (ctx.launch_count() - n0) / 200
    t0 = time.perf_counter()
    for k in range(100):
        ctx.helio_step_pl(W.GMSUN, d["dt"], LOOP_AUTO, True, lfirst=False, want_nfail=False)
        ctx.pl_encounter_check(d["dt"], fetch=False)
    ctx.synchronize()
    t_both = (time.perf_counter() - t0) / 100
    res["whole_step"] = {"helio_step_pl_ms": t_step * 1e3, "kernels_per_step": lps,
                         "graph_replays": ctx.step_graph_replays() - r0,
                         "helio_step_pl_plus_sweep_ms": t_both * 1e3,
                         "pairs_per_s_whole_step": 2.0 * pairs / t_step,
                         "note": "r1 / first half of r2: 0.255 ms per step stream-ordered + 0.222 ms sweep"}
    return res


def side_legs(ctx, args, d, hbm_peak, peak_src):
    """Sweep on the same disk and the WHM tp configuration; timed outside the K headline steps."""
    from swiftest_b200 import PL, TP, LOOP_AUTO, workloads as W
    from swiftest_b200.context import FAM_PLTP, FAM_DRIFT, FAM_SWEEP
    ex = {}
    n = d["n"]
    ctx.enable_kernel_timing(True)
    ctx.body_put(PL, r=d["rh"], v=d["vh"])
    ctx.pl_set_renc(0)
    ms = []
    for it in range(6):
        ctx.flush_l2()
        nenc = ctx.pl_encounter_check(d["dt"], fetch=False)
        if it >= 2:
            ms.append(ctx.last_kernel_ms(FAM_SWEEP))
    st = ctx.encounter_stats()
    b = SWEEP_BYTES
    bytes_alg = n * b["body"] + 2 * n * b["sort"] / 2 + 2 * n * 56.0 + st["nbox_total"] * b["cand"] + nenc * b["out"]
    t = float(np.mean(ms)) * 1e-3
    ex["sweep_plpl"] = {"npl": n, "nenc": int(nenc), "nbox_total": int(st["nbox_total"]), "ms": t * 1e3,
                        "algorithmic_bytes": bytes_alg, "roofline": {"bound": "hbm", "achieved": bytes_alg / t / 1e9,
                                                                     "peak": hbm_peak, "unit": "GB/s",
                                                                     "frac": bytes_alg / t / 1e9 / hbm_peak,
                                                                     "peak_source": peak_src}}
    # WHM: Sun + 8 planets + ntp test particles (BASELINE.json configs[1])
    p = W.planets8_year_units()
    ntp = args.ntp
    tp = W.tp_cloud(ntp, seed=123)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=2)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, p["cb_Gmass"]), generation=3)
    kms, dms = [], []
    for it in range(8):
        ctx.flush_l2()
        ctx.body_zero_accel(TP)
        ctx.tp_accel_int()
        ctx.body_kick_velocity(TP, 0.01)
        ctx.body_drift(TP, 0.01, want_nfail=False)
        if it >= 3:
            kms.append(ctx.last_kernel_ms(FAM_PLTP))
            dms.append(ctx.last_kernel_ms(FAM_DRIFT))
    tk, td = float(np.mean(kms)) * 1e-3, float(np.mean(dms)) * 1e-3
    kick_bytes = ntp * 76.0 + 8 * 32.0
    ex["whm_tp"] = {"npl": 8, "ntp": ntp,
                    "pltp_kick": {"ms": tk * 1e3, "evals_per_s": 8.0 * ntp / tk, "gflops_algorithmic": FLOP_PER_TPEVAL * 8 * ntp / tk / 1e9,
                                  "roofline": {"bound": "hbm", "achieved": kick_bytes / tk / 1e9, "peak": hbm_peak,
                                               "unit": "GB/s", "frac": kick_bytes / tk / 1e9 / hbm_peak}},
                    "drift": {"ms": td * 1e3, "bodies_per_s": ntp / td,
                              "roofline": {"bound": "hbm", "achieved": DRIFT_BYTES_PER_BODY * ntp / td / 1e9,
                                           "peak": hbm_peak, "unit": "GB/s",
                                           "frac": DRIFT_BYTES_PER_BODY * ntp / td / 1e9 / hbm_peak}}}
    # next row: the fused WHM tp step (kick dt/2, drift, new ah, kick dt/2 in one pass: 152 B per tp)
    fms = []
    ah0 = np.zeros(3)
    for i in range(8):
        r2 = float(p["rh"][i] @ p["rh"][i])
        ah0 -= p["Gmass"][i] / (r2 * np.sqrt(r2)) * p["rh"][i]
    ctx.body_put(TP, r=tp["rh"], v=tp["vh"])
    ctx.body_zero_accel(TP)
    ctx.tp_accel_int()
    for it in range(8):
        ctx.flush_l2()
        ctx.whm_tp_step(0.01, ah0, want_nfail=False)
        if it >= 3:
            fms.append(ctx.last_kernel_ms(FAM_DRIFT))
    tf = float(np.mean(fms)) * 1e-3
    ex["whm_tp"]["fused_step"] = {"ms": tf * 1e3, "tp_steps_per_s": ntp / tf, "unfused_ms": (tk + td) * 1e3 + 2 * 0.02,
                                  "roofline": {"bound": "hbm", "achieved": 152.0 * ntp / tf / 1e9, "peak": hbm_peak,
                                               "unit": "GB/s", "frac": 152.0 * ntp / tf / 1e9 / hbm_peak,
                                               "bytes_per_tp": 152}}
    # BASELINE configs[1] as a run executes it: planets AND test particles resident, swcu_whm_step_pl (one launch for a
    # small system) + swcu_whm_tp_step(ah0 from the device) per step, nothing crosses PCIe; wall clock over 200 steps
    ctx.body_put(TP, r=tp["rh"], v=tp["vh"])
    ctx.body_put(PL, r=p["rh"], v=p["vh"])
    ctx.whm_tp_first_accel()
    nw = 200
    for rep in range(2):
        ctx.synchronize()
        n0 = ctx.launch_count()
        t0 = time.perf_counter()
        for k in range(nw):
            ctx.whm_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=(rep == 0 and k == 0), want_nfail=False)
            ctx.whm_tp_step(0.01, None, want_nfail=False)
        ctx.synchronize()
        tw = (time.perf_counter() - t0) / nw
        lps = (ctx.launch_count() - n0) / nw
    pms, tms = [], []
    for it in range(8):
        ctx.whm_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=False, want_nfail=False)
        pms.append(ctx.last_kernel_ms(FAM_DRIFT))
        ctx.whm_tp_step(0.01, None, want_nfail=False)
        tms.append(ctx.last_kernel_ms(FAM_DRIFT))
    ex["whm_tp"]["whole_step"] = {"ms": tw * 1e3, "tp_steps_per_s": ntp / tw, "launches_per_step": lps,
                                  "whm_step_pl_8_planets_ms": float(np.mean(pms[3:])), "tp_kernel_ms": float(np.mean(tms[3:])),
                                  "note": "planet step + fused tp step, device resident, no L2 flush (the tp arrays are 1.2 x L2)"}
    ctx.body_put(TP, r=tp["rh"], v=tp["vh"])
    ctx.pl_set_renc(0)

    def sweep_ms():
        ms = []
        for it in range(6):
            ctx.flush_l2()
            nenc = ctx.tp_encounter_check(0.01, fetch=False)
            if it >= 2:
                ms.append(ctx.last_kernel_ms(FAM_SWEEP))
        return float(np.mean(ms)), int(nenc), ctx.encounter_stats()

    nd0, nf0 = ctx.encounter_direct_count()
    t_direct, nenc, st = sweep_ms()
    nd1, nf1 = ctx.encounter_direct_count()
    os.environ["SWCU_PLTP_DIRECT_MAX"] = "0"   # the sort path of round 1 (2(npl+ntp)-key radix sort, gather, chunked sweep)
    t_sort, nenc_sort, st_sort = sweep_ms()
    del os.environ["SWCU_PLTP_DIRECT_MAX"]
    b_direct = ntp * 48.0 + 8 * 56.0 + 9.0 * nenc          # r, v of every particle once; planets; the pair list out
    nn = ntp + 8
    b_survey = nn * 56.0 + 2 * nn * 24.0 + 2 * nn * 56.0 + st_sort["nbox_total"] * 56.0 + 9.0 * nenc
    ex["sweep_pltp"] = {"npl": 8, "ntp": ntp, "nenc": nenc, "nbox_total": int(st["nbox_total"]), "ms": t_direct * 1.0,
                        "path": "sort-free pass over the particles (npl <= 128): %d of %d calls, %d fell back to the sort path"
                                % (nd1 - nd0 - (nf1 - nf0), nd1 - nd0, nf1 - nf0),
                        "sort_path_ms": t_sort, "sort_path_same_result": bool(nenc_sort == nenc and st_sort["nbox_total"] == st["nbox_total"]),
                        "algorithmic_bytes": b_survey,
                        "roofline": {"bound": "hbm", "achieved": b_survey / (t_direct * 1e-3) / 1e9, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": b_survey / (t_direct * 1e-3) / 1e9 / hbm_peak,
                                     "note": "SURVEY 8(d) bytes of the sort-and-sweep (body read + sort in/out + gather + candidate "
                                             "stream + pair list; the unit sweep_plpl is quoted in) over the time of the sort-free "
                                             "pass, read-back of the pair count included; the same formula gives the sort path "
                                             "%.0f GB/s" % (b_survey / (t_sort * 1e-3) / 1e9)},
                        "bytes_moved": {"bytes": b_direct, "GB/s": b_direct / (t_direct * 1e-3) / 1e9,
                                        "frac_of_hbm": b_direct / (t_direct * 1e-3) / 1e9 / hbm_peak,
                                        "note": "what the sort-free pass itself needs: 48 B per particle + 9 B per pair; it is bound "
                                                "by instruction issue, not HBM (profiles/r02_pltp_direct_ncu.txt)"}}
    ctx.enable_kernel_timing(False)
    return ex


def next_row_legs(ctx, args, d, hbm_peak, fp64_peak):
    """the device-resident democratic-heliocentric step and the energy sums."""
    from swiftest_b200 import PL, TP, LOOP_AUTO, workloads as W
    from swiftest_b200.context import FAM_DRIFT, FAM_PLPL
    ex = {}
    n = d["n"]
    GMcb = W.GMSUN
    # (1) helio_step_pl at the headline size: two kicks + drift + all O(N) glue per call, nothing crosses PCIe
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=d["Gmass"], radius=d["radius"], rhill=d["rhill"],
                  mu=np.full(n, GMcb), generation=11)
    ctx.helio_step_pl(GMcb, d["dt"], LOOP_AUTO, True, lfirst=True)
    l0 = ctx.launch_count()
    ctx.timer_start()
    reps = 3
    for _ in range(reps):
        ctx.helio_step_pl(GMcb, d["dt"], LOOP_AUTO, True, lfirst=False, want_nfail=False)
    ms = ctx.timer_stop() / reps
    ex["helio_step_pl"] = {"npl": n, "ms_per_step": ms, "kicks_per_step": 2, "launches_per_step": (ctx.launch_count() - l0) / reps,
                           "pair_interactions_per_s": 2 * n * (n - 1) / 2.0 / (ms * 1e-3)}
    # (2) potential energy + KE + L of the same disk through the host-pointer call (upload included in e2e_ms)
    mass = d["Gmass"] / GMcb
    ctx.enable_kernel_timing(True)
    ctx.util_get_potential_energy(n, None, GMcb, d["Gmass"], mass, d["rh"])
    kms, ems = [], []
    for _ in range(3):
        ctx.flush_l2()
        t0 = time.perf_counter()
        pe = ctx.util_get_potential_energy(n, None, GMcb, d["Gmass"], mass, d["rh"])
        ems.append((time.perf_counter() - t0) * 1e3)
        kms.append(ctx.last_kernel_ms(FAM_PLPL))
    ctx.enable_kernel_timing(False)
    tk = float(np.mean(kms)) * 1e-3
    pairs = n * (n - 1) / 2.0
    ex["potential_energy"] = {"npl": n, "pe": pe, "kernel_ms": tk * 1e3, "e2e_ms": float(np.mean(ems)), "pairs_per_s": pairs / tk,
                              "flop_per_pair": FLOP_PER_PE_PAIR,
                              "roofline": {"bound": "fp64", "achieved": pairs * FLOP_PER_PE_PAIR / tk / 1e12,
                                           "peak": fp64_peak, "unit": "TFLOP/s",
                                           "frac": pairs * FLOP_PER_PE_PAIR / tk / 1e12 / fp64_peak}}
    # (3) helio_step_tp as one kernel: Sun + 8 planets + ntp test particles
    p = W.planets8_year_units()
    ntp = args.ntp
    tp = W.tp_cloud(ntp, seed=123)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=np.full(8, p["cb_Gmass"]), generation=12)
    ctx.body_sync(TP, ntp, r=tp["rh"], v=tp["vh"], mu=np.full(ntp, p["cb_Gmass"]), generation=13)
    ctx.enable_kernel_timing(True)
    fms, pms = [], []
    for it in range(8):
        ctx.flush_l2()
        ctx.timer_start()
        ctx.helio_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=(it == 0), want_nfail=False)
        pms.append(ctx.timer_stop())
        ctx.helio_step_tp(p["cb_Gmass"], 0.01, lfirst=(it == 0), want_nfail=False)
        if it >= 3:
            fms.append(ctx.last_kernel_ms(FAM_DRIFT))
    ctx.enable_kernel_timing(False)
    tf = float(np.mean(fms)) * 1e-3
    # the whole resident HELIO step (planets in one launch + particles in one launch), host wall clock, no L2 flush
    for rep in range(2):
        ctx.synchronize()
        t0 = time.perf_counter()
        for k in range(300):
            ctx.helio_step_pl(p["cb_Gmass"], 0.01, LOOP_AUTO, True, lfirst=False, want_nfail=False)
            ctx.helio_step_tp(p["cb_Gmass"], 0.01, lfirst=False, want_nfail=False)
        ctx.synchronize()
        t_whole = (time.perf_counter() - t0) / 300
    ex["helio_step_tp"] = {"npl": 8, "ntp": ntp, "ms": tf * 1e3, "tp_steps_per_s": ntp / tf,
                           "helio_step_pl_8_planets_ms": float(np.mean(pms[3:])),
                           "whole_step_ms": t_whole * 1e3,
                           "note": "ms / helio_step_pl_8_planets_ms are per-launch times after an L2 flush (cold planets: the one-CTA "
                                   "planet kernel then pays DRAM latency on every dependent load); whole_step_ms is planets + "
                                   "particles back to back, device resident, no flush",
                           "roofline": {"bound": "hbm", "achieved": HELIO_TP_BYTES * ntp / tf / 1e9, "peak": hbm_peak,
                                        "unit": "GB/s", "frac": HELIO_TP_BYTES * ntp / tf / 1e9 / hbm_peak,
                                        "bytes_per_tp": HELIO_TP_BYTES}}
    return ex


def conservation_run(ctx, nsteps):
    """Energy / angular momentum of Sun + 8 planets over nsteps helio steps: GPU path vs CPU oracle (north star)."""
    from oracle import load
    from swiftest_b200 import workloads as W
    from tests.helio import GpuBackend, HelioSystem, OracleBackend
    p = W.planets8_year_units()
    a = HelioSystem(p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], OracleBackend(load()))
    b = HelioSystem(p["cb_Gmass"], p["Gmass"], p["rh"], p["vh"], p["radius"], GpuBackend(ctx))
    E0, L0 = a.energy_and_momentum()
    every = max(1, nsteps // 20)
    ts, dEa, dEb, dLa, dLb = [], [], [], [], []
    for k in range(1, nsteps + 1):
        a.step(0.01)
        b.step(0.01)
        if k % every == 0 or k == nsteps:
            Ea, La = a.energy_and_momentum()
            Eb, Lb = b.energy_and_momentum()
            ts.append(k * 0.01)
            dEa.append((Ea - E0) / abs(E0))
            dEb.append((Eb - E0) / abs(E0))
            dLa.append(float(np.linalg.norm(La - L0) / np.linalg.norm(L0)))
            dLb.append(float(np.linalg.norm(Lb - L0) / np.linalg.norm(L0)))
    return {"system": "Sun + 8 planets (reference fixture 8pl_0tp)", "steps": nsteps, "dt": 0.01, "years": nsteps * 0.01,
            "dE_cpu": dEa[-1], "dE_gpu": dEb[-1], "dL_cpu": dLa[-1], "dL_gpu": dLb[-1],
            "max_abs_dE_gpu": float(np.max(np.abs(dEb))), "max_abs_dE_cpu": float(np.max(np.abs(dEa))),
            "gpu_minus_cpu_dE": float(np.max(np.abs(np.array(dEa) - np.array(dEb)))),
            "gpu_minus_cpu_dL": float(np.max(np.abs(np.array(dLa) - np.array(dLb)))),
            "dE_slope_per_year_gpu": _slope_per_year(ts, np.abs(dEb)), "dL_slope_per_year_gpu": _slope_per_year(ts, dLb),
            "dE_slope_per_year_cpu": _slope_per_year(ts, np.abs(dEa)), "dL_slope_per_year_cpu": _slope_per_year(ts, dLa),
            "reference_limits": "tests/test_swiftest.py:119-121: |dE/E0| slope < 1e-8 /y, |dL/L0| slope < 1e-10 /y",
            "max_position_difference": float(np.max(np.abs(a.rh - b.rh)))}


def conservation_reference_test(ctx, nsteps, nout=1000, budget_s=50.0):
    """The reference's own quantitative test, at its own length (tests/test_swiftest.py:112-169: Sun + 8 planets, dt = 0.01 y,
    tstop = 1e4 y = 1e6 steps, 1000 outputs, linear fits of E_error and L_error against time: |slope| < 1e-8 /y and
    1e-10 /y, |dGM/GM| < 1e-14).  The reference runs it with SyMBA; without close encounters among the planets a SyMBA step IS
    helio_step_pl, and the run checks that no encounter occurs.  GPU: the planets stay resident, one launch per step
    (helio_step_pl_small_kernel); CPU: the oracle's C stepper (bit-identical to the interpreted Fortran).  Both states go
    through the same energy / momentum evaluation, so their difference is the difference of the trajectories.  Each arm stops
    after budget_s seconds of stepping (steps_done says where)."""
    import ctypes as C
    from oracle import load
    from swiftest_b200 import PL, LOOP_TRIANGULAR, workloads as W
    o = load()
    p = W.planets8_year_units()
    n, GMcb, Gm, rad, dt = 8, p["cb_Gmass"], np.ascontiguousarray(p["Gmass"]), np.ascontiguousarray(p["radius"]), 0.01
    mass = Gm / GMcb  # any constant G: only ratios enter dE/E0 and dL/L0
    every = max(1, nsteps // nout)

    def el(rh, vh):
        rb, vb, rbcb, vbcb = o.coord_h2b_pl(GMcb, Gm, rh, vh)
        e = o.get_energy_and_momentum(GMcb, 1.0, rbcb, vbcb, Gm, mass, rad, rb, vb, None, True)
        return e["te"], e["L_orbit"], e["GMtot"]

    def e_true(rh, vh):
        """The physical energy (central-body term with |rb_i - rb_cb|).  The reference's own sum uses |rb_i|
        (swiftest_util.f90:1377), so its E_error carries the Sun's barycentric wobble, ~3e-4, whatever the integrator does."""
        gt = GMcb + Gm.sum()
        vcb = -(Gm[:, None] * vh).sum(0) / gt
        vb = vh + vcb
        d = rh[:, None, :] - rh[None, :, :]
        iu = np.triu_indices(n, 1)
        return (0.5 * (Gm * (vb ** 2).sum(1)).sum() + 0.5 * GMcb * (vcb ** 2).sum() -
                (GMcb * Gm / np.linalg.norm(rh, axis=1)).sum() - (Gm[iu[0]] * Gm[iu[1]] / np.linalg.norm(d, axis=2)[iu]).sum())

    E0, L0, GM0 = el(p["rh"], p["vh"])
    Et0 = e_true(p["rh"], p["vh"])
    Etg = []
    # ---- GPU arm
    ctx.body_sync(PL, n, nplm=n, r=p["rh"], v=p["vh"], Gmass=Gm, radius=rad, rhill=p["rhill"], mu=np.full(n, GMcb),
                  generation=777001)
    ctx.pl_set_renc(0)
    step = ctx._L.swcu_helio_step_pl
    h, cGM, cdt = ctx._h, C.c_double(GMcb), C.c_double(dt)
    tg, Eg, Lg, nenc_seen, done_g = [], [], [], 0, 0
    t0 = time.perf_counter()
    l0 = ctx.launch_count()
    for k in range(1, nsteps + 1):
        rc = step(h, cGM, cdt, LOOP_TRIANGULAR, 1, 1 if k == 1 else 0, None)
        if rc != 0:
            ctx._ck(rc)  # raises with the library's message
        if k % every == 0 or k == nsteps:
            out = ctx.body_get(PL, a=False)
            E, L, GM = el(out["r"], out["v"])
            tg.append(k * dt), Eg.append((E - E0) / E0), Lg.append(float(np.linalg.norm(L - L0) / np.linalg.norm(L0)))
            Etg.append((e_true(out["r"], out["v"]) - Et0) / abs(Et0))
            if len(tg) % 50 == 0:
                nenc_seen += int(ctx.pl_encounter_check(dt, fetch=False))
            done_g = k
            if time.perf_counter() - t0 > budget_s:
                break
    t_gpu = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    # ---- CPU arm (the oracle's stepper, called without the Python wrapper's per-call allocations)
    st = {k2: np.ascontiguousarray(p[k2]).copy() for k2 in ("rh", "vh")}
    st.update(vb=np.zeros((n, 3)), ah=np.zeros((n, 3)), rbeg=np.zeros((n, 3)), rend=np.zeros((n, 3)), ptbeg=np.zeros(3),
              ptend=np.zeros(3), vbcb=np.zeros(3))
    lf, iflag = C.c_int32(1), np.zeros(n, np.int32)
    ptr = lambda a: a.ctypes.data  # noqa: E731
    args_c = (n, GMcb, ptr(Gm), ptr(rad), 0, None, C.byref(lf), dt, ptr(st["rh"]), ptr(st["vh"]), ptr(st["vb"]), ptr(st["ah"]),
              ptr(st["rbeg"]), ptr(st["rend"]), ptr(st["ptbeg"]), ptr(st["ptend"]), ptr(st["vbcb"]), ptr(iflag))
    fn = o.lib.swo_helio_step_pl
    tc, Ec, Lc, done_c = [], [], [], 0
    t0 = time.perf_counter()
    for k in range(1, done_g + 1):
        fn(*args_c)
        if k % every == 0 or k == nsteps:
            E, L, GMc = el(st["rh"], st["vh"])
            tc.append(k * dt), Ec.append((E - E0) / E0), Lc.append(float(np.linalg.norm(L - L0) / np.linalg.norm(L0)))
            done_c = k
            if time.perf_counter() - t0 > budget_s:
                break
    t_cpu = time.perf_counter() - t0
    m = min(len(tg), len(tc))
    Eg_a, Ec_a, Lg_a, Lc_a = np.array(Eg), np.array(Ec), np.array(Lg), np.array(Lc)
    res = {"test": "tests/test_swiftest.py:112-169 (test_conservation), Sun + 8 planets of the reference fixture 8pl_0tp, dt 0.01 y",
           "steps_requested": nsteps, "steps_done_gpu": done_g, "steps_done_cpu": done_c, "years_gpu": done_g * dt,
           "outputs": len(tg), "seconds_gpu": t_gpu, "seconds_cpu": t_cpu, "us_per_step_gpu": 1e6 * t_gpu / max(done_g, 1),
           "kernel_launches_per_step_gpu": launches / max(done_g, 1), "encounters_seen": nenc_seen,
           "E_slope_per_year_gpu": _slope_per_year(tg, Eg_a), "L_slope_per_year_gpu": _slope_per_year(tg, Lg_a),
           "E_slope_per_year_cpu": _slope_per_year(tc, Ec_a), "L_slope_per_year_cpu": _slope_per_year(tc, Lc_a),
           "limits": {"E_slope": 1e-8, "L_slope": 1e-10, "GM": 1e-14},
           "GM_error_final": float((GM - GM0) / GM0),
           "max_abs_E_error_gpu": float(np.max(np.abs(Eg_a))), "max_L_error_gpu": float(np.max(Lg_a)),
           "max_abs_true_energy_error_gpu": float(np.max(np.abs(Etg))),
           "true_energy_slope_per_year_gpu": _slope_per_year(tg, np.array(Etg)),
           "note": "E_error as the reference defines it (its potential energy takes |rb_i| for the central-body term, "
                   "swiftest_util.f90:1377: the Sun's wobble shows as a ~3e-4 oscillation); true_energy = the same states in "
                   "the physical energy",
           "gpu_minus_cpu_E_error_max": float(np.max(np.abs(Eg_a[:m] - Ec_a[:m]))) if m else None,
           "gpu_minus_cpu_L_error_max": float(np.max(np.abs(Lg_a[:m] - Lc_a[:m]))) if m else None}
    res["passes_reference_limits"] = bool(abs(res["E_slope_per_year_gpu"]) < 1e-8 and abs(res["L_slope_per_year_gpu"]) < 1e-10 and
                                          abs(res["GM_error_final"]) < 1e-14 and nenc_seen == 0)
    return res


def _slope_per_year(t, y):
    """least-squares slope of y(t)"""
    t, y = np.asarray(t, float), np.asarray(y, float)
    return float(np.polyfit(t, y, 1)[0]) if len(t) > 1 else 0.0


def conservation_disk_run(ctx, nsteps, n=10000):
    """SyMBA-style disk of 1e4 fully interacting bodies, democratic-heliocentric steps (helio_step.f90:37-78):
    the device-resident swcu_helio_step_pl against the CPU oracle's swo_helio_step_pl (full-row kick through its OpenMP
    row loop), energy / angular momentum sampled every nsteps/10 steps -- the GPU state through
    swcu_util_get_energy_and_momentum, the CPU state through the oracle's restatement of
    swiftest_util_get_energy_and_momentum_system (swiftest_util.f90:1222-1394)."""
    from oracle import load
    from swiftest_b200 import PL, LOOP_AUTO, workloads as W
    o = load(native=True)
    o.omp_set_threads()
    o.use_omp_kick(True)
    d = W.disk(n, seed=3031179)
    GMcb, dt = W.GMSUN, d["dt"]
    Gm, rad = d["Gmass"], d["radius"]
    mass = Gm / GMcb

    def energy_gpu(rh, vh):
        rb, vb, rbcb, vbcb = o.coord_h2b_pl(GMcb, Gm, rh, vh)
        e = ctx.util_get_energy_and_momentum(n, None, GMcb, 1.0, rbcb, vbcb, Gm, mass, rad, rb, vb, True)
        return e["te"], e["L_orbit"] + np.cross(rbcb, vbcb)

    def energy_cpu(rh, vh):
        rb, vb, rbcb, vbcb = o.coord_h2b_pl(GMcb, Gm, rh, vh)
        e = o.get_energy_and_momentum(GMcb, 1.0, rbcb, vbcb, Gm, mass, rad, rb, vb, None, True)
        return e["te"], e["L_orbit"] + np.cross(rbcb, vbcb)

    st = {"rh": d["rh"].copy(), "vh": d["vh"].copy(), "vb": np.zeros((n, 3)), "lfirst": True}
    ctx.body_sync(PL, n, nplm=n, r=d["rh"], v=d["vh"], Gmass=Gm, radius=rad, rhill=d["rhill"], mu=np.full(n, GMcb),
                  generation=51)
    E0g, L0g = energy_gpu(d["rh"], d["vh"])
    E0c, L0c = energy_cpu(d["rh"], d["vh"])
    every = max(1, nsteps // 10)
    ts, dEg, dEc, dLg, dLc, dpos = [], [], [], [], [], []
    t_gpu = t_cpu = 0.0
    nfail_gpu = nfail_cpu = 0
    for k in range(1, nsteps + 1):
        t0 = time.perf_counter()
        nfail_gpu += ctx.helio_step_pl(GMcb, dt, LOOP_AUTO, True, lfirst=(k == 1), want_nfail=(k % every == 0))
        t1 = time.perf_counter()
        nfail_cpu += int(np.count_nonzero(o.helio_step_pl(st, GMcb, Gm, rad, dt)))
        t_cpu += time.perf_counter() - t1
        t_gpu += t1 - t0
        if k % every == 0 or k == nsteps:
            g = ctx.body_get(PL, a=False)
            Eg, Lg = energy_gpu(g["r"], g["v"])
            Ec, Lc = energy_cpu(st["rh"], st["vh"])
            ts.append(k * dt)
            dEg.append((Eg - E0g) / abs(E0g))
            dEc.append((Ec - E0c) / abs(E0c))
            dLg.append(float(np.linalg.norm(Lg - L0g) / np.linalg.norm(L0g)))
            dLc.append(float(np.linalg.norm(Lc - L0c) / np.linalg.norm(L0c)))
            dpos.append(float(np.max(np.linalg.norm(g["r"] - st["rh"], axis=1) / np.linalg.norm(st["rh"], axis=1))))
    o.use_omp_kick(False)
    return {"npl": n, "steps": nsteps, "dt": dt, "years": nsteps * dt, "samples_at_years": ts,
            "dE_over_E0_gpu": dEg, "dE_over_E0_cpu": dEc, "dL_over_L0_gpu": dLg, "dL_over_L0_cpu": dLc,
            "gpu_minus_cpu_dE": float(np.max(np.abs(np.array(dEg) - np.array(dEc)))),
            "gpu_minus_cpu_dL": float(np.max(np.abs(np.array(dLg) - np.array(dLc)))),
            "max_rel_position_difference": dpos, "E0_gpu_vs_cpu_rel": abs(E0g - E0c) / abs(E0c),
            "dE_slope_per_year_gpu": _slope_per_year(ts, dEg), "dL_slope_per_year_gpu": _slope_per_year(ts, dLg),
            "reference_limits": "tests/test_swiftest.py:119-121: |dE/E0| slope < 1e-8 /y, |dL/L0| slope < 1e-10 /y "
                                "(set for Sun + 8 planets; the disk run has no close-encounter handling, like HELIO)",
            "drift_failures": {"gpu": int(nfail_gpu), "cpu": int(nfail_cpu)},
            "seconds": {"gpu_steps": t_gpu, "cpu_steps": t_cpu}}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
