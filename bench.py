#!/usr/bin/env python
"""bench.py -- headline benchmark of the force-and-drift hot path (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU path on the host cores
  torchrun ... bench.py --gpus N ...                        N>1: one rank per GPU, NCCL inside the library

Metric: FP64 pair-interactions/s of the SyMBA planetesimal disk, npl = 1e5 fully interacting massive bodies
(BASELINE.json configs[3]; fits one GPU).  One STEP = one pass of the hot path over the resident system:
   ah = 0 ; pl%accel_int (pl-pl gravity, radius-checked, all N(N-1)/2 pairs) ; vb += ah*dt ; pl%drift (Kepler drift)
   ; [N>1: third-law kernel: allreduce of the partial ah inside accel_int; full-row kernel: allgather of the drifted
   slices] .
`value` = N(N-1)/2 pairs per step / step time with everything resident in HBM (strong scaling: the system is fixed,
block pairs -- or rows -- are split over the ranks).  `e2e` = the same step with that step's positions and velocities copied from pinned
host memory and the accelerations/positions/velocities read back, every step.
The sort-and-sweep encounter check and the WHM test-particle configuration (8 planets + 1e6 tp: pl->tp gravity and tp
drift) are HBM-bound side legs of the same path; they are timed outside the K steps and reported under "extra" with
their own roofline figures.
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--conservation", type=int, default=0, help="run an n-step energy/L tracking run (extra)")
    return ap.parse_args()


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
def cpu_kick_sample(o, d, rows, reps=1):
    """Time `rows` rows of the full-row pl-pl loop (swiftest_kick.f90:219-240 shape, schedule(static), all threads).
    Returns seconds per call."""
    n = d["n"]
    acc = np.zeros((n, 3))
    best = 1e300
    for _ in range(reps):
        t0 = time.perf_counter()
        o.omp_kick_tri_rad_pl_rows(d["rh"], d["Gmass"], d["radius"], acc, n, 0, rows)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline(d, seconds):
    from oracle import load
    o = load(native=True)
    n = d["n"]
    threads = o.omp_threads()
    probe_rows = min(n, 8 * threads)
    t = cpu_kick_sample(o, d, probe_rows)
    rows = int(min(n, max(probe_rows, probe_rows * seconds / max(t, 1e-6))))
    rows = max(threads, rows - rows % threads)
    t = cpu_kick_sample(o, d, rows)
    t_full = t * n / rows
    pairs = n * (n - 1) / 2.0
    return {"value": pairs / t_full, "unit": "pair-interactions/s", "cores": threads, "kind": "port",
            "sample": f"{rows} of {n} rows of the full-row pl-pl loop (kick.f90:219-240 shape, OpenMP schedule(static), "
                      f"{threads} threads, gcc -O3 -march=x86-64-v3 strict IEEE), {t:.2f} s measured, scaled to all rows",
            "seconds_full_evaluation_est": t_full}


def run_reference(args):
    """--impl reference: the CPU path (oracle port; the Fortran reference cannot be built in this image) on the host
    cores.  A step = a bounded row sample of the pl-pl kick scaled to the whole system + the full serial drift."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import load
    from swiftest_b200 import workloads as W
    o = load(native=True)
    d = W.disk(args.npl, seed=3031179)
    n = d["n"]
    threads = o.omp_threads()
    rows = min(n, 8 * threads)
    t = cpu_kick_sample(o, d, rows)
    rows = int(min(n, max(rows, rows * 2.0 / max(t, 1e-6))))  # ~2 s of CPU work per step
    rows = max(threads, rows - rows % threads)
    pairs = n * (n - 1) / 2.0
    times = []
    for it in range(args.warmup + args.steps):
        tk = cpu_kick_sample(o, d, rows)
        t0 = time.perf_counter()
        x, v, fl = o.drift_all(d["mu"], d["rh"], d["vh"], d["dt"])  # serial, as in the reference (drift.f90:99-103)
        td = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(tk * n / rows + td)
    step = float(np.mean(times))
    val = pairs / step
    sample = (f"per step: {rows} of {n} rows of the full-row pl-pl loop on {threads} OpenMP threads scaled to all rows "
              f"+ serial Kepler drift of all {n} bodies")
    print(json.dumps({
        "impl": "reference", "metric": "FP64 pair-interactions/s (pl-pl N=1e5)", "value": val,
        "unit": "pair-interactions/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": f"symba_disk_npl{n}_fully_interacting", "npl": n,
                                        "loop": "triangular full-row, radius-checked", "host": "cpu"},
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

    def step():
        ctx.flush_l2()
        if use_p2p:
            ctx.pl_kick_drift_p2p(dt, True, want_nfail=False)
            return
        ctx.body_zero_accel(PL)
        ctx.pl_accel_int(variant, True)
        ctx.body_kick_velocity(PL, dt)
        ctx.body_drift(PL, dt, want_nfail=False)
        if world > 1 and variant == LOOP_TRIANGULAR:
            ctx.pl_allgather(with_v=True)

    def reset_state():
        ctx.body_put(PL, r=d["rh"], v=d["vh"])

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    barrier()
    ctx.enable_kernel_timing(2)  # an event pair per launch group, read after the loop: nothing stalls the stream
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms_total = ctx.timer_stop()
    barrier()
    fam_ms = {name: ctx.kernel_ms_accumulated(f) for name, f in
              (("gravity", FAM_PLPL), ("drift", FAM_DRIFT), ("collective", FAM_ALLGATHER))}
    ctx.enable_kernel_timing(0)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_total / args.steps)
    value = pairs / (ms_step * 1e-3)
    kick_ms_avg = max_over_ranks(fam_ms["gravity"][0] / max(1, fam_ms["gravity"][1]))
    breakdown = {k: (max_over_ranks(v[0] / args.steps) if v[1] else 0.0) for k, v in fam_ms.items()}
    fp64_peak = ctx.probe_fp64_peak()

    # ---------------- end-to-end: host buffers in, results out, every step ----------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_r, h_v = pin(d["rh"]), pin(d["vh"])
    out = {"r": pin(np.zeros((n, 3))), "v": pin(np.zeros((n, 3))), "a": pin(np.zeros((n, 3)))}

    def step_e2e():
        ctx.body_put(PL, r=h_r, v=h_v)
        step()
        ctx.body_get(PL, out=out)

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e = {"value": pairs / e2e_s, "unit": "pair-interactions/s", "h2d_bytes_per_step": int(2 * 3 * n * 8 * world),
           "d2h_bytes_per_step": int(3 * 3 * n * 8 * world), "ms_per_step": e2e_s * 1e3,
           "path": "swcu_body_put(r,v) -> zero/accel_int/kick/drift[/allgather] -> swcu_body_get(r,v,a), pinned host arrays"}

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
                "note": "kernel_ms brackets the gravity launch group (memset/max-radius, kernel, +allreduce at N>1, +ah update); "
                        "FP64 instructions hold the B200 issue port 2 cycles and nothing co-issues, so the bound is "
                        "2*N_fp64 + N_other issue cycles per pair, not the DFMA peak (profiles/r01_fp64_pipe.md)"}

    extra = {}
    if not args.no_extra and rank == 0 and world == 1:
        try:
            extra = side_legs(ctx, args, d, hbm_peak, peak_src)
            extra["next_rows"] = next_row_legs(ctx, args, d, hbm_peak, fp64_peak)
        except Exception as e:  # side legs are reported figures, never a dependency of the headline number
            extra["error"] = str(e)
    cpu = None
    if not args.no_extra and rank == 0 and world == 1:
        try:
            cpu = cpu_baseline(d, args.cpu_seconds)
        except Exception as e:  # the CPU baseline is a reported figure, never a dependency of the GPU number
            cpu = {"value": None, "unit": "pair-interactions/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    if args.conservation and rank == 0 and world == 1:
        extra["conservation"] = conservation_run(ctx, args.conservation)

    if world > 1 and not args.no_extra:
        try:  # a side leg must never cost the headline line
            extra["whm_tp_sharded"] = tp_sharded_leg(ctx, args, rank, world, barrier, max_over_ranks, hbm_peak)
        except Exception as e:
            extra["whm_tp_sharded"] = {"error": str(e)}

    if rank == 0:
        line = {
            "metric": "FP64 pair-interactions/s (pl-pl N=1e5)", "value": value, "unit": "pair-interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"symba_disk_npl{n}_fully_interacting", "npl": n, "nplm": n,
                       "loop": "triangular full-row" if variant == LOOP_TRIANGULAR else "flat third-law",
                       "lclose": True,
                       "step": "zero_accel+accel_int+kick_velocity+drift" +
                               ((("[reduce-scatter+kick+drift+allgather fused over NVLink peer memory]" if use_p2p
                                  else "+ncclAllReduce(ah)") if variant == LOOP_FLAT else "+allgather(r,v)") if world > 1 else ""),
                       "sharding": (f"block-pair runs over {world} rank(s)" if variant == LOOP_FLAT
                                    else f"i-slices over {world} rank(s), allgather of drifted r,v"),
                       "collective": ("p2p-fused" if use_p2p else ("nccl" if world > 1 else "none")),
                       "l2": "flushed between steps (160 MiB write = 1.33 x L2, inside the timed region)",
                       "seed": 3031179},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "breakdown_ms_per_step": breakdown,
            "peaks": {"fp64_tflops_measured": fp64_peak, "hbm_gbs": hbm_peak, "hbm_source": peak_src},
            "extra": extra}
        print(json.dumps(line))
    if dist:
        dist.barrier()
        if use_p2p:
            ctx.p2p_close()
        ctx.comm_finalize()
        dist.destroy_process_group()
    ctx.close()


def tp_sharded_leg(ctx, args, rank, world, barrier, max_over_ranks, hbm_peak):
    """WHM test particles over several GPUs: block partition like swiftest_coarray_distribute_system
    (swiftest_coarray.f90:705-711), planets replicated, no per-step communication.  Fused tp step, all ranks."""
    from swiftest_b200 import PL, TP, shard, workloads as W
    p = W.planets8_year_units()
    ntp = args.ntp
    tp = W.tp_cloud(ntp, seed=123)
    t0, t1 = shard.tp_block_partition(ntp, world, rank)
    ctx.body_sync(PL, 8, nplm=8, r=p["rh"], v=p["vh"], Gmass=p["Gmass"], radius=p["radius"], rhill=p["rhill"],
                  mu=p["cb_Gmass"] + p["Gmass"], generation=902)
    ctx.body_sync(TP, t1 - t0, r=tp["rh"][t0:t1], v=tp["vh"][t0:t1], mu=np.full(t1 - t0, p["cb_Gmass"]), generation=903)
    ah0 = np.zeros(3)
    for i in range(8):
        r2 = float(p["rh"][i] @ p["rh"][i])
        ah0 -= p["Gmass"][i] / (r2 * np.sqrt(r2)) * p["rh"][i]
    ctx.body_zero_accel(TP)
    ctx.tp_accel_int()
    for _ in range(3):
        ctx.whm_tp_step(0.01, ah0, want_nfail=False)
    nsteps = 20
    barrier()
    ctx.timer_start()
    for _ in range(nsteps):
        ctx.flush_l2()
        ctx.whm_tp_step(0.01, ah0, want_nfail=False)
    ms = max_over_ranks(ctx.timer_stop() / nsteps)
    barrier()
    return {"ntp_total": ntp, "ranks": world, "ms_per_step": ms, "tp_steps_per_s": ntp / (ms * 1e-3),
            "partition": "block (coarray_distribute shape), planets replicated, no per-step communication",
            "note": "includes the 160 MiB L2 flush between steps"}


def side_legs(ctx, args, d, hbm_peak, peak_src):
    """Sweep on the same disk and the WHM tp configuration; timed outside the K headline steps."""
    from swiftest_b200 import PL, TP, workloads as W
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
    ms = []
    ctx.body_put(TP, r=tp["rh"], v=tp["vh"])
    ctx.pl_set_renc(0)
    for it in range(5):
        ctx.flush_l2()
        nenc = ctx.tp_encounter_check(0.01, fetch=False)
        if it >= 2:
            ms.append(ctx.last_kernel_ms(FAM_SWEEP))
    st = ctx.encounter_stats()
    ex["sweep_pltp"] = {"npl": 8, "ntp": ntp, "nenc": int(nenc), "nbox_total": int(st["nbox_total"]),
                        "ms": float(np.mean(ms))}
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
    ex["helio_step_tp"] = {"npl": 8, "ntp": ntp, "ms": tf * 1e3, "tp_steps_per_s": ntp / tf,
                           "helio_step_pl_8_planets_ms": float(np.mean(pms[3:])),
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
    for _ in range(nsteps):
        a.step(0.01)
        b.step(0.01)
    Ea, La = a.energy_and_momentum()
    Eb, Lb = b.energy_and_momentum()
    return {"steps": nsteps, "dt": 0.01, "dE_cpu": (Ea - E0) / abs(E0), "dE_gpu": (Eb - E0) / abs(E0),
            "dL_cpu": float(np.linalg.norm(La - L0) / np.linalg.norm(L0)),
            "dL_gpu": float(np.linalg.norm(Lb - L0) / np.linalg.norm(L0)),
            "max_position_difference": float(np.max(np.abs(a.rh - b.rh)))}


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
