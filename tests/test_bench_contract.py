"""bench.py contract checks that run on CPU: the reference arm prints one JSON line with the required keys, and the
committed round-1 GPU line (profiles/r01_bench_final.json) carries roofline / cpu_baseline / e2e / clocks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_one_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--npl", "3000",
                                   "--steps", "2", "--warmup", "1"], text=True, cwd=ROOT)
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "pair-interactions/s" and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_ignores_torchrun_thread_limit_and_times_whole_steps():
    """torchrun exports OMP_NUM_THREADS=1 to its workers (the reference arm ran single-threaded at N > 1):
    the arm sets its thread count from the CPU affinity set, and nothing in its line is extrapolated."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--npl", "2000",
                                   "--steps", "1", "--warmup", "0", "--gpus", "2"], text=True, cwd=ROOT, env=env)
    d = json.loads([l for l in out.strip().splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "all 2000 rows" in d["cpu_baseline"]["sample"] and "nothing extrapolated" in d["cpu_baseline"]["sample"]
    assert d["n_gpus"] == 2


def test_both_arms_print_the_same_config():
    import bench
    ours = bench.common_config(100000)
    assert ours["workload"] == "symba_disk_npl100000_fully_interacting" and "model" not in ours
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": common_config(n)') == 2      # the reference line and our line


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--npl", "2000",
                                   "--steps", "1", "--warmup", "0"], text=True, cwd=ROOT, env=env)
    assert out.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    path = os.path.join(ROOT, "profiles", "r01_bench_final.json")
    d = json.loads(open(path).read().strip().splitlines()[-1])
    assert REQUIRED | {"roofline", "gpu_launches", "clocks"} <= set(d)
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["config"]["workload"].startswith("symba_disk_npl100000")
