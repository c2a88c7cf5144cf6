"""bench.py's driver contract, as far as it can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--batch", "2048", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_ranks_exit_quietly():
    r = _run("--impl", "reference", "--batch", "512", "--steps", "1", "--warmup", "1", "--gpus", "2",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must still report (and use) every host thread
    r = _run("--impl", "reference", "--batch", "2048", "--steps", "1", "--warmup", "1", env={"OMP_NUM_THREADS": "1"})
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run("--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_both_arms_print_the_same_config():
    """`config` is built by one function for both arms (the driver compares them); the rotating-set rule keeps every set
    on one stream (R a multiple of S) with a footprint above the 126 MB L2."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    for wl_name, streams in (("qp_diag_n8", 4), ("qcqp_n24", 4), ("qcqp_n32", 3)):
        a = argparse.Namespace(workload=wl_name, batch=0, streams=streams)
        kind, B, N, _, _ = bench.WORKLOADS[wl_name]
        R, S, nbytes = bench.rotating_sets(kind, B, N, streams)
        assert R % S == 0 and R >= S >= 1 and (R * nbytes > 126e6 or R == 16)
        c = bench.common_config(a, 1)
        assert set(c) == {"workload", "name", "B_per_gpu", "N", "eps", "max_iter", "sharding", "l2_policy"}
    r = _run("--impl", "reference", "--batch", "1024", "--steps", "1", "--warmup", "1")
    d = json.loads(r.stdout.strip().splitlines()[-1])
    a = argparse.Namespace(workload="qp_diag_n8", batch=1024, streams=4)
    assert d["config"] == bench.common_config(a, 1)
