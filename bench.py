#!/usr/bin/env python
"""bench.py -- headline benchmark of the QPFn2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one forward + one backward pass of the solver over one batch of synthetic problems.
At N=1 the default workload is BASELINE.json configs[1]: B=65536, N=8, diagonal-P QP, fp64,
P = diag_embed(rand), q = 2 rand - 1, grad_l = 2 rand - 1, eps=1e-7, max_iter=1000.  With --gpus N
(launched under torchrun, one rank per GPU) every rank owns its own shard of B problems -- problems
are independent, so there is no collective on the data path (weak scaling).

Timed region (``value``): inputs resident in HBM, K steps back to back over R rotating input sets
whose total footprint exceeds the 126 MB L2 (so no step finds its inputs cached from the previous
use), bracketed by barrier + synchronize, CUDA events on the launching stream, max over ranks.
``e2e``: the same metric through the host-buffer C-ABI entry point (dq_qp_solve_host /
dq_qcqp_solve_host), host->device and device->host copies inside the timed region.
``roofline``: the dominant kernel's algorithmic bytes / its measured duration, against the measured
HBM peak in MEASURED_PEAKS.json.  ``cpu_baseline``: the CPU oracle (or the reference-source build
in oracle/_ref when present) timed on this box's host cores on a bounded sample.

``--impl reference`` times the reference's CPU path on the same workload (bounded sample per step).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback

WORKLOADS = {
    # name: (kind, B, N, generator kwargs, description)
    "qp_diag_n8": ("qp", 65536, 8, dict(gen="qp_diag"), "B=65536 N=8 diagonal-P QP fp64 fwd+bwd (BASELINE configs[1])"),
    "qp_dense_n8": ("qp", 65536, 8, dict(gen="qp_dense"), "B=65536 N=8 dense-P QP fp64 fwd+bwd"),
    "qcqp_n24": ("qcqp", 65536, 24, dict(gen="qcqp_dense"), "B=65536 N=24 QCQP (12 contacts) fp64 fwd+bwd (BASELINE configs[2])"),
    "qcqp_n16": ("qcqp", 262144, 16, dict(gen="qcqp_dense"), "B=262144/GPU N=16 QCQP (8 contacts) fp64 fwd+bwd (BASELINE configs[4] shard)"),
    "qp_dense_n32": ("qp", 131072, 32, dict(gen="qp_dense"), "B=131072 N=32 dense-P QP fp64 fwd+bwd (BASELINE configs[3] QP half)"),
    "qcqp_n32": ("qcqp", 131072, 32, dict(gen="qcqp_dense"), "B=131072 N=32 QCQP fp64 fwd+bwd (BASELINE configs[3] QCQP half)"),
}
EPS, MAX_ITER, MU_PROX = 1e-7, 1000, 1e-7


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qp_diag_n8", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override B (per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def make_inputs(kind, gen, B, N, seed):
    from diffqcqp_b200 import workloads as wl

    if kind == "qp":
        P, q, g = getattr(wl, gen)(B, N, seed=seed)
        return dict(P=P, q=q, g=g)
    P, q, l_n, mu, g = getattr(wl, gen)(B, N, seed=seed)
    return dict(P=P, q=q, l_n=l_n, mu=mu, g=g)


def alg_bytes(kind, N):
    from diffqcqp_b200 import workloads as wl

    f = wl.qp_bytes if kind == "qp" else wl.qcqp_bytes
    return f(N, True, False), f(N, False, True)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# ------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """nvidia-smi polling during the timed region (B200_PROFILING.md's clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for f in self.samples:
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------- CPU baseline
def cpu_run(kind, inp, n, threads=0, use_ref=True):
    """One forward+backward over the first n problems on the host; returns (seconds, kind, cores)."""
    import numpy as np
    from oracle import pyoracle as orc

    a = {k: np.ascontiguousarray(v[:n].numpy()) for k, v in inp.items()}
    ref = None
    if use_ref:
        try:
            from oracle import pyref
            if pyref.available():
                ref = pyref
        except Exception:
            ref = None
    eng = ref if ref is not None else orc
    cores = eng.max_threads() if threads <= 0 else threads
    t0 = time.perf_counter()
    if kind == "qp":
        x = eng.qp_forward(a["P"], a["q"], None, EPS, MAX_ITER, MU_PROX, threads=threads)
        eng.qp_backward(a["P"], a["q"], x, a["g"], threads=threads)
    else:
        x = eng.qcqp_forward(a["P"], a["q"], a["l_n"], a["mu"], None, EPS, MAX_ITER, MU_PROX, threads=threads)
        eng.qcqp_backward(a["P"], a["q"], a["l_n"], a["mu"], x, a["g"], threads=threads)
    dt = time.perf_counter() - t0
    return dt, ("reference" if ref is not None else "port"), cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, B, N, kw, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    inp = make_inputs(kind, kw["gen"], B, N, seed=0)
    # bounded sample: calibrate to ~1.5 s per step so K+W steps stay within minutes
    n = min(B, 4096)
    dt, ckind, cores = cpu_run(kind, inp, n)
    rate = n / dt
    n = int(min(B, max(4096, rate * 1.5)))
    if args.cpu_sample:
        n = min(B, args.cpu_sample)
    for _ in range(args.warmup):
        cpu_run(kind, inp, n)
    ts = []
    for _ in range(args.steps):
        ts.append(cpu_run(kind, inp, n)[0])
    total = sum(ts)
    value = n * args.steps / total
    line = {
        "impl": "reference", "metric": "QP/QCQP fwd+bwd solves/sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "B": B, "N": N, "eps": EPS, "max_iter": MAX_ITER,
                   "step": f"fwd+bwd over the first {n} of {B} problems (bounded sample)"},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": ckind,
                         "sample": f"first {n} of {B} problems per step, OpenMP over problems"},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from diffqcqp_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kind, B, N, kw, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    nc = N // 2
    L = _lib.load()
    fb, bb = alg_bytes(kind, N)
    # one set = inputs (P, q, grad_l [, l_n, mu]) + outputs (x, grad_P, grad_q [, grad_l_n, grad_mu])
    in_bytes_per_set = 8 * B * (2 * N * N + 4 * N + (4 * nc if kind == "qcqp" else 0))
    R = max(2, int(-(-200e6 // in_bytes_per_set)) + 1)  # rotating sets: total footprint > 126 MB L2
    R = min(R, 16)
    sets = []
    host0 = None
    for r in range(R):
        inp = make_inputs(kind, kw["gen"], B, N, seed=1000 * rank + r)
        if r == 0:
            host0 = inp
        d = {k: v.to(dev) for k, v in inp.items()}
        d["x"] = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        d["gP"] = torch.empty((B, N, N), dtype=torch.float64, device=dev)
        d["gq"] = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        if kind == "qcqp":
            d["gl"] = torch.empty((B, nc, 1), dtype=torch.float64, device=dev)
            d["gm"] = torch.empty((B, nc, 1), dtype=torch.float64, device=dev)
        sets.append(d)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def fwd(d):
        if kind == "qp":
            rc = L.dq_qp_forward(d["P"].data_ptr(), d["q"].data_ptr(), None, d["x"].data_ptr(), None, B, N, EPS,
                                 MU_PROX, MAX_ITER, 1, sp)
        else:
            rc = L.dq_qcqp_forward(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                   None, d["x"].data_ptr(), None, B, N, EPS, MU_PROX, MAX_ITER, 1, sp)
        _lib.check(rc, "forward")

    def bwd(d):
        if kind == "qp":
            rc = L.dq_qp_backward(d["P"].data_ptr(), d["q"].data_ptr(), d["x"].data_ptr(), d["g"].data_ptr(),
                                  d["gP"].data_ptr(), d["gq"].data_ptr(), B, N, sp)
        else:
            rc = L.dq_qcqp_backward(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                    d["x"].data_ptr(), d["g"].data_ptr(), d["gP"].data_ptr(), d["gq"].data_ptr(),
                                    d["gl"].data_ptr(), d["gm"].data_ptr(), B, N, sp)
        _lib.check(rc, "backward")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up
    for w in range(max(args.warmup, 3)):
        d = sets[w % R]
        fwd(d); bwd(d)
    barrier()

    # ---- timed region: exactly K steps, events on the launching stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = L.dq_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for k in range(args.steps):
        d = sets[k % R]
        fwd(d); bwd(d)
    ev1.record(stream)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = int(L.dq_launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel durations (separate pass, same rotation, events around each launch)
    fwd_ms, bwd_ms = [], []
    for k in range(max(args.steps, 10)):
        d = sets[k % R]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream); fwd(d); e[1].record(stream); bwd(d); e[2].record(stream)
        torch.cuda.synchronize(dev)
        fwd_ms.append(e[0].elapsed_time(e[1])); bwd_ms.append(e[1].elapsed_time(e[2]))
    fwd_avg, bwd_avg = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)

    # ---- e2e through the host-buffer C-ABI entry point
    e2e = None
    if not args.no_e2e:
        h = {k: v.pin_memory() for k, v in host0.items()}
        hx = torch.empty((B, N, 1), dtype=torch.float64).pin_memory()
        hgP = torch.empty((B, N, N), dtype=torch.float64).pin_memory()
        hgq = torch.empty((B, N, 1), dtype=torch.float64).pin_memory()
        hgl = torch.empty((B, max(nc, 1), 1), dtype=torch.float64).pin_memory()
        hgm = torch.empty((B, max(nc, 1), 1), dtype=torch.float64).pin_memory()

        def e2e_step():
            if kind == "qp":
                rc = L.dq_qp_solve_host(h["P"].data_ptr(), h["q"].data_ptr(), hx.data_ptr(), h["g"].data_ptr(),
                                        hgP.data_ptr(), hgq.data_ptr(), B, N, EPS, MU_PROX, MAX_ITER, local_rank)
            else:
                rc = L.dq_qcqp_solve_host(h["P"].data_ptr(), h["q"].data_ptr(), h["l_n"].data_ptr(), h["mu"].data_ptr(),
                                          hx.data_ptr(), h["g"].data_ptr(), hgP.data_ptr(), hgq.data_ptr(),
                                          hgl.data_ptr(), hgm.data_ptr(), B, N, EPS, MU_PROX, MAX_ITER, local_rank)
            _lib.check(rc, "solve_host")

        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()  # returns when the outputs are in host memory
        t_e2e = time.perf_counter() - t0
        if distributed:
            t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_e2e = float(t.item())
        h2d = 8 * B * (N * N + 2 * N + (2 * nc if kind == "qcqp" else 0))
        d2h = 8 * B * (N * N + 2 * N + (2 * nc if kind == "qcqp" else 0))
        e2e = {"value": B * world * args.steps / t_e2e, "unit": "solves/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e / args.steps,
               "api": "dq_qp_solve_host" if kind == "qp" else "dq_qcqp_solve_host"}

    if distributed:
        t = torch.tensor([total_ms, fwd_avg, bwd_avg], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, fwd_avg, bwd_avg = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        dom = "admm_fwd_kernel" if fwd_avg >= bwd_avg else ("qp_bwd_kernel" if kind == "qp" else "qcqp_bwd_kernel")
        dom_ms = max(fwd_avg, bwd_avg)
        dom_bytes = (fb if fwd_avg >= bwd_avg else bb) * B
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_achieved = (fb + bb) * B / (total_ms / args.steps * 1e-3) / 1e9
        line = {
            "metric": "QP/QCQP fwd+bwd solves/sec", "value": B * world * args.steps / (total_ms * 1e-3),
            "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "name": args.workload, "B_per_gpu": B, "N": N, "eps": EPS,
                       "max_iter": MAX_ITER, "sharding": f"batch-sharded x{world}, no data-path collective",
                       "l2_policy": f"{R} rotating input sets, {R * in_bytes_per_set / 1e6:.0f} MB footprint > 126 MB L2"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "alg_bytes_per_solve": {"fwd": fb, "bwd": bb},
                         "kernel_ms": {"fwd": fwd_avg, "bwd": bwd_avg},
                         "step_achieved_gbs": step_achieved, "step_frac": step_achieved / peak},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world >= 1:
            n = args.cpu_sample or min(B, 16384)
            dt, ckind, cores = cpu_run(kind, host0, n)
            if not args.cpu_sample:  # re-size to about 10 s of CPU work
                n = int(min(B, max(n, n / dt * 10)))
                dt, ckind, cores = cpu_run(kind, host0, n)
            line["cpu_baseline"] = {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": ckind,
                                    "sample": f"one fwd+bwd over the first {n} of {B} problems, OpenMP over problems, {dt:.2f} s"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
