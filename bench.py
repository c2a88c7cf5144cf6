#!/usr/bin/env python
"""bench.py -- headline benchmark of the QPFn2 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one forward + one backward pass of the solver over one batch of synthetic problems.
At N=1 the default workload is BASELINE.json configs[1]: B=65536, N=8, diagonal-P QP, fp64,
P = diag_embed(rand), q = 2 rand - 1, grad_l = 2 rand - 1, eps=1e-7, max_iter=1000.  With --gpus N
(launched under torchrun, one rank per GPU) every rank owns its own shard of B problems -- problems
are independent, so there is no collective on the data path (weak scaling).

Timed region (``value``): inputs resident in HBM, K steps over R rotating input sets whose total
footprint exceeds the 126 MB L2 (so no step finds its inputs cached from the previous use), issued
round-robin on ``--streams`` CUDA streams (default 4: fwd -> bwd of one batch stay ordered, independent
batches overlap), bracketed by barrier + synchronize, CUDA events on the launching stream (the side
streams fork from and join into it), max over ranks.  ``detail.single_stream_ms_per_step`` is the same
loop on one stream.  ``config`` is identical for both arms (common_config).
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
    "qcqp_n8": ("qcqp", 65536, 8, dict(gen="qcqp_dense"), "B=65536 N=8 dense-P QCQP (4 contacts) fp64 fwd+bwd (the QCQP half of BASELINE's metric)"),
    "qcqp_diag_n8": ("qcqp", 65536, 8, dict(gen="qcqp_diag"), "B=65536 N=8 diagonal-P QCQP (4 contacts) fp64 fwd+bwd"),
    "qcqp_n24": ("qcqp", 65536, 24, dict(gen="qcqp_dense"), "B=65536 N=24 QCQP (12 contacts) fp64 fwd+bwd (BASELINE configs[2])"),
    "qcqp_n16": ("qcqp", 262144, 16, dict(gen="qcqp_dense"), "B=262144/GPU N=16 QCQP (8 contacts) fp64 fwd+bwd (BASELINE configs[4] shard)"),
    "qp_dense_n32": ("qp", 131072, 32, dict(gen="qp_dense"), "B=131072 N=32 dense-P QP fp64 fwd+bwd (BASELINE configs[3] QP half)"),
    "qcqp_n32": ("qcqp", 131072, 32, dict(gen="qcqp_dense"), "B=131072 N=32 QCQP fp64 fwd+bwd (BASELINE configs[3] QCQP half)"),
}
EPS, MAX_ITER, MU_PROX = 1e-7, 1000, 1e-7


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--streams", type=int, default=4,
                    help="CUDA streams the timed steps are issued on round-robin (independent batches overlap)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qp_diag_n8", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override B (per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU-baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="(N=1, default workload) skip the short runs of BASELINE configs[2] and [3]")
    ap.add_argument("--no-cfg5", action="store_true", help="(N>1) skip the BASELINE configs[4] section (B=262144/GPU N=16 QCQP, "
                                                          "shard-resident and scatter-inclusive)")
    ap.add_argument("--scatter", action="store_true",
                    help="(N>1) also time the scatter-inclusive step: rank 0 holds all N*B problems, NCCL scatter of "
                         "(P,q[,l_n,mu],grad_l), solve, NCCL gather of x* and grad_q")
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


def rotating_sets(kind, B, N, streams):
    """R input/output sets the GPU arm rotates through: total input footprint > 126 MB L2, at least one per stream, a
    multiple of the stream count (a set is then always used on the same stream: no two streams ever share buffers)."""
    nc = N // 2
    in_bytes_per_set = 8 * B * (2 * N * N + 4 * N + (4 * nc if kind == "qcqp" else 0))
    R = max(2, int(-(-200e6 // in_bytes_per_set)) + 1)
    S = max(1, min(streams, 16))
    R = min(max(R, S), 16)
    S = min(S, R)
    R = -(-R // S) * S
    return R, S, in_bytes_per_set


def common_config(args, world):
    """The part of `config` both arms (--impl b200 / reference) print identically: what is solved, how it is sharded and
    how the GPU arm keeps its inputs out of L2."""
    kind, B, N, kw, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    R, S, in_bytes = rotating_sets(kind, B, N, args.streams)
    return {"workload": desc, "name": args.workload, "B_per_gpu": B, "N": N, "eps": EPS, "max_iter": MAX_ITER,
            "sharding": f"batch-sharded x{world}, no data-path collective",
            "l2_policy": f"GPU arm: {R} rotating input sets, {R * in_bytes / 1e6:.0f} MB footprint > 126 MB L2; "
                         "CPU arm: a bounded sample of the same batch per step"}


FP64_PEAK_WARP_INST_PER_CLK_PER_SM = 59.6 / 32  # measured: 59.6 DFMA lane-ops/clk/SM (profiles/r01_micro_fp64_latency.txt, r02_micro_mufu64.txt)


def fp64_roofline(workload, kernel, ms, sm_mhz, sms=148):
    """FP64 side of the roofline for one kernel launch of `ms` milliseconds, from the instruction counts an ncu pass over
    the same command recorded in profiles/fp64_ops.json (thread-level DADD/DMUL/DFMA and FP64-pipe warp instructions per
    launch).  peak = measured 59.6 FP64 lane-ops/clk/SM x SMs x the SM clock sampled during the timed region."""
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_ops.json")) as fh:
            ops = json.load(fh).get(workload, {}).get(kernel)
    except Exception:
        ops = None
    if not ops or not ms:
        return None
    clk = (sm_mhz or 1965.0) * 1e6
    lane_ops = ops["dadd"] + ops["dmul"] + ops["dfma"]
    flops = ops["dadd"] + ops["dmul"] + 2 * ops["dfma"]
    peak_lane = 32 * FP64_PEAK_WARP_INST_PER_CLK_PER_SM * sms * clk
    peak_warp = FP64_PEAK_WARP_INST_PER_CLK_PER_SM * sms * clk
    t = ms * 1e-3
    return {"unit": "TFLOP/s (FMA = 2)", "achieved": flops / t / 1e12, "peak": 2 * peak_lane / 1e12,
            "frac_lane_ops": lane_ops / t / peak_lane,          # useful FP64 lane-operations against the pipe's lane rate
            "frac_pipe": ops["fp64_warp_inst"] / t / peak_warp,  # FP64-pipe issue slots used (partially filled warps count whole)
            "fp64_lane_ops_per_launch": lane_ops, "fp64_warp_inst_per_launch": ops["fp64_warp_inst"],
            "warp_inst_per_launch": ops.get("warp_inst"), "source": "profiles/fp64_ops.json (ncu counters of the same command)"}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# ------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock / throttle-reason polling during the timed region (B200_PROFILING.md's clocks line).
    NVML when importable (1 ms period), else nvidia-smi (one query takes tens of ms)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    @staticmethod
    def _physical_index(index):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        flags = []
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("sw_power_cap", 0x4)):
            flags.append("Active" if (r & bit) else "Not Active")
        return [str(sm), str(self.max_sm), "0"] + flags

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    self.samples.append(self._sample_nvml())
                    self.stop_flag.wait(0.001)
                    continue
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------- CPU baseline
CPU_BUILD_NOTE = {
    "reference": "the reference's own qcqplib/Solver.cpp, unmodified, compiled with -O3 -fopenmp against oracle/eigen_standin "
                 "(a plain-loop stand-in for the Eigen API it uses: real Eigen is absent here, so no packetised Eigen kernels)",
    "port": "oracle/dq_oracle.c, the line-by-line C restatement of Solver.cpp (-O3 -fopenmp)",
}


def cpu_engine(use_ref=True):
    """The CPU implementation timed as the baseline: the reference's own Solver.cpp build (oracle/_ref,
    kind "reference") when that library travelled with the repo, else the oracle restatement ("port")."""
    from oracle import pyoracle as orc

    if use_ref:
        try:
            from oracle import pyref
            if pyref.available():
                pyref.lib()
                return pyref, "reference"
        except Exception:
            pass
    return orc, "port"


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs ask for these explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_pass(eng, kind, a, threads=0):
    """One forward + backward over the numpy batch `a` on the host cores; returns seconds."""
    threads = threads or host_threads()
    t0 = time.perf_counter()
    if kind == "qp":
        x = eng.qp_forward(a["P"], a["q"], None, EPS, MAX_ITER, MU_PROX, threads=threads)
        eng.qp_backward(a["P"], a["q"], x, a["g"], threads=threads)
    else:
        x = eng.qcqp_forward(a["P"], a["q"], a["l_n"], a["mu"], None, EPS, MAX_ITER, MU_PROX, threads=threads)
        eng.qcqp_backward(a["P"], a["q"], a["l_n"], a["mu"], x, a["g"], threads=threads)
    return time.perf_counter() - t0


def as_shipped_loop(kind, inp, eng, n=2000):
    """The reference's shipped call shape (qcqp.py:29-31, :45-47): a Python loop over the batch calling the
    per-problem binding, one core (the GIL is held across the call).  Small sample, reported for context."""
    import numpy as np

    n = min(n, inp["P"].shape[0])
    a = {k: np.ascontiguousarray(v[:n].numpy()) for k, v in inp.items()}
    N = a["P"].shape[1]
    ws = np.zeros(N)
    t0 = time.perf_counter()
    if kind == "qp":
        for i in range(n):
            x = eng.solveQP(a["P"][i], a["q"][i], ws, EPS, MU_PROX, MAX_ITER, True)
            eng.solveDerivativesQP(a["P"][i], a["q"][i], x, a["g"][i])
    else:
        for i in range(n):
            x = eng.solveQCQP(a["P"][i], a["q"][i], a["l_n"][i], a["mu"][i], ws, EPS, MU_PROX, MAX_ITER, True)
            eng.solveDerivativesQCQP(a["P"][i], a["q"][i], a["l_n"][i], a["mu"][i], x, a["g"][i])
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "solves/s", "cores": 1,
            "sample": f"Python per-item loop over the first {n} problems (the shipped shape of qcqp.py), ctypes binding"}


def cpu_sample(kind, inp, B, budget_s, sample=0):
    """Pick a sample of the workload worth about `budget_s` seconds of CPU work: the first n problems, or the
    whole batch repeated `reps` times when one pass is shorter than the budget."""
    import numpy as np

    eng, ckind = cpu_engine()
    n0 = min(B, 4096)
    a0 = {k: np.ascontiguousarray(v[:n0].numpy()) for k, v in inp.items()}
    cpu_pass(eng, kind, a0)  # page in the library / thread pool
    rate = n0 / cpu_pass(eng, kind, a0)
    n = int(min(B, max(n0, rate * budget_s))) if not sample else min(B, sample)
    reps = max(1, int(rate * budget_s / n)) if not sample else 1
    a = {k: np.ascontiguousarray(v[:n].numpy()) for k, v in inp.items()}
    return eng, ckind, a, n, reps


def run_reference_arm(args):
    """The reference's CPU implementation of the path on this box's host cores, all threads, same workload;
    each step is a bounded sample sized so that steps + warmup finish within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, B, N, kw, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    inp = make_inputs(kind, kw["gen"], B, N, seed=0)
    budget = min(1.0, 120.0 / max(1, args.steps + args.warmup))  # seconds of CPU work per step
    eng, ckind, a, n, reps = cpu_sample(kind, inp, B, budget, args.cpu_sample)
    cores = host_threads()
    for _ in range(min(args.warmup, 3)):
        cpu_pass(eng, kind, a)
    ts = []
    for _ in range(args.steps):
        ts.append(sum(cpu_pass(eng, kind, a) for _ in range(reps)))
    total = sum(ts)
    value = n * reps * args.steps / total
    sample = (f"each step = {reps} x (fwd+bwd over the first {n} of {B} problems), OpenMP over problems, "
              f"{'oracle/_ref: the reference Solver.cpp build' if ckind == 'reference' else 'oracle port'}")
    line = {
        "impl": "reference", "metric": "QP/QCQP fwd+bwd solves/sec", "value": value, "unit": "solves/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args, args.gpus),
        "detail": {"step": sample},
        "cpu_baseline": {"value": value, "unit": "solves/s", "cores": cores, "kind": ckind, "sample": sample,
                         "build": CPU_BUILD_NOTE[ckind]},
        "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- BASELINE configs[2] and [3] at N = 1
def time_workload(L, dev, kind, gen, B, N, steps, warmup, nsets=2, seed0=9000):
    """fwd + bwd of one batch shape on one stream: ms per step (CUDA events), `nsets` rotating input sets."""
    import torch
    from diffqcqp_b200 import _lib

    nc = N // 2
    sets, host0 = [], None
    for r in range(nsets):
        inp = make_inputs(kind, gen, B, N, seed=seed0 + r)
        host0 = host0 or inp
        d = {k: v.to(dev) for k, v in inp.items()}
        d.update(x=torch.empty((B, N, 1), dtype=torch.float64, device=dev), st=torch.empty((B, N, 1), dtype=torch.float64, device=dev),
                 gP=torch.empty((B, N, N), dtype=torch.float64, device=dev), gq=torch.empty((B, N, 1), dtype=torch.float64, device=dev))
        if kind == "qcqp":
            d.update(gl=torch.empty((B, nc, 1), dtype=torch.float64, device=dev), gm=torch.empty((B, nc, 1), dtype=torch.float64, device=dev))
        sets.append(d)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def step(d):
        if kind == "qp":
            _lib.check(L.dq_qp_forward_ex(d["P"].data_ptr(), d["q"].data_ptr(), None, d["x"].data_ptr(), None, d["st"].data_ptr(),
                                          B, N, EPS, MU_PROX, MAX_ITER, 1, sp), "forward")
            _lib.check(L.dq_qp_backward_ex(d["P"].data_ptr(), d["q"].data_ptr(), d["x"].data_ptr(), d["g"].data_ptr(), d["st"].data_ptr(),
                                           d["gP"].data_ptr(), d["gq"].data_ptr(), B, N, sp), "backward")
        else:
            _lib.check(L.dq_qcqp_forward_ex(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(), None,
                                            d["x"].data_ptr(), None, d["st"].data_ptr(), B, N, EPS, MU_PROX, MAX_ITER, 1, sp), "forward")
            _lib.check(L.dq_qcqp_backward_ex2(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                              d["x"].data_ptr(), d["g"].data_ptr(), d["st"].data_ptr(), d["gP"].data_ptr(),
                                              d["gq"].data_ptr(), d["gl"].data_ptr(), d["gm"].data_ptr(), None, None, B, N, sp), "backward")

    for k in range(warmup):
        step(sets[k % nsets])
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(steps):
        step(sets[k % nsets])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    del sets
    torch.cuda.empty_cache()
    return ms, host0


def run_other_configs(L, dev, cpu_budget_s=2.0):
    """BASELINE configs[2] (B=65536 N=24 QCQP) and configs[3] (B=262144 N=32 mixed: 131072 QPs + 131072 QCQPs) on one B200:
    a few steps each, device-resident, with a bounded sample of the reference's CPU path beside each."""
    peak, _ = hbm_peak()
    out = {}
    parts = {}
    for name in ("qcqp_n24", "qp_dense_n32", "qcqp_n32"):
        kind, B, N, kw, desc = WORKLOADS[name]
        ms, host0 = time_workload(L, dev, kind, kw["gen"], B, N, steps=6, warmup=3)
        fb, bb = alg_bytes(kind, N)
        ent = {"workload": desc, "B": B, "N": N, "ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "solves/s",
               "step_frac_hbm": (fb + bb) * B / (ms * 1e-3) / 1e9 / peak, "timing": "6 steps after 3 warm-ups, one stream, 2 rotating input sets"}
        eng, ckind, a, n, reps = cpu_sample(kind, host0, B, cpu_budget_s)
        dt = sum(cpu_pass(eng, kind, a) for _ in range(reps))
        ent["cpu_baseline"] = {"value": n * reps / dt, "unit": "solves/s", "cores": host_threads(), "kind": ckind,
                               "sample": f"{reps} x (fwd+bwd over the first {n} of {B} problems), {dt:.1f} s"}
        parts[name] = ent
        del host0
    out["configs[2] B=65536 N=24 QCQP"] = parts["qcqp_n24"]
    t4 = parts["qp_dense_n32"]["ms_per_step"] + parts["qcqp_n32"]["ms_per_step"]
    b4 = parts["qp_dense_n32"]["B"] + parts["qcqp_n32"]["B"]
    cpu4 = b4 / (parts["qp_dense_n32"]["B"] / parts["qp_dense_n32"]["cpu_baseline"]["value"]
                 + parts["qcqp_n32"]["B"] / parts["qcqp_n32"]["cpu_baseline"]["value"])
    out["configs[3] B=262144 N=32 mixed QP/QCQP"] = {
        "ms_per_step": t4, "value": b4 / (t4 * 1e-3), "unit": "solves/s", "B": b4,
        "cpu_baseline": {"value": cpu4, "unit": "solves/s", "cores": host_threads(), "kind": parts["qcqp_n32"]["cpu_baseline"]["kind"]},
        "note": "warm_start is dead in the reference (Solver.cpp:70 -> :80): the warm-started solve is the same computation",
        "qp_half": parts["qp_dense_n32"], "qcqp_half": parts["qcqp_n32"]}
    return out


# ------------------------------------------------------------------------------- BASELINE configs[4] at N > 1
def run_cfg5(L, dev, rank, world, barrier, steps=6, scatter_steps=3):
    import torch
    import torch.distributed as dist
    from diffqcqp_b200 import _lib, shard, workloads as wl

    B5, N5 = 262144, 16
    nc = N5 // 2
    keys = ["P", "q", "l_n", "mu", "g"]
    sets = []
    for r in range(2):  # two rotating sets: 2 x 0.6 GB of inputs, far beyond L2
        d = dict(zip(keys, (t.to(dev) for t in wl.qcqp_dense(B5, N5, seed=7000 + 10 * rank + r))))
        d.update(x=torch.empty((B5, N5, 1), dtype=torch.float64, device=dev), st=torch.empty((B5, N5, 1), dtype=torch.float64, device=dev),
                 gP=torch.empty((B5, N5, N5), dtype=torch.float64, device=dev), gq=torch.empty((B5, N5, 1), dtype=torch.float64, device=dev),
                 gl=torch.empty((B5, nc, 1), dtype=torch.float64, device=dev), gm=torch.empty((B5, nc, 1), dtype=torch.float64, device=dev))
        sets.append(d)
    stream = torch.cuda.current_stream(dev)

    def step(d):
        sp = stream.cuda_stream
        _lib.check(L.dq_qcqp_forward_ex(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(), None,
                                        d["x"].data_ptr(), None, d["st"].data_ptr(), B5, N5, EPS, MU_PROX, MAX_ITER, 1, sp), "forward")
        _lib.check(L.dq_qcqp_backward_ex2(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                          d["x"].data_ptr(), d["g"].data_ptr(), d["st"].data_ptr(), d["gP"].data_ptr(),
                                          d["gq"].data_ptr(), d["gl"].data_ptr(), d["gm"].data_ptr(), None, None, B5, N5, sp), "backward")

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(n):
            fn(k)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for k in range(2):
        step(sets[k % 2])
    resident_ms = timed(lambda k: step(sets[k % 2]), steps)

    # scatter-inclusive: rank 0 holds all world * B5 problems
    full = [torch.cat([sets[0][k]] * world, 0) for k in keys] if rank == 0 else None
    trailing = [tuple(sets[0][k].shape[1:]) for k in keys]
    d = sets[1]
    t_sc = [0.0]

    def scatter_step(_k):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        parts = shard.scatter_batch(full, B5 * world, src=0, device=dev, trailing=trailing)
        a1.record(stream)
        loc = dict(zip(keys, parts))
        loc.update(x=d["x"], st=d["st"], gP=d["gP"], gq=d["gq"], gl=d["gl"], gm=d["gm"])
        step(loc)
        shard.gather_batch(loc["x"], B5 * world, dst=0)
        shard.gather_batch(loc["gq"], B5 * world, dst=0)
        torch.cuda.synchronize(dev)
        t_sc[0] += a0.elapsed_time(a1)

    scatter_step(0)
    t_sc[0] = 0.0
    scatter_ms = timed(scatter_step, scatter_steps)
    sc_only = torch.tensor([t_sc[0] / scatter_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(sc_only, op=dist.ReduceOp.MAX)
    # the same, software-pipelined: every rank's shard in PIPE_CHUNKS pieces, piece c solved while piece c+1 is on the wire
    # and piece c-1's results travel back (shard.solve_sharded_pipelined)
    PIPE_CHUNKS = 4

    def pipe_fn(P, q, l_n, mu, g, lo, hi):
        n, sp = hi - lo, stream.cuda_stream
        o = {k: d[k][lo:hi] for k in ("x", "st", "gP", "gq", "gl", "gm")}
        if n:
            _lib.check(L.dq_qcqp_forward_ex(P.data_ptr(), q.data_ptr(), l_n.data_ptr(), mu.data_ptr(), None, o["x"].data_ptr(), None,
                                            o["st"].data_ptr(), n, N5, EPS, MU_PROX, MAX_ITER, 1, sp), "forward")
            _lib.check(L.dq_qcqp_backward_ex2(P.data_ptr(), q.data_ptr(), l_n.data_ptr(), mu.data_ptr(), o["x"].data_ptr(), g.data_ptr(),
                                              o["st"].data_ptr(), o["gP"].data_ptr(), o["gq"].data_ptr(), o["gl"].data_ptr(),
                                              o["gm"].data_ptr(), None, None, n, N5, sp), "backward")
        return o["x"], o["gq"]

    def pipe_step(_k):
        shard.solve_sharded_pipelined(pipe_fn, full, B5 * world, chunks=PIPE_CHUNKS, src=0, device=dev, trailing=trailing)
        torch.cuda.synchronize(dev)

    pipe_step(0)
    pipe_ms = timed(pipe_step, scatter_steps)
    sweep = {}
    for c in [int(v) for v in os.environ.get("DQ_CFG5_CHUNKS", "").split(",") if v]:  # experiment hook (scripts/cfg5_pipe.py)
        PIPE_CHUNKS_SAVED, PIPE_CHUNKS = PIPE_CHUNKS, c
        pipe_step(0)
        sweep[str(c)] = timed(pipe_step, scatter_steps)
        PIPE_CHUNKS = PIPE_CHUNKS_SAVED
    egress = sum(int(f.numel()) * 8 for f in full) * (world - 1) // world if rank == 0 else 0
    eg = torch.tensor([float(egress)], dtype=torch.float64, device=dev)
    dist.all_reduce(eg, op=dist.ReduceOp.MAX)
    egress = int(eg.item())
    fb = 8 * (N5 * N5 + 3 * N5 + 2 * nc) + 8 * N5  # forward incl. the hand-off write
    bb = 8 * (2 * N5 * N5 + 4 * N5 + 4 * nc)
    peak, _ = hbm_peak()
    out = {"workload": f"B={B5 * world} N={N5} QCQP (8 contacts) fp64 batch-sharded across {world} x B200 (BASELINE configs[4]: 2097152 at 8)",
           "B_per_gpu": B5, "B_total": B5 * world,
           "shard_resident": {"ms_per_step": resident_ms, "value": B5 * world / (resident_ms * 1e-3), "unit": "solves/s",
                              "step_frac_hbm": (fb + bb) * B5 / (resident_ms * 1e-3) / 1e9 / peak},
           "scatter_inclusive": {"ms_per_step": scatter_ms, "value": B5 * world / (scatter_ms * 1e-3), "unit": "solves/s",
                                 "scatter_ms": float(sc_only.item())},
           "scatter_pipelined": {"ms_per_step": pipe_ms, "value": B5 * world / (pipe_ms * 1e-3), "unit": "solves/s", "chunks": PIPE_CHUNKS,
                                 "api": "diffqcqp_b200.shard.solve_sharded_pipelined", **({"sweep_ms": sweep} if sweep else {})},
           "root_egress_bytes": egress,
           "nccl_gbs": egress / (float(sc_only.item()) * 1e-3) / 1e9 if sc_only.item() > 0 else None,
           "nccl_reference_gbs": {"nvlink5_nominal_per_direction": 900.0, "measured_peer_copy_per_direction": 770.0},
           "note": "scatter = grouped NCCL send/recv (ncclGroup of isend/irecv, diffqcqp_b200/shard.py) of (P, q, l_n, mu, grad_l) from "
                   "rank 0; gather of x* and grad_q (grad_P stays sharded).  Every byte bound for another rank leaves rank 0 exactly "
                   "once, so its egress -- (world-1)/world of the batch -- is the floor for any scatter schedule (a tree moves the "
                   "same bytes out of the root); the achieved rate against the 900 GB/s per direction of its NVLink ports is nccl_gbs.  "
                   "scatter_pipelined overlaps that egress with the solves (pieces of the shard), the same bytes and results."}
    del sets, full
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------- GPU arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from diffqcqp_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not os.path.exists(_lib.LIB_PATH):  # a checkout without the built artefact: rank 0 compiles it, the others wait
        if rank == 0:
            from diffqcqp_b200 import build as dq_build
            dq_build.build()
        else:
            t_wait = time.time()
            while not os.path.exists(_lib.LIB_PATH) and time.time() - t_wait < 300:
                time.sleep(1.0)
            time.sleep(2.0)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's kernels on a high-priority stream: in cfg5's pipelined scatter the root's send kernels then get SM slots ahead of
        # the solve's CTAs queued behind them (8 x B200: 11.8 -> 9.3 ms per step; nothing else in this file overlaps NCCL with compute)
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=dev)

    kind, B, N, kw, desc = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    nc = N // 2
    L = _lib.load()
    fb, bb = alg_bytes(kind, N)
    # QPFn2 / QCQPFn2 hand diag(P) from the forward to the backward (dq_*_forward_ex / dq_*_backward_ex*): for a diagonal
    # P the backward does not re-read it, so the compulsory traffic is forward +8N (the hand-off) and backward -8N^2 +8N
    handoff = kw["gen"] in ("qp_diag", "qcqp_diag")
    if handoff:
        fb, bb = fb + 8 * N, bb - 8 * N * N + 8 * N
    # one set = inputs (P, q, grad_l [, l_n, mu]) + outputs (x, grad_P, grad_q [, grad_l_n, grad_mu])
    R, S, in_bytes_per_set = rotating_sets(kind, B, N, args.streams)
    sets = []
    host0 = None
    for r in range(R):
        inp = make_inputs(kind, kw["gen"], B, N, seed=1000 * rank + r)
        if r == 0:
            host0 = inp
        d = {k: v.to(dev) for k, v in inp.items()}
        d["x"] = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        d["st"] = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        d["gP"] = torch.empty((B, N, N), dtype=torch.float64, device=dev)
        d["gq"] = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        if kind == "qcqp":
            d["gl"] = torch.empty((B, nc, 1), dtype=torch.float64, device=dev)
            d["gm"] = torch.empty((B, nc, 1), dtype=torch.float64, device=dev)
        sets.append(d)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream  # the stream fwd()/bwd() launch on (step_on switches it)

    def fwd(d):
        if kind == "qp":
            rc = L.dq_qp_forward_ex(d["P"].data_ptr(), d["q"].data_ptr(), None, d["x"].data_ptr(), None,
                                    d["st"].data_ptr(), B, N, EPS, MU_PROX, MAX_ITER, 1, sp)
        else:
            rc = L.dq_qcqp_forward_ex(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                      None, d["x"].data_ptr(), None, d["st"].data_ptr(), B, N, EPS, MU_PROX, MAX_ITER, 1, sp)
        _lib.check(rc, "forward")

    def bwd(d):
        if kind == "qp":
            rc = L.dq_qp_backward_ex(d["P"].data_ptr(), d["q"].data_ptr(), d["x"].data_ptr(), d["g"].data_ptr(),
                                     d["st"].data_ptr(), d["gP"].data_ptr(), d["gq"].data_ptr(), B, N, sp)
        else:
            rc = L.dq_qcqp_backward_ex2(d["P"].data_ptr(), d["q"].data_ptr(), d["l_n"].data_ptr(), d["mu"].data_ptr(),
                                        d["x"].data_ptr(), d["g"].data_ptr(), d["st"].data_ptr(), d["gP"].data_ptr(),
                                        d["gq"].data_ptr(), d["gl"].data_ptr(), d["gm"].data_ptr(), None, None, B, N, sp)
        _lib.check(rc, "backward")

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- streams: step k runs on stream k % S (fwd then bwd of one batch stay ordered on their stream; independent
    # batches overlap, which hides each forward launch's long-iteration tail behind the next batch's work)
    streams = [stream] + [torch.cuda.Stream(dev) for _ in range(S - 1)]
    sps = [st.cuda_stream for st in streams]

    def step_on(k, si):
        nonlocal sp
        sp = sps[si]
        d = sets[k % R]
        fwd(d); bwd(d)
        sp = sps[0]

    def run_steps(n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for k in range(n):
            step_on(k, k % S)
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            streams[0].wait_event(e)
        ev1.record(streams[0])
        return ev0, ev1

    # ---- warm-up
    W = max(args.warmup, 3)
    run_steps(W)
    barrier()

    # ---- timed region: exactly K steps, events on the launching stream (the side streams fork from / join into it)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.005)
    n0 = L.dq_launch_count()
    barrier()
    ev0, ev1 = run_steps(args.steps)
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = int(L.dq_launch_count() - n0)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the same K steps on ONE stream (no overlap between steps): reported next to the headline
    S_saved, S = S, 1
    barrier()
    e0, e1 = run_steps(min(args.steps, 200))
    barrier()
    serial_ms_per_step = e0.elapsed_time(e1) / min(args.steps, 200)
    S = S_saved

    # ---- per-kernel durations (separate pass, same rotation, events around each launch)
    fwd_ms, bwd_ms = [], []
    for k in range(min(max(args.steps, 10), 100)):
        d = sets[k % R]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream); fwd(d); e[1].record(stream); bwd(d); e[2].record(stream)
        torch.cuda.synchronize(dev)
        fwd_ms.append(e[0].elapsed_time(e[1])); bwd_ms.append(e[1].elapsed_time(e[2]))
    fwd_avg, bwd_avg = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)

    # ---- sustained per-launch cost of each kernel: the same launches back to back, round-robin on the S streams (what a
    # launch costs when the next batch's work fills the GPU behind its stragglers, as in the timed region)
    def sustained(fn, n=200):
        nonlocal sp
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(e0)
        for k in range(n):
            sp = sps[k % S]
            fn(sets[k % R])
        sp = sps[0]
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            streams[0].wait_event(e)
        e1.record(streams[0])
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    sustained(fwd, 20)
    fwd_sus, bwd_sus = sustained(fwd), sustained(bwd)

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

    # ---- the same step through the autograd surface a user of the reference calls (README.md:45-56): QPFn2.apply(...) and
    # .backward(), with CUDA tensors (device-resident, for comparison with `value`) and with pinned CPU tensors (host
    # buffers in, host buffers out: the chunked copy/compute pipeline of diffqcqp_b200/qcqp.py)
    e2e_autograd = None
    if not args.no_e2e:
        import qcqp as surface

        def autograd_step(t):
            leaves = [v.detach().requires_grad_(True) for v in t[:-1]]
            if kind == "qp":
                x = surface.QPFn2.apply(leaves[0], leaves[1], None, EPS, MAX_ITER)
            else:
                x = surface.QCQPFn2.apply(leaves[0], leaves[1], leaves[2], leaves[3], None, EPS, MAX_ITER)
            x.backward(t[-1])
            return leaves[0].grad

        keys = ["P", "q"] + (["l_n", "mu"] if kind == "qcqp" else []) + ["g"]
        e2e_autograd = {"unit": "solves/s"}
        nst = min(args.steps, 200)
        for label, tens in (("cuda_tensors", [sets[0][k] for k in keys]), ("cpu_pinned_tensors", [h[k] for k in keys])):
            for _ in range(3):
                autograd_step(tens)
            barrier()
            t0 = time.perf_counter()
            for _ in range(nst):
                autograd_step(tens)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            if distributed:
                tt = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            e2e_autograd[label] = {"value": B * world * nst / dt, "ms_per_step": 1e3 * dt / nst}
        e2e_autograd["api"] = ("QPFn2" if kind == "qp" else "QCQPFn2") + ".apply(...) + .backward(grad_l), wall clock incl. autograd and allocator"

    # ---- BASELINE configs[4] (N>1): B = 2,097,152 / 8 = 262144 QCQPs of N = 16 per GPU, shard-resident and with the
    # NCCL scatter of (P, q, l_n, mu, grad_l) from rank 0 and the gather of x*, grad_q back
    cfg5 = None
    if distributed and not args.no_cfg5:
        cfg5 = run_cfg5(L, dev, rank, world, barrier)

    # ---- scatter-inclusive step (N>1, --scatter): rank 0 owns all world*B problems; NCCL point-to-point scatter of the
    # inputs, the sharded solve, gather of x* and grad_q (grad_P stays sharded: it is 8N^2 bytes per problem)
    scatter_line = None
    if distributed and args.scatter:
        from diffqcqp_b200 import shard

        keys = ["P", "q", "g"] + (["l_n", "mu"] if kind == "qcqp" else [])
        full = None
        if rank == 0:
            full = [torch.cat([sets[0][k]] * world, 0) for k in keys]
        trailing = [tuple(sets[0][k].shape[1:]) for k in keys]
        d = sets[0]

        def scatter_step():
            parts = shard.scatter_batch(full, B * world, src=0, device=dev, trailing=trailing)
            loc = dict(zip(keys, parts))
            loc.update(x=d["x"], gP=d["gP"], gq=d["gq"])
            if kind == "qcqp":
                loc.update(gl=d["gl"], gm=d["gm"])
            fwd(loc); bwd(loc)
            shard.gather_batch(loc["x"], B * world, dst=0)
            shard.gather_batch(loc["gq"], B * world, dst=0)

        for _ in range(3):
            scatter_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nsc = 20
        e0.record(stream)
        for _ in range(nsc):
            scatter_step()
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / nsc], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sc_ms = float(t.item())
        in_bytes = sum(int(f.numel()) * 8 for f in full) * (world - 1) // world if rank == 0 else 0
        scatter_line = {"ms_per_step": sc_ms, "value": B * world / (sc_ms * 1e-3), "unit": "solves/s",
                        "root_egress_bytes_per_step": in_bytes,
                        "note": "rank 0 scatters (P,q,grad_l[,l_n,mu]) over NCCL send/recv, every rank solves its shard, x* and grad_q are gathered on rank 0"}

    if distributed:
        t = torch.tensor([total_ms, fwd_avg, bwd_avg, fwd_sus, bwd_sus], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, fwd_avg, bwd_avg, fwd_sus, bwd_sus = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = hbm_peak()
        # launch_admm_fwd's dispatch (csrc/admm_fwd.cu): N == 8 QPs -> thread-per-problem kernel from 65536 problems, the
        # persistent tile kernel below that; everything else the generic kernel
        if N == 8 and B >= 65536:
            fwd_name = "admm_fwd_tpp8_kernel"
        elif N == 8 and kind == "qp":
            fwd_name = "admm_fwd_diag8_kernel"
        else:
            fwd_name = "admm_fwd_kernel"
        sm_mhz = (clocks or {}).get("sm_mhz")
        dom = fwd_name if fwd_avg >= bwd_avg else ("qp_bwd_kernel" if kind == "qp" else "qcqp_bwd_kernel")
        dom_ms = max(fwd_avg, bwd_avg)
        dom_bytes = (fb if fwd_avg >= bwd_avg else bb) * B
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        step_achieved = (fb + bb) * B / (total_ms / args.steps * 1e-3) / 1e9
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the ncu capture
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
                traffic = json.load(fh).get(args.workload, {}).get(dom) if not args.batch else None
        except Exception:
            traffic = None
        line = {
            "metric": "QP/QCQP fwd+bwd solves/sec", "value": B * world * args.steps / (total_ms * 1e-3),
            "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": common_config(args, world),
            "detail": {"streams": S, "single_stream_ms_per_step": serial_ms_per_step, "fwd_bwd_handoff": handoff,
                       "forward_kernel": fwd_name},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "alg_bytes_per_solve": {"fwd": fb, "bwd": bb},
                         "kernel_ms": {"fwd": fwd_avg, "bwd": bwd_avg},
                         "kernel_frac": {"fwd": fb * B / (fwd_avg * 1e-3) / 1e9 / peak, "bwd": bb * B / (bwd_avg * 1e-3) / 1e9 / peak},
                         "kernel_ms_sustained": {"fwd": fwd_sus, "bwd": bwd_sus},
                         "kernel_frac_sustained": {"fwd": fb * B / (fwd_sus * 1e-3) / 1e9 / peak,
                                                   "bwd": bb * B / (bwd_sus * 1e-3) / 1e9 / peak},
                         "fp64": {"fwd": fp64_roofline(args.workload, fwd_name, fwd_avg, sm_mhz),
                                  "fwd_sustained": fp64_roofline(args.workload, fwd_name, fwd_sus, sm_mhz),
                                  "bwd": fp64_roofline(args.workload, "qp_bwd_kernel" if kind == "qp" else "qcqp_bwd_kernel", bwd_avg, sm_mhz)}
                         if not args.batch else None,
                         "step_achieved_gbs": step_achieved, "step_frac": step_achieved / peak,
                         "overlapped_step_ms": total_ms / args.steps,
                         "note": "the forward kernel is FP64-issue/latency bound, not HBM bound (DESIGN.md section 5.1); "
                                 "achieved / frac / kernel_ms are isolated launches on one stream (they include the launch's "
                                 "straggler tail); kernel_ms_sustained are the same launches back to back on the timed "
                                 "region's streams, where the next batch's work fills the GPU behind the stragglers (for the "
                                 "HBM-bound backward that figure also overlaps one launch's write-back with the next launch, "
                                 "so its fraction can exceed 1)"},
            "e2e": e2e, "e2e_autograd": e2e_autograd, "gpu_launches": launches, "clocks": clocks,
        }
        if cfg5 is not None:
            line["cfg5"] = cfg5
        if world == 1 and not args.batch and args.workload == "qp_diag_n8" and not args.no_other_configs:
            line["other_configs"] = run_other_configs(L, dev)
        if not args.no_cpu_baseline and world == 1:  # reported at N=1 only (the reference arm covers N>1)
            eng, ckind, a, n, reps = cpu_sample(kind, host0, B, 10.0, args.cpu_sample)
            dt = sum(cpu_pass(eng, kind, a) for _ in range(reps))
            line["cpu_baseline"] = {"value": n * reps / dt, "unit": "solves/s", "cores": host_threads(), "kind": ckind,
                                    "build": CPU_BUILD_NOTE[ckind],
                                    "sample": f"{reps} x (fwd+bwd over the first {n} of {B} problems), OpenMP over problems, {dt:.1f} s"}
            line["cpu_baseline"]["as_shipped"] = as_shipped_loop(kind, host0, eng)
        if scatter_line is not None:
            line["scatter_inclusive"] = scatter_line
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
