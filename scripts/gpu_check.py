"""Verbose GPU-vs-oracle comparison used while developing (prints stats; pytest holds the asserts)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from diffqcqp_b200 import qcqp as dq  # noqa: E402
from diffqcqp_b200 import workloads as wl  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def stats(name, a, b, tol=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b).reshape(a.shape[0], -1).max(1)
    s = dict(max=float(np.nanmax(d)), mean=float(np.nanmean(d)), nan=int(np.isnan(d).sum()))
    if tol is not None:
        s["frac_gt_tol"] = float((d > tol).mean())
    rel = d / (np.abs(b).reshape(b.shape[0], -1).max(1) + 1e-300)
    s["rel_median"] = float(np.median(rel)); s["rel_p99"] = float(np.percentile(rel, 99)); s["rel_max"] = float(rel.max())
    print(f"  {name}: {s}", flush=True)
    out[name] = s
    return s


def run_qp(tag, P, q, g, eps=1e-7, max_iter=1000):
    print(f"[QP {tag}] B={P.shape[0]} N={P.shape[1]} eps={eps}", flush=True)
    t = time.time()
    xo, ito = orc.qp_forward(P.numpy(), q.numpy(), None, eps, max_iter, return_iters=True)
    gPo, gqo = orc.qp_backward(P.numpy(), q.numpy(), xo, g.numpy())
    t_or = time.time() - t
    Pd, qd, gd = P.to(dev), q.to(dev), g.to(dev)
    x, it = dq.qp_forward(Pd, qd, eps, max_iter, return_iters=True)
    torch.cuda.synchronize()
    stats(f"qp_{tag}_x", x.cpu().numpy(), xo, 10 * eps)
    it = it.cpu().numpy()
    print(f"  iters: gpu mean {it.mean():.2f} max {it.max()}  oracle mean {ito.mean():.2f} max {ito.max()}  mismatches {(it != ito).sum()}")
    out[f"qp_{tag}_iter_mismatch"] = int((it != ito).sum())
    # backward on the ORACLE's x so both sides differentiate the same point
    xod = torch.from_numpy(xo).to(dev)
    gP, gq = dq.qp_backward(Pd, qd, xod, gd)
    torch.cuda.synchronize()
    stats(f"qp_{tag}_gq", gq.cpu().numpy(), gqo)
    stats(f"qp_{tag}_gP", gP.cpu().numpy(), gPo)
    print(f"  oracle time {t_or:.2f}s", flush=True)


def run_qcqp(tag, P, q, ln, mu, g, eps=1e-7, max_iter=1000):
    print(f"[QCQP {tag}] B={P.shape[0]} N={P.shape[1]} eps={eps}", flush=True)
    xo, ito = orc.qcqp_forward(P.numpy(), q.numpy(), ln.numpy(), mu.numpy(), None, eps, max_iter, return_iters=True)
    gPo, gqo, glo, gmo = orc.qcqp_backward(P.numpy(), q.numpy(), ln.numpy(), mu.numpy(), xo, g.numpy())
    Pd, qd, ld, md, gd = P.to(dev), q.to(dev), ln.to(dev), mu.to(dev), g.to(dev)
    x, it = dq.qcqp_forward(Pd, qd, ld, md, eps, max_iter, return_iters=True)
    torch.cuda.synchronize()
    stats(f"qcqp_{tag}_x", x.cpu().numpy(), xo, 10 * eps)
    it = it.cpu().numpy()
    print(f"  iters: gpu mean {it.mean():.2f} max {it.max()}  oracle mean {ito.mean():.2f} max {ito.max()}  mismatches {(it != ito).sum()}")
    out[f"qcqp_{tag}_iter_mismatch"] = int((it != ito).sum())
    xod = torch.from_numpy(xo).to(dev)
    gP, gq, gl, gm = dq.qcqp_backward(Pd, qd, ld, md, xod, gd)
    torch.cuda.synchronize()
    stats(f"qcqp_{tag}_gq", gq.cpu().numpy(), gqo)
    stats(f"qcqp_{tag}_gP", gP.cpu().numpy(), gPo)
    stats(f"qcqp_{tag}_gln", gl.cpu().numpy(), glo)
    stats(f"qcqp_{tag}_gmu", gm.cpu().numpy(), gmo)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    run_qp("readme", *wl.qp_diag(10, 8, seed=1, signed_q=False))
    run_qp("diag8_small", *wl.qp_diag(37, 8, seed=2))
    run_qp("diag8", *wl.qp_diag(65536, 8, seed=0))
    run_qp("diag8_e10", *wl.qp_diag(8192, 8, seed=3), eps=1e-10)
    run_qp("dense8", *wl.qp_dense(4096, 8, seed=4))
    run_qp("dense5", *wl.qp_dense(1001, 5, seed=5))
    run_qp("dense16", *wl.qp_dense(2048, 16, seed=6))
    run_qp("dense32", *wl.qp_dense(1024, 32, seed=7))
    run_qp("dense24", *wl.qp_dense(515, 24, seed=8))
    run_qp("diag32", *wl.qp_diag(4096, 32, seed=9))
    run_qcqp("dense16", *wl.qcqp_dense(2048, 16, seed=10))
    run_qcqp("dense24", *wl.qcqp_dense(1024, 24, seed=11))
    run_qcqp("dense32", *wl.qcqp_dense(1024, 32, seed=12))
    run_qcqp("dense8", *wl.qcqp_dense(4099, 8, seed=13))
    run_qcqp("diag16", *wl.qcqp_dense(2048, 16, seed=14, diag=True))
    run_qcqp("dense6", *wl.qcqp_dense(333, 6, seed=15))
    # quick timing of the headline config (device resident)
    P, q, g = wl.qp_diag(65536, 8, seed=0)
    Pd, qd, gd = P.to(dev), q.to(dev), g.to(dev)
    for _ in range(3):
        x = dq.qp_forward(Pd, qd, 1e-7, 1000); dq.qp_backward(Pd, qd, x, gd)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); x = dq.qp_forward(Pd, qd, 1e-7, 1000); e[1].record(); dq.qp_backward(Pd, qd, x, gd); e[2].record()
    torch.cuda.synchronize()
    print(f"headline: fwd {e[0].elapsed_time(e[1])*1e3:.1f} us  bwd {e[1].elapsed_time(e[2])*1e3:.1f} us", flush=True)
    out["headline_fwd_us"] = e[0].elapsed_time(e[1]) * 1e3
    out["headline_bwd_us"] = e[1].elapsed_time(e[2]) * 1e3
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gpu_check.json", "w"), indent=1)
