"""Where the pinned-CPU-tensor autograd step spends its time (forward / backward wall clock, pieces inside)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import qcqp
from diffqcqp_b200 import qcqp as dq, workloads as wl

B, N = 65536, 8
P, q, g = wl.qp_diag(B, N, seed=0)
P, q, g = P.pin_memory(), q.pin_memory(), g.pin_memory()
dev = torch.device("cuda", 0)


def step():
    Pl, ql = P.detach().requires_grad_(True), q.detach().requires_grad_(True)
    t0 = time.perf_counter()
    x = qcqp.QPFn2.apply(Pl, ql, None, 1e-7, 1000)
    t1 = time.perf_counter()
    x.backward(g)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


for _ in range(5):
    step()
for chunks in (1, 2, 6):
    dq.HOST_PIPE_CHUNKS = chunks
    for _ in range(3):
        step()
    f = b = 0.0
    n = 100
    for _ in range(n):
        a, c = step()
        f += a; b += c
    print(f"chunks {chunks}: forward {1e3 * f / n:.3f} ms, backward {1e3 * b / n:.3f} ms, step {1e3 * (f + b) / n:.3f} ms")
# bare copies for reference
Pd = torch.empty_like(P, device=dev); Ph = torch.empty_like(P).pin_memory()
torch.cuda.synchronize()
for name, fn in (("H2D 33.5 MB", lambda: Pd.copy_(P, non_blocking=True)), ("D2H 33.5 MB", lambda: Ph.copy_(Pd, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    print(f"{name}: {1e3 * (time.perf_counter() - t0) / 50:.3f} ms")
t0 = time.perf_counter()
for _ in range(50):
    t = torch.empty((B, N, N), dtype=torch.float64, pin_memory=True)
print(f"pinned alloc (cached) 33.5 MB: {1e3 * (time.perf_counter() - t0) / 50:.3f} ms")
