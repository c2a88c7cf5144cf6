"""PCIe copy rates with pinned host memory: H2D alone, D2H alone, both at once (what bounds bench.py's e2e)."""
import time, torch
n = 42 * 1024 * 1024 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory(); h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.randn(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for name, a, b in (("H2D 42 MB", 1, 0), ("D2H 42 MB", 0, 1), ("both 42+42 MB", 1, 1)):
    run(a, b, 3); t = run(a, b)
    print(f"{name}: {t*1e3:.3f} ms  -> {(a+b)*n*8/t/1e9:.1f} GB/s total")
