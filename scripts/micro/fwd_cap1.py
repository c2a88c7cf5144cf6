import os, sys
sys.path.insert(0, os.getcwd())
import torch
from diffqcqp_b200 import _lib, workloads as wl
L = _lib.load(); dev = torch.device("cuda", 0)
B, N = 65536, 8
d = [x.to(dev) for x in wl.qp_diag(B, N, seed=0)]
x = torch.empty(B, N, 1, dtype=torch.float64, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
for k in range(4):
    L.dq_qp_forward(d[0].data_ptr(), d[1].data_ptr(), None, x.data_ptr(), None, B, N, 1e-7, 1e-7, 1, 1, sp)
torch.cuda.synchronize()
