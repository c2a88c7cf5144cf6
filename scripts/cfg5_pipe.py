"""The cfg5 section of bench.py alone (B=262144 N=16 QCQPs per GPU: shard-resident, scatter-inclusive, pipelined), under
torchrun:   DQ_CFG5_CHUNKS=2,8 python -m torch.distributed.run --nproc-per-node 8 ... scripts/cfg5_pipe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from diffqcqp_b200 import _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
L = _lib.load()


def barrier():
    dist.barrier()
    torch.cuda.synchronize(dev)


out = bench.run_cfg5(L, dev, rank, world, barrier)
if rank == 0:
    out.pop("note", None)
    out["env"] = {k: os.environ.get(k) for k in ("TORCH_NCCL_HIGH_PRIORITY", "DQ_CFG5_CHUNKS", "NCCL_MAX_NCHANNELS")}
    print(json.dumps(out))
dist.destroy_process_group()
