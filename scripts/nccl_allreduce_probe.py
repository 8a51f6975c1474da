"""Time the flat-gradient all-reduce (170.6 M fp32 = 682 MB) on its own: torchrun --nproc-per-node N this_file.
Prints algorithm bandwidth and bus bandwidth; environment (NCCL_*) is whatever the caller exported."""
import os

import torch
import torch.distributed as dist

rank = int(os.environ["RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
n = 170_600_000
g = torch.randn(n, device="cuda")
for _ in range(3):
    dist.all_reduce(g)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters):
    dist.all_reduce(g)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
t = torch.tensor([ms], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gb = n * 4 / 1e9
    print({"world": world, "ms": round(t.item(), 3), "algbw_GBs": round(gb / t.item() * 1e3, 1),
           "busbw_GBs": round(gb / t.item() * 1e3 * 2 * (world - 1) / world, 1),
           "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}, flush=True)
dist.destroy_process_group()
