"""Two (or more) ranks through vvae_comm_* WITHOUT torch.distributed (ddp.NativeComm, csrc/comm.cu).

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_native_comm.py OUT.json

torchrun only spawns the processes and sets RANK / LOCAL_RANK / WORLD_SIZE; the rendezvous token travels through a file.
Checks sum / mean / broadcast in fp32 and bf16 against closed forms, then times the production gradient exchange
(170.6 M fp32 values = 682 MB, in place) with CUDA events on the launching stream."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import ddp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
path = os.path.join(tempfile.gettempdir(), "vvae_comm_token_" + os.environ.get("MASTER_PORT", "0"))
comm = ddp.NativeComm.from_env(path, timeout_s=120.0)
n = 1 << 22
base = torch.arange(n, dtype=torch.float32, device="cuda") % 1024
x = base * (rank + 1)
comm.all_reduce(x)
tri = world * (world + 1) // 2
ok_sum = torch.equal(x, base * tri)
y = base * (rank + 1)
comm.all_reduce(y, average=True)
ok_mean = torch.allclose(y, base * (tri / world), rtol=1e-6)
z = torch.full((n,), float(rank + 7), device="cuda")
comm.broadcast(z, world - 1)
ok_bcast = bool((z == float(world - 1 + 7)).all())
w = (base * (rank + 1)).to(torch.bfloat16) % 64     # integers below 64: their sums over the ranks are exact in bf16
expect = sum(((base * (r + 1)).to(torch.bfloat16) % 64).float() for r in range(world)).to(torch.bfloat16)
comm.all_reduce(w)
ok_bf16 = torch.equal(w, expect)
# the production exchange: 170.63 M fp32 gradients in place
g = torch.ones(170_630_000, dtype=torch.float32, device="cuda")
for _ in range(2):
    comm.all_reduce(g, average=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    comm.all_reduce(g, average=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
ok_g = bool((g == 1.0).all())
comm.close()
res = dict(rank=rank, world=world, sum=ok_sum, mean=ok_mean, broadcast=ok_bcast, bf16_sum=ok_bf16, grad_mean_unchanged=ok_g,
           allreduce_682MB_ms=ms, algbw_GBps=g.numel() * 4 / ms / 1e6, busbw_GBps=g.numel() * 4 / ms / 1e6 * 2 * (world - 1) / world)
print(json.dumps(res), flush=True)
if len(sys.argv) > 1:
    with open(f"{sys.argv[1]}.rank{rank}", "w") as f:
        json.dump(res, f)
if rank == 0 and os.path.exists(path):
    os.remove(path)
assert ok_sum and ok_mean and ok_bcast and ok_bf16 and ok_g, res
