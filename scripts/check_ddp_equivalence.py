#!/usr/bin/env python
"""2-GPU correctness check of the data-parallel path (reference analogue
claude_distributed/test_distributed.py:159-163), run under torchrun:

  * the all-reduced (mean) gradient of 2 ranks x 2 clips equals the single-process gradient of the same 4 clips;
  * after k clip+Adam steps on the reduced gradients every rank's parameters are bit-identical;
  * FlatParams.broadcast replicates rank 0's weights.

fp32 model (generic kernels, tolerance 1e-4) and bf16 model on the tensor-core path (bf16 tolerance).  Rank 0 prints
one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_vae_b200 as V  # noqa: E402
from video_vae_b200.ddp import FlatAdam, FlatParams  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = (64, 64, 3, 16, 2, 2, 256, 2, 128, 32, 8, 4)
per_rank, T, hw, lat = 2, 8, 16, 96
G = world * per_rank
out = {"world": world}
for name, dtype, tol in (("fp32", torch.float32, 1e-4), ("bf16", torch.bfloat16, 2e-2)):
    model = V.VideoVAE(*cfg, V.Rngs(2 + rank), dtype=dtype, device=dev)      # different init per rank ...
    with torch.no_grad():
        model.decoder.unet.final_conv.kernel.normal_(0.0, 0.05, generator=torch.Generator(device=dev).manual_seed(7 + rank))
    flat = FlatParams(model)
    flat.broadcast(src=0)                                                    # ... until the broadcast
    both = [torch.empty_like(flat.flat) for _ in range(world)]
    dist.all_gather(both, flat.flat)
    out[f"{name}_broadcast_identical"] = all(torch.equal(both[0], b) for b in both[1:])
    g = torch.Generator().manual_seed(99)                                    # the same global batch on every rank
    video = torch.rand(G, T, 64, 64, 3, generator=g).to(dev)
    mask = torch.ones(G, T, dtype=torch.bool)
    mask[1, 5:] = False
    mask[2, 2:] = False
    mask = mask.to(dev)
    noise = torch.randn(G, T, hw, lat, generator=g).to(dev)
    u = torch.rand(G, T, 1, generator=g).to(dev)
    hp = dict(V.DEFAULT_HPARAMS, gamma4=0.1)

    def grads(sl):
        flat.zero_grad()
        loss, _ = V.loss_fn(model, video[sl], mask[sl][:, None, None, :], mask[sl], V.Rngs(0), hp, train=True,
                            noise=noise[sl], gumbel_u=u[sl])
        loss.backward()
        return loss.detach()

    loss_full = grads(slice(0, G))
    g_full = flat.grad.clone()
    loss_mine = grads(slice(rank * per_rank, (rank + 1) * per_rank))
    dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
    flat.grad.div_(world)
    dist.all_reduce(loss_mine, op=dist.ReduceOp.SUM)
    err = ((flat.grad - g_full).norm() / g_full.norm()).item()
    out[f"{name}_grad_rel_l2"] = err
    out[f"{name}_loss_full_vs_mean_of_ranks"] = [loss_full.item(), (loss_mine / world).item()]
    out[f"{name}_grad_ok"] = err < (5 * tol if dtype == torch.bfloat16 else tol)
    # k optimizer steps on the reduced gradients, per-rank data: replicas must stay bit-identical
    opt = FlatAdam(flat, lr=1e-3)
    for step in range(3):
        grads(slice(rank * per_rank, (rank + 1) * per_rank))
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
        opt.step(grad_scale=1.0 / world)
    dist.all_gather(both, flat.flat)
    out[f"{name}_replicas_bit_identical_after_3_steps"] = all(torch.equal(both[0], b) for b in both[1:])
    del model, flat, opt
ok = all(v for k, v in out.items() if k.endswith(("_ok", "_identical", "_steps")))
out["ok"] = ok
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
