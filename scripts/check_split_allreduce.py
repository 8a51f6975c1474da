#!/usr/bin/env python
"""N-GPU check (torchrun): the decoder-first split all-reduce (ddp.SplitAllReduce, overlapped with the encoder's backward
through an external event recorded inside the step graph) gives the same gradients as one all-reduce after the graph,
and how long a production-size step takes both ways.  Prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import video_vae_b200 as V  # noqa: E402
from video_vae_b200.ddp import FlatParams, SplitAllReduce  # noqa: E402
from video_vae_b200.graph import GraphedTrainStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
small = "--small" in sys.argv
if small:
    cfg, B, T, S = (64, 64, 3, 16, 2, 2, 256, 2, 128, 32, 8, 4), 2, 8, 64
else:
    cfg, B, T, S = (256, 256, 3, 16, 9, 12, 1536, 8, 512, 64, 8, 4), 8, 16, 256
model = V.VideoVAE(*cfg, V.Rngs(2), dtype=torch.bfloat16, device=dev)
with torch.no_grad():
    model.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(7))
flat = FlatParams(model)
flat.enable_bf16_shadow()
g = torch.Generator().manual_seed(1234 + rank)
video = torch.rand(B, T, S, S, 3, generator=g).to(torch.bfloat16).to(dev)
mask = torch.ones(B, T, dtype=torch.bool, device=dev)
graphed = GraphedTrainStep(model, flat, video, mask, V.DEFAULT_HPARAMS, mark_decoder_done=True)
split = SplitAllReduce(flat, model)


def step(mode, seed):
    graphed(video, mask, V.Rngs(seed))
    if mode == "split":
        split(graphed.decoder_done)
    else:
        dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)


out = {"world": world, "small": small}
step("single", 5)
torch.cuda.synchronize()
ref = flat.grad.clone()
for i in range(3):
    step("split", 5)
    torch.cuda.synchronize()
    # same draws, same data: the reduced gradients must agree up to the fp32-atomics noise of the step itself
    err = ((flat.grad - ref).abs().max() / ref.abs().max()).item()
    out[f"split_vs_single_relerr_{i}"] = err
for mode in ("single", "split", "single", "split"):
    for _ in range(2):
        step(mode, 7)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        step(mode, 7)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 8], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.setdefault(f"ms_{mode}", []).append(round(t.item(), 3))
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
