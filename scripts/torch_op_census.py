"""Which torch (non-libvvae) GPU kernels run inside one production training step, and from which source line.
Run on the B200 box:  python scripts/torch_op_census.py > gpurun_out/torch_ops.txt"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import video_vae_b200 as V  # noqa: E402
from video_vae_b200.ddp import FlatParams  # noqa: E402
from scripts.run_configs import build  # noqa: E402

dev = torch.device("cuda", 0)
m, flat = build(256, dev)
g = torch.Generator().manual_seed(1)
video = torch.rand(8, 16, 256, 256, 3, generator=g).to(dev)
mask = torch.ones(8, 16, dtype=torch.bool, device=dev)
hp = dict(V.DEFAULT_HPARAMS, gamma4=0.05)


def step():
    flat.zero_grad()
    loss, _ = V.loss_fn(m, video, mask[:, None, None, :], mask, V.Rngs(3), hp, train=True)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for it in prof.key_averages(group_by_stack_n=12):
    us = getattr(it, "self_device_time_total", None)
    if us is None:
        us = getattr(it, "self_cuda_time_total", 0.0)
    if not it.key.startswith("aten::") or us <= 0:
        continue
    where = "?"
    for fr in it.stack or []:
        if "video_vae_b200/" in fr:
            where = fr.split("video_vae_b200/")[-1].strip()
            break
    agg[(it.key, where)][0] += it.count
    agg[(it.key, where)][1] += us
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in agg.values())
print(f"torch ops with device time in one step: {sum(v[0] for v in agg.values())} calls, {tot / 1e3:.3f} ms")
for (name, where), (n, us) in rows[:45]:
    print(f"{us / 1e3:8.3f} ms  {n:4d}x  {name:28s} {where}")
