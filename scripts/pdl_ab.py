"""A/B of programmatic dependent launch inside ONE process: two captured step graphs (with / without the PDL launch
attribute, vvae_debug_set(11)), replayed alternately so that both see the same clocks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_vae_b200 as V
from video_vae_b200 import _ffi
from video_vae_b200.ddp import FlatParams
from video_vae_b200.graph import GraphedTrainStep

prod = (256, 256, 3, 16, 9, 12, 1536, 8, 512, 64, 8, 4)
m = V.VideoVAE(*prod, V.Rngs(2), dtype=torch.bfloat16)
with torch.no_grad():
    m.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator(device="cuda").manual_seed(7))
flat = FlatParams(m)
flat.enable_bf16_shadow()
g = torch.Generator().manual_seed(1234)
video = torch.rand(8, 16, 256, 256, 3, generator=g).to(torch.bfloat16).cuda()
mask = torch.ones(8, 16, dtype=torch.bool).cuda()
hp = dict(V.DEFAULT_HPARAMS, gamma4=0.1)
graphs = {}
for name, flag in (("pdl", 0), ("serialized", 1)):
    _ffi.lib.vvae_debug_set(11, flag)
    graphs[name] = GraphedTrainStep(m, flat, video, mask, hp)
_ffi.lib.vvae_debug_set(11, 0)
rngs = V.Rngs(3)
res = {k: [] for k in graphs}
for rnd in range(4):
    for name, gs in graphs.items():
        for _ in range(2):
            gs(video, mask, rngs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            gs(video, mask, rngs)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(round(e0.elapsed_time(e1) / 8, 3))
print(json.dumps({"what": "cfg2 step graph (fwd+loss+bwd, no optimizer) ms/step, alternating", **res,
                  "mean": {k: round(sum(v) / len(v), 3) for k, v in res.items()}}))
