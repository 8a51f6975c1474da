"""Can GraphedTrainStep be captured AFTER an eager backward ran in the process?  Diagnostics."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_vae_b200 as V
from video_vae_b200.ddp import FlatParams
from video_vae_b200 import graph as G, functional as F_

variant = sys.argv[1] if len(sys.argv) > 1 else "default"
cfg = (64, 64, 3, 16, 1, 1, 256, 2, 128, 32, 8, 4)
m = V.VideoVAE(*cfg, V.Rngs(2), dtype=torch.bfloat16)
with torch.no_grad():
    m.decoder.unet.final_conv.kernel.normal_(0.0, 0.05, generator=torch.Generator(device="cuda").manual_seed(7))
flat = FlatParams(m)
flat.enable_bf16_shadow()
g = torch.Generator().manual_seed(9)
video = torch.rand(2, 4, 64, 64, 3, generator=g).to(torch.bfloat16).cuda()
mask = torch.ones(2, 4, dtype=torch.bool).cuda()

def eager():
    flat.zero_grad()
    loss, _ = V.loss_fn(m, video, mask[:, None, None, :], mask, V.Rngs(5), V.DEFAULT_HPARAMS, train=True)
    loss.backward()
    return loss

if variant == "eager_on_side":
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        loss = eager()
    torch.cuda.current_stream().wait_stream(s)
else:
    loss = eager()
torch.cuda.synchronize()
print("variant", variant, "eager loss", loss.item(), flush=True)
if variant == "trace":
    for name in dir(F_):
        cls = getattr(F_, name)
        if isinstance(cls, type) and issubclass(cls, torch.autograd.Function) and cls is not torch.autograd.Function:
            orig = cls.backward
            def wrap(ctx, *a, _o=orig, _n=name):
                print("  bwd", _n, "stream", torch.cuda.current_stream().cuda_stream, "capturing", torch.cuda.is_current_stream_capturing(), flush=True)
                return _o(ctx, *a)
            cls.backward = staticmethod(wrap)
try:
    gs = G.GraphedTrainStep(m, flat, video, mask, V.DEFAULT_HPARAMS)
    l1 = gs(video, mask, V.Rngs(5)).item()
    torch.cuda.synchronize()
    print(f"variant {variant}: capture OK, loss {l1}")
except Exception as e:  # noqa: BLE001
    print(f"variant {variant}: FAILED {type(e).__name__}: {str(e)[:200]}")
