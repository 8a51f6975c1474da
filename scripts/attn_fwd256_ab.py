"""Unmasked L = 256 attention forward: the persistent kernel (attn_fwd256_sm100_kernel, default) against the one-tile-per-CTA
kernel (vvae_debug_set(18, 1)) and a float64 torch reference, at the production spatial shape (128 sequences x 8 heads) and
at ragged unit counts; cold-L2 CUDA-event timing of both.  Writes gpurun_out/attn_fwd256_ab.jsonl."""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
from video_vae_b200.ops import AttnGeom
_ffi.require_device()
H, HD = 8, 64
Q = H * HD
sc = 1.0 / math.sqrt(HD)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = []


def ref(q, k, v, nseq, L):
    q = q.double().view(nseq, L, H, HD).transpose(1, 2)
    k = k.double().view(nseq, L, H, HD).transpose(1, 2)
    v = v.double().view(nseq, L, H, HD).transpose(1, 2)
    s = q @ k.transpose(-1, -2) * sc
    p = torch.softmax(s, -1)
    return (p @ v).transpose(1, 2).reshape(nseq * L, Q), torch.logsumexp(s, -1)


def run(nseq, timed):
    g = torch.Generator(device="cuda").manual_seed(nseq)
    L = 256
    N = nseq * L
    qk = (torch.randn(N, 2 * Q, device="cuda", generator=g) * 1.5).bfloat16()
    v = torch.randn(N, 3 * Q, device="cuda", generator=g).bfloat16()[:, 2 * Q:]     # strided rows, as in the model
    geom = AttnGeom(nseq, 1, L, L, 0, 1)
    res = {}
    for old in (1, 0):
        _ffi.lib.vvae_debug_set(18, old)
        o, lse = ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], v, None, sc)
        torch.cuda.synchronize()
        res[old] = (o.float().clone(), lse.clone())
    o_ref, lse_ref = ref(qk[:, :Q], qk[:, Q:], v, nseq, L)
    rec = {"nseq": nseq, "units": nseq * H}
    for old, name in ((1, "tile_per_cta"), (0, "persistent")):
        o, lse = res[old]
        rec[name + "_o_err"] = float((o.double() - o_ref).abs().max() / o_ref.abs().max())
        rec[name + "_lse_err"] = float((lse.double().view_as(lse_ref) - lse_ref).abs().max())
    rec["o_new_vs_old"] = float((res[0][0] - res[1][0]).abs().max())
    rec["lse_new_vs_old"] = float((res[0][1] - res[1][1]).abs().max())
    if timed:
        for old, name in ((1, "tile_per_cta"), (0, "persistent"), (0, "persistent_lockstep")):
            _ffi.lib.vvae_debug_set(18, old)
            _ffi.lib.vvae_debug_set(10, 32 if name.endswith("lockstep") else 0)
            ts = []
            for _ in range(8):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], v, None, sc)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            rec[name + "_us"] = round(ts[len(ts) // 2], 2)
            rec[name + "_tflops"] = round(4.0 * nseq * H * L * L * HD / (ts[len(ts) // 2] * 1e-6) / 1e12, 1)
    _ffi.lib.vvae_debug_set(18, 0)
    _ffi.lib.vvae_debug_set(10, 0)
    print(json.dumps(rec), flush=True)
    out.append(rec)


for nseq, timed in ((1, False), (3, False), (19, False), (37, False), (128, True), (512, True)):
    run(nseq, timed)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/attn_fwd256_ab.jsonl", "w") as f:
    for r in out:
        f.write(json.dumps(r) + "\n")
