"""A/B of the LayerNorm kernels at BASELINE configs[1] shapes ([32768, 768] bf16), cold L2 (256 MB flush between launches):
register-staged (vvae_debug_set(7, 1)) vs bulk-async staged streaming kernels; checks both give the same result."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_vae_b200 import _ffi, ops
_ffi.require_device()
lib = _ffi.lib
g = torch.Generator(device="cuda").manual_seed(0)
N, D = 32768, 768
x = torch.randn(N, D, device="cuda", generator=g).bfloat16()
dy = torch.randn(N, D, device="cuda", generator=g).bfloat16()
dres = torch.randn(N, D, device="cuda", generator=g).bfloat16()
gamma = torch.randn(D, device="cuda", generator=g); beta = torch.randn(D, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


out = {}
res = {}
for mode in (1, 0, 2):
    lib.vvae_debug_set(7, mode * 17)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta)
    dgam = torch.zeros(D, device="cuda"); dbet = torch.zeros(D, device="cuda")
    dx = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet)
    dx2 = ops.layernorm_bwd(dy, x, mean, rstd, gamma, None, None, None)
    res[mode] = (y.float(), dx.float(), dgam.clone(), dbet.clone(), dx2.float())
    out[f"mode{mode}_fwd_us"] = round(timed(lambda: ops.layernorm_fwd(x, gamma, beta)), 2)
    out[f"mode{mode}_bwd_us"] = round(timed(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet)), 2)
    out[f"mode{mode}_bwd_nores_us"] = round(timed(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, None, dgam, dbet)), 2)
lib.vvae_debug_set(7, 0)
# size-matched streaming references (same cold-L2 protocol): a 50 MB -> 50 MB copy and a 3-read / 1-write add
yy = torch.empty_like(x)
out["copy_100MB_us"] = round(timed(lambda: yy.copy_(x)), 2)
out["add3_200MB_us"] = round(timed(lambda: torch.add(torch.add(x, dy, out=yy), dres, out=yy)), 2)
flush_r = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed_clean(fn, reps=10):
    ts = []
    for _ in range(reps + 2):
        flush.zero_(); s_ = flush_r.sum()            # leave L2 full of CLEAN lines (no write-backs during fn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]
out["clean_copy_100MB_us"] = round(timed_clean(lambda: yy.copy_(x)), 2)
out["clean_fwd_us"] = round(timed_clean(lambda: ops.layernorm_fwd(x, gamma, beta)), 2)
out["clean_bwd_us"] = round(timed_clean(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dgam, dbet)), 2)
for mode in (0, 2):
    for i, nm in enumerate(("y", "dx", "dgamma", "dbeta", "dx_nores")):
        a, b = res[mode][i], res[1][i]
        out[f"mode{mode}_{nm}_maxrel"] = float(((a - b).abs().max() / b.abs().max()).item())
out["fwd_bytes"] = N * D * 4; out["bwd_bytes"] = N * D * 8
print(json.dumps(out))

# ---- QK-norm + RoPE backward at [32768, 1536] (8 heads x 64): register-staged (vvae_debug_set(7, 0x100)) vs streaming
H, HD = 8, 64
qkv = torch.randn(N, 3 * H * HD, device="cuda", generator=g).bfloat16()
dq0 = torch.randn(N, 3 * H * HD, device="cuda", generator=g).bfloat16()
qs = torch.randn(HD, device="cuda", generator=g); ks = torch.randn(HD, device="cuda", generator=g)
pos = torch.arange(256, device="cuda").float()[:, None] * torch.exp(-torch.arange(0, HD, 2, device="cuda").float() / HD * 9.2)[None]
emb = torch.cat([pos, pos], -1)
cos, sin = torch.cos(emb).bfloat16().contiguous(), torch.sin(emb).bfloat16().contiguous()
qres = {}
qout = {}
for mode in (0x100, 0):
    lib.vvae_debug_set(7, mode)
    d = dq0.clone()
    dqs = torch.zeros(HD, device="cuda"); dks = torch.zeros(HD, device="cuda"); dbqk = torch.zeros(2 * H * HD, device="cuda")
    ops.qknorm_rope_bwd_(d, qkv, qs, ks, cos, sin, dqs, dks, H, HD, 1, 256, dbias_qk=dbqk)
    torch.cuda.synchronize()
    qres[mode] = (d, dqs, dks, dbqk)
    d2 = dq0.clone()
    qout["qknorm_bwd_%s_us" % ("stream" if mode == 0 else "regs")] = round(timed(
        lambda: ops.qknorm_rope_bwd_(d2, qkv, qs, ks, cos, sin, dqs, dks, H, HD, 1, 256, dbias_qk=dbqk)), 2)
lib.vvae_debug_set(7, 0)
qout["dqkv_bit_identical"] = bool(torch.equal(qres[0][0], qres[0x100][0]))
for i, nm in ((1, "dq_scale"), (2, "dk_scale"), (3, "dbias")):
    qout[nm + "_maxrel"] = float(((qres[0][i] - qres[0x100][i]).abs().max() / qres[0x100][i].abs().max()).item())
for mode, nm in ((0, "stream_16x2_nobias"), (0x200, "stream_8x4_nobias"), (0x100, "regs_nobias")):
    lib.vvae_debug_set(7, mode)
    d2 = dq0.clone()
    ops.qknorm_rope_bwd_(d2, qkv, qs, ks, cos, sin, dqs, dks, H, HD, 1, 256)
    qout[nm + "_same"] = bool(torch.equal(d2, qres[0x100][0]))
    d2 = dq0.clone()
    qout["qknorm_bwd_%s_us" % nm] = round(timed(lambda: ops.qknorm_rope_bwd_(d2, qkv, qs, ks, cos, sin, dqs, dks, H, HD, 1, 256)), 2)
lib.vvae_debug_set(7, 0)
qout["bytes"] = N * 1024 * 2 * 3
print(json.dumps(qout))
