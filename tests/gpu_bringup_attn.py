"""Bring-up: tcgen05 attention (fwd, bwd) vs the generic kernel on the same bf16 data. Run on the B200 box."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_vae_b200 import _ffi, ops  # noqa: E402
from video_vae_b200.ops import AttnGeom, AttnMask  # noqa: E402

H, HD = 8, 64
Q = H * HD


def make_case(name, b, t, hw, temporal, masked, timed=False, all_masked_row=False):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    N = b * t * hw
    qkv = torch.randn(N, 3 * Q, device=dev, generator=g).bfloat16()
    qk = (torch.randn(N, 2 * Q, device=dev, generator=g) * 1.5).bfloat16()
    d_o = torch.randn(N, Q, device=dev, generator=g).bfloat16()
    if temporal:
        geom = AttnGeom(b, hw, t, t * hw, 1, hw)
        L = t
    else:
        geom = AttnGeom(b * t, 1, hw, hw, 0, 1)
        L = hw
    mask = None
    if masked:
        keep = torch.randint(1, L + 1, (geom.n_outer if temporal else geom.n_seq,), device=dev, generator=g)
        m = (torch.arange(L, device=dev)[None, :] < keep[:, None])
        if all_masked_row:
            m[0] = False
        m = m.to(torch.uint8).contiguous()
        mask = AttnMask(m, hw if temporal else 1, L, 0, 0, 1)
    res = {"name": name, "L": L, "n_seq": geom.n_seq}
    outs = {}
    for backend in (_ffi.BACKEND_SIMT, _ffi.BACKEND_AUTO):
        ops.ATTN_BACKEND = backend
        o, lse = ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], mask, 1.0 / math.sqrt(HD))
        dqkv = torch.zeros(N, 3 * Q, device=dev, dtype=torch.bfloat16)
        ops.attn_bwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], o, lse, d_o, dqkv[:, :Q], dqkv[:, Q:2 * Q],
                     dqkv[:, 2 * Q:], mask, 1.0 / math.sqrt(HD))
        torch.cuda.synchronize()
        outs[backend] = (o.float(), lse.clone(), dqkv.float())
        if timed and backend == _ffi.BACKEND_AUTO:
            for kind in ("fwd", "bwd"):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    if kind == "fwd":
                        ops.attn_fwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], mask, 1.0 / math.sqrt(HD))
                    else:
                        ops.attn_bwd(geom, H, HD, qk[:, :Q], qk[:, Q:], qkv[:, 2 * Q:], o, lse, d_o, dqkv[:, :Q],
                                     dqkv[:, Q:2 * Q], dqkv[:, 2 * Q:], mask, 1.0 / math.sqrt(HD))
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                fl = (4.0 if kind == "fwd" else 10.0) * geom.n_seq * H * L * L * HD
                res[kind + "_ms"] = ms
                res[kind + "_tflops"] = fl / ms / 1e9
    ops.ATTN_BACKEND = _ffi.BACKEND_AUTO
    (o_s, lse_s, d_s), (o_t, lse_t, d_t) = outs[_ffi.BACKEND_SIMT], outs[_ffi.BACKEND_AUTO]
    res["o_err"] = ((o_t - o_s).abs().max() / o_s.abs().max()).item()
    fin = lse_s > -1e30
    res["lse_err"] = (lse_t[fin] - lse_s[fin]).abs().max().item() if fin.any() else 0.0
    for nm, sl in (("dq", slice(0, Q)), ("dk", slice(Q, 2 * Q)), ("dv", slice(2 * Q, 3 * Q))):
        a, bref = d_t[:, sl], d_s[:, sl]
        res[nm + "_err"] = ((a - bref).abs().max() / bref.abs().max().clamp_min(1e-9)).item()
    return res


CASES = [
    ("spatial_L256", 1, 3, 256, False, False),
    ("spatial_L128", 1, 2, 128, False, False),
    ("spatial_L256_masked", 1, 2, 256, False, True),
    ("spatial_L64_pack2", 1, 5, 64, False, False),
    ("spatial_L16_pack8", 2, 5, 16, False, False),
    ("temporal_L16_hw16_masked", 2, 16, 16, True, True),
    ("temporal_L16_hw12_tail", 1, 16, 12, True, True),
    ("temporal_L8_hw64", 1, 8, 64, True, False),
    ("temporal_L32_hw8_masked", 3, 32, 8, True, True),
    ("temporal_L64_hw4", 1, 64, 4, True, True),
    ("temporal_allmasked_row", 2, 16, 16, True, True),
    ("temporal_L12_hw20_masked", 2, 12, 20, True, True),
    ("temporal_L5_hw7", 3, 5, 7, True, False),
    ("spatial_L9_masked", 1, 6, 9, False, True),
    ("temporal_L17_hw5_masked", 2, 17, 5, True, True),
    ("temporal_L24_hw9", 1, 24, 9, True, False),
    ("temporal_L40_hw6_masked", 2, 40, 6, True, True),
    ("temporal_L48_hw3", 1, 48, 3, True, False),
    ("temporal_L64_hw5_masked", 2, 64, 5, True, True),
    ("spatial_L36_masked", 1, 3, 36, False, True),
    ("long_L384", 1, 2, 384, False, False),
    ("long_L512_masked", 1, 3, 512, False, True),
    ("long_L1024", 1, 1, 1024, False, False),
]

if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    sel = sys.argv[1:]
    with open("gpurun_out/attn_bringup.jsonl", "a") as f:
        for c in CASES:
            if sel and c[0] not in sel:
                continue
            try:
                r = make_case(*c, all_masked_row=(c[0] == "temporal_allmasked_row"))
            except Exception as e:  # noqa: BLE001
                r = {"name": c[0], "error": repr(e)[:300]}
            print(json.dumps(r), flush=True)
            f.write(json.dumps(r) + "\n")
        for c in [("prod_spatial", 8, 16, 256, False, False), ("prod_temporal", 8, 16, 256, True, True),
                  ("prod_temporal_tcgen05", 8, 16, 256, True, True), ("prod_cfg5_spatial", 1, 64, 1024, False, False),
                  ("prod_rl_temporal_L32", 4, 32, 256, True, True), ("prod_rl_temporal_L32_tcgen05", 4, 32, 256, True, True),
                  ("prod_cfg5_temporal_L64", 1, 64, 1024, True, False),
                  ("prod_cfg5_temporal_L64_tcgen05", 1, 64, 1024, True, False)]:
            if sel and c[0] not in sel:
                continue
            # key 9: keep L <= 16 on the packed-tile tcgen05 kernels instead of the one-warp-per-sequence kernel
            _ffi.lib.vvae_debug_set(9, 1 if c[0].endswith("_tcgen05") else 0)
            try:
                r = make_case(*c, timed=True)
            except Exception as e:  # noqa: BLE001
                r = {"name": c[0], "error": repr(e)[:300]}
            print(json.dumps(r), flush=True)
            f.write(json.dumps(r) + "\n")
