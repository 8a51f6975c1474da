#!/usr/bin/env python
"""Headline benchmark: train clips/sec of the video-VAE hot path (fwd + loss + bwd [+ all-reduce] + Adam) at
16x256x256, bf16, batch 8 per GPU (BASELINE.json configs[1]), production hyper-parameters
(train/rl_nonadversarial.py:234-236).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg4|cfg5]

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Rank 0 prints ONE JSON line.  `--impl reference` times the
CPU oracle (the reference is JAX/Flax and cannot be installed offline) on the host cores, on the SAME configuration.

--config selects the BASELINE.json row (same JSON schema; `config.workload` names the row):
  cfg2 (default)  configs[1]: training step, 16x256x256, batch 8 per GPU           -> the headline metric
  cfg3            configs[2]: encode-only latent extraction, 32x256x256, batch 32 (chunks of 8), eval mode, no grad
  cfg4            configs[3]: cfg2 with prefix masks keeping 4/8/12/16 of 16 frames (two clips each; --keep K = all K)
  cfg5            configs[4]: long clip 64x512x512, batch 1 per GPU (run with --gpus 8 for the BASELINE row)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROD = dict(patch_size=16, encoder_depth=9, decoder_depth=12, mlp_dim=1536, num_heads=8, qkv_features=512,
            max_temporal_len=64, spatial_compression_rate=8, unembedding_upsample_rate=4)
METRIC = "train clips/sec at 16x256x256 fwd+bwd"
FLOP_PER_CLIP = {256: 5.10e12, 128: 1.25e12}   # SURVEY.md section 8(d), fwd+bwd = 3x fwd, remat not counted
# BASELINE.json rows other than the headline: (frames, size, batch/GPU, metric, algorithmic FLOP per clip, workload text)
CONFIGS = {
    "cfg2": dict(frames=16, size=256, batch=8, metric=METRIC, flop=5.10e12, row="configs[1]"),
    "cfg3": dict(frames=32, size=256, batch=32, metric="encode-only clips/sec at 32x256x256 (latent extraction)",
                 flop=1.215e12, row="configs[2]"),
    "cfg4": dict(frames=16, size=256, batch=8, metric="train clips/sec at 16x256x256 fwd+bwd, 25-100% frames kept",
                 flop=5.10e12, row="configs[3]"),
    "cfg5": dict(frames=64, size=512, batch=1, metric="train clips/sec at 64x512x512 fwd+bwd", flop=88.5e12,
                 row="configs[4]"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default: the BASELINE row's)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--keep", type=int, default=None, help="cfg4: frames kept in every clip (default 4/8/12/16 mix)")
    ap.add_argument("--chunk", type=int, default=8, help="cfg3: clips per encoder call (data_prep/save_latents.py:183-206)")
    ap.add_argument("--enc-depth", type=int, default=PROD["encoder_depth"])
    ap.add_argument("--dec-depth", type=int, default=PROD["decoder_depth"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-optimizer", action="store_true")
    ap.add_argument("--split-allreduce", action="store_true",
                    help="N > 1: reduce the decoder's gradients on a side stream while the encoder's backward still runs "
                         "inside the graph (ddp.SplitAllReduce); combine with NCCL_MAX_CTAS=<n> in the environment to keep "
                         "the collective off most SMs")
    ap.add_argument("--graph-allreduce", action="store_true", help="N > 1, EXPERIMENTAL: capture the bucketed all-reduce "
                    "inside the graph instead of one all-reduce after it (hung at N=2 with NCCL 2.28.9 in round 1)")
    ap.add_argument("--grad-comm", default="fp32", choices=["fp32", "bf16"], help="N > 1: dtype of the gradient all-reduce "
                    "(bf16 halves the bytes on NVLink; the fp32 default matches the reference's XLA all-reduce)")
    ap.add_argument("--recompute", action="store_true", help="per-layer activation recompute in every FactoredAttention "
                    "(the reference's @nnx.remat, train/layers.py:209): one extra forward per layer, ~1.1 GB/layer less")
    ap.add_argument("--debug-set", action="append", default=[], metavar="KEY=VALUE",
                    help="A/B switch: vvae_debug_set(KEY, VALUE) before the run (keys in include/vvae.h); not a bench line")
    ap.add_argument("--wgrad-lane", action="store_true",
                    help="A/B switch: weight-gradient kernels on a second stream (measured: no gain; off by default)")
    ap.add_argument("--no-pdl", action="store_true", help="launch every kernel fully serialized (vvae_debug_set(11, 1)) "
                    "instead of with programmatic dependent launch")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python each step (eager) instead "
                    "of replaying the captured CUDA graph; N > 1 then overlaps the bucketed all-reduce with backward")
    ap.add_argument("--profile-kernels", action="store_true", help="print the per-kernel-class time table to stderr")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    a.custom_shape = any(v is not None for v in (a.batch, a.frames, a.size))
    a.batch = c["batch"] if a.batch is None else a.batch
    a.frames = c["frames"] if a.frames is None else a.frames
    a.size = c["size"] if a.size is None else a.size
    return a


def keep_per_clip(args):
    """cfg4: frames kept per clip of the batch (prefix masks, train/dataloader.py:232-234)."""
    if args.config != "cfg4":
        return [args.frames] * args.batch
    if args.keep is not None:
        return [args.keep] * args.batch
    q = args.frames // 4
    return [(q * (1 + i % 4)) for i in range(args.batch)]


def workload_text(args, impl_note=""):
    c = CONFIGS[args.config]
    shape = f"{args.frames}x{args.size}x{args.size} RGB clips, batch {args.batch}/GPU"
    hyper = f"enc {args.enc_depth}/dec {args.dec_depth}, mlp 1536, 8 heads x 64, latent 96"
    if args.config == "cfg3":
        w = f"Encoder forward (eval, no grad, chunks of {args.chunk}) on {shape}, {hyper}"
    else:
        w = f"VideoVAE train step (fwd+loss+bwd{'' if args.no_optimizer else '+clip+Adam'}) on {shape}, {hyper}"
        if args.config == "cfg4":
            w += f", prefix masks keeping {sorted(set(keep_per_clip(args)))} of {args.frames} frames"
    return w + f" (BASELINE.json {c['row']}{', non-default shape' if args.custom_shape else ''}){impl_note}"


# ----------------------------------------------------------------------------------------------- CPU oracle arm
def oracle_step_time(args, steps, warmup, threads, budget_s=None):
    """One clip of the arm's configuration per step on the host cores through the CPU oracle (fp32): fwd + loss + bwd +
    clip + Adam for the training rows, the eval-mode Encoder forward for cfg3.  With `budget_s` the loop stops early
    (never before one timed step) so that a slow host still ends in time; returns (seconds per step, steps timed)."""
    import torch
    from oracle import Rngs as ORngs
    from oracle.losses import DEFAULT_HPARAMS, loss_fn
    from oracle.model import VideoVAE as OVAE
    from oracle.optim import ClipAdam
    torch.set_num_threads(threads)
    size, frames = args.size, args.frames
    m = OVAE(size, size, 3, PROD["patch_size"], args.enc_depth, args.dec_depth, PROD["mlp_dim"], PROD["num_heads"],
             PROD["qkv_features"], PROD["max_temporal_len"], PROD["spatial_compression_rate"],
             PROD["unembedding_upsample_rate"], ORngs(2))
    with torch.no_grad():
        m.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator().manual_seed(7))
    params = list(m.parameters())
    opt = None if (args.no_optimizer or args.config == "cfg3") else ClipAdam(params, lr=5e-5, clip=1.0)
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(1, frames, size, size, 3, generator=g)
    mask = torch.arange(frames)[None, :] < keep_per_clip(args)[0]
    hp = dict(DEFAULT_HPARAMS, gamma4=0.1)
    def one(i):
        t0 = time.perf_counter()
        if args.config == "cfg3":
            with torch.no_grad():
                m.encoder(x, mask[:, None, None, :], ORngs(0), train=False)
        else:
            for p in params:
                p.grad = None
            loss, _ = loss_fn(m, x, mask[:, None, None, :], mask, ORngs(i), hp)
            loss.backward()
            if opt is not None:
                opt.step([p.grad if p.grad is not None else torch.zeros_like(p) for p in params])
        return time.perf_counter() - t0

    t_start, times = time.perf_counter(), []
    for i in range(warmup):
        dt = one(i)
        if budget_s is not None and (time.perf_counter() - t_start) + (steps + 1) * dt > budget_s:
            break                               # host slower than planned: stop warming up, start timing
    for i in range(steps):
        dt = one(warmup + i)
        times.append(dt)
        if budget_s is not None and (time.perf_counter() - t_start) + dt > budget_s:
            break
    return sum(times) / len(times), len(times)


def cpu_sample_text(args):
    what = "eval-mode Encoder forward" if args.config == "cfg3" else \
        ("fwd+loss+bwd" + ("" if args.no_optimizer else "+clip+Adam"))
    return f"1 clip of {args.frames}x{args.size}x{args.size} per step (the arm's own clip shape), fp32, {what}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = cpu_sample_text(args)
    t, n_timed = oracle_step_time(args, args.steps, args.warmup, threads, budget_s=280.0)
    value = 1.0 / t
    if n_timed < args.steps:
        sample += f"; host time budget reached: {n_timed} of {args.steps} steps timed"
    line = {
        "impl": "reference", "metric": CONFIGS[args.config]["metric"], "value": value, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "same_config": True, "steps_timed": n_timed,
        "config": {"workload": workload_text(args, "; reference arm = CPU oracle port of the JAX/Flax reference (JAX is "
                                                   "not installable offline), fp32, all host cores, one clip per step"),
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, val in zip(names, c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import video_vae_b200 as V
    from video_vae_b200 import _ffi, ops
    from video_vae_b200.ddp import FlatAdam, FlatParams, GradAllReducer
    from video_vae_b200.graph import GraphedTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; keep stdout for the ONE JSON line.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    _ffi.require_device()
    if args.no_pdl:
        _ffi.lib.vvae_debug_set(11, 1)
    if args.wgrad_lane:
        from video_vae_b200 import functional as _F
        _F.WGRAD_LANE = True
    for kv in args.debug_set:
        k, v = kv.split("=")
        _ffi.lib.vvae_debug_set(int(k), int(v))

    S, Tn, B = args.size, args.frames, args.batch
    model = V.VideoVAE(S, S, 3, PROD["patch_size"], args.enc_depth, args.dec_depth, PROD["mlp_dim"], PROD["num_heads"],
                       PROD["qkv_features"], PROD["max_temporal_len"], PROD["spatial_compression_rate"],
                       PROD["unembedding_upsample_rate"], V.Rngs(2), dtype=torch.bfloat16, device=dev)
    with torch.no_grad():   # exercise the UNet backward (the reference zero-inits final_conv: SURVEY.md 7.3)
        model.decoder.unet.final_conv.kernel.normal_(0.0, 0.02, generator=torch.Generator(device=dev).manual_seed(7))
    if args.recompute:
        model.set_recompute(True)
    flat = FlatParams(model)
    flat.enable_bf16_shadow()
    flat.broadcast(src=0)     # rank 0's weights everywhere (distributed_train.py:339 broadcast_one_to_all); untimed
    encode_only = args.config == "cfg3"
    reducer = GradAllReducer(flat) if (world > 1 and not encode_only) else None
    opt = None if (args.no_optimizer or encode_only) else FlatAdam(flat, lr=5e-5)
    hp = dict(V.DEFAULT_HPARAMS, gamma4=0.1)   # MSE + selection + KL + MAE terms

    g = torch.Generator().manual_seed(1234 + rank)
    host_video = torch.rand(B, Tn, S, S, 3, generator=g).to(torch.bfloat16).pin_memory()   # reference casts to bf16 on host
    keep = torch.tensor(keep_per_clip(args))
    host_mask = (torch.arange(Tn)[None, :] < keep[:, None]).contiguous().pin_memory()
    video = host_video.to(dev, non_blocking=True)
    mask = host_mask.to(dev, non_blocking=True)
    rngs = V.Rngs(3 + rank)

    graphed, ar_in_graph = None, False
    if not args.no_graph and not encode_only:
        if reducer is not None and args.graph_allreduce:
            try:      # bucketed all-reduce captured inside the graph: overlaps the rest of backward
                graphed = GraphedTrainStep(model, flat, video, mask, hp, reducer=reducer)
                ar_in_graph = True
            except Exception as e:  # noqa: BLE001  (capture of NCCL work refused: fall back to one all-reduce after the graph)
                print(f"[bench] all-reduce capture failed ({type(e).__name__}: {str(e)[:200]}); using a post-graph all-reduce",
                      file=sys.stderr)
                torch.cuda.synchronize()
        if graphed is None:
            graphed = GraphedTrainStep(model, flat, video, mask, hp, mark_decoder_done=bool(world > 1 and args.split_allreduce))
    split_ar = None
    if world > 1 and args.split_allreduce and graphed is not None and not ar_in_graph:
        from video_vae_b200.ddp import SplitAllReduce
        split_ar = SplitAllReduce(flat, model)

    def eager_step(v, m):
        flat.zero_grad()
        if reducer:
            reducer.start_step()
        loss, aux = V.loss_fn(model, v, m[:, None, None, :], m, rngs, hp, train=True)
        loss.backward()
        if reducer:
            reducer.finish_step()
        return loss

    def encode_step(v, m):
        """cfg3: the loop of data_prep/save_latents.py:183-206 -- chunked eval-mode Encoder calls, no grad; returns a
        device scalar (sum of the latents) standing in for the step's result that is read back."""
        acc = None
        with torch.no_grad():
            for i in range(0, B, args.chunk):
                mean, _, _ = model.encoder(v[i:i + args.chunk], m[i:i + args.chunk, None, None, :], V.Rngs(0), train=False)
                part = mean.float().sum()
                acc = part if acc is None else acc + part
        return acc

    def step(v, m):
        """One training step.  Default: replay the captured graph (zero-grad + fwd + loss + bwd: one launch), then the
        gradient all-reduce (N > 1) and the fused clip + Adam kernels."""
        if encode_only:
            return encode_step(v, m)
        if graphed is not None:
            loss = graphed(v, m, rngs)
            if world > 1 and split_ar is not None:
                split_ar(graphed.decoder_done)
            elif world > 1 and not ar_in_graph:
                flat.all_reduce_grads(torch.bfloat16 if args.grad_comm == "bf16" else torch.float32)
        else:
            loss = eager_step(v, m)
        if opt:
            opt.step(grad_scale=1.0 / world)
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(video, mask)
    sync_all()

    # ---- device-resident timed region, with per-launch CUDA events on the dominant kernel class (GEMM)
    clocks = ClockSampler(local) if rank == 0 else None
    ops.PROFILE = [] if graphed is None else None
    calls0 = _ffi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        step(video, mask)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # CPU time to ENQUEUE a step (no sync inside)
    e1.record()
    sync_all()
    calls = _ffi.launch_count - calls0
    prof, ops.PROFILE = ops.PROFILE, None
    clock_info = clocks.stop() if clocks else None
    kernels_per_step = None
    if graphed is not None:
        # The graph hides individual launches from CUDA events: time the SAME kernels with per-launch events in one
        # eager pass over the same batch (kernel durations do not depend on how they were enqueued).
        ops.PROFILE = []
        calls_e0 = _ffi.launch_count
        eager_step(video, mask)      # NB: runs after the graph was captured (capture must precede eager backward)
        torch.cuda.synchronize()
        kernels_per_step = _ffi.launch_count - calls_e0
        prof, ops.PROFILE = ops.PROFILE, None
        prof_steps = 1
        if args.profile_kernels and rank == 0:       # device time of every C-ABI entry point over one eager step
            with _ffi.AbiProfile() as ap:
                eager_step(video, mask)
                if opt:
                    opt.step(grad_scale=1.0 / world)
            torch.cuda.synchronize()
            tot = sum(t for _, _, t in ap.table())
            for name, calls_, t in ap.table():
                print(json.dumps({"abi_entry": name, "calls_per_step": calls_, "ms_per_step": round(t, 4),
                                  "share_of_abi_time": round(t / tot, 4)}), file=sys.stderr)
    else:
        prof_steps = args.steps
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = tms.item()
    value = world * B / (ms * 1e-3)

    # ---- end-to-end: pinned host batch -> device, step, loss -> host, every step
    e0.record()
    last = 0.0
    for _ in range(args.steps):
        if graphed is not None:        # pinned host -> the graph's static input buffers, replay, loss -> host
            last = step(host_video, host_mask).item()
        else:
            v = host_video.to(dev, non_blocking=True)
            m = host_mask.to(dev, non_blocking=True)
            last = step(v, m).item()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    if world > 1:
        tms = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = tms.item()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class
    classes = {}
    for name, flops, nbytes, ev0, ev1 in prof:
        t = ev0.elapsed_time(ev1) * 1e-3
        c = classes.setdefault(name, [0.0, 0.0, 0.0, 0])
        c[0] += t; c[1] += flops; c[2] += nbytes; c[3] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    traffic, traffic_detail = None, None
    try:   # DRAM bytes of one representative launch of the dominant kernel, from the committed `ncu --set full` capture
        nc = json.load(open(os.path.join(ROOT, "profiles", "r02r_gemm_qkv_fwd_ncu.json")))
        # bytes (dram__bytes_read.sum + dram__bytes_write.sum) of that one launch; the capture reports Mbyte
        traffic = (float(nc["dram__bytes_read.sum"][0]) + float(nc["dram__bytes_write.sum"][0])) * 1e6
        traffic_detail = {
            "launch": "QKV projection forward, M=32768 N=1536 K=768 (algorithmic 153.4 MB: A 50.3 + B 2.4 + C 100.7; "
                      "half of C is still L2-resident when the launch ends); in the step this launch also writes the "
                      "67 MB of rope(LN(q))|rope(LN(k)) from its epilogue",
            "algorithmic_bytes_per_launch": 153.4e6,
            "tensor_pipe_active_pct": float(nc["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]),
            "source": "profiles/r02r_gemm_qkv_fwd_ncu.json"}
    except (OSError, KeyError, ValueError):
        traffic, traffic_detail = None, None
    roofline = None
    table = []
    for name, (t, fl, nb, n) in sorted(classes.items(), key=lambda kv: -kv[1][0]):
        table.append({"kernel": name, "launches_per_step": n / prof_steps, "ms_per_step": t * 1e3 / prof_steps,
                      "share_of_step": t * 1e3 / prof_steps / ms, "tflops": fl / t / 1e12 if t > 0 else 0.0})
    if args.profile_kernels:
        for row in table:
            print(json.dumps(row), file=sys.stderr)
    if "gemm_tcgen05" in classes:
        t, fl, nb, n = classes["gemm_tcgen05"]
        ach = fl / t / 1e12
        roofline = {"kernel": "gemm_sm100_kernel (tcgen05 bf16 GEMM: every Linear fwd/dgrad/wgrad)", "bound": "tensor",
                    "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback", "traffic": traffic, "traffic_detail": traffic_detail,
                    "launches_per_step": n / prof_steps,
                    "share_of_step": t * 1e3 / prof_steps / ms,
                    "timing": "CUDA events around every launch of this kernel in one eager pass after the timed region"
                    if graphed is not None else "CUDA events around every launch inside the timed region"}

    flop_per_clip = None if args.custom_shape else CONFIGS[args.config]["flop"]
    if args.custom_shape and args.config == "cfg2" and Tn == 16:
        flop_per_clip = FLOP_PER_CLIP.get(S)
    line = {
        "metric": CONFIGS[args.config]["metric"], "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_text(args), "global_batch": world * B, "parallelism": f"dp{world}",
                   "l2": "inputs+activations per step (tens of GB) >> 126 MB L2; no flush needed",
                   "algorithmic_flop_per_clip": flop_per_clip,
                   "frames_kept_per_clip": keep_per_clip(args) if args.config == "cfg4" else None},
        "model_tflops": (value * flop_per_clip / 1e12 / world) if flop_per_clip else None,
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "clips/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": host_video.numel() * host_video.element_size() + host_mask.numel(),
                "d2h_bytes_per_step": 4, "last_loss": last},
        "host_enqueue_ms_per_step": host_ms,
        "gpu_launches": (kernels_per_step * args.steps + calls) if graphed is not None else calls,
        "gpu_launches_note": ("libvvae kernel-launching C-ABI calls executed in the timed region: those captured in the "
                              "replayed CUDA graph (counted in one eager pass) x steps + those enqueued directly")
        if graphed is not None else "libvvae C-ABI kernel-launching calls in the timed region",
        "execution": ("cuda-graph replay (zero-grad+fwd+loss+bwd" + ("+bucketed all-reduce" if ar_in_graph else "") +
                      ") + eager " + ("all-reduce + " if (world > 1 and not ar_in_graph) else "") + "optimizer")
        if graphed is not None else "eager",
        "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30,
        "grad_comm": (("split-fp32" if args.split_allreduce else args.grad_comm) if world > 1 else None), "recompute": bool(args.recompute),
        "clocks": clock_info, "roofline": roofline, "kernel_classes": table[:6],
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t, _ = oracle_step_time(args, 1, 0, threads)
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": "clips/s", "cores": threads, "kind": "port",
                                "sample": cpu_sample_text(args) + " (1 step, no warm-up)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
