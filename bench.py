#!/usr/bin/env python
"""bench.py — frame-stacks/sec of the MultiDimStacker FULL forward (15 x 1280 x 736 grayscale stack -> logits).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One "step" = one pass of the hot path (stem -> encoder x5 -> 3D blocks -> GeM -> classifier) over one batch of B
synthetic uint8 frame-stacks (BASELINE.json configs[1]: batch=4, fp16 storage / fp32 accumulate).  N>1: launched by
torchrun, one rank per GPU, every rank runs its own batch (weak scaling) and the per-step logits are all-gathered over
NCCL (the path's only exchange).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "frame-stacks/sec (15x1280x736)"
UNIT = "frame-stacks/s"
H, W, STORED_H, FRAMES = 736, 1280, 720, 15


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4, help="frame-stacks per GPU per step (configs[1] = 4, configs[2] = 32)")
    ap.add_argument("--chunk-images", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="full", choices=["full", "sweep"],
                    help="full: FULL forward (configs[1]/[2], the headline); sweep: SLIDING synthetic-match sweep (configs[3])")
    ap.add_argument("--sweep-frames", type=int, default=67500, help="frames per half (45 min x 25 fps)")
    ap.add_argument("--tta", type=int, default=0)
    ap.add_argument("--tail-mode", type=int, default=3, choices=[0, 1, 2, 3],
                    help="MBConv tails: 3 = TMA depthwise + SE kernel + GEMM (default), 2 = depthwise+SE kernel and gating GEMM, 1 = one fused launch, 0 = round-1 path")
    ap.add_argument("--conv-mode", type=int, default=2, choices=[1, 2],
                    help="dense 3x3 blocks: 2 conv_tc_kernel with the expanded tensor in TMEM, 1 the same with smem staging (A/B)")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2, 3, 4], help="encoder on one stream or as equal parts of the images on several streams (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3)
    a = ap.parse_args()
    a.steps_given = a.steps is not None
    if a.steps is None:
        a.steps = 20
    return a


# ------------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port (the reference is Python + un-vendored timm and cannot travel)
# ------------------------------------------------------------------------------------------------------------------
def cpu_forward_timer(steps: int, warmup: int):
    import torch
    from oracle import mds_oracle as O           # bench.py's cpu_baseline / --impl reference legs only
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=1234)
    x = torch.rand((1, FRAMES, H, W), generator=torch.Generator().manual_seed(0))
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.forward(sd, x, cfg)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return times, cores


def train_step_cpu_baseline(sd, frames: int, b: int, fhw, steps: int):
    """cpu_baseline leg of tools/bench_train.py (BASELINE.json configs[4]): the training-step oracle (reference modules
    restated, torch CPU fp32 autograd + SGD) on the box's host cores, same batch shape as the GPU step."""
    import torch
    from oracle import mds_oracle as O                  # baseline legs only
    from oracle import mds_train_oracle as TO
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.ModelConfig(num_frames=frames)
    enc, tg = TO.make_case(cfg, b, tuple(fhw), seed=7)
    dp, do = TO.make_masks(cfg, b, 0.2, 0.2, 11)
    bufs, times = {}, []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        _, _, grads, stats = TO.loss_and_grads(sd, enc, tg, cfg, dp, do)
        params = {k: sd[k] for k in grads}
        TO.sgd_nesterov_step(params, grads, bufs, 1e-3)
        sd.update(params)
        sd.update(stats)
        if i > 0:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return {"ms_per_step": ms, "frame_stacks_per_s": 1e3 * b / ms, "cores": cores, "kind": "port",
            "sample": f"{len(times)} timed + 1 warm-up 3D-only steps (fwd+bwd+SGD) of the same batch, torch CPU fp32 autograd"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    steps, warmup = min(steps, 5), min(warmup, 2)       # bounded: ~3 s per stack on 8 cores
    times, cores = cpu_forward_timer(steps, warmup)
    total = sum(times)
    value = len(times) / total
    sample = f"{len(times)} timed + {warmup} warm-up forwards of one (1,15,736,1280) fp32 stack, torch CPU, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
            "warmup": warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FULL forward, one 15x1280x736 stack per step (bounded sample of the B200 arm's workload)",
                       "frames": FRAMES, "height": H, "width": W},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for ln in out.splitlines():
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # keep samples under load (upper half) for the median
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ball_action_spotting_b200 import MultiDimStacker
    from ball_action_spotting_b200 import accounting as acc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = args.batch, args.steps, max(3, args.warmup)
    T = FRAMES // 3

    torch.manual_seed(1234)
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=FRAMES, stack_size=3, num_3d_blocks=4,
                          expansion_3d_ratio=3, se_reduce_3d_ratio=24, chunk_images=args.chunk_images)
    net.init_random_(seed=1234)
    net.to(dev).eval()
    eng = net.engine(dev)
    eng.lib.mds_set_tail_mode(args.tail_mode)
    eng.lib.mds_set_streams(args.streams)
    eng.lib.mds_set_conv_mode(args.conv_mode)

    # synthetic uint8 frames; R rotating batches so that consecutive steps never re-read the same input from L2
    bytes_in = B * FRAMES * STORED_H * W
    R = max(2, -(-160_000_000 // bytes_in))                 # R * bytes_in > L2 (126 MB)
    R = min(R, 8)
    g = torch.Generator(device="cpu").manual_seed(rank)
    host = [torch.randint(0, 256, (B, FRAMES, STORED_H, W), dtype=torch.uint8, generator=g).pin_memory() for _ in range(min(R, 3))]
    devb = [host[i % len(host)].to(dev, non_blocking=True) for i in range(R)]
    logits = torch.empty((B, 2), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * B, 2), dtype=torch.float32, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    def desc_of(t):
        return eng.frames_desc(t, H, W, 3 * STORED_H * W, STORED_H * W)

    def step_local(i):
        eng.forward(desc_of(devb[i % R]), B, out=logits)

    def step_resident(i):
        step_local(i)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()            # sampled every 50 ms from the warm-up to the end of the e2e pass
    for i in range(Wm):
        step_resident(i)
    eng.launch_count(reset=True)
    torch.cuda.profiler.start()        # `ncu --profile-from-start off` then lists exactly the timed launches
    ms = timed(step_resident, K)
    torch.cuda.profiler.stop()
    launches = eng.launch_count()
    value = world * B * K / (ms / 1e3)

    # ---- e2e: host (pinned) buffers -> H2D on a copy stream (NBUF device stages) -> forward -> D2H logits, every step ----
    # The copy of step i + 2 is queued as soon as the forward of step i - 1 (the last reader of that stage) has finished, so
    # every copy has two whole steps to complete: with 8 ranks sharing one host the copy takes ~80 % of a step.
    NBUF = 3
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [torch.empty_like(devb[0]) for _ in range(NBUF)]
    host_out = torch.empty((B, 2), dtype=torch.float32).pin_memory()
    ready = [torch.cuda.Event() for _ in range(NBUF)]
    done = [torch.cuda.Event() for _ in range(NBUF)]

    def upload(i):
        s = i % NBUF
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s])               # the forward that last read stage[s] has finished
            stage[s].copy_(host[i % len(host)], non_blocking=True)
            ready[s].record(copy_stream)

    state = {"next": 0, "limit": 1}

    def step_e2e(i):
        s = i % NBUF
        while state["next"] <= min(i + NBUF - 1, state["limit"] - 1):   # keep NBUF - 1 batches in flight ahead of the one being computed
            upload(state["next"]); state["next"] += 1
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready[s])
        eng.forward(desc_of(stage[s]), B, out=logits)
        done[s].record(cur)
        if world > 1:
            dist.all_gather_into_tensor(gathered, logits)
        host_out.copy_(logits, non_blocking=True)

    for ev in done:
        ev.record(torch.cuda.current_stream(dev))
    for i in range(2):
        state["next"], state["limit"] = 0, 1
        step_e2e(0)
        torch.cuda.synchronize()
    state["next"], state["limit"] = 0, K                  # exactly K uploads inside the timed region
    ms_e2e = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e / 1e3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event pass (same workload, same stream) for the roofline of the dominant kernel ----
    roofline, by_kind = None, {}
    if rank == 0:
        torch.cuda.synchronize()
        eng.profile_begin()
        PK = min(K, 5)
        for i in range(PK):
            step_local(i)          # rank 0 only: no collective inside this pass
        recs = eng.profile_end()
        per_kind_ms = {}
        for kind, tag, t in recs:
            per_kind_ms[kind] = per_kind_ms.get(kind, 0.0) + t
        tot = acc.per_stack_totals(H, W, STORED_H, T, tail_mode=0 if args.tail_mode == 3 else args.tail_mode)
        peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            j = json.loads(pk.read_text())
            peaks = {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "src": "measured"}
        traffic = {}
        tj = ROOT / "profiles" / "ncu_traffic.json"
        if tj.exists():
            traffic = json.loads(tj.read_text())
        total_ms = sum(per_kind_ms.values())
        for kind, kms in sorted(per_kind_ms.items(), key=lambda kv: -kv[1]):
            name = acc.KINDS[kind]
            a = tot.get(name, {"bytes": 0, "flops": 0, "launches": 1})
            sec = kms / 1e3 / (PK * B)                    # seconds per stack spent in this kernel kind
            gbs = a["bytes"] / sec / 1e9 if sec > 0 else 0.0
            tfs = a["flops"] / sec / 1e12 if sec > 0 else 0.0
            ai = a["flops"] / a["bytes"] if a["bytes"] else 0.0
            tensor_bound = name == "conv3x3"              # AI 250-640 FLOP/B, above the ~215 FLOP/B ridge
            entry = {"bound": "tensor" if tensor_bound else "hbm",
                     "achieved": tfs if tensor_bound else gbs,
                     "peak": peaks["bf16_tflops"] if tensor_bound else peaks["hbm_gbs"],
                     "unit": "TFLOP/s" if tensor_bound else "GB/s",
                     "share_of_step": kms / total_ms if total_ms else 0.0,
                     "ms_per_step": kms / PK, "algorithmic_bytes_per_stack": a["bytes"], "flops_per_stack": a["flops"],
                     "gbs": gbs, "tflops": tfs, "traffic": traffic.get(name)}
            entry["frac"] = entry["achieved"] / entry["peak"] if entry["peak"] else 0.0
            by_kind[name] = entry
        top = max(by_kind.items(), key=lambda kv: kv[1]["share_of_step"])
        roofline = {"kernel": top[0], "bound": top[1]["bound"], "achieved": top[1]["achieved"], "peak": top[1]["peak"],
                    "unit": top[1]["unit"], "frac": top[1]["frac"], "traffic": top[1]["traffic"],
                    "peak_source": peaks["src"], "timing": f"per-launch CUDA events on the launching stream, {PK} steps"}

    # ---- depthwise stage alone (SURVEY.md 8(d) "DW-only" bytes, the north_star's >= 60 % figure) ----
    roofline_dw = None
    if rank == 0 and by_kind:
        dwb = acc.depthwise_only_bytes(H, W, T)
        if any(k in by_kind for k in ("tail2d", "tail3d")):
            # fused build: the depthwise items of the fused tail kernels are timed by running the same launches with their
            # projection items switched off (mds_set_tail_dw_only); outputs of these passes are discarded
            eng.lib.mds_set_tail_dw_only(1)
            for i in range(2):
                step_local(i)
            torch.cuda.synchronize()
            eng.profile_begin()
            for i in range(PK):
                step_local(i)
            recs_dw = eng.profile_end()
            eng.lib.mds_set_tail_dw_only(0)
            step_local(0)
            torch.cuda.synchronize()
            ms2 = sum(t for k, _, t in recs_dw if k == 11) / PK
            ms3 = sum(t for k, _, t in recs_dw if k == 12) / PK
            how = "depthwise + SE items of the fused tail kernels, projection items off (mds_set_tail_dw_only)"
        else:
            ms2, ms3 = by_kind["dwconv2d"]["ms_per_step"], by_kind["dwconv3d"]["ms_per_step"]
            how = {3: "dwconv_tma kernels (TMA-staged depthwise + squeeze partials)", 2: "dwconv_tma kernels (depthwise + squeeze + SE MLP of the last CTA)",
                   0: "round-1 cp.async dwconv kernels"}[args.tail_mode]
        g2 = dwb["dwconv2d"] * B / (ms2 / 1e3) / 1e9 if ms2 > 0 else 0.0
        g3 = dwb["dwconv3d"] * B / (ms3 / 1e3) / 1e9 if ms3 > 0 else 0.0
        roofline_dw = {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_source": peaks["src"],
                       "achieved": g2, "frac": g2 / peaks["hbm_gbs"], "ms_per_step": ms2,
                       "bytes_per_stack": dwb["dwconv2d"], "what": "16 depthwise 3x3 layers x 5 images: read mid x H x W + write mid x Ho x Wo, fp16",
                       "dw3d": {"achieved": g3, "frac": g3 / peaks["hbm_gbs"], "ms_per_step": ms3, "bytes_per_stack": dwb["dwconv3d"]},
                       "timing": how + f", per-launch CUDA events, {PK} steps"}

    # ---- platform ceiling of the e2e leg: the same pinned host -> device copies alone, all ranks at once ----
    def h2d_only(i):
        s_ = i % NBUF
        with torch.cuda.stream(copy_stream):
            stage[s_].copy_(host[i % len(host)], non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(copy_stream)
    for i in range(2):
        h2d_only(i)
    ms_h2d = timed(h2d_only, K)
    h2d_gbs = world * bytes_in * K / (ms_h2d / 1e3) / 1e9

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        times, cores = cpu_forward_timer(args.cpu_steps, 1)
        cpu_baseline = {"value": len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{len(times)} timed + 1 warm-up forwards of one (1,15,736,1280) fp32 stack (oracle port, torch CPU)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
                "data": "synthetic",
                "config": {"workload": f"FULL MultiDimStacker forward, batch={B} stacks/GPU/step (BASELINE.json configs[1] is batch=4), "
                                       "uint8 15x720x1280 frames -> pad 736 -> logits", "batch_per_gpu": B, "frames": FRAMES,
                           "height": H, "width": W, "sharding": f"stacks x{world}, NCCL all-gather of logits" if world > 1 else "single GPU",
                           "l2": f"{R} rotating input batches ({R * bytes_in / 1e6:.0f} MB) + >1 GB of intermediates per step: inputs larger than L2",
                           "chunk_images": eng.cfg.chunk_images or 160},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": B * 2 * 4,
                        "ms_per_step": ms_e2e / K, "h2d_ceiling_gbs": h2d_gbs,
                        "h2d_ceiling_stacks_per_s": world * B * K / (ms_h2d / 1e3),
                        "h2d_ceiling_note": "the same pinned-host -> device copies with no compute, all ranks at once"},
                "gpu_launches": launches,
                "bias_correction": bool(net._bias_correction),
                "roofline": roofline, "roofline_dw": roofline_dw, "roofline_by_kind": by_kind, "cpu_baseline": cpu_baseline,
                "encoder_streams": args.streams,
                "mbconv_tail": {2: "dwconv_tma (depthwise + SE) + gating GEMM: 2 launches", 1: "mbconv_tail: 1 launch",
                                0: "round-1 path: dwconv + se_fc + gated GEMM: 3 launches",
                                3: "dwconv_tma + se_fc + pre-gated GEMM: 3 launches"}[args.tail_mode]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: full synthetic-match sweep (SLIDING semantics), sharded over the GPUs, one all-gather per half
# ------------------------------------------------------------------------------------------------------------------
def synthetic_video(dev, h=STORED_H, w=W, seed=1234, block=64):
    """Deterministic synthetic video: frame i is generated on the device from (seed, i // block), so a 45-minute half
    (62 GB of uint8 frames) never exists as a whole and every rank can produce exactly the frames of its shard."""
    import torch
    cache = {}

    def block_frames(bi):
        if bi not in cache:
            if len(cache) > 40:
                cache.pop(next(iter(cache)))
            g = torch.Generator(device=dev).manual_seed(seed * 100003 + bi)
            cache[bi] = torch.randint(0, 256, (block, h, w), dtype=torch.uint8, device=dev, generator=g)
        return cache[bi]

    def source(i0, i1):
        parts = []
        for bi in range(i0 // block, (i1 - 1) // block + 1):
            lo, hi = max(i0, bi * block), min(i1, (bi + 1) * block)
            parts.append(block_frames(bi)[lo - bi * block: hi - bi * block])
        return torch.cat(parts, 0).contiguous()
    return source


def run_sweep(args):
    """One step = one half of a match (args.sweep_frames frames, scripts/ball_action/predict.py:29-55 semantics: every
    frame index in [clip(0), clip(frame_count)] gets one prediction from its 15-frame window).  The prediction range of a
    half is split into `world` contiguous shards (28-frame halo), one all-gather of the (n, 2) probabilities per half."""
    import torch
    import torch.distributed as dist
    from ball_action_spotting_b200 import MultiDimStacker
    from ball_action_spotting_b200 import postprocess as PP
    from ball_action_spotting_b200.sweep import SlidingSweep, prediction_bounds, shard_range, sweep_video

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    halves = max(1, min(args.steps, 2)) if args.steps_given else 2
    Wm = max(3, args.warmup)
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=FRAMES, stack_size=3, num_3d_blocks=4,
                          expansion_3d_ratio=3, se_reduce_3d_ratio=24).init_random_(seed=1234).to(dev).eval()
    eng = net.engine(dev)
    sweep = SlidingSweep(net, FRAMES, 2, (W, H), tta=bool(args.tta), max_stacks=128)
    sources = [synthetic_video(dev, seed=1234 + hf) for hf in range(halves)]
    F = args.sweep_frames
    lo, hi = prediction_bounds(sweep.gen, F, 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(Wm):                                   # warm-up: short excerpts (also builds the NCCL communicator)
        sweep_video(sweep, sources[0], min(F, 400), 1, rank, world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    results = [sweep_video(sweep, sources[hf], F, 1, rank, world) for hf in range(halves)]
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    launches = eng.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    n_pred = (hi - lo + 1) * halves
    if rank == 0:
        frame_indexes, preds = results[0]
        # ---- sharded == unsharded, checked where it can differ: windows that straddle every shard boundary ----
        # (at world == 1 the same check runs against the boundaries an 8-way split would have: different batch composition)
        vw = world if world > 1 else 8
        checks, max_diff = [], 0.0
        for r in range(1, vw):
            bnd = shard_range(lo, hi, r, vw)[0]
            a, b = max(lo, bnd - 24), min(hi + 1, bnd + 24)
            f0, f1 = a - sweep.gen.behind, b - 1 + sweep.gen.ahead
            single = sweep.predict_range(sources[0](f0, f1 + 1), f0, a, b)
            d = float((single - preds[a - lo: b - lo]).abs().max())
            max_diff = max(max_diff, d)
            checks.append([a, b, d])
        actions = PP.raw_predictions_to_actions(frame_indexes, preds, {"PASS": 0, "DRIVE": 1},
                                                {"gauss_sigma": 3.0, "height": 0.2, "distance": 15}, device=str(dev))
        host_preds = preds.cpu()
        value = n_pred / (ms / 1e3)
        line = {"metric": "SLIDING predictions/sec (15x1280x736 window, step 2)", "value": value, "unit": "predictions/s",
                "n_gpus": world, "steps": halves, "warmup": Wm, "ms_per_step": ms / halves, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
                "config": {"workload": f"BASELINE.json configs[3]: synthetic match, {halves} halves x {F} uint8 720x1280 frames generated on the "
                                       f"device per block, SLIDING semantics, tta={bool(args.tta)}", "frames_per_half": F,
                           "predictions_per_half": hi - lo + 1, "sharding": f"prediction range x{world}, 28-frame halo, one all-gather of (n, 2) per half",
                           "l2": "every frame is read once; 2048-frame buffers (1.9 GB) exceed L2"},
                "video_fps_equivalent": F * halves / (ms / 1e3), "clocks": clocks,
                "e2e": None, "e2e_note": "configs[3] synthesises frames on the device, so this mode has no host->device leg; "
                                          f"the gathered predictions ({host_preds.numel() * 4} bytes per half) are copied to the host after the timed region",
                "gpu_launches": launches,
                "sharded_equals_unsharded": {"boundaries_checked": len(checks), "window": 48, "max_abs_diff": max_diff,
                                             "bit_identical": max_diff == 0.0,
                                             "how": "rank 0 recomputes, unsharded, the 48 predictions around every shard boundary "
                                                    + ("of this run" if world > 1 else "an 8-way split would have")},
                "spots": {k: len(v[0]) for k, v in actions.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sweep":
        run_sweep(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
