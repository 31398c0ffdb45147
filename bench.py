#!/usr/bin/env python
"""Headline benchmark: images/sec of one ViT-B/32 KAdaptation fine-tune step (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [...]                         # reference algorithm on the host CPU

A step = zero_grad + forward + CrossEntropy + backward + (N>1: one NCCL all-reduce of the flat
adapter-gradient buffer) + SGD, on one batch of synthetic 224x224 images per GPU (weak scaling,
256 images per GPU = BASELINE configs[1]).  One JSON line is printed by rank 0.

  value    : whole-job images/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e      : same step driven from pinned HOST buffers: per-step H2D copy of the image batch and
             labels (prefetched on a copy stream, inside the timed region) and a D2H read of the loss
  roofline : dominant kernel class of the step (by device time, CUDA events around every launch of
             this library inside the timed region) against the measured peaks in MEASURED_PEAKS.json
  cpu_baseline : the oracle (CPU restatement of the reference algorithm) timed on the host cores
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "images/sec ViT-B/32 KAdaptation step at 1/2/4/8 B200; logits max-abs-err"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="kadaptation", choices=["kadaptation", "lora", "adapter", "compacter"])
    ap.add_argument("--model", default="vit_b32", choices=["vit_b32", "vit_b16", "vit_l14"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fixed images per step over ALL GPUs (per-GPU batch = global / N); 0 = weak "
                         "scaling with --batch images per GPU (the driver's default)")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="images per CPU-baseline step (0 = the full --batch: same configuration as the GPU arm)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: cpu = the bench contract's CPU arm; cuda = the same reference algorithm as "
                         "eager PyTorch on one B200 (north_star's >=5x denominator; diagnostics, not the contract arm)")
    ap.add_argument("--ref-autocast", action="store_true", help="--ref-device cuda under torch.autocast(bfloat16)")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--no-parity-probe", action="store_true")
    ap.add_argument("--no-text-tower", action="store_true")
    ap.add_argument("--no-peer-exchange", action="store_true",
                    help="N > 1: NCCL all-reduce + SGD kernel instead of the fused one-shot peer-memory all-reduce + SGD")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--pixels", default="bf16", choices=["f32", "bf16", "u8"],
                    help="format of the image batch on the host and in HBM: f32 = the reference data loader's output; "
                         "bf16 = the same values rounded on the host (encode_image accepts any float dtype, "
                         "model.py:1152; the stem's GEMM operand is bf16 either way: bit-identical step, half the H2D "
                         "bytes); u8 = raw pixels, ToTensor + Normalize evaluated inside the stem kernel")
    return ap.parse_args()


def shape_of(name):
    from pevit_b200 import synth
    return {"vit_b32": synth.VIT_B32, "vit_b16": synth.VIT_B16, "vit_l14": synth.VIT_L14}[name]


def workload(args, shape):
    # names the workload (tier contract: a "workload" description, not the ML-style model / seq_len keys)
    return {"workload": f"{args.model} CLIP + {args.method} fine-tune step, synthetic 224x224, "
                        f"batch {args.batch}/GPU", "backbone": args.model, "peft": args.method,
            "images_per_gpu_step": args.batch, "images_per_step": args.batch * args.gpus, "tokens": shape.tokens,
            "width": shape.vision_width, "layers": shape.vision_layers, "parallelism": f"dp{args.gpus}",
            "pixels": {"f32": "fp32 (reference data loader output)",
                       "bf16": "bf16 on the host and in HBM (any float dtype is the reference's contract, model.py:1152; "
                               "the stem rounds pixels to bf16 anyway: bit-identical step)",
                       "u8": "uint8 on the host and in HBM, ToTensor + Normalize inside the stem kernel"}[args.pixels],
            "l2_policy": f"inputs larger than L2 ({args.batch * 3 * shape.image_resolution ** 2 * PIXEL_BYTES[args.pixels] / 1e6:.0f} MB "
                         f"{args.pixels} image batch and GBs of saved activations per step stream through the 126 MB L2)"}


PIXEL_BYTES = {"f32": 4, "bf16": 2, "u8": 1}
CLIP_NORM = ((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))   # CLIP's Normalize


def to_pixels(images, kind):
    """fp32 normalised synthetic images -> the requested host / HBM format (u8: de-normalised and quantised)."""
    import torch
    if kind == "f32":
        return images
    if kind == "bf16":
        return images.to(torch.bfloat16)
    mean = torch.tensor(CLIP_NORM[0], device=images.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_NORM[1], device=images.device).view(1, 3, 1, 1)
    return ((images * std + mean).clamp(0, 1) * 255).round().to(torch.uint8)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm (oracle)
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core it may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def oracle_step_fn(args, shape, device="cpu", autocast=False):
    """One fwd + CE + bwd step of the reference algorithm (oracle/pevit_oracle.py: materialised Kronecker sums, dense
    delta -- kadaptation_clip.py:347-352) on `device`; returns seconds per call (device-synchronised)."""
    from oracle import pevit_oracle as O
    use_all_host_threads()
    from pevit_b200 import synth
    dev = torch.device(device)
    p = dict(synth.clip_state_dict(shape, seed=0))
    O.init_adapters(p, args.method, seed=0)
    synth.randomize_adapters(p.items(), seed=1)
    p = {k: v.to(dev) for k, v in p.items()}
    g = torch.Generator().manual_seed(4)
    hw = (torch.randn(10, shape.embed_dim, generator=g) * shape.embed_dim ** -0.5).to(dev)
    hb = torch.zeros(10, device=dev)

    def step(n, seed):
        img = synth.images(n, shape.image_resolution, seed=seed).to(dev)
        lab = synth.labels(n, 10, seed=seed + 1).to(dev)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
            _, loss, _ = O.train_step_grads(img, lab, p, hw, hb, args.method)
        float(loss)  # the reference reads the loss back every step (kadaptation_clip.py:354)
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
        return time.perf_counter() - t0
    return step


def cpu_batch_of(args) -> int:
    return args.cpu_batch if args.cpu_batch > 0 else args.batch


def cpu_baseline(args, shape, steps=2, warmup=1) -> dict:
    step = oracle_step_fn(args, shape)
    nb = cpu_batch_of(args)
    for i in range(warmup):
        step(min(4, nb), 100 + i)
    ts = [step(nb, 200 + i) for i in range(steps)]
    return {"value": nb / statistics.median(ts), "unit": "images/s", "cores": torch.get_num_threads(),
            "kind": "port", "host_cpus": os.cpu_count(), "same_config": nb == args.batch,
            "sample": f"{steps} fwd+bwd step(s) of {nb} images, {args.model} {args.method}, fp32, "
                      "oracle/pevit_oracle.py (reference algorithm: materialised Kronecker sums)"}


def gpu_eager_baseline(args, shape, dev, steps=3, warmup=2) -> dict:
    """north_star's denominator ("the reference's own 1-GPU images/sec"): the reference ALGORITHM as eager PyTorch on
    this B200 -- fp32 with TF32 off (the reference's default) and under autocast(bf16).  The Python reference itself
    cannot travel to the GPU box; the oracle is its restatement (same op sequence, same intermediates)."""
    out = {"kind": "port (oracle/pevit_oracle.py as eager PyTorch on cuda)", "images_per_step": args.batch,
           "steps": steps, "unit": "images/s"}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for key, autocast in (("fp32", False), ("autocast_bf16", True)):
            step = oracle_step_fn(args, shape, device=dev, autocast=autocast)
            for i in range(warmup):
                step(args.batch, 100 + i)
            ts = [step(args.batch, 200 + i) for i in range(steps)]
            out[key] = args.batch / statistics.median(ts)
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out


def inference_throughput(args, tuner, images, iters=10, warmup=3) -> dict:
    """the validate / feature-extraction path (kadaptation_clip.py:376-410 ``validate``; feature.py): the
    same fused blocks under ``torch.no_grad()`` -- ``save = 0`` in the block descriptor, no activation is kept."""
    with torch.no_grad():
        for _ in range(warmup):
            tuner(images)
        torch.cuda.synchronize()
        mem0 = torch.cuda.memory_allocated()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = tuner(images)
        e1.record()
        torch.cuda.synchronize()
        kept = torch.cuda.memory_allocated() - mem0
    ms = e0.elapsed_time(e1) / iters
    return {"images_per_s": images.shape[0] / ms * 1e3, "ms_per_batch": ms, "batch": int(images.shape[0]),
            "bytes_kept_after_forward": int(kept), "logits_bytes": int(out.numel() * out.element_size()),
            "what": "Classifier forward (backbone.encode_image + head) under torch.no_grad(), eager launches"}


def text_tower_throughput(dev, n=1024, iters=5, warmup=2) -> dict:
    """``encode_text`` of the text tower every OpenAI CLIP ViT-B checkpoint carries (width 512, 8 heads, 12
    layers, context 77) on the fused blocks (method plain + causal mask), beside the same module on its stock PyTorch
    path (fp32 nn.MultiheadAttention: what the reference runs) on the same GPU, and the difference of the two."""
    import pevit_b200
    from pevit_b200 import synth
    shape = synth.TEXT_B32
    model = pevit_b200.build_model(dict(synth.clip_state_dict(shape, seed=7))).to(dev)
    text = synth.prompts(n, shape.context_length, shape.vocab_size, seed=5).to(dev)

    def timed():
        with torch.no_grad():
            for _ in range(warmup):
                out = model.encode_text(text)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                out = model.encode_text(text)
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters, out.float()
    ms_fused, feat = timed()
    for blk in model.transformer.resblocks:
        blk._pevit_causal = 0            # stock nn.MultiheadAttention path
    ms_stock, ref = timed()
    return {"prompts_per_s": n / ms_fused * 1e3, "ms_per_batch": ms_fused, "batch": n, "context": shape.context_length,
            "width": shape.transformer_width, "layers": shape.transformer_layers,
            "stock_pytorch_fp32_prompts_per_s": n / ms_stock * 1e3, "speedup_vs_stock": ms_stock / ms_fused,
            "features_rel_err_vs_stock_fp32": ((feat - ref).abs().max() / ref.abs().max()).item()}


def parity_probe(args, shape, tuner, dev, n=16) -> dict:
    """BASELINE.json's second metric: logits max-abs-err of this path vs the reference algorithm (oracle, fp32, CPU) on
    the same weights and the same n-image batch (F4 couples the samples of a batch, so both sides see the same n)."""
    from oracle import pevit_oracle as O
    from pevit_b200 import synth
    p = {k: v.detach().float() for k, v in tuner.backbone.named_parameters()}
    p.update({k: v.detach().float() for k, v in tuner.backbone.named_buffers()})
    hw, hb = tuner.head.weight.detach().float(), tuner.head.bias.detach().float()
    img = synth.images(n, shape.image_resolution, seed=77).to(dev)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = O.classifier_logits(img, p, hw, hb, args.method).float()         # reference algorithm, fp32
            with torch.autocast("cuda", dtype=torch.bfloat16):
                ref16 = O.classifier_logits(img, p, hw, hb, args.method).float()   # the reference's own bf16 noise floor
            got = tuner(img).float()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    scale = ref.abs().max().item()
    err, floor = (got - ref).abs().max().item(), (ref16 - ref).abs().max().item()
    l2 = ((got - ref).norm() / ref.norm()).item()
    return {"max_abs_err": err, "rel_err": err / scale, "rel_l2": l2, "ref_max_abs": scale,
            "reference_autocast_bf16_max_abs_err": floor, "reference_autocast_bf16_rel_err": floor / scale,
            "ratio_to_reference_bf16_floor": err / floor if floor > 0 else None,
            "sample": f"{n} images, weights of this run after the timed steps; reference algorithm (oracle) in fp32 on "
                      "the same GPU with TF32 off vs this bf16 path, and vs the reference under autocast(bf16)",
            "tolerance": "SURVEY 7.6: rel <= 1e-2 and <= 1.5 x the reference-autocast-bf16 error"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = shape_of(args.model)
    nb = cpu_batch_of(args)
    on_gpu = args.ref_device == "cuda"
    if on_gpu:
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False  # reference default
    step = oracle_step_fn(args, shape, device=args.ref_device, autocast=args.ref_autocast)
    for i in range(args.warmup):
        step(nb, 100 + i)
    ts = [step(nb, 200 + i) for i in range(args.steps)]
    total = sum(ts)
    ips = nb * args.steps / total
    cfg = workload(args, shape)
    cfg["sample"] = (f"each step = the full {nb}-image batch" if nb == args.batch else
                     f"each step = {nb} images (bounded sample of the {args.batch}-image batch)")
    cfg["reference_pixels"] = "fp32 (what the oracle consumes; `pixels` above is the b200 arm's transport format of the same batch)"
    cfg["reference_device"] = ("cuda eager PyTorch, " + ("autocast bf16" if args.ref_autocast else "fp32, TF32 off")
                               if on_gpu else "host CPU")
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16-autocast" if (on_gpu and args.ref_autocast) else "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                             "host_cpus": os.cpu_count(), "same_config": nb == args.batch,
                             "sample": f"{args.steps} fwd+bwd steps of {nb} images on "
                                       f"{'one B200 (eager PyTorch)' if on_gpu else 'the host CPU'}, "
                                       "oracle/pevit_oracle.py (the Python reference cannot travel to the GPU box)"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- B200 arm
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return d, "measured (MEASURED_PEAKS.json)"
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


def gemm_flops(cls: str, M: int, D: int, r2: int) -> float:
    W3 = 3 * D + r2
    dims = {"gemm_qkv": (W3, D), "gemm_out": (D, D), "gemm_fc": (4 * D, D), "gemm_proj": (D, 4 * D),
            "gemm_dproj": (4 * D, D), "gemm_dfc": (D, 4 * D), "gemm_dout": (D, D), "gemm_dqkv": (D, W3),
            "gemm_dT": (r2 // 2, D), "gemm_delta": (D, r2), "gemm_stem": (D, 3072)}
    n, k = dims[cls]
    return 2.0 * M * n * k


def attn_bytes(cls: str, L: int, NB: int, D: int, H: int, r: int) -> float:
    # per image per layer, bf16 q,k,v,o in HBM
    per_img = (8 * L * D + 4 * L * r + 4 * L * H) if cls == "attn_fwd" else (16 * L * D + 8 * L * r + 4 * L * H)
    return float(per_img) * NB


def run_b200(args):
    import torch.distributed as dist
    from pevit_b200 import _lib, engine
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    _lib.check(lib.pevit_check_device(), "pevit_check_device")
    shape = shape_of(args.model)
    tuner = engine.FineTuner(args.method, shape, device=dev, distributed=world > 1, seed=0,
                             pixel_norm=CLIP_NORM if args.pixels == "u8" else None,
                             peer_exchange=False if args.no_peer_exchange else None)
    N, R = args.batch, shape.image_resolution
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    images = to_pixels(torch.randn(N, 3, R, R, device=dev, generator=g), args.pixels)
    labels = torch.randint(0, 10, (N,), device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        tuner.step(images, labels)
    torch.cuda.synchronize()
    # whole-step CUDA graph (eager fallback if capture is not possible in this configuration)
    graphed = False
    if not args.no_graph:
        try:
            tuner.capture(images, labels, slots=2)   # two graphs over one pool: ping-pong input buffers for e2e
            for _ in range(2):
                tuner.step_graphed()
            torch.cuda.synchronize()
            graphed = True
        except Exception as exc:  # pragma: no cover - depends on the runtime (e.g. NCCL capture support)
            print(f"[bench] CUDA-graph capture unavailable, timing the eager step: {exc!r}", file=sys.stderr)
            torch.cuda.synchronize()
    run_step = (lambda: tuner.step_graphed()) if graphed else (lambda: tuner.step(images, labels))
    clocks = ClockSampler(local) if rank == 0 else None

    # ---- device-resident timing (value)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_cpu0 = time.perf_counter()
    for _ in range(args.steps):
        run_step()
    cpu_ms_per_step = 1e3 * (time.perf_counter() - t_cpu0) / args.steps  # host time to ENQUEUE a step
    e1.record()
    torch.cuda.synchronize()
    ms_total = reduce_max(e0.elapsed_time(e1))
    barrier()

    # ---- per-kernel-class device time: the same K steps again, eagerly, every launch of this library bracketed by
    # CUDA events on its stream (a graph replay has no per-kernel events; this pass is not the reported value)
    ncls = lib.pevit_prof_num_classes()
    lib.pevit_prof_reset()
    lib.pevit_prof_enable(1)
    launches0 = lib.pevit_launch_count()
    ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ep0.record()
    for _ in range(args.steps):
        tuner.step(images, labels)
    ep1.record()
    torch.cuda.synchronize()
    lib.pevit_prof_enable(0)
    launches = lib.pevit_launch_count() - launches0
    ms_eager = reduce_max(ep0.elapsed_time(ep1))
    barrier()
    t_wall1 = time.time()
    ms_arr, cnt_arr = (C.c_double * ncls)(), (C.c_int64 * ncls)()
    _lib.check(lib.pevit_prof_read(ms_arr, cnt_arr, ncls), "pevit_prof_read")

    # ---- end-to-end timing: host-resident inputs, H2D every step (prefetched), loss read back every step
    host_img = [to_pixels(torch.randn(N, 3, R, R), args.pixels).pin_memory() for _ in range(2)]
    host_lab = [torch.randint(0, 10, (N,)).pin_memory() for _ in range(2)]
    if graphed:   # H2D lands directly in the static input buffers of the two captured graphs (no staging copy)
        dev_img, dev_lab = [tuner.input_buffers(b)[0] for b in range(2)], [tuner.input_buffers(b)[1] for b in range(2)]
    else:
        dev_img = [torch.empty_like(images) for _ in range(2)]
        dev_lab = [torch.empty_like(labels) for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            dev_img[b].copy_(host_img[b], non_blocking=True)
            dev_lab[b].copy_(host_lab[b], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_loop(steps):
        for b in range(2):
            consumed[b].record()
        prefetch(0)
        loss_host = 0.0
        for i in range(steps):
            b = i & 1
            if i + 1 < steps:
                prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[b])
            loss = tuner.step_graphed(slot=b) if graphed else tuner.step(dev_img[b], dev_lab[b])
            consumed[b].record()
            loss_host = loss.item()  # D2H of the step's result (kadaptation_clip.py:354)
        return loss_host

    e2e_loop(2)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last_loss = e2e_loop(args.steps)
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = reduce_max(e2.elapsed_time(e3))
    barrier()
    t_wall2 = time.time()
    if clocks is not None:
        clocks.stop()
    if rank != 0:
        tuner.release_graph()
        torch.cuda.synchronize()
        finish(world)
        return

    ms_step = ms_total / args.steps
    value = N * world * args.steps / (ms_total / 1e3)
    e2e_value = N * world * args.steps / (ms_e2e / 1e3)
    pk, pk_src = peaks()
    L_, D, H = shape.tokens, shape.vision_width, shape.heads
    r = {"kadaptation": 32, "lora": 4}.get(args.method, 0)
    M = L_ * N
    kernels = {}
    for i in range(ncls):
        if cnt_arr[i]:
            name = lib.pevit_prof_class_name(i).decode()
            kernels[name] = {"ms_per_step": ms_arr[i] / args.steps, "launches_per_step": cnt_arr[i] / args.steps,
                             "avg_us": 1e3 * ms_arr[i] / cnt_arr[i]}
    own_ms = sum(k["ms_per_step"] for k in kernels.values())
    ms_step_eager = ms_eager / args.steps
    for k in kernels.values():
        # per-launch event brackets add a little to every class (own_ms > ms_step): the share against the TIMED step
        # is an upper bound, the share of the summed kernel time is the distribution
        k["share_of_step"] = k["ms_per_step"] / ms_step
        k["share_of_kernel_time"] = k["ms_per_step"] / own_ms
    traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = {}
    if os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            traffic = json.load(fh).get(f"{args.model}_{args.method}_b{args.batch}", {})

    def roofline_of(name):
        k = kernels[name]
        sec = k["avg_us"] * 1e-6
        if name.startswith("attn"):
            alg = attn_bytes(name, L_, N, D, H, r)
            ach, peak, unit, bound = alg / sec / 1e9, pk["hbm_gbs"], "GB/s", "hbm"
        else:
            alg = gemm_flops(name, M, D, 2 * r)
            ach, peak, unit, bound = alg / sec / 1e12, pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "TFLOP/s", "tensor"
        return {"kernel": name, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": traffic.get(name), "algorithmic_per_launch": alg, "avg_launch_us": k["avg_us"],
                "share_of_step": k["share_of_step"], "peak_source": pk_src}

    rated = [n for n in kernels if n.startswith("attn") or (n.startswith("gemm") and n not in
                                                              ("gemm_other", "gemm_bottleneck"))]
    dominant = max(rated, key=lambda n: kernels[n]["ms_per_step"])
    roof = roofline_of(dominant)
    # whole step against the tensor roofline: SURVEY 8(d) algorithmic FLOPs per image (dense base path, dgrad only)
    p_, E_ = shape.vision_patch_size, shape.embed_dim
    fwd_flops = shape.vision_layers * (24 * L_ * D * D + 4 * L_ * L_ * D) + 2 * (L_ - 1) * 3 * p_ * p_ * D + 2 * D * E_
    bwd_flops = shape.vision_layers * (24 * L_ * D * D + 10 * L_ * L_ * D) + 2 * D * E_
    sustained = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    step_tflops = (fwd_flops + bwd_flops) * N / (ms_step * 1e-3) / 1e12
    whole_step = {"flops_per_image": fwd_flops + bwd_flops, "achieved_tflops": step_tflops, "peak_tflops": sustained,
                  "frac": step_tflops / sustained, "peak_source": pk_src,
                  "ceiling_images_per_s_per_gpu": sustained * 1e12 / (fwd_flops + bwd_flops)}
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload(args, shape),
        "clocks": clocks.summary(t_wall0, t_wall2) if clocks else None,
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": (images.numel() * images.element_size() + labels.numel() * 8) * world,
                "d2h_bytes_per_step": 4 * world, "last_loss": last_loss},
        "gpu_launches": int(launches),  # this library's kernels per K steps (counted in the eager pass; the graph replays the same launches)
        "roofline": roof,
        "roofline_attn": {n: roofline_of(n) for n in ("attn_fwd", "attn_bwd") if n in kernels},
        "whole_step_tensor_frac": whole_step["frac"], "whole_step": whole_step,
        "kernels": kernels, "own_kernel_ms_per_step": own_ms,
        "trainable_params": tuner.trainable_numel(), "host_enqueue_ms_per_step": cpu_ms_per_step,
        "cuda_graph": graphed, "eager_profiled_ms_per_step": ms_step_eager,
        "kernel_timing": "per-launch CUDA events in a separate eager pass of the same K steps right after the timed region",
    }
    if world > 1:
        line["exchange"] = {"kind": "one-shot all-reduce over CUDA-IPC peer memory fused with the SGD update, 1 launch "
                                    "per step (pevit_allreduce_sgd)" if tuner.peer is not None else
                                    "ncclAllReduce over the flat gradient buffer + SGD kernel(s)",
                            "floats": int(tuner.flat_grad.numel()),
                            "timed_out": bool(tuner.peer.timed_out()) if tuner.peer is not None else None}
    tuner.release_graph()
    torch.cuda.synchronize()
    if world == 1:
        line["inference"] = inference_throughput(args, tuner, images)
    if world == 1 and not args.no_text_tower:
        try:
            line["text_tower"] = text_tower_throughput(dev)
        except Exception as exc:  # pragma: no cover
            line["text_tower"] = {"unavailable": repr(exc)[:200]}
    if world == 1 and not args.no_parity_probe:
        probe = parity_probe(args, shape, tuner, dev)
        line["logits_max_abs_err"] = probe["max_abs_err"]
        line["logits_parity"] = probe
    if world == 1 and not args.no_gpu_eager_baseline:
        del tuner
        torch.cuda.empty_cache()
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(args, shape, dev)
            line["gpu_eager_baseline"]["speedup_vs_fp32"] = value / line["gpu_eager_baseline"]["fp32"]
            line["gpu_eager_baseline"]["speedup_vs_autocast_bf16"] = value / line["gpu_eager_baseline"]["autocast_bf16"]
        except Exception as exc:  # pragma: no cover - e.g. out of memory at a large configuration
            line["gpu_eager_baseline"] = {"unavailable": repr(exc)[:200]}
            torch.cuda.empty_cache()
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, shape)
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world: int) -> None:
    """Leave through the NORMAL interpreter exit (atexit handlers and the driver's process hooks must run).  With
    N > 1 the step graph held NCCL work and tearing the communicator down has been seen to block: the graphs are
    already reset, destroy gets a bounded chance, and a daemon watchdog ends the process only if the regular shutdown
    that follows is still stuck a minute later (results are printed by then)."""
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        import torch.distributed as dist
        t = threading.Thread(target=lambda: dist.is_initialized() and dist.destroy_process_group(), daemon=True)
        t.start()
        t.join(20.0)

        def watchdog():
            time.sleep(60.0)
            os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()


def resolve_scaling(args) -> None:
    """--global-batch G: strong scaling (global 2048 on ViT-B/32, N = G / world per GPU)."""
    args.scaling = "weak"
    if args.global_batch > 0:
        if args.global_batch % args.gpus:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by --gpus {args.gpus}")
        args.batch = args.global_batch // args.gpus
        args.scaling = "strong"


def main():
    import signal

    def on_alarm(signum, frame):  # never hang a GPU box: a stuck collective or teardown ends the process instead
        print("[bench] watchdog: run exceeded 25 minutes, aborting", file=sys.stderr, flush=True)
        os._exit(3)
    signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(1500)
    args = parse()
    resolve_scaling(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
