#!/usr/bin/env python
"""bench.py — the headline benchmark: dehazed images/s at 1024x2048 for the adaptive HDEN -> Light/Medium/Complex mix.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--hden densenet121|resnet18]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm (oracle port) on the host cores

A step = one pass of the hot path over one batch of B synthetic hazy images per GPU (BASELINE.json configs[3]: batch 256,
beta in {0.03,0.06,0.09} round-robin, 1024x2048): HDEN classifies the full-resolution batch, the device-side router
buckets it, each branch dehazes its bucket.  A random-init HDEN sends every image to one class, so the mix is injected
through the reference's own `HardRouter.forward(x, intensity=labels)` parameter while HDEN still runs inside the timed
region.  Work is sharded by image: every rank owns B images, no data-path collective (weak scaling).

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: the same step through the public module API from
pinned HOST buffers, H2D of the batch and D2H of the dehazed batch inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1024, 2048
METRIC = "dehazed images/sec @1024x2048 adaptive mix (HDEN -> Light/Medium/Complex)"
UNIT = "images/s"

CFG = {
    "classifier": {"model": "densenet121", "num_classes": 3, "pretrained": False},
    "dehazing": {"low": {"model_type": "lightweight", "channels": 32, "blocks": 3},
                 "medium": {"model_type": "standard", "channels": 64, "blocks": 6},
                 "high": {"model_type": "complex", "channels": 96, "blocks": 9}},
    "routing": {"type": "hard", "temperature": 0.5},
    "device": "cuda",
}

# algorithmic conv/linear FLOPs (2*MAC) per image at 1024x2048 from the reference graph
TFLOP_PER_IMAGE = {"low": 0.278, "medium": 3.591, "high": 8.059, "densenet121": 0.237, "resnet18": 0.152}



def quiet_nccl_stdout():
    """Keep stdout to the ONE JSON line under torchrun: NCCL writes its log (the "NCCL version ..." line at NCCL_DEBUG=VERSION
    or above) to stdout unless NCCL_DEBUG_FILE names a file — and it honours NCCL_DEBUG_FILE only above the VERSION level."""
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--hden", default="densenet121", choices=["densenet121", "resnet18"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the stock-PyTorch-eager comparator leg")
    ap.add_argument("--no-skew", action="store_true", help="skip the skewed-mix / re-balance record (N > 1 only)")
    ap.add_argument("--no-train", action="store_true", help="skip the embedded BASELINE configs[4] training record")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train: BASELINE configs[4], the joint training step (batch 16/GPU at 512x512, gradient all-reduce)")
    return ap.parse_args()


def build_models(cfg, device):
    """Random-init (seed 42) branches + HDEN through the drop-in factories, eval mode."""
    import torch
    from adam_dehaze_b200.models.classifier import create_classifier
    from adam_dehaze_b200.models.dehazing.high_intensity import create_high_intensity_model
    from adam_dehaze_b200.models.dehazing.low_intensity import create_low_intensity_model
    from adam_dehaze_b200.models.dehazing.medium_intensity import create_medium_intensity_model
    torch.manual_seed(42)
    branches = {"low": create_low_intensity_model(cfg), "medium": create_medium_intensity_model(cfg),
                "high": create_high_intensity_model(cfg)}
    clf = create_classifier(cfg)
    for m in list(branches.values()) + [clf]:
        m.eval().to(device)
    return branches, clf


def synth_batch_on_device(n, h, w, dev, seed):
    """The synthetic hazy recipe of SURVEY.md §8d / utils/helpers.py:241-258, generated on the device:
    I = clip(J*t + 0.8(1-t)), t = exp(-beta*d), beta = {0.03,0.06,0.09}[i % 3].  Returns (hazy fp32 NCHW, labels int64)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h, device=dev), torch.linspace(0, 1, w, device=dev), indexing="ij")
    d = (0.3 + 0.7 * torch.sqrt((xx - 0.5) ** 2 + (yy - 0.2) ** 2)) * 100.0
    labels = torch.arange(n, device=dev) % 3
    betas = torch.tensor([0.03, 0.06, 0.09], device=dev)
    hazy = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
    for i in range(n):   # image by image: keeps the generator's scratch small
        t = torch.exp(-betas[labels[i]] * d)
        hazy[i] = torch.clamp(torch.rand((3, h, w), generator=g, device=dev) * t + 0.8 * (1 - t), 0, 1)
    return hazy, labels


# --------------------------------------------------------------------------- CPU leg (oracle port of the reference path)
def cpu_sample(hden, h=512, w=1024, threads=None, repeats=1, warm=0):
    """One image per branch + HDEN on the three, fp32 on the host cores, through the oracle (a port of the reference
    arithmetic).  Returns (images/s scaled to 1024x2048, seconds per sample, description).  FLOPs are linear in pixels
    (SURVEY.md §8), so the rate at h x w is scaled by (h*w)/(1024*2048)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import adam_oracle as oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    branches, clf = build_models(dict(CFG, classifier=dict(CFG["classifier"], model=hden)), device="cpu")
    sds = {k: m.state_dict() for k, m in branches.items()}
    csd = clf.state_dict()
    hazy, _, labels = oracle.synth_hazy(3, h, w, device="cpu")
    times = []
    with torch.no_grad():
        for it in range(warm + repeats):
            t0 = time.perf_counter()
            oracle.classifier_forward(csd, hazy, hden)
            oracle.hard_route(sds, hazy, intensity=labels)
            if it >= warm:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    scale = (h * w) / float(H * W)
    return 3.0 * scale / sec, sec, f"1 image per branch + HDEN({hden}) on the 3, {h}x{w} fp32 (rate scaled by pixel ratio {scale:.4f}), {threads} threads, oracle port"


def gpu_eager_baseline(torch, branches, clf, hden, h, w, dev, nimg=2, reps=3):
    """Stock PyTorch eager (cuDNN / cuBLAS) on the same B200, same synthetic recipe and weights: the on-box comparator
    SURVEY.md 8(d) and BASELINE.md 4 ask for.  Runs the oracle's functional restatement of the reference forwards (plain
    F.conv2d / F.batch_norm calls, i.e. exactly what the reference's nn.Modules dispatch to) in fp32 with TF32 off and in
    bf16 autocast + channels_last.  Baseline leg only: nothing here is on the measured path."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import adam_oracle as oracle
    x, _, _ = oracle.synth_hazy(nimg, h, w, device=dev)
    fwd = {"low": oracle.light_forward, "medium": oracle.medium_forward, "high": oracle.complex_forward}
    out = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        for mode in ("fp32", "bf16_autocast_channels_last"):
            torch.backends.cudnn.allow_tf32 = False
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.benchmark = True
            cl = mode != "fp32"
            xin = x.contiguous(memory_format=torch.channels_last) if cl else x

            def prep(sd):
                return {k: (v.detach().contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v.detach())
                        for k, v in sd.items()}
            sds = {k: prep(m.state_dict()) for k, m in branches.items()}
            csd = prep(clf.state_dict())
            res = {}
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=cl):
                jobs = [(k, (lambda k=k: fwd[k](sds[k], xin))) for k in ("low", "medium", "high")]
                jobs.append((hden, lambda: oracle.classifier_forward(csd, xin, hden)))
                for name, fn in jobs:
                    fn(); fn()                                    # warm-up (cuDNN algorithm search included)
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    for _ in range(reps):
                        fn()
                    b.record()
                    torch.cuda.synchronize()
                    res[name] = a.elapsed_time(b) / reps / nimg
            per_img = (res["low"] + res["medium"] + res["high"]) / 3.0 + res[hden]
            out[mode] = {"ms_per_image": {k: round(v, 3) for k, v in res.items()}, "mix_images_per_s": 1000.0 / per_img}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    out["what"] = (f"torch {torch.__version__} eager, cuDNN {torch.backends.cudnn.version()}, {nimg} images at {h}x{w} per call, "
                   f"mean of {reps} after 2 warm-ups, CUDA events; mix = 1/3 Light + 1/3 Medium + 1/3 Complex + HDEN({hden}) on every image; "
                   "fp32 = TF32 off; bf16 = torch.autocast(bfloat16) + channels_last inputs and weights; cudnn.benchmark on")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    vals = []
    desc = ""
    # warm-up + timed steps, each step one bounded sample
    total = args.warmup + args.steps
    full = None
    for it in range(total):
        if it == 0:      # one un-extrapolated sample at the metric's own resolution, untimed for `value`
            fv, fsec, fdesc = cpu_sample(args.hden, h=args.height, w=args.width, threads=threads)
            full = {"value": fv, "seconds": fsec, "sample": fdesc}
        v, sec, desc = cpu_sample(args.hden, threads=threads)
        if it >= args.warmup:
            vals.append((v, sec))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1000.0 * sum(s for _, s in vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"adaptive mix 1/3 Light, 1/3 Medium, 1/3 Complex + HDEN {args.hden}, 1024x2048 equivalent", "hden": args.hden},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "extrapolated_from": "512x1024", "full_resolution_check": full},
        "extrapolated_from": "512x1024 (rate scaled by the pixel ratio; conv FLOPs are linear in pixels)",
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1]); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "power_w_max": max(power) if power else None, "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "train":
        return run_train(args)

    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "oracle"))   # synthetic-input recipe only (never the measured path)
    from adam_dehaze_b200 import _lib, ops
    from adam_dehaze_b200.models.routing import create_router

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl_stdout()
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().adb_device_check(), "adb_device_check")
    torch.set_grad_enabled(False)

    cfg = dict(CFG, classifier=dict(CFG["classifier"], model=args.hden))
    branches, clf = build_models(cfg, device=dev)
    router = create_router(branches, clf, cfg).eval()
    B, Hh, Ww = args.batch, args.height, args.width
    hazy, labels = synth_batch_on_device(B, Hh, Ww, dev, seed=42 + rank)

    launches = {"n": 0}
    kernels_per_call = {"adb_conv2d": 1, "adb_stem_pack": 1, "adb_attn_pool": 2, "adb_attn_gate_stats": 2, "adb_attn_apply": 2,
                        "adb_maxpool3x3s2": 1, "adb_global_avgpool": 3, "adb_head_mlp": 1, "adb_route": 1, "adb_blend3": 2,
                        "adb_affine_relu": 1, "adb_avgpool2x2": 1, "adb_linear": 1, "adb_nchw_to_nhwc_bf16": 1, "adb_nhwc_bf16_to_nchw": 1}
    _orig_call = _lib.call

    def counting_call(name, *a):
        launches["n"] += kernels_per_call.get(name, 1)
        return _orig_call(name, *a)

    _lib.call = counting_call
    ops._lib.call = counting_call

    def step(x, lab=labels):
        logits, _ = clf(x)                       # HDEN on the full-resolution batch (timed)
        out, info = router(x, intensity=lab)     # device-side bucketing + the three branches
        return out, logits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        step(hazy)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches["n"] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push("adb_timed")      # ncu --nvtx --nvtx-include "adb_timed/" isolates the timed region
    for _ in range(args.steps):
        out, logits = step(hazy)
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches_per_step = launches["n"] // max(1, args.steps)
    _lib.call("adb_kernel_error_flag")
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step / 1000.0)

    # ---- per-branch ms / image and the conv kernel's roofline (instrumented pass, CUDA events around every conv launch)
    per_branch, roof = {}, None
    if rank == 0:
        per_branch, roof = instrumented_pass(torch, ops, _lib, branches, clf, hazy, args.hden)

    # ---- e2e: pinned host buffers -> H2D -> public API -> D2H, all inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dist, world, dev, step, hazy, args, B)

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sec, desc = cpu_sample(args.hden, h=Hh, w=Ww, threads=os.cpu_count() or 1, repeats=1, warm=0)   # ~10-30 s of host work
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port", "sample": desc + ", one sample, no extrapolation",
               "seconds": sec}

    # ---- shard imbalance (N > 1): a 70/20/10 class mix rotated by rank, with and without the class re-balance
    skew = None
    if world > 1 and not args.no_skew:
        skew = run_skew(torch, dist, world, rank, dev, clf, router, hazy, B, args)

    # ---- on-box comparator: stock PyTorch eager (cuDNN) on this GPU (rank 0, N=1 only)
    eager = None
    if rank == 0 and world == 1 and not args.no_eager:
        try:
            eager = gpu_eager_baseline(torch, branches, clf, args.hden, Hh, Ww, dev)
        except Exception as e:  # noqa: BLE001  (a baseline leg must not take the measured line down with it)
            eager = {"error": f"{type(e).__name__}: {e}"[:300]}

    # ---- BASELINE configs[4] in the same run: the joint training step at this N (every rank takes part)
    train = None
    if not args.no_train:
        del out, logits
        for m in branches.values():
            eng = m.__dict__.get("_adb_engine") or getattr(m, "engine", None)
            if eng is not None and hasattr(eng, "release_buffers"):
                eng.release_buffers()
        del hazy, router, branches, clf
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        _lib.call = _orig_call
        ops._lib.call = _orig_call
        targs = argparse.Namespace(**vars(args))
        targs.steps, targs.warmup, targs.batch, targs.height, targs.width = max(10, args.steps), max(3, args.warmup), 16, 512, 512
        train = run_train(targs, embedded=True)
        torch.set_grad_enabled(False)

    if rank == 0:
        mix = (TFLOP_PER_IMAGE["low"] + TFLOP_PER_IMAGE["medium"] + TFLOP_PER_IMAGE["high"]) / 3 + TFLOP_PER_IMAGE[args.hden]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[3]: adaptive mix HDEN({args.hden}) -> Light/Medium/Complex, {B} images/GPU/step at {Hh}x{Ww}, "
                                   "beta round-robin {0.03,0.06,0.09}, image-sharded, random-init weights seed 42",
                       "images_per_gpu_per_step": B, "height": Hh, "width": Ww, "hden": args.hden,
                       "mix": "labels i%3 injected via HardRouter.forward(x, intensity=labels); HDEN timed",
                       "l2": f"inputs {B * 3 * Hh * Ww * 4 / 2**30:.1f} GiB per step (> 126 MB L2)",
                       "algorithmic_tflop_per_image": mix},
            "clocks": clocks, "gpu_launches": launches_per_step, "per_branch_ms_per_image": per_branch,
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "model_tflops_per_gpu": value / world * mix,
            "gpu_eager_baseline": eager,
            "shard_imbalance": skew,
            "train": train,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_skew(torch, dist, world, rank, dev, clf, router, hazy, B, args):
    """"load imbalance (a shard heavy in Complex images costs far more than a Light shard) is the only scaling
    risk".  Every rank gets a 70 / 20 / 10 class mix whose heavy class is (rank % 3); measured (a) as sharded by image count
    alone, (b) with adam_dehaze_b200.sharding.Exchange: all-gather of the class ids, one NCCL all-to-all of whole images so
    that every rank holds an equal share of every class, dehaze, all-to-all back.  HDEN runs on the local shard first in
    both (the class ids come from it in production; here they are the injected labels).  Max over ranks, whole job."""
    from adam_dehaze_b200.sharding import Exchange
    heavy = rank % 3
    n70, n20 = int(round(0.7 * B)), int(round(0.2 * B))
    lab = torch.empty(B, dtype=torch.int64)
    lab[:n70], lab[n70:n70 + n20], lab[n70 + n20:] = heavy, (heavy + 1) % 3, (heavy + 2) % 3
    lab = lab[torch.randperm(B, generator=torch.Generator().manual_seed(100 + rank))].to(dev)

    def plain():
        clf(hazy)
        return router(hazy, intensity=lab)[0]

    state = {}

    def rebalanced():
        clf(hazy)
        gathered = torch.empty(world * B, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, lab)                       # the only extra collective besides the image trade
        by_rank = gathered.view(world, B).cpu().tolist()                 # (host read of world*B class ids)
        ex = Exchange(by_rank, rank)
        x, l2 = ex.forward(hazy)
        y = router(x, intensity=l2)[0]
        state["moved"], state["held"] = ex.moved, x.shape[0]
        return ex.backward(y)

    def timed(fn, steps):
        fn()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps
    steps = max(2, min(3, args.steps))
    ms_plain = timed(plain, steps)
    ms_reb = timed(rebalanced, steps)
    moved = torch.tensor([float(state["moved"])], device=dev)
    dist.all_reduce(moved)
    return {"mix": "each rank: 70 % of class (rank % 3), 20 % of the next, 10 % of the third, shuffled",
            "images_per_s_by_count_only": world * B / (ms_plain / 1000.0), "ms_per_step_by_count_only": ms_plain,
            "images_per_s_rebalanced": world * B / (ms_reb / 1000.0), "ms_per_step_rebalanced": ms_reb,
            "images_moved_per_step": int(moved.item()), "bytes_moved_per_step": int(moved.item()) * 2 * hazy[0].numel() * 4,
            "exchange": "all_gather of class ids + NCCL all_to_all_single of whole fp32 images (there and back), inside the timed region",
            "steps": steps}


# --------------------------------------------------------------------------- training step (BASELINE configs[4])
TRAIN_METRIC = "training samples/sec, joint step (HDEN + SoftRouter over Light/Medium/Complex, JointLoss, Adam) @512x512"


def run_train(args, embedded=False):
    """One step = train_joint.py:129-150 on 16 synthetic samples per GPU at 512x512: HDEN logits, SoftRouter forward of the
    three branches in train() mode (batch-statistics BatchNorm), DehazingLoss, backward (dgrad + wgrad kernels), one NCCL
    all-reduce of the flat gradient bucket, one fused Adam launch.  Prints ONE JSON line (rank 0)."""
    import torch
    import torch.distributed as dist
    from adam_dehaze_b200 import _lib, ops
    from adam_dehaze_b200.models.routing import SoftRouter
    from adam_dehaze_b200.training.loss import DehazingLoss
    from adam_dehaze_b200.training.optim import FlatAdam
    import adam_dehaze_b200.training.autograd as ag

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not embedded:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_nccl_stdout()
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().adb_device_check(), "adb_device_check")
    torch.set_grad_enabled(True)
    B = 16 if args.batch == 256 else args.batch
    Hh, Ww = (512, 512) if (args.height, args.width) == (H, W) else (args.height, args.width)
    # the joint step trains HDEN through the router (train_joint.py:80-88,117-121)
    hden = args.hden
    cfg = dict(CFG, classifier=dict(CFG["classifier"], model=hden), routing={"type": "soft", "temperature": 0.5})
    branches, clf = build_models(cfg, device=dev)
    router = SoftRouter(branches, classifier=clf, temperature=0.5).to(dev).train()
    from adam_dehaze_b200.training.loss import JointLoss
    crit = JointLoss(1.0, 0.2, 0.5, dehazing_loss=DehazingLoss(lambda_l1=1.0, lambda_content=TRAIN_LAMBDAS[0],
                                                                 lambda_perceptual=TRAIN_LAMBDAS[1])).to(dev)
    opt = FlatAdam(router.parameters(), lr=5e-5, weight_decay=1e-4)
    nparams = sum(p.numel() for p in router.parameters())
    hazy, labels = synth_batch_on_device(B, Hh, Ww, dev, seed=42 + rank)
    clear = torch.rand((B, 3, Hh, Ww), generator=torch.Generator(device=dev).manual_seed(7 + rank), device=dev)

    counts = {"n": 0}
    inner = _lib.call

    def counting(name, *a):
        counts["n"] += 1
        return inner(name, *a)

    def set_call(fn):
        _lib.call = fn
        ops._lib.call = fn
        ag._lib.call = fn

    def step():
        opt.zero_grad()
        logits, _ = clf(hazy)                          # HDEN in train() mode: its logits weight the blend and feed the CE term
        out, info = router(hazy, logits)
        loss, parts = crit(out, clear, logits, labels)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    import gc
    torch.cuda.synchronize()
    m0 = torch.cuda.memory_allocated()
    found = gc.collect()           # untimed: anything only the cycle collector can free would otherwise be released at a random later step
    if rank == 0:
        sys.stderr.write(f"gc.collect() after warm-up: {found} unreachable objects, {(m0 - torch.cuda.memory_allocated()) / 2**20:.0f} MiB of device memory released\n")
    if os.environ.get("ADB_PROFILE_HOST") and world == 1:    # developer aid: where the host time of a step goes
        import cProfile
        import pstats
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        for _ in range(3):
            step()
        pr.disable()
        host_s = (time.perf_counter() - t0) / 3
        torch.cuda.synchronize()
        sys.stderr.write(f"host time per step (launch only, no sync): {host_s * 1e3:.1f} ms\n")
        st_ = pstats.Stats(pr, stream=sys.stderr).sort_stats("tottime")
        st_.print_stats(28)
        st_.print_callers("masked_select|built-in method torch.tensor|run_backward")
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    set_call(counting)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push("adb_timed")
    torch.cuda.profiler.start()       # ncu --profile-from-start off: loss.backward() runs in autograd's worker thread,
    host_t = [time.perf_counter()]
    for _ in range(args.steps):       # which a per-thread NVTX range would miss
        loss = step()
        host_t.append(time.perf_counter())
    torch.cuda.profiler.stop()
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    set_call(inner)
    clocks = sampler.stop()
    _lib.call("adb_kernel_error_flag")
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = world * B / (ms_step / 1000.0)

    # ---- e2e: the same step with the batch (hazy + clear) copied from pinned host memory and the loss read back
    host_h = torch.empty(hazy.shape, dtype=torch.float32, pin_memory=True).copy_(hazy)
    host_c = torch.empty(clear.shape, dtype=torch.float32, pin_memory=True).copy_(clear)

    from adam_dehaze_b200.training.train_dehazing import LossMeter
    meter = LossMeter()

    def e2e_step():
        hazy.copy_(host_h, non_blocking=True)
        clear.copy_(host_c, non_blocking=True)
        meter.push(step())         # train_dehazing.py:95 reads the loss every step; LossMeter reads step i's value from
                                   # pinned memory once step i+1 is queued, so the host never drains the launch queue

    e2e_step()
    meter.flush()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    host_e = [time.perf_counter()]
    for _ in range(args.steps):
        e2e_step()
        host_e.append(time.perf_counter())
    e2e_loss = meter.flush()       # the last step's loss is on the host before the clock stops
    b.record()
    if rank == 0:                  # host queueing time per step (stderr): a host-bound phase shows up here, not in the kernels
        sys.stderr.write("host ms per step, resident loop: " + " ".join(f"{(y - x) * 1e3:.0f}" for x, y in zip(host_t, host_t[1:])) +
                         " | e2e loop: " + " ".join(f"{(y - x) * 1e3:.0f}" for x, y in zip(host_e, host_e[1:])) + "\n")
    barrier()
    t2 = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = t2.item() / args.steps

    # ---- instrumented step: CUDA events around every C-ABI call -> tensor-pipe kernels' achieved TFLOP/s.
    #      Every rank runs it (the step holds the gradient all-reduce, so the ranks must stay in lock-step); rank 0 reports.
    detail = None
    if True:
        calls = []

        def timed(name, *a):
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            r = inner(name, *a)
            eb.record()
            fl = 0.0
            if name == "adb_conv2d":
                fl = float(_lib.load().adb_conv2d_flops(a[0]))
            elif name == "adb_wgrad":
                fl = float(_lib.load().adb_wgrad_flops(a[0]))
            calls.append((name, ea, eb, fl))
            return r
        set_call(timed)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        opt.allreduce_events = []
        s0.record(); step(); s1.record()
        torch.cuda.synchronize()
        set_call(inner)
        ar_ms = [a_.elapsed_time(b_) for a_, b_ in opt.allreduce_events]
        opt.allreduce_events = None
        by, fl = {}, {}
        slow = []
        for nm, ea, eb, f in calls:
            ms_ = ea.elapsed_time(eb)
            by[nm] = by.get(nm, 0.0) + ms_
            fl[nm] = fl.get(nm, 0.0) + f
            slow.append((ms_, nm))
        if os.environ.get("ADB_PROFILE_HOST") and rank == 0:
            slow.sort(reverse=True)
            sys.stderr.write("slowest calls: " + ", ".join(f"{n}:{m:.2f}" for m, n in slow[:25]) + "\n")
            ms_ = torch.cuda.memory_stats()
            sys.stderr.write(f"allocator: reserved {ms_['reserved_bytes.all.peak'] / 2**30:.1f} GiB, alloc retries {ms_['num_alloc_retries']}, "
                             f"cudaMalloc segments {ms_['segment.all.allocated']}, freed {ms_['segment.all.freed']}\n")
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        mma_ms = by.get("adb_conv2d", 0.0) + by.get("adb_wgrad", 0.0)
        mma_fl = fl.get("adb_conv2d", 0.0) + fl.get("adb_wgrad", 0.0)
        detail = {
            "step_ms_instrumented": s0.elapsed_time(s1),
            "ms_by_entry_point": {k.replace("adb_", ""): round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])},
            "conv2d_fwd_dgrad_tflops": fl.get("adb_conv2d", 0.0) / max(1e-9, by.get("adb_conv2d", 0.0) * 1e-3) / 1e12,
            "wgrad_tflops": fl.get("adb_wgrad", 0.0) / max(1e-9, by.get("adb_wgrad", 0.0) * 1e-3) / 1e12,
            "launch_flops_tflop_per_step": mma_fl / 1e12,
        }
        n_mma = sum(1 for nm, *_ in calls if nm in ("adb_conv2d", "adb_wgrad"))
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel + conv_wgrad_kernel", "achieved": mma_fl / max(1e-9, mma_ms * 1e-3) / 1e12,
                "peak": peak, "unit": "TFLOP/s", "frac": mma_fl / max(1e-9, mma_ms * 1e-3) / 1e12 / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                "how": f"launch FLOPs (padded-channel 2*MAC of every conv fwd/dgrad/wgrad launch) / CUDA-event time over the {n_mma} tensor-pipe launches of one step",
                "flops_per_launch_avg": mma_fl / max(1, n_mma), "ms_per_launch_avg": mma_ms / max(1, n_mma)}
    # ---- the same step as ONE CUDA graph replay (training/graphed.py): removes the host's launch queueing from the step
    graphed = None
    loss_value = float(loss.item())
    # (autograd keeps the AccumulateGrad nodes of the eager steps alive through any surviving loss tensor; they belong to the
    # default stream and would invalidate a capture on another stream)
    del loss
    meter = None
    gc.collect()
    if (world == 1 or os.environ.get("ADB_TRAIN_GRAPH") == "1") and not os.environ.get("ADB_NO_TRAIN_GRAPH"):
        try:
            from adam_dehaze_b200.training.graphed import GraphedStep
            gs = GraphedStep(step, warmup=2)
            for _ in range(2):
                gs()
            barrier()
            ga, gb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ga.record()
            for _ in range(args.steps):
                gloss = gs()
            gb.record()
            barrier()
            tg = torch.tensor([ga.elapsed_time(gb)], device=dev)
            if world > 1:
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            g_ms = tg.item() / args.steps
            graphed = {"ms_per_step": g_ms, "samples_per_s": world * B / (g_ms / 1000.0), "loss": float(gloss.item()),
                       "what": "the identical step (forward, JointLoss, backward, gradient all-reduce, Adam) captured once in a torch.cuda.CUDAGraph "
                               "and replayed: one host call per step instead of ~1.7 k launches from Python"}
            del gs
        except Exception as e:  # noqa: BLE001  (an optional leg must not take the measured line down)
            graphed = {"error": f"{type(e).__name__}: {e}"[:300]}
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
    if rank == 0:
        line = {
            "metric": TRAIN_METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: joint training step, {B} samples/GPU at {Hh}x{Ww}, SoftRouter(T=0.5) over "
                                   f"Light+Medium+Complex + HDEN({hden}), all in train() mode, JointLoss(1.0, 0.2, 0.5) over DehazingLoss "
                                   f"lambdas (1.0, {TRAIN_LAMBDAS[0]}, {TRAIN_LAMBDAS[1]}) + CE, FlatAdam(lr 5e-5, wd 1e-4), one flat "
                                   f"{nparams * 4 / 2**20:.0f} MiB gradient all-reduce per step",
                       "samples_per_gpu_per_step": B, "height": Hh, "width": Ww, "trainable_params": nparams,
                       "l2": f"activations {B}x{Hh}x{Ww} per layer (> 126 MB L2 for every full-resolution map)"},
            "clocks": clocks, "gpu_launches": counts["n"] // max(1, args.steps), "loss": loss_value,
            "roofline": roof, "train_detail": detail, "cpu_baseline": None, "graphed_step": graphed,
            "allreduce": {"collective": "NCCL all_reduce(sum) of the flat fp32 gradient bucket, one per step" if world > 1 else "none (1 GPU)",
                          "bytes": int(opt.grad.numel()) * 4, "ms": (sum(ar_ms) / len(ar_ms)) if ar_ms else None,
                          "bus_GBs": (2.0 * (world - 1) / world * opt.grad.numel() * 4 / (sum(ar_ms) / len(ar_ms) * 1e-3) / 1e9) if ar_ms else None,
                          "how": "CUDA events on the step's stream around dist.all_reduce in the instrumented step"},
            "e2e": {"value": world * B / (e2e_ms / 1000.0), "unit": "samples/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": 2 * B * 3 * Hh * Ww * 4, "d2h_bytes_per_step": 4,
                    "api": "SoftRouter.forward + DehazingLoss + loss.backward() + FlatAdam.step() from pinned host batches; every step's loss read on the host through training.train_dehazing.LossMeter (pinned 4-byte D2H per step, read one step later)"},
        }
        if not embedded:
            print(json.dumps(line), flush=True)
    else:
        line = None
    del router, opt, crit, branches, clf, hazy, clear
    if world > 1 and not embedded:
        dist.barrier()
        dist.destroy_process_group()
    return line


TRAIN_LAMBDAS = (0.1, 0.1)   # (content, perceptual) terms of DehazingLoss in the training bench


def instrumented_pass(torch, ops, _lib, branches, clf, hazy, hden):
    """Per-branch device ms per image, and for the dominant kernel (conv_igemm_kernel) the achieved TFLOP/s: true
    conv FLOPs of every launch / CUDA-event time of that launch, summed over one pass of each branch + HDEN."""
    from adam_dehaze_b200 import ops as ops_mod
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        peaks = json.load(fh) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    rec = []
    orig = ops_mod.conv2d

    def timed_conv(spec, src0, src1=None, **kw):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = orig(spec, src0, src1, **kw)
        b.record()
        n = kw.get("n") or src0.shape[0]
        rec.append((a, b, conv_flops(spec, src0, src1, n, kw)))
        return r

    # every model is measured at the launch shape the timed step gives it: a branch walks its bucket 16 images at a time
    # (engine micro-batch), HDEN runs the chunk its engine cuts a 256-image batch into (capped at 32 here to bound the pass)
    B_all = hazy.shape[0]
    n_of = {}
    for name, m in branches.items():
        n_of[name] = min(B_all, m._branch_engine().micro_batch(hazy.shape[2], hazy.shape[3], max(1, B_all // 3)))
    n_of[hden] = min(B_all, 32, clf._hden_engine().chunk_size(B_all, hazy.shape[2], hazy.shape[3]))
    per_branch = {}
    import adam_dehaze_b200.engine as eng
    eng.ops.conv2d = timed_conv
    ops_mod.conv2d = timed_conv
    # every C-ABI call by entry point (CUDA events around each call) -> where a branch's non-conv time goes
    calls = []
    inner_call = _lib.call

    def timed_call(name, *a):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        r = inner_call(name, *a)
        eb.record()
        calls.append((name, ea, eb))
        return r

    try:
        for name, m in list(branches.items()) + [(hden, clf)]:
            nimg = n_of[name]
            x = hazy[:nimg].contiguous()
            m(x)  # warm
            torch.cuda.synchronize()
            rec.clear()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.nvtx.range_push(f"adb_roofline_{name}")     # the launches tools/conv_traffic.py reads dram bytes for
            s.record(); m(x); e.record()
            torch.cuda.synchronize()
            torch.cuda.nvtx.range_pop()
            conv_ms = sum(a.elapsed_time(b) for a, b, _ in rec)
            flops = sum(f for _, _, f in rec)
            per_branch[name] = {"ms": s.elapsed_time(e) / nimg, "conv_ms": conv_ms / nimg, "conv_launches": len(rec),
                                "conv_tflops": flops / (conv_ms * 1e-3) / 1e12 if conv_ms else None,
                                "tflop_per_image": flops / nimg / 1e12, "images": nimg}
            eng.ops.conv2d = orig
            ops_mod.conv2d = orig
            _lib.call = timed_call
            ops_mod._lib.call = timed_call
            calls.clear()
            m(x)
            torch.cuda.synchronize()
            _lib.call = inner_call
            ops_mod._lib.call = inner_call
            eng.ops.conv2d = timed_conv
            ops_mod.conv2d = timed_conv
            by = {}
            for nm, ea, eb in calls:
                by[nm] = by.get(nm, 0.0) + ea.elapsed_time(eb)
            per_branch[name]["ms_by_entry_point"] = {k.replace("adb_", ""): round(v / nimg, 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1])}
    finally:
        eng.ops.conv2d = orig
        ops_mod.conv2d = orig
        _lib.call = inner_call
        ops_mod._lib.call = inner_call
    tot_ms = sum(v["conv_ms"] for k, v in per_branch.items() if k in ("low", "medium", "high")) + per_branch[hden]["conv_ms"] * 3
    tot_fl = sum(v["tflop_per_image"] for k, v in per_branch.items() if k in ("low", "medium", "high")) + per_branch[hden]["tflop_per_image"] * 3
    achieved = tot_fl / (tot_ms * 1e-3)
    peak = peaks.get("bf16_tflops_sustained")
    src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)"
    if not peak:
        peak, src = 1400.0, "fallback sustained figure, B200_PROFILING.md (of fallback)"
    launches = sum(v["conv_launches"] for v in per_branch.values())
    # one average launch of the equal-thirds mix: every image runs HDEN, a third of the images run each branch, so a model's
    # launches, FLOPs and time all enter with weight 1 (branches) or 3 (HDEN)
    wts = {k: (3.0 if k == hden else 1.0) for k in per_branch}
    w_launch = sum(wts[k] * v["conv_launches"] for k, v in per_branch.items())
    w_flops = sum(wts[k] * v["tflop_per_image"] * v["images"] for k, v in per_branch.items()) * 1e12
    w_ms = sum(wts[k] * v["conv_ms"] * v["images"] for k, v in per_branch.items())
    per_model = {k: {"conv_launches": v["conv_launches"], "images_per_launch": v["images"], "achieved_tflops": v["conv_tflops"],
                     "frac": (v["conv_tflops"] / peak) if v["conv_tflops"] else None,
                     "flops_per_launch": v["tflop_per_image"] * v["images"] * 1e12 / max(1, v["conv_launches"]),
                     "ms_per_launch": v["conv_ms"] * v["images"] / max(1, v["conv_launches"]), "mix_weight": wts[k]}
                 for k, v in per_branch.items()}
    # dram bytes per conv launch of this same pass, from the committed ncu capture (tools/conv_traffic.py); same weighting
    # and same 8-image pass as flops_per_launch_avg, so bytes/launch and FLOPs/launch describe the same average launch
    traffic, traffic_src = None, "profiles/r*_conv_traffic.json missing"
    tpath = next((q for q in (os.path.join(ROOT, "profiles", f) for f in ("r2_conv_traffic.json", "r1_conv_traffic.json")) if os.path.exists(q)), "")
    if tpath:
        with open(tpath) as fh:
            tj = json.load(fh)
        pm = tj.get("per_model", {})
        if tj.get("hden") == hden and (tj.get("height"), tj.get("width")) == (hazy.shape[2], hazy.shape[3]) and all(
                k in pm and pm[k]["launches"] == per_branch[k]["conv_launches"] for k in per_branch):
            # DRAM bytes of a conv launch scale with its image count (weights are < 1 % of the traffic): the capture's
            # bytes per image times this pass's images per launch, mix-weighted like flops_per_launch_avg
            tot = sum(wts[k] * (pm[k]["dram_read_bytes"] + pm[k]["dram_write_bytes"]) / pm[k].get("images", tj.get("images", 8)) * per_branch[k]["images"]
                      for k in per_branch)
            traffic, traffic_src = tot / max(1.0, w_launch), tj["source"]
            # the same launches against the HBM roofline: Light and the DenseNet trunk sit below the ridge (arithmetic intensity
            # 104 and 122 FLOP/B against 219), so for them this — not the tensor fraction — is the bound that matters
            hbm = peaks.get("hbm_gbs") or 6450.9
            for k in per_branch:
                gb_img = (pm[k]["dram_read_bytes"] + pm[k]["dram_write_bytes"]) / pm[k].get("images", tj.get("images", 8)) / 1e9
                per_model[k]["dram_gb_per_image"] = gb_img
                per_model[k]["achieved_gbs"] = gb_img / (per_branch[k]["conv_ms"] * 1e-3) if per_branch[k]["conv_ms"] else None
                per_model[k]["hbm_frac"] = per_model[k]["achieved_gbs"] / hbm if per_model[k]["achieved_gbs"] else None
                per_model[k]["bound"] = "hbm" if (per_branch[k]["tflop_per_image"] * 1e12 / (gb_img * 1e9)) < (peak * 1e12) / (hbm * 1e9) else "tensor"
        else:
            traffic_src = os.path.basename(tpath) + " was captured on a different pass (model, resolution or launch count)"
    roof = {"bound": "tensor", "kernel": "conv_igemm_kernel + conv_roll_kernel (adb_conv2d)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
            "peak_source": src,
            "traffic_ref": "ncu --set full captures of representative conv shapes: profiles/r2/README.md (DRAM traffic ~= algorithmic bytes)",
            "how": f"sum of true conv FLOPs / sum of CUDA-event durations over the {launches} conv launches of one pass of "
                   f"Light+Medium+Complex+HDEN at {hazy.shape[2]}x{hazy.shape[3]}, each at its own launch size ({n_of}) (equal-thirds mix weighting)",
            "flops_per_launch_avg": w_flops / max(1.0, w_launch),
            "ms_per_launch_avg": w_ms / max(1.0, w_launch),
            "launch_weighting": "mix-weighted over the same pass: each model's launches, FLOPs and milliseconds enter with weight 1 "
                                "(Light/Medium/Complex) or 3 (HDEN); achieved = flops_per_launch_avg / ms_per_launch_avg",
            "per_model": per_model}
    return per_branch, roof


def conv_flops(spec, src0, src1, n, kw):
    """True (unpadded) 2*MAC of one launch: stems count their real 3-channel taps, padded channels count nothing."""
    from adam_dehaze_b200 import _lib
    _, h, w, p0 = src0.shape
    cin = (kw.get("c0") or p0) + ((kw.get("c1") or src1.shape[3]) if src1 is not None else 0)
    if spec.kind == _lib.CONVT_4X4S2:
        return 2.0 * n * h * w * 16 * cin * spec.cout
    if spec.kind == _lib.CONV_S2:
        return 2.0 * n * (h // 2) * (w // 2) * spec.kh * spec.kw * cin * spec.cout
    k = spec.kh * spec.kw * cin
    if spec.kw == 1 and spec.kh in (3, 7) and cin in (16, 32):   # horizontally packed stem: real K = kh*kh*3
        k = spec.kh * spec.kh * 3
    if spec.kh == 1 and cin == 160:                               # full-im2col stem: real K = 147
        k = 147
    if spec.kind == _lib.CONV_K4_S2D:                             # 7x7 stride-2 stem over the space-to-depth image: real K = 147
        k = 147
    return 2.0 * n * h * w * k * spec.cout


def run_e2e(torch, dist, world, dev, step, hazy, args, B):
    """Public-API step from pinned host memory.  The batch streams through in chunks on three CUDA streams (H2D, compute,
    D2H) with two device buffers each way, so copies overlap the kernels; every byte of the batch crosses PCIe in both
    directions inside the timed region."""
    chunk = min(B, 48)      # 16 images per branch per chunk = one full micro-batch each

    from adam_dehaze_b200.data.pipeline import u8_to_tensor
    Hh, Ww = hazy.shape[2], hazy.shape[3]

    def pinned(nimg, dtype, shape):
        return torch.empty((nimg,) + shape, dtype=dtype, pin_memory=True)
    # inputs live on the host as DECODED IMAGES (uint8 HWC, BGR like cv2.imread returns them, dataset.py:76): 3 bytes per pixel
    # cross PCIe instead of the 12 of the reference loader's float tensor; BGR->RGB + ToTensor run on the device
    # (adb_image_u8_to_f32) inside the timed region.  Outputs go back as the fp32 tensors the reference returns.
    ring = B                # images the pinned staging buffers hold: the whole batch, or (if the host refuses that much
    try:                    # pinned memory per rank, e.g. 8 ranks on one box) a ring of four chunks walked modulo its size
        host_in, host_out = pinned(ring, torch.uint8, (Hh, Ww, 3)), pinned(ring, torch.float32, (3, Hh, Ww))
    except RuntimeError:
        host_in = host_out = None
        ring = min(B, 4 * chunk)
        try:
            host_in, host_out = pinned(ring, torch.uint8, (Hh, Ww, 3)), pinned(ring, torch.float32, (3, Hh, Ww))
        except RuntimeError:
            return {"value": None, "unit": UNIT, "error": "pinned host allocation failed"}
    for s0 in range(0, ring, 16):        # quantise the synthetic batch to 8 bits, channel-swapped to BGR
        blk = hazy[s0:min(ring, s0 + 16)]
        host_in[s0:s0 + blk.shape[0]].copy_((blk.flip(1).permute(0, 2, 3, 1) * 255.0 + 0.5).clamp(0, 255).to(torch.uint8))
    labels_full = (torch.arange(B, device=dev) % 3)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    s_cmp = torch.cuda.current_stream()
    ubuf = [torch.empty((chunk, Hh, Ww, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
    xbuf = [torch.empty((chunk,) + tuple(hazy.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
    obuf = [torch.empty((chunk,) + tuple(hazy.shape[1:]), dtype=torch.float32, device=dev) for _ in range(2)]
    x_ready = [torch.cuda.Event() for _ in range(2)]
    x_free = [torch.cuda.Event() for _ in range(2)]
    o_ready = [torch.cuda.Event() for _ in range(2)]
    o_free = [torch.cuda.Event() for _ in range(2)]
    from adam_dehaze_b200.models.routing import HardRouter  # noqa: F401  (the public API `step` drives)

    for e in x_free + o_free:
        e.record(s_cmp)

    def e2e_step():
        # buffer hand-offs carry over from one step to the next (x_free / o_free of the previous step's last two chunks),
        # so the upload of step k+1's first chunks overlaps the kernels of step k's last ones; every step still moves its
        # whole batch host -> device and its whole result device -> host
        for i, s in enumerate(range(0, B, chunk)):
            n, b = min(chunk, B - s), i & 1
            hs = s % ring                                      # == s when the staging buffers hold the whole batch
            with torch.cuda.stream(s_in):
                s_in.wait_event(x_free[b])                     # compute finished reading this input buffer
                ubuf[b][:n].copy_(host_in[hs:hs + n], non_blocking=True)
                x_ready[b].record(s_in)
            s_cmp.wait_event(x_ready[b])
            s_cmp.wait_event(o_free[b])                        # D2H finished reading this output buffer
            u8_to_tensor(ubuf[b][:n], out=xbuf[b][:n])         # BGR uint8 HWC -> RGB fp32 NCHW / 255 (the loader's ToTensor)
            out, _ = step(xbuf[b][:n], labels_full[s:s + n])
            obuf[b][:n].copy_(out, non_blocking=True)
            x_free[b].record(s_cmp)
            o_ready[b].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(o_ready[b])
                host_out[hs:hs + n].copy_(obuf[b][:n], non_blocking=True)
                o_free[b].record(s_out)

    for _ in range(max(1, min(2, args.warmup))):
        e2e_step()
    s_cmp.wait_stream(s_out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        e2e_step()
    s_cmp.wait_stream(s_out)       # the last step's results are in host memory before the clock stops
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / args.steps
    nbytes = B * hazy[0].numel() * 4
    return {"value": world * B / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": B * hazy[0].numel(),
            "d2h_bytes_per_step": nbytes, "input_format": "uint8 HWC BGR decoded images (3 B/pixel), transformed on the device by adb_image_u8_to_f32", "chunk_images": chunk, "host_staging_images": ring, "streams": "H2D / compute / D2H, double-buffered, uploads of the next step overlap the tail of the current one",
            "api": "FogIntensityClassifier.forward + HardRouter.forward(x, intensity=labels) on host-resident batches"}


if __name__ == "__main__":
    main()
