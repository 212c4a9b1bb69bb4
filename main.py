#!/usr/bin/env python
"""Entry points of the drop-in: python main.py --mode {evaluate, demo, train_dehazing, train_joint} [--config ...]
[--exp_name ...] [--data_dir ...] [--device ...] [--seed ...]   (same flags as the reference's main.py:29-56).

The hot path (HDEN -> router -> branches) runs through adam_dehaze_b200.  When --data_dir (or config['dataset'][...]) names a
directory laid out like the reference's dataset (<root>/<split>/<low|medium|high>/{hazy,clear,dehazed}/*.png|jpg) the modes
read it through the device input pipeline (adam_dehaze_b200/data/pipeline.py); when it does not exist (the corpus is private)
they run on the synthetic hazy recipe of SURVEY.md §8d so that they are exercisable offline.  The detection sweep and the
matplotlib figures are out of scope.
"""
import argparse
import json
import os
import random
import time

import numpy as np
import torch
import yaml

from adam_dehaze_b200.models.classifier import create_classifier
from adam_dehaze_b200.models.dehazing.high_intensity import create_high_intensity_model
from adam_dehaze_b200.models.dehazing.low_intensity import create_low_intensity_model
from adam_dehaze_b200.models.dehazing.medium_intensity import create_medium_intensity_model
from adam_dehaze_b200.models.routing import create_router
from adam_dehaze_b200.training.loss import get_dehazing_loss

MODES = ["preprocess", "train_classifier", "train_dehazing", "train_joint", "train_all", "evaluate", "demo"]


def parse_args():
    p = argparse.ArgumentParser(description="Adaptive fog-intensity dehazing — B200 path")
    p.add_argument("--config", type=str, default="config/config.yaml")
    p.add_argument("--mode", type=str, default="evaluate", choices=MODES)
    p.add_argument("--exp_name", type=str, default=None)
    p.add_argument("--data_dir", type=str, default=None)
    p.add_argument("--device", type=str, default=None)
    p.add_argument("--resume", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--synthetic", type=int, default=12, help="images to synthesise when no dataset directory exists")
    p.add_argument("--size", type=int, nargs=2, default=[256, 256], metavar=("H", "W"))
    p.add_argument("--epochs", type=int, default=2, help="epochs per branch for --mode train_dehazing on synthetic data")
    return p.parse_args()


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def experiment_dir(config, name):
    name = name or f"experiment_{time.strftime('%Y%m%d_%H%M%S')}"
    root = os.path.join("experiments", name)
    ck = os.path.join(root, "checkpoints")
    for sub in ("checkpoints", "logs", os.path.join("results", "metrics")):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    for section, leaf in (("classifier", "classifier"), ("dehazing", "dehazing"), ("routing", "routing"), ("joint_training", "joint")):
        config[section]["checkpoint_dir"] = os.path.join(ck, leaf)
    config["evaluation"]["results_dir"] = os.path.join(root, "results", "metrics")
    with open(os.path.join(root, "config.yaml"), "w") as fh:
        yaml.safe_dump(config, fh)
    return root


def synth_hazy(n, h, w, device, seed):
    """I = clip(J t + 0.8 (1 - t)), t = exp(-beta d); beta round-robin over {0.03, 0.06, 0.09} (labels 0/1/2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    clear = torch.rand(n, 3, h, w, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    d = (0.3 + 0.7 * torch.sqrt((xx - 0.5) ** 2 + (yy - 0.2) ** 2)) * 100.0
    labels = torch.arange(n) % 3
    t = torch.exp(-torch.tensor([0.03, 0.06, 0.09])[labels].view(n, 1, 1, 1) * d)
    return torch.clamp(clear * t + 0.8 * (1 - t), 0, 1).to(device), clear.to(device), labels.to(device)


def load_if_present(module, path, key):
    """Missing checkpoints are not errors: print and continue with the current weights (evaluate.py:21-30)."""
    if os.path.exists(path):
        module.load_state_dict(torch.load(path, map_location="cpu")[key])
        print(f"loaded {path}")
    else:
        print(f"checkpoint {path} not found — continuing with random-init weights")


def build(config, device):
    branches = {"low": create_low_intensity_model(config), "medium": create_medium_intensity_model(config),
                "high": create_high_intensity_model(config)}
    clf = create_classifier(config)
    for name, m in branches.items():
        load_if_present(m, os.path.join(config["dehazing"]["checkpoint_dir"], f"{name}_best.pth"), "model_state_dict")
    load_if_present(clf, os.path.join(config["classifier"]["checkpoint_dir"], "classifier_best.pth"), "model_state_dict")
    router = create_router(branches, clf, config)
    return branches, clf, router.eval().to(device)


def psnr(a, b):
    """Batch-mean PSNR from the on-device metrics kernel (evaluation/metrics.py:13-36 without the per-image host copy)."""
    from adam_dehaze_b200.evaluation.metrics import image_metrics
    return image_metrics(a, b)[0].mean().item()


def ssim(a, b):
    from adam_dehaze_b200.evaluation.metrics import image_metrics
    return image_metrics(a, b)[1].mean().item()


def dataset_loader(config, split, device, keys=("hazy", "clear")):
    """The device input pipeline over config['dataset'][<split>_path]/<split>/... when that directory exists, else None."""
    ds = config["dataset"]
    root = ds["train_path"] if split == "train" else ds["val_path"] if split == "val" else ds["test_path"]
    if not os.path.isdir(os.path.join(root, split)):
        return None
    from adam_dehaze_b200.data.pipeline import get_dataloader
    return get_dataloader(config, split, device=device, keys=keys)


def evaluate(config, args, device, root, loader=None):
    branches, clf, router = build(config, device)
    loader = loader if loader is not None else dataset_loader(config, "test", device)
    if loader is not None:
        # evaluation/evaluate.py:33-175 over the real test split (--data_dir / config['dataset']['test_path'])
        from adam_dehaze_b200.evaluation.evaluate import evaluate_baseline_models, evaluate_joint_model
        results = {"baseline": evaluate_baseline_models(branches, loader, config, device),
                   "joint": evaluate_joint_model(router, clf, loader, config, device)}
        print(json.dumps(results))
        return results
    hazy, clear, labels = synth_hazy(args.synthetic, args.size[0], args.size[1], device, config["seed"])
    results = {}
    with torch.no_grad():
        for name, m in branches.items():
            results[f"branch_{name}_psnr"] = psnr(m(hazy), clear)
        logits, _ = clf(hazy)
        kind = config["routing"]["type"]
        if kind == "hard":
            out, info = router(hazy)                       # classifier-driven routes
            results["routes"] = info["intensity"].tolist()
            out_lab, _ = router(hazy, intensity=labels)    # ground-truth routes
            results["joint_psnr_gt_routes"] = psnr(out_lab, clear)
        elif kind == "soft":
            out, _ = router(hazy, logits)
        else:
            out, _ = router(hazy)
        results["joint_psnr"] = psnr(out, clear)
        results["joint_ssim"] = ssim(out, clear)
        results["classifier_accuracy"] = (logits.argmax(1) == labels).float().mean().item()
    path = os.path.join(config["evaluation"]["results_dir"], "evaluation.json")
    with open(path, "w") as fh:
        json.dump(results, fh, indent=1)
    print(json.dumps(results))
    return results


def main():
    args = parse_args()
    with open(args.config) as fh:
        config = yaml.safe_load(fh)
    if args.data_dir:
        for k in ("train_path", "val_path", "test_path"):
            config["dataset"][k] = args.data_dir
    if args.device:
        config["device"] = args.device
    if args.seed is not None:
        config["seed"] = args.seed
    seed_everything(config["seed"])
    root = experiment_dir(config, args.exp_name)
    device = torch.device(config["device"])
    if device.type != "cuda":
        raise RuntimeError("this build runs on B200 (sm_100a) only: pass --device cuda[:i]")
    if device.index is not None:
        torch.cuda.set_device(device)

    if args.mode == "evaluate":
        evaluate(config, args, device, root)
    elif args.mode == "demo":
        demo_dir = os.path.join(root, "demo")
        os.makedirs(demo_dir, exist_ok=True)
        _, _, router = build(config, device)
        hazy, _, _ = synth_hazy(3, args.size[0], args.size[1], device, config["seed"])
        with torch.no_grad():
            out, _ = router(hazy) if config["routing"]["type"] != "soft" else router(hazy, None)
        torch.save({"hazy": hazy.cpu(), "dehazed": out.cpu()}, os.path.join(demo_dir, "demo.pt"))
        print(f"demo outputs written to {demo_dir}")
    elif args.mode == "train_dehazing":
        # training/train_dehazing.py:16-223 per intensity level, on synthetic batches when no dataset directory exists
        from adam_dehaze_b200.training.train_dehazing import synthetic_loader, train_dehazing_model
        makers = {"low": create_low_intensity_model, "medium": create_medium_intensity_model, "high": create_high_intensity_model}
        bs = max(3, args.synthetic)
        train = synthetic_loader(2, bs, args.size[0], args.size[1], device, seed=config["seed"])
        val = synthetic_loader(1, bs, args.size[0], args.size[1], device, seed=config["seed"] + 1)
        real_train, real_val = dataset_loader(config, "train", device), dataset_loader(config, "val", device)
        if real_train is not None:
            train, val = real_train, real_val
        for level, mk in makers.items():
            print(f"Training {level} intensity dehazing model...")
            train_dehazing_model(mk(config), level, config, train_loader=train, val_loader=val, epochs=args.epochs, resume=args.resume)
    elif args.mode == "train_joint":
        # training/train_joint.py:29-318 — real loaders when the dataset directory exists, synthetic batches otherwise
        from adam_dehaze_b200.training.train_dehazing import synthetic_loader
        from adam_dehaze_b200.training.train_joint import train_joint_model
        train, val = dataset_loader(config, "train", device), dataset_loader(config, "val", device)
        if train is None:
            bs = max(3, args.synthetic)
            train = synthetic_loader(2, bs, args.size[0], args.size[1], device, seed=config["seed"])
            val = synthetic_loader(1, bs, args.size[0], args.size[1], device, seed=config["seed"] + 1)
        train_joint_model(config, train_loader=train, val_loader=val, epochs=args.epochs)
    else:
        raise NotImplementedError(f"--mode {args.mode} is outside the B200 hot path (SURVEY.md §2: out of scope)")


if __name__ == "__main__":
    main()
