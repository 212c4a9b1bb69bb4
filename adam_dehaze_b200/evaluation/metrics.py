"""Drop-in for the image-quality half of evaluation/metrics.py (reference metrics.py:13-120): `calculate_image_metrics`
and `ImageQualityMetrics`, computed on the device by adb_image_metrics (PSNR, SSIM with skimage's defaults) and the
LPIPS-alex trunk of training/loss.py — no per-image device->host copy, no skimage.  Results come back as Python floats
only when `compute_averages()` / `.item()` is asked for (one sync per report instead of one per image).

The COCO detection metrics of the reference (metrics.py:122-330, pycocotools) are outside the hot path.
"""
from collections import defaultdict

import torch

from .. import _lib


def image_metrics(pred, target):
    """pred/target: [B,3,H,W] (or [3,H,W]) fp32 CUDA in [0,1] -> (psnr [B], ssim [B]) fp32 device tensors."""
    if pred.dim() == 3:
        pred, target = pred.unsqueeze(0), target.unsqueeze(0)
    if not pred.is_cuda:
        raise RuntimeError("image_metrics: expected CUDA tensors — this package runs on B200 (sm_100a) only and has no CPU path")
    pred, target = pred.contiguous().float(), target.contiguous().float()
    n, c, h, w = pred.shape
    if c != 3 or target.shape != pred.shape:
        raise ValueError(f"image_metrics: expected two [B,3,H,W] batches, got {tuple(pred.shape)} and {tuple(target.shape)}")
    scratch = torch.empty(2 * n, dtype=torch.float64, device=pred.device)
    psnr = torch.empty(n, dtype=torch.float32, device=pred.device)
    ssim = torch.empty(n, dtype=torch.float32, device=pred.device)
    _lib.call("adb_image_metrics", _lib.ptr(pred), _lib.ptr(target), n, h, w, _lib.ptr(scratch), _lib.ptr(psnr), _lib.ptr(ssim),
              _lib.current_stream())
    return psnr, ssim


def calculate_image_metrics(pred, target):
    """metrics.py:13-36 for one image ([3,H,W] CUDA tensors): {'psnr', 'ssim'} as floats."""
    psnr, ssim = image_metrics(pred, target)
    return {"psnr": psnr[0].item(), "ssim": ssim[0].item()}


class ImageQualityMetrics:
    """metrics.py:38-120: accumulate PSNR / SSIM / LPIPS per category; averages are reduced on the device."""

    def __init__(self, device="cuda", with_lpips=True):
        self.device = device
        self.results = defaultdict(list)      # category -> list of (psnr[B], ssim[B], lpips[B]|None) device tensors
        self.lpips_fn = None
        if with_lpips:
            from ..training.loss import PerceptualLoss
            self.lpips_fn = PerceptualLoss().to(device)

    def add_sample(self, pred, target, category=None):
        """pred/target: [3,H,W] or [B,3,H,W] CUDA tensors (the reference takes one image at a time; batches are welcome)."""
        if pred.dim() == 3:
            pred, target = pred.unsqueeze(0), target.unsqueeze(0)
        psnr, ssim = image_metrics(pred, target)
        lp = None
        if self.lpips_fn is not None:
            with torch.no_grad():
                lp = self.lpips_fn(pred.contiguous().float(), target.contiguous().float()).reshape(-1)
        self.results[category or "all"].append((psnr, ssim, lp))

    def compute_averages(self):
        out = {}
        for cat, items in self.results.items():
            if not items:
                continue
            psnr = torch.cat([i[0] for i in items])
            ssim = torch.cat([i[1] for i in items])
            res = {"psnr": psnr.mean().item(), "ssim": ssim.mean().item()}
            if items[0][2] is not None:
                res["lpips"] = torch.cat([i[2] for i in items]).mean().item()
                if not getattr(self.lpips_fn, "pretrained", True):
                    # LPIPS needs its trained AlexNet + lin weights; without them the number only orders images consistently
                    res["lpips_weights"] = "random-init (not comparable with published LPIPS)"
            res["samples"] = int(psnr.numel())
            out[cat] = res
        return out

    def print_results(self):
        print("Image Quality Evaluation Results:")
        for cat, m in sorted(self.compute_averages().items()):
            print(f"\n{cat.upper()} ({m['samples']} samples):")
            for k in ("psnr", "ssim", "lpips"):
                if k in m:
                    note = f"  [{m['lpips_weights']}]" if k == "lpips" and "lpips_weights" in m else ""
                    print(f"  {k.upper()}: {m[k]:.4f}{note}")
