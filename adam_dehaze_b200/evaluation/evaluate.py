"""Drop-in for the image-quality half of evaluation/evaluate.py (reference evaluate.py:33-92 `evaluate_baseline_models`,
:94-175 `evaluate_joint_model`): the same loops over a test loader, with the loader injectable (reference batch dicts), the
per-image model calls of the baseline loop replaced by one routed batch pass (`HardRouter.forward(x, intensity=labels)` is
exactly "each image through the branch of its ground-truth level"), and the metrics accumulated on the device.  The detection
sweep (evaluate.py:177-330, pycocotools) is out of scope."""
import json
import os

import torch

from ..models.routing import GatedRouter, HardRouter
from .metrics import ImageQualityMetrics

CATEGORY = ("low_intensity", "medium_intensity", "high_intensity")


def _add_by_category(metrics, out, clear, labels):
    lab = labels.tolist()          # labels come from the loader (host-side metadata in the reference too)
    for k, cat in enumerate(CATEGORY):
        idx = [i for i, v in enumerate(lab) if v == k]
        if idx:
            sel = torch.tensor(idx, device=out.device)
            metrics.add_sample(out.index_select(0, sel), clear.index_select(0, sel), cat)


def evaluate_baseline_models(models, loader, config, device, with_lpips=False):
    """Each branch on the images of its own level (evaluate.py:63-86).  models: {'low','medium','high'} -> nn.Module."""
    router = HardRouter(models, classifier=None).to(device).eval()
    metrics = ImageQualityMetrics(device=device, with_lpips=with_lpips)
    with torch.no_grad():
        for batch in loader:
            hazy, clear, labels = batch["hazy"].to(device), batch["clear"].to(device), batch["intensity"].to(device)
            out, _ = router(hazy, intensity=labels)
            _add_by_category(metrics, out, clear, labels)
    results = metrics.compute_averages()
    _save(results, config, "baseline_results.json")
    return results


def evaluate_joint_model(router, classifier, loader, config, device, with_lpips=False):
    """classifier -> router on every batch (evaluate.py:145-168); returns per-category and overall averages plus the
    classifier's accuracy against the loader's labels."""
    metrics = ImageQualityMetrics(device=device, with_lpips=with_lpips)
    correct = torch.zeros((), dtype=torch.int64, device=device)
    total = 0
    with torch.no_grad():
        for batch in loader:
            hazy, clear, labels = batch["hazy"].to(device), batch["clear"].to(device), batch["intensity"].to(device)
            logits, _ = classifier(hazy)
            if isinstance(router, HardRouter):
                out, _ = router(hazy)                  # (evaluate.py:156 passes the logits positionally: the `intensity` trap)
            elif isinstance(router, GatedRouter):
                out, _ = router(hazy)
            else:
                out, _ = router(hazy, logits)
            _add_by_category(metrics, out, clear, labels)
            metrics.add_sample(out, clear, "all")
            correct += (logits.argmax(1) == labels).sum()
            total += hazy.size(0)
    results = metrics.compute_averages()
    results["classifier_accuracy"] = correct.item() / max(1, total)
    _save(results, config, "joint_results.json")
    return results


def _save(results, config, name):
    d = config.get("evaluation", {}).get("results_dir")
    if d:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as fh:
            json.dump(results, fh, indent=1)
