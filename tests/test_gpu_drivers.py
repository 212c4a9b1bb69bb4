"""GPU: the reference-facing drivers end to end on tiny inputs — training/train_joint.py:29-318, evaluation/evaluate.py:33-175
and main.py --data_dir reading a dataset directory through the device input pipeline."""
import copy
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import CONFIG, ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _grad_on():
    torch.set_grad_enabled(True)       # (tests/test_oracle_golden.py switches autograd off at import time)
    yield
    from adam_dehaze_b200 import _lib
    torch.cuda.synchronize()
    _lib.call("adb_kernel_error_flag")


def _config(tmp_path, routing="soft"):
    cfg = copy.deepcopy(CONFIG)
    cfg["routing"] = {"type": routing, "temperature": 0.5}
    cfg["device"] = "cuda:0"
    cfg["seed"] = 42
    cfg["dataset"] = {"train_path": str(tmp_path / "data"), "val_path": str(tmp_path / "data"), "test_path": str(tmp_path / "data"),
                      "img_size": 64, "batch_size": 6, "num_workers": 2}
    cfg["classifier"]["checkpoint_dir"] = str(tmp_path / "ck" / "classifier")
    cfg["dehazing"]["checkpoint_dir"] = str(tmp_path / "ck" / "dehazing")
    for lvl in ("low", "medium", "high"):
        cfg["dehazing"][lvl]["learning_rate"] = 1e-4
    cfg["joint_training"].update({"learning_rate": 5e-5, "epochs": 2, "checkpoint_dir": str(tmp_path / "ck" / "joint")})
    cfg["evaluation"] = {"results_dir": str(tmp_path / "results")}
    return cfg


def _write_dataset(root, splits=("train", "val", "test"), per_level=2, shape=(80, 96)):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for split in splits:
        for level in ("low", "medium", "high"):
            for kind in ("hazy", "clear", "dehazed"):
                d = root / split / level / kind
                d.mkdir(parents=True)
                for i in range(per_level):
                    cv2.imwrite(str(d / f"{level}_{i}.png"), rng.integers(0, 256, shape + (3,), dtype=np.uint8))


def test_train_joint_driver_synthetic_batches(tmp_path):
    from adam_dehaze_b200.training.loss import DehazingLoss, JointLoss
    from adam_dehaze_b200.training.train_dehazing import synthetic_loader
    from adam_dehaze_b200.training.train_joint import train_joint_model
    cfg = _config(tmp_path)
    torch.manual_seed(42)
    train = synthetic_loader(2, 3, 64, 64, "cuda:0", seed=1)
    val = synthetic_loader(1, 3, 64, 64, "cuda:0", seed=2)
    crit = JointLoss(1.0, 0.2, 0.5, dehazing_loss=DehazingLoss(1.0, 0.0, 0.0))
    router, models, clf = train_joint_model(cfg, train_loader=train, val_loader=val, epochs=2, criterion=crit)
    ck = torch.load(os.path.join(cfg["joint_training"]["checkpoint_dir"], "best_model.pth"), map_location="cpu")
    assert set(ck) == {"epoch", "router_state_dict", "low_model_state_dict", "medium_model_state_dict", "high_model_state_dict",
                       "classifier_state_dict", "optimizer_state_dict", "val_psnr", "val_ssim", "val_loss"}      # train_joint.py:268-279
    assert np.isfinite(ck["val_loss"]) and ck["val_psnr"] > 0
    # the checkpoint loads back with strict=True (evaluate.py:117-122)
    router.load_state_dict(ck["router_state_dict"])
    clf.load_state_dict(ck["classifier_state_dict"])
    for k in ("low", "medium", "high"):
        models[k].load_state_dict(ck[f"{k}_model_state_dict"])


def test_evaluate_and_train_from_a_dataset_directory(tmp_path):
    """--data_dir is read: evaluate over <root>/test through the device pipeline (baseline + joint loops), and one epoch of
    train_dehazing on <root>/train."""
    _write_dataset(tmp_path / "data")
    cfg_path = os.path.join(ROOT, "config", "config.yaml")
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), "--mode", "evaluate", "--config", cfg_path, "--data_dir",
                          str(tmp_path / "data"), "--exp_name", "t_eval", "--device", "cuda:0"], cwd=str(tmp_path), env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert set(res) == {"baseline", "joint"}
    assert set(res["baseline"]) == {"low_intensity", "medium_intensity", "high_intensity"}
    assert all(res["baseline"][k]["samples"] == 2 for k in res["baseline"])
    assert res["joint"]["all"]["samples"] == 6 and 0.0 <= res["joint"]["classifier_accuracy"] <= 1.0
    assert "Loaded 6 samples for test split" in out.stdout


def test_evaluate_loops_with_an_injected_loader(tmp_path):
    from helpers import make_branch, make_classifier
    from adam_dehaze_b200.evaluation.evaluate import evaluate_baseline_models, evaluate_joint_model
    from adam_dehaze_b200.models.routing import create_router
    from adam_dehaze_b200.training.train_dehazing import synthetic_loader
    import adam_oracle as oracle
    cfg = _config(tmp_path, routing="hard")
    branches = {k: make_branch(k).cuda() for k in ("low", "medium", "high")}
    clf = make_classifier().cuda()
    router = create_router(branches, clf, cfg).eval()
    loader = synthetic_loader(2, 6, 64, 64, "cuda:0", seed=5)
    base = evaluate_baseline_models(branches, loader, cfg, "cuda:0")
    assert all(base[c]["samples"] == 4 for c in ("low_intensity", "medium_intensity", "high_intensity"))
    # the baseline numbers are the oracle's: each image through the branch of its own level
    sds = {k: m.state_dict() for k, m in branches.items()}
    want = {0: [], 1: [], 2: []}
    for batch in loader:
        ref, _, _ = oracle.hard_route(sds, batch["hazy"], intensity=batch["intensity"])
        for i, lab in enumerate(batch["intensity"].tolist()):
            mse = torch.mean((ref[i] - batch["clear"][i]) ** 2).item()
            want[lab].append(10 * np.log10(1.0 / mse))
    for k, cat in enumerate(("low_intensity", "medium_intensity", "high_intensity")):
        assert abs(base[cat]["psnr"] - float(np.mean(want[k]))) <= 0.05
    joint = evaluate_joint_model(router, clf, loader, cfg, "cuda:0")
    assert joint["all"]["samples"] == 12 and os.path.exists(os.path.join(cfg["evaluation"]["results_dir"], "joint_results.json"))
