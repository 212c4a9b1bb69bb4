"""Drop-in for training/train_joint.py (reference train_joint.py:29-318): fine-tune HDEN + router + the three branches
together.  One step = train_joint.py:129-150: classifier(x) -> router(x, logits) -> JointLoss -> backward -> Adam.

Differences from the reference, all on the host side of the same step:
  * loaders are injectable (`train_loader` / `val_loader` yield the reference's batch dicts); by default the device input
    pipeline (adam_dehaze_b200/data/pipeline.get_dataloader) reads the directories named in config['dataset'], and synthetic
    batches stand in when they do not exist;
  * the optimizer is FlatAdam (one fused launch + one NCCL all-reduce per step); the reference's parameter list holds every
    branch parameter twice (train_joint.py:80-82) — de-duplicated here;
  * the loss is read through LossMeter (no per-step host stall) and validation PSNR / SSIM come from adb_image_metrics;
  * TensorBoard logging is out of scope.
Checkpoints carry the reference's keys (train_joint.py:268-279), so either side can load the other's files.
"""
import os

import torch

from ..models.classifier import create_classifier
from ..models.dehazing.high_intensity import create_high_intensity_model
from ..models.dehazing.low_intensity import create_low_intensity_model
from ..models.dehazing.medium_intensity import create_medium_intensity_model
from ..models.routing import create_router
from .loss import get_joint_loss
from .optim import FlatAdam
from .train_dehazing import LossMeter, synthetic_loader


def load_pretrained_model(model, path):
    """train_joint.py:16-27: a missing checkpoint is not an error."""
    if os.path.exists(path):
        model.load_state_dict(torch.load(path, map_location="cpu")["model_state_dict"])
        print(f"Loaded pretrained weights from {path}")
    else:
        print(f"Pretrained weights not found at {path}. Starting from scratch.")
    return model


def _route(router, hazy, logits):
    """train_joint.py:141-144 calls router(hazy, logits) positionally.  For SoftRouter that is `classifier_logits`; for
    HardRouter it would bind the logits to `intensity` — there the class ids are passed instead; GatedRouter
    takes no second argument."""
    from ..models.routing import GatedRouter, HardRouter
    if isinstance(router, HardRouter):
        return router(hazy, logits.argmax(1))
    if isinstance(router, GatedRouter):
        return router(hazy)
    return router(hazy, logits)


def joint_step(classifier, router, criterion, optimizer, hazy, clear, labels):
    """One optimisation step (train_joint.py:129-150).  Returns (loss tensor, components dict)."""
    optimizer.zero_grad()
    logits, _ = classifier(hazy)
    out, _ = _route(router, hazy, logits)
    loss, parts = criterion(out, clear, logits, labels)
    loss.backward()
    optimizer.step()
    return loss, parts


def _checkpoint(epoch, router, models, classifier, optimizer, val_psnr, val_ssim, val_loss):
    return {"epoch": epoch, "router_state_dict": router.state_dict(), "low_model_state_dict": models["low"].state_dict(),
            "medium_model_state_dict": models["medium"].state_dict(), "high_model_state_dict": models["high"].state_dict(),
            "classifier_state_dict": classifier.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
            "val_psnr": val_psnr, "val_ssim": val_ssim, "val_loss": val_loss}


def train_joint_model(config, train_loader=None, val_loader=None, epochs=None, criterion=None):
    from ..evaluation.metrics import image_metrics
    device = torch.device(config["device"])
    if device.type != "cuda":
        raise RuntimeError("train_joint_model: this build trains on B200 (sm_100a) only — config['device'] must be cuda")
    print("Creating classifier model...")
    classifier = create_classifier(config)
    print("Creating dehazing models...")
    models = {"low": create_low_intensity_model(config), "medium": create_medium_intensity_model(config),
              "high": create_high_intensity_model(config)}
    load_pretrained_model(classifier, os.path.join(config["classifier"]["checkpoint_dir"], "best_model.pth"))
    for level, m in models.items():
        load_pretrained_model(m, os.path.join(config["dehazing"]["checkpoint_dir"], level, "best_model.pth"))
    print("Creating routing mechanism...")
    router = create_router(models, classifier, config).to(device)
    seen, params = set(), []
    for p in list(router.parameters()) + [p for m in models.values() for p in m.parameters()]:
        if id(p) not in seen:
            seen.add(id(p))
            params.append(p)
    optimizer = FlatAdam(params, lr=config["joint_training"]["learning_rate"], weight_decay=0.0001)
    criterion = (criterion if criterion is not None else get_joint_loss(config)).to(device)
    if train_loader is None:
        ds = config.get("dataset", {})
        if os.path.isdir(os.path.join(ds.get("train_path", ""), "train")):
            from ..data.pipeline import get_dataloader
            train_loader = get_dataloader(config, "train", device=device, keys=("hazy", "clear"))
            val_loader = get_dataloader(config, "val", device=device, keys=("hazy", "clear"))
        else:
            s = ds.get("img_size", 256)
            train_loader = synthetic_loader(4, ds.get("batch_size", 16), s, s, device, seed=config.get("seed", 42))
            val_loader = synthetic_loader(1, ds.get("batch_size", 16), s, s, device, seed=config.get("seed", 42) + 1)
    ck_dir = config["joint_training"]["checkpoint_dir"]
    os.makedirs(ck_dir, exist_ok=True)
    best, bad_epochs, history = 0.0, 0, []
    epochs = config["joint_training"]["epochs"] if epochs is None else epochs
    for epoch in range(epochs):
        router.train()                                   # classifier and branches are sub-modules of the router
        meter, dmeter, cmeter = LossMeter(), LossMeter(), LossMeter()
        for batch in train_loader:
            hazy, clear, labels = batch["hazy"].to(device), batch["clear"].to(device), batch["intensity"].to(device)
            loss, parts = joint_step(classifier, router, criterion, optimizer, hazy, clear, labels)
            meter.push(loss); dmeter.push(parts["dehazing"]); cmeter.push(parts["classification"])
        for m_ in (meter, dmeter, cmeter):
            m_.flush()
        nb = max(1, meter.count)
        router.eval()
        acc = torch.zeros(5, dtype=torch.float64, device=device)          # loss, dehaze, class, psnr, ssim sums (device)
        vs = 0
        with torch.no_grad():
            for batch in (val_loader or []):
                hazy, clear, labels = batch["hazy"].to(device), batch["clear"].to(device), batch["intensity"].to(device)
                logits, _ = classifier(hazy)
                out, _ = _route(router, hazy, logits)
                loss, parts = criterion(out, clear, logits, labels)
                psnr, ssim = image_metrics(out, clear)
                n = hazy.size(0)
                acc += torch.stack([loss.double() * n, parts["dehazing"].double() * n, parts["classification"].double() * n,
                                    psnr.double().sum(), ssim.double().sum()])
                vs += n
        vl, vd, vc, vp, vssim = (acc / max(1, vs)).tolist()                # one host read per epoch
        if history and vl >= min(history):                                 # ReduceLROnPlateau(mode='min', factor=0.5, patience=3)
            bad_epochs += 1
            if bad_epochs > 3:
                optimizer.lr *= 0.5
                bad_epochs = 0
        else:
            bad_epochs = 0
        history.append(vl)
        print(f"Epoch {epoch + 1}/{epochs}:")
        print(f"  Train Loss: {meter.total / nb:.4f} (Dehaze: {dmeter.total / nb:.4f}, Class: {cmeter.total / nb:.4f})")
        print(f"  Val Loss: {vl:.4f} (Dehaze: {vd:.4f}, Class: {vc:.4f})")
        print(f"  Val PSNR: {vp:.2f} dB, Val SSIM: {vssim:.4f}")
        if vp > best or epoch == 0:
            best = vp
            torch.save(_checkpoint(epoch, router, models, classifier, optimizer, vp, vssim, vl), os.path.join(ck_dir, "best_model.pth"))
            print(f"Saved best model with validation PSNR: {vp:.2f} dB")
        if (epoch + 1) % 5 == 0:
            torch.save(_checkpoint(epoch, router, models, classifier, optimizer, vp, vssim, vl),
                       os.path.join(ck_dir, f"checkpoint_epoch_{epoch + 1}.pth"))
    bestck = torch.load(os.path.join(ck_dir, "best_model.pth"), map_location="cpu")
    router.load_state_dict(bestck["router_state_dict"])
    return router, models, classifier
