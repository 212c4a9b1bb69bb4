"""Drop-in for training/train_dehazing.py (reference train_dehazing.py:16-223): `train_dehazing_model(model, level, config)`.

Same loop — filter the batch by intensity level, zero_grad, forward in train() mode, DehazingLoss, backward, Adam
(weight_decay 1e-4), validation PSNR, `best_model.pth` with the reference's checkpoint keys — but every tensor op is a
libadb200 kernel: batch-statistics BatchNorm forward, dgrad/wgrad on tcgen05, one flat gradient all-reduce when
torch.distributed is initialised, one fused Adam launch.  The reference's cv2 dataset, TensorBoard writer and skimage
metrics are outside the hot path; loaders are injectable and default to the synthetic hazy recipe of
SURVEY.md 8d, and validation PSNR is computed on the device.
"""
import os

import torch

from .loss import get_dehazing_loss
from .optim import FlatAdam

_LEVEL = {"low": 0, "medium": 1, "high": 2}


def synthetic_loader(n_batches, batch, h, w, device, seed=42):
    """Batches {'hazy','clear','intensity'} like data/dataset.py:97-124: I = clip(J t + 0.8 (1 - t)), t = exp(-beta d),
    beta = {0.03, 0.06, 0.09}[intensity] (utils/helpers.py:241-258)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    d = (0.3 + 0.7 * torch.sqrt((xx - 0.5) ** 2 + (yy - 0.2) ** 2)) * 100.0
    betas = torch.tensor([0.03, 0.06, 0.09])
    out = []
    for _ in range(n_batches):
        clear = torch.rand(batch, 3, h, w, generator=g)
        labels = torch.arange(batch) % 3
        t = torch.exp(-betas[labels].view(batch, 1, 1, 1) * d)
        hazy = torch.clamp(clear * t + 0.8 * (1 - t), 0, 1)
        out.append({"hazy": hazy.to(device), "clear": clear.to(device), "intensity": labels.to(device)})
    return out


def batch_psnr(pred, target):
    """Mean over the batch of 10 log10(1 / mse_i) (skimage peak_signal_noise_ratio, data_range=1; train_dehazing.py:146-159),
    on the device through adb_image_metrics — the same kernel evaluation/metrics.py reports from."""
    from ..evaluation.metrics import image_metrics
    psnr, _ = image_metrics(pred.detach(), target)
    return psnr.mean()


class LossMeter:
    """Every step's loss value on the host without draining the launch queue.

    The reference reads `loss.item()` right after each step (train_dehazing.py:95), which makes the host wait for the whole
    step before it can queue the next one.  Here step i's scalar is copied to pinned host memory on the stream (4 bytes
    D2H per step) and read once step i+1 has been queued; `flush()` reads the last one.  `total` / `count` / `last` hold
    the same numbers the reference accumulates, one step later."""

    def __init__(self):
        self._slots = [torch.empty(1, dtype=torch.float32, pin_memory=True) for _ in range(2)]
        self._events = [torch.cuda.Event() for _ in range(2)]
        self._pending, self._next = None, 0
        self.total, self.count, self.last = 0.0, 0, None

    def _read(self, i):
        self._events[i].synchronize()
        self.last = float(self._slots[i][0])
        self.total += self.last
        self.count += 1

    def push(self, loss):
        """Queue the D2H copy of this step's loss; returns the previous step's value (None on the first call)."""
        i = self._next
        self._next ^= 1
        self._slots[i].copy_(loss.detach().reshape(1), non_blocking=True)
        self._events[i].record()
        prev, self._pending = self._pending, i
        if prev is not None:
            self._read(prev)
        return self.last

    def flush(self):
        if self._pending is not None:
            self._read(self._pending)
            self._pending = None
        return self.last


def train_step(model, criterion, optimizer, hazy, clear):
    """train_dehazing.py:86-96: zero_grad, forward, loss, backward, step.  Returns (loss tensor, components)."""
    optimizer.zero_grad()
    out = model(hazy)
    loss, parts = criterion(out, clear)
    loss.backward()
    optimizer.step()
    return loss, parts


def train_dehazing_model(model, intensity_level, config, train_loader=None, val_loader=None, epochs=30, criterion=None,
                         resume=False):
    device = torch.device(config["device"])
    if device.type != "cuda":
        raise RuntimeError("train_dehazing_model: this build trains on B200 (sm_100a) only — config['device'] must be cuda")
    model = model.to(device)
    lr = config["dehazing"][intensity_level]["learning_rate"]
    optimizer = FlatAdam(model.parameters(), lr=lr, weight_decay=0.0001)
    criterion = (criterion if criterion is not None else get_dehazing_loss(config)).to(device)
    if train_loader is None:
        size = config.get("dataset", {}).get("image_size", [256, 256])
        bs = config.get("dataset", {}).get("batch_size", 16)
        train_loader = synthetic_loader(4, bs, size[0], size[1], device, seed=config.get("seed", 42))
        val_loader = synthetic_loader(1, bs, size[0], size[1], device, seed=config.get("seed", 42) + 1)
    ck_dir = os.path.join(config["dehazing"]["checkpoint_dir"], intensity_level)
    os.makedirs(ck_dir, exist_ok=True)
    k = _LEVEL[intensity_level]
    best, bad_epochs, history = 0.0, 0, []
    first_epoch = 0
    last_path = os.path.join(ck_dir, "last_checkpoint.pth")
    if resume and os.path.exists(last_path):
        # main.py:50 parses --resume but the reference never acts on it; here the last epoch's model,
        # Adam moments / step count / learning rate and the plateau scheduler's history are restored and training continues
        ck = torch.load(last_path, map_location=device)
        model.load_state_dict(ck["model_state_dict"])
        optimizer.load_state_dict(ck["optimizer_state_dict"])
        best, bad_epochs, history = ck["best_val_psnr"], ck["bad_epochs"], list(ck["val_loss_history"])
        first_epoch = ck["epoch"] + 1
        print(f"Resumed {intensity_level} from {last_path} at epoch {first_epoch + 1}")
    for epoch in range(first_epoch, epochs):
        model.train()
        meter = LossMeter()
        for batch in train_loader:
            sel = batch["intensity"] == k
            if not bool(sel.any()):
                # train_dehazing.py:71-75 skips a batch that holds no sample of this level.  With replicas every rank must
                # still join the step's gradient all-reduce (FlatAdam: zero gradients + "no gradient here" flags), or the
                # ranks that did find samples would wait for this one forever.
                if getattr(optimizer, "world_size", lambda: 1)() > 1:
                    optimizer.zero_grad()
                    optimizer.step()
                continue
            hazy, clear = batch["hazy"][sel].to(device), batch["clear"][sel].to(device)
            loss, _ = train_step(model, criterion, optimizer, hazy, clear)
            meter.push(loss)
        meter.flush()
        tot, nb = meter.total, meter.count
        model.eval()
        vl, vp, vs = 0.0, 0.0, 0
        with torch.no_grad():
            for batch in (val_loader or []):
                sel = batch["intensity"] == k
                if not bool(sel.any()):
                    continue
                hazy, clear = batch["hazy"][sel].to(device), batch["clear"][sel].to(device)
                out = model(hazy)
                loss, _ = criterion(out, clear)
                vl += loss.item() * hazy.size(0)
                vp += batch_psnr(out, clear).item() * hazy.size(0)
                vs += hazy.size(0)
        vl, vp = (vl / vs, vp / vs) if vs else (0.0, 0.0)
        # ReduceLROnPlateau(mode='min', factor=0.5, patience=5), train_dehazing.py:40-42
        if history and vl >= min(history):
            bad_epochs += 1
            if bad_epochs > 5:
                optimizer.lr *= 0.5
                bad_epochs = 0
        else:
            bad_epochs = 0
        history.append(vl)
        print(f"Epoch {epoch + 1}/{epochs}:\n  Train Loss: {tot / max(1, nb):.4f}\n  Val Loss: {vl:.4f}, Val PSNR: {vp:.2f}")
        if vp > best or epoch == 0:
            best = vp
            torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                        "val_psnr": vp, "val_ssim": None, "val_loss": vl}, os.path.join(ck_dir, "best_model.pth"))
        torch.save({"epoch": epoch, "model_state_dict": model.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                    "val_psnr": vp, "val_ssim": None, "val_loss": vl, "best_val_psnr": best, "bad_epochs": bad_epochs,
                    "val_loss_history": history}, last_path)
    best_ck = torch.load(os.path.join(ck_dir, "best_model.pth"), map_location="cpu")
    model.load_state_dict(best_ck["model_state_dict"])
    return model
