"""Input pipeline — the reference loader's per-sample work (data/dataset.py:9-124, 243-258) re-cut for a GPU that consumes
> 1.5 k images/s: files are decoded on host threads (cv2.imread, as the reference), the decoded uint8 HWC image is copied
into pinned memory and uploaded AS IT IS (3 bytes per pixel instead of the 12 of a float tensor, no CPU resize, no
CPU ToTensor), and one libadb200 launch per image group does BGR->RGB, cv2-exact bilinear resize, /255 and the training
flips (adb_image_u8_to_f32), writing straight into the NCHW fp32 batch the models read.

`HazyImageFolder` lists samples exactly like `HazyImageDataset.__init__` (dataset.py:21-56); `DeviceLoader` yields the same
batch dicts as the reference's DataLoader ('hazy', 'clear', 'dehazed', 'intensity', 'name') with the tensors already on the
device.  Train-split augmentation: the flips are applied (same draw for the three images of a sample, dataset.py:102-115);
ColorJitter(0.1, 0.1) is not reproduced (its draws depend on torchvision's RNG call sequence) — pass `augment=False` for
bit-exact agreement with the reference's eval/test transform.
"""
import os
import random
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .. import _lib

INTENSITY = {"low": 0, "medium": 1, "high": 2}


def u8_to_tensor(src, size=None, bgr=True, flips=None, out=None):
    """src: uint8 CUDA tensor [n, hs, ws, 3] (decoded images of one size) -> float32 NCHW [n, 3, hd, wd] in [0, 1].
    size = (hd, wd) resizes like cv2.resize(INTER_LINEAR) (bit-exact); flips: uint8 CUDA [n], bit 0 horizontal, bit 1 vertical."""
    if not (src.is_cuda and src.dtype == torch.uint8 and src.dim() == 4 and src.shape[3] == 3 and src.is_contiguous()):
        raise RuntimeError("u8_to_tensor: expected a contiguous uint8 CUDA tensor [n, h, w, 3] — this package has no CPU path")
    n, hs, ws, _ = src.shape
    hd, wd = (hs, ws) if size is None else (int(size[0]), int(size[1]))
    if out is None:
        out = torch.empty((n, 3, hd, wd), dtype=torch.float32, device=src.device)
    assert out.is_contiguous() and tuple(out.shape) == (n, 3, hd, wd) and out.dtype == torch.float32
    _lib.call("adb_image_u8_to_f32", _lib.ptr(src), n, hs, ws, hs * ws * 3, int(bool(bgr)), _lib.ptr(flips), _lib.ptr(out), hd, wd,
              _lib.current_stream())
    return out


class HazyImageFolder:
    """Sample list of `<root>/<split>/<low|medium|high>/{hazy,clear,dehazed}/<name>` (dataset.py:21-56): a sample exists when
    all three files do."""

    def __init__(self, root_dir, split="train", img_size=256):
        self.root_dir = os.path.join(root_dir, split)
        self.split, self.img_size = split, img_size
        self.samples = []
        for intensity in ("low", "medium", "high"):
            hazy_dir = os.path.join(self.root_dir, intensity, "hazy")
            if not os.path.isdir(hazy_dir):
                continue
            for name in sorted(os.listdir(hazy_dir)):
                if not (name.endswith(".jpg") or name.endswith(".png")):
                    continue
                paths = {k: os.path.join(self.root_dir, intensity, k, name) for k in ("hazy", "clear", "dehazed")}
                if all(os.path.exists(p) for p in paths.values()):
                    self.samples.append(dict(paths, intensity=INTENSITY[intensity], name=name))
        print(f"Loaded {len(self.samples)} samples for {split} split")

    def __len__(self):
        return len(self.samples)


def _imread(path):
    import cv2
    img = cv2.imread(path)                  # BGR uint8 HWC, as dataset.py:76
    if img is None:
        raise FileNotFoundError(path)
    return img


class DeviceLoader:
    """Iterates `folder` in batches; decode on `workers` host threads, upload uint8, transform on the device.

    keys: which of the three images of a sample to produce (the hot path needs 'hazy' (+ 'clear' for metrics / training))."""

    def __init__(self, folder, batch_size, device="cuda", shuffle=False, augment=None, workers=8, keys=("hazy", "clear", "dehazed"),
                 seed=0, drop_last=False):
        self.folder, self.batch_size, self.device = folder, int(batch_size), torch.device(device)
        self.shuffle, self.keys, self.drop_last = shuffle, tuple(keys), drop_last
        self.augment = (folder.split == "train") if augment is None else bool(augment)
        self.pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self.rng = random.Random(seed)
        self._pinned = {}

    def __len__(self):
        n = len(self.folder)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _staging(self, key, shape):
        buf = self._pinned.get((key, shape))
        if buf is None:
            buf = self._pinned[(key, shape)] = torch.empty(shape, dtype=torch.uint8).pin_memory()
        return buf

    def _convert(self, imgs, out, flips_host):
        """imgs: list of decoded HWC arrays (any sizes) -> rows of `out`; images of one size share an upload and a launch."""
        size = out.shape[2:]
        groups = {}
        for i, im in enumerate(imgs):
            groups.setdefault(im.shape[:2], []).append(i)
        for (h, w), idx in groups.items():
            stage = self._staging(("u8", h, w), (self.batch_size, h, w, 3))
            for j, i in enumerate(idx):
                stage[j].copy_(torch.from_numpy(imgs[i]))
            dev = stage[:len(idx)].to(self.device, non_blocking=True)
            flips = None
            if flips_host is not None:
                flips = torch.tensor([flips_host[i] for i in idx], dtype=torch.uint8).to(self.device)
            if len(groups) == 1:
                u8_to_tensor(dev, size=size, flips=flips, out=out)
            else:
                tmp = u8_to_tensor(dev, size=size, flips=flips)
                out.index_copy_(0, torch.tensor(idx, device=self.device), tmp)

    def __iter__(self):
        order = list(range(len(self.folder)))
        if self.shuffle:
            self.rng.shuffle(order)
        s = self.folder.img_size
        size = (s, s) if isinstance(s, int) else tuple(s)
        for b0 in range(0, len(order), self.batch_size):
            ids = order[b0:b0 + self.batch_size]
            if self.drop_last and len(ids) < self.batch_size:
                break
            samples = [self.folder.samples[i] for i in ids]
            decoded = {k: list(self.pool.map(_imread, [smp[k] for smp in samples])) for k in self.keys}
            flips = [self.rng.randrange(4) for _ in samples] if self.augment else None      # one draw per sample, all three images
            batch = {"intensity": torch.tensor([smp["intensity"] for smp in samples], dtype=torch.long, device=self.device),
                     "name": [smp["name"] for smp in samples]}
            for k in self.keys:
                out = torch.empty((len(samples), 3) + size, dtype=torch.float32, device=self.device)
                self._convert(decoded[k], out, flips)
                batch[k] = out
            yield batch


def get_dataloader(config, split="train", device=None, **kw):
    """Factory with the reference's config keys (dataset.py:243-258)."""
    ds = config["dataset"]
    root = ds["train_path"] if split == "train" else ds["val_path"] if split == "val" else ds["test_path"]
    folder = HazyImageFolder(root, split=split, img_size=ds["img_size"])
    return DeviceLoader(folder, ds["batch_size"], device=device or config.get("device", "cuda"), shuffle=(split == "train"),
                        workers=ds.get("num_workers", 8), **kw)
