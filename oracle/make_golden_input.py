"""Golden vectors for the input transform from the real thing: cv2 (the reference's own dependency) + torchvision's ToTensor,
the exact call chain of /root/reference/data/dataset.py:76-99.  Run in the build container (cv2 present):

    python oracle/make_golden_input.py        ->  tests/golden/input_pipeline.pt
"""
import os

import cv2
import numpy as np
import torch
from torchvision import transforms

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [  # (source h, w) -> img_size
    ((37, 53), 32), ((20, 24), 32), ((64, 64), 32), ((32, 32), 32), ((33, 31), 32), ((100, 60), 48), ((16, 200), 64), ((96, 96), 64),
]


def main():
    rng = np.random.default_rng(42)
    out = []
    for (h, w), size in CASES:
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)                              # dataset.py:77
        if img.shape[0] != size or img.shape[1] != size:
            img = cv2.resize(img, (size, size))                                 # dataset.py:86-87
        t = transforms.ToTensor()(img)                                          # dataset.py:96
        out.append({"bgr": torch.from_numpy(bgr), "size": size, "tensor": t,
                    "hflip": transforms.functional.hflip(t), "vflip": transforms.functional.vflip(t)})
    path = os.path.join(ROOT, "tests", "golden", "input_pipeline.pt")
    torch.save({"cv2": cv2.__version__, "cases": out}, path)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
