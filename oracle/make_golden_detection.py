"""Generate tests/golden/detection_handoff.pt from the UNMODIFIED reference `IntegratedDetectionSystem`
(/root/reference/models/detection.py:75-127) on CPU.  Build container only:

    python oracle/make_golden_detection.py

The dehazing model and the detector are stubs outside the arithmetic being pinned (the per-image ImageNet normalisation
between them, detection.py:109-121): the dehazer returns (0.5 x + 0.25, {}), the detector records what it is given.
"""
import os
import sys

import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
SEED, SHAPE = 3, (3, 40, 72)


class Dehazer(nn.Module):
    def forward(self, x):
        return x * 0.5 + 0.25, {}


class Detector(nn.Module):
    def __init__(self):
        super().__init__()
        self.w = nn.Parameter(torch.zeros(1))
        self.seen = None

    def forward(self, images, targets=None):
        self.seen = images
        return [{} for _ in images]


def main():
    sys.path.insert(0, REF)
    from models.detection import create_integrated_system
    x = torch.rand(SHAPE[0], 3, SHAPE[1], SHAPE[2], generator=torch.Generator().manual_seed(SEED))
    system = create_integrated_system(Dehazer(), Detector()).eval()
    with torch.no_grad():
        results, dehazed = system(x)
    seen = system.detection_model.seen
    assert isinstance(seen, list) and len(seen) == SHAPE[0] and len(results) == SHAPE[0]
    assert all(not p.requires_grad for p in system.detection_model.parameters())
    torch.save({"seed": SEED, "shape": SHAPE, "dehazed": dehazed.clone(), "normalized": torch.stack(list(seen)).clone()},
               os.path.join(OUT, "detection_handoff.pt"))
    print("detection hand-off:", torch.stack(list(seen)).shape, float(torch.stack(list(seen)).abs().max()))


if __name__ == "__main__":
    main()
