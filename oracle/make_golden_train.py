"""Generate tests/golden/train_*.pt: one training step of the UNMODIFIED reference branch models (model.train(), L1 loss,
loss.backward(); training/train_dehazing.py:66-92) so the oracle's train_mode() restatement is pinned.

Run in the build container only (the GPU box has no /root/reference):  python oracle/make_golden_train.py

Light stores every gradient; Medium / Complex (7.2 M / 16.3 M parameters) store per-parameter summaries
(L2 norm, sum, first 8 values) to keep the fixtures small.  Weights are not stored: models are built right after
torch.manual_seed(42) exactly like tests/helpers.make_branch, and the state_dict fingerprint is recorded.
"""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import CONFIG, OUT, REF, SEED, fingerprint, rand_image  # noqa: E402


def main():
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    sys.path.insert(0, REF)
    from models.dehazing.high_intensity import create_high_intensity_model
    from models.dehazing.low_intensity import create_low_intensity_model
    from models.dehazing.medium_intensity import create_medium_intensity_model
    mk = {"low": create_low_intensity_model, "medium": create_medium_intensity_model, "high": create_high_intensity_model}
    for name, (n, h, w) in {"low": (2, 32, 48), "medium": (2, 64, 64), "high": (2, 64, 64)}.items():
        torch.manual_seed(SEED)
        m = mk[name](CONFIG).train()
        fp = fingerprint(m.state_dict())
        x, tgt = rand_image(n, h, w, 5), rand_image(n, h, w, 6)
        out = m(x)
        loss = torch.nn.L1Loss()(out, tgt)
        loss.backward()
        grads = {}
        for k, p in m.named_parameters():
            g = p.grad.detach()
            grads[k] = g.clone() if name == "low" else {"norm": g.norm().item(), "sum": g.double().sum().item(),
                                                        "head": g.flatten()[:8].clone(), "shape": tuple(g.shape)}
        stats = {k: v.clone() for k, v in m.state_dict().items() if "running_" in k or "num_batches" in k}
        if name != "low":   # keep a handful of running statistics only
            keep = sorted(stats)[:6]
            stats = {k: stats[k] for k in keep}
        torch.save({"fingerprint": fp, "shape": (n, h, w), "x_seed": 5, "target_seed": 6, "out": out.detach().clone(),
                    "loss": loss.detach().clone(), "grads": grads, "stats_after": stats},
                   os.path.join(OUT, f"train_{name}.pt"))
        print(name, float(loss), os.path.getsize(os.path.join(OUT, f"train_{name}.pt")))


if __name__ == "__main__":
    main()
