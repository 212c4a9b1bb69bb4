"""Generate tests/golden/*.pt by importing the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/make_golden.py

Stubs, all outside the arithmetic being pinned:
  * `timm` (absent offline) is an empty module — only the efficientnet arm of classifier.py uses it;
  * `lpips` (absent offline) is a stand-in whose LPIPS(x, t) returns mean((x-t)^2) per image as [B,1,1,1], so that
    DehazingLoss/JointLoss's *combination* arithmetic is pinned; LPIPS itself stays "parity unpinned";
  * torchvision.models.vgg16(pretrained=True) cannot download: it is redirected to seeded random init.
Weights are never stored: each model is constructed right after torch.manual_seed(SEED), and a sha256 fingerprint of its
state_dict is recorded so the drop-in modules can prove they initialise identically.
"""
import hashlib
import os
import random
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
SEED = 42


def seed_everything(seed):
    """The reference's recipe, utils/helpers.py:10-19."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def fingerprint(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rand_image(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, h, w, generator=g)


CONFIG = {
    "classifier": {"model": "resnet18", "num_classes": 3, "pretrained": False},
    "dehazing": {
        "low": {"model_type": "lightweight", "channels": 32, "blocks": 3},
        "medium": {"model_type": "standard", "channels": 64, "blocks": 6},
        "high": {"model_type": "complex", "channels": 96, "blocks": 9},
    },
    "routing": {"type": "hard", "temperature": 0.5},
    "joint_training": {"lambda_dehazing": 1.0, "lambda_classification": 0.2, "lambda_detection": 0.5},
    "device": "cpu",
}


def main():
    sys.path.insert(0, REF)
    sys.modules.setdefault("timm", types.ModuleType("timm"))
    lp = types.ModuleType("lpips")

    class LPIPS(torch.nn.Module):
        def __init__(self, net="alex"):
            super().__init__()

        def forward(self, x, t):
            return ((x - t) ** 2).mean(dim=(1, 2, 3), keepdim=True)

    lp.LPIPS = LPIPS
    sys.modules["lpips"] = lp
    import torchvision.models as tvm
    _vgg16 = tvm.vgg16
    tvm.vgg16 = lambda pretrained=False, **kw: _vgg16(weights=None)

    from models.dehazing.low_intensity import create_low_intensity_model
    from models.dehazing.medium_intensity import create_medium_intensity_model
    from models.dehazing.high_intensity import create_high_intensity_model
    from models.dehazing.base_model import AttentionBlock, ResidualBlock
    from models.classifier import create_classifier
    from models.routing import create_router
    from training.loss import ContentLoss, get_dehazing_loss, get_joint_loss

    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    makers = {"low": create_low_intensity_model, "medium": create_medium_intensity_model, "high": create_high_intensity_model}

    # --- branches: config 1 (Light, 1x3x256x256) and small cases for all three
    models = {}
    for name, mk in makers.items():
        seed_everything(SEED)
        m = mk(CONFIG).eval()
        models[name] = m
        gold = {"fingerprint": fingerprint(m.state_dict()), "keys": list(m.state_dict().keys()),
                "info": m.get_info(), "cases": []}
        shapes = [(2, 64, 64, 1), (1, 32, 96, 2)] + ([(1, 256, 256, 3)] if name == "low" else [])
        for (n, h, w, s) in shapes:
            x = rand_image(n, h, w, s)
            gold["cases"].append({"shape": (n, h, w), "seed": s, "out": m(x).clone()})
        torch.save(gold, os.path.join(OUT, f"branch_{name}.pt"))
        print(name, gold["fingerprint"][:16], len(gold["keys"]), "tensors")

    # --- non-default variants: the factories select them through model_type
    vcfg = {"low_unet": ("low", "enhanced", create_low_intensity_model), "corun": ("medium", "corun", create_medium_intensity_model),
            "dual_branch": ("high", "dual_branch", create_high_intensity_model)}
    for name, (lvl, mtype, mk) in vcfg.items():
        cfg = {**CONFIG, "dehazing": {**CONFIG["dehazing"], lvl: {**CONFIG["dehazing"][lvl], "model_type": mtype}}}
        seed_everything(SEED)
        m = mk(cfg).eval()
        gold = {"fingerprint": fingerprint(m.state_dict()), "keys": list(m.state_dict().keys()), "info": m.get_info(),
                "model_type": mtype, "level": lvl, "cases": []}
        for (n, h, w, s) in [(2, 64, 64, 11), (1, 32, 96, 12)]:
            x = rand_image(n, h, w, s)
            gold["cases"].append({"shape": (n, h, w), "seed": s, "out": m(x).clone()})
        torch.save(gold, os.path.join(OUT, f"branch_{name}.pt"))
        print(name, gold["fingerprint"][:16], len(gold["keys"]), "tensors")

    # --- building blocks (base_model.py)
    seed_everything(SEED)
    rb = ResidualBlock(32).eval()
    ab = AttentionBlock(96).eval()
    xr = torch.randn(1, 32, 16, 24, generator=torch.Generator().manual_seed(5))
    xa = torch.randn(1, 96, 16, 24, generator=torch.Generator().manual_seed(6)).relu()
    torch.save({"res_fp": fingerprint(rb.state_dict()), "res_out": rb(xr.clone()), "attn_fp": fingerprint(ab.state_dict()),
                "attn_out": ab(xa)}, os.path.join(OUT, "blocks.pt"))

    # --- classifier (resnet18, pretrained=False) — config 2's model at a CPU-sized input
    seed_everything(SEED)
    clf = create_classifier(CONFIG).eval()
    x = rand_image(4, 64, 64, 7)
    logits, feats = clf(x)
    torch.save({"fingerprint": fingerprint(clf.state_dict()), "keys": list(clf.state_dict().keys()), "seed": 7,
                "shape": (4, 64, 64), "logits": logits.clone(), "features": feats.clone(),
                "feature_dim": clf.feature_dim}, os.path.join(OUT, "classifier_resnet18.pt"))
    print("classifier", logits)

    # --- routers: hard (natural logits, crafted logits incl. ties, given intensity) and soft
    x = rand_image(6, 32, 32, 8)
    hard = create_router(models, clf, CONFIG).eval()
    out_nat, info_nat = hard(x)
    crafted = torch.tensor([[0.5, 0.5, 0.5], [0.1, 0.7, 0.7], [2.0, -1.0, 0.3], [-3.0, -2.0, -1.0],
                            [0.0, 1.0, 0.5], [9.0, 9.0, 8.0]])
    inten = torch.argmax(crafted, dim=1)
    out_cr, info_cr = hard(x, intensity=inten)
    soft_cfg = dict(CONFIG, routing={"type": "soft", "temperature": 0.5})
    soft = create_router(models, clf, soft_cfg).eval()
    out_soft, info_soft = soft(x, crafted)
    torch.save({"seed": 8, "shape": (6, 32, 32), "router_keys": list(hard.state_dict().keys()),
                "natural_out": out_nat.clone(), "natural_intensity": info_nat["intensity"].clone(),
                "crafted_logits": crafted, "crafted_intensity": inten, "crafted_out": out_cr.clone(),
                "crafted_masks": torch.stack([info_cr["low_mask"], info_cr["medium_mask"], info_cr["high_mask"]]),
                "soft_out": out_soft.clone(), "soft_weights": info_soft["weights"].clone()},
               os.path.join(OUT, "routing.pt"))

    # --- losses
    seed_everything(SEED)
    closs = ContentLoss().eval()
    vgg_sd = {k: v.clone() for k, v in closs.model.state_dict().items()}
    pred, tgt = rand_image(2, 64, 64, 9), rand_image(2, 64, 64, 10)
    seed_everything(SEED)
    dl = get_dehazing_loss(CONFIG).eval()   # same seed -> same VGG init as `closs`
    total, parts = dl(pred, tgt)
    seed_everything(SEED)
    jl = get_joint_loss(CONFIG).eval()
    labels = torch.tensor([2, 0])
    lg = torch.tensor([[0.2, -0.4, 1.0], [0.3, 0.1, -0.2]])
    jtotal, jparts = jl(pred, tgt, lg, labels)
    torch.save({"vgg_fingerprint": fingerprint(vgg_sd), "pred_seed": 9, "target_seed": 10, "shape": (2, 64, 64),
                "content": closs(pred, tgt).clone(), "dehazing_total": total.clone(),
                "dehazing_parts": {k: v.clone() for k, v in parts.items()},
                "joint_total": jtotal.clone(), "joint_ce": jparts["classification"].clone(), "joint_logits": lg,
                "joint_labels": labels}, os.path.join(OUT, "losses.pt"))
    print("losses", float(total), float(jtotal))


if __name__ == "__main__":
    main()
