"""Shared test helpers: golden fixtures, seeded model construction, the reference config."""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

SEED = 42

CONFIG = {
    "classifier": {"model": "resnet18", "num_classes": 3, "pretrained": False},
    "dehazing": {
        "low": {"model_type": "lightweight", "channels": 32, "blocks": 3},
        "medium": {"model_type": "standard", "channels": 64, "blocks": 6},
        "high": {"model_type": "complex", "channels": 96, "blocks": 9},
    },
    "routing": {"type": "hard", "temperature": 0.5},
    "joint_training": {"lambda_dehazing": 1.0, "lambda_classification": 0.2, "lambda_detection": 0.5},
    "device": "cuda",
}


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def fingerprint(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rand_image(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, h, w, generator=g)


def make_branch(name, seed=SEED):
    from adam_dehaze_b200.models.dehazing.low_intensity import create_low_intensity_model
    from adam_dehaze_b200.models.dehazing.medium_intensity import create_medium_intensity_model
    from adam_dehaze_b200.models.dehazing.high_intensity import create_high_intensity_model
    mk = {"low": create_low_intensity_model, "medium": create_medium_intensity_model, "high": create_high_intensity_model}
    # non-default variants are selected through the reference's own config key `model_type`
    variants = {"low_unet": ("low", "enhanced"), "corun": ("medium", "corun"), "dual_branch": ("high", "dual_branch")}
    cfg = CONFIG
    if name in variants:
        lvl, mtype = variants[name]
        cfg = {**CONFIG, "dehazing": {**CONFIG["dehazing"], lvl: {**CONFIG["dehazing"][lvl], "model_type": mtype}}}
        name = lvl
    torch.manual_seed(seed)
    return mk[name](cfg).eval()


def make_classifier(model="resnet18", seed=SEED):
    from adam_dehaze_b200.models.classifier import create_classifier
    cfg = dict(CONFIG, classifier=dict(CONFIG["classifier"], model=model))
    torch.manual_seed(seed)
    return create_classifier(cfg).eval()


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()


def randomize_bn(module, seed=7):
    """Give BatchNorm layers non-trivial running statistics/affine so folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return module


class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def oracle_bf16_storage(fn, *args):
    """Run an oracle forward with every conv / BatchNorm output (and the gradient flowing back through it) rounded to
    bf16 and conv weights rounded to bf16 — the fp32 oracle restricted to the storage format the north_star prescribes
    (bf16 tensors, fp32 accumulate).  Used to bound what ANY bf16-storage implementation can achieve against fp32."""
    import torch.nn.functional as F
    orig_conv, orig_convt, orig_bn = F.conv2d, F.conv_transpose2d, F.batch_norm

    def conv2d(inp, w, b=None, **kw):
        return _RoundBf16.apply(orig_conv(_RoundBf16.apply(inp), w.to(torch.bfloat16).float(), b, **kw))

    def convt(inp, w, b=None, **kw):
        return _RoundBf16.apply(orig_convt(_RoundBf16.apply(inp), w.to(torch.bfloat16).float(), b, **kw))

    def bn(inp, *a, **kw):
        return _RoundBf16.apply(orig_bn(inp, *a, **kw))
    F.conv2d, F.conv_transpose2d, F.batch_norm = conv2d, convt, bn
    try:
        return fn(*args)
    finally:
        F.conv2d, F.conv_transpose2d, F.batch_norm = orig_conv, orig_convt, orig_bn
