"""Diagnose train-step gradient error: ours vs fp32 oracle vs a bf16-rounding simulation of the oracle."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import torch
import torch.nn.functional as F
from helpers import make_branch, rand_image
import adam_oracle as oracle
from adam_dehaze_b200.training.loss import DehazingLoss


class Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).float()


def sim(name, sd, x):
    """oracle forward with activations/gradients rounded to bf16 at every layer boundary"""
    orig_conv, orig_bn = F.conv2d, F.batch_norm

    def conv2d(inp, w, b=None, **kw):
        return Round.apply(orig_conv(Round.apply(inp), w.to(torch.bfloat16).float(), b, **kw))

    def bn(inp, *a, **kw):
        return Round.apply(orig_bn(inp, *a, **kw))
    F.conv2d, F.batch_norm = conv2d, bn
    try:
        fwd = {"low": oracle.light_forward, "medium": oracle.medium_forward, "high": oracle.complex_forward}[name]
        with oracle.train_mode():
            return fwd(sd, x)
    finally:
        F.conv2d, F.batch_norm = orig_conv, orig_bn


name = sys.argv[1] if len(sys.argv) > 1 else "low"
n, h, w = 2, int(sys.argv[2]) if len(sys.argv) > 2 else 32, int(sys.argv[3]) if len(sys.argv) > 3 else 48
smooth = len(sys.argv) > 4
m = make_branch(name).cuda().train()
x = rand_image(n, h, w, 5).cuda()
tgt = rand_image(n, h, w, 6).cuda()
if smooth:
    tgt = F.avg_pool2d(x, 5, 1, 2) * 0.5 + 0.4
sd = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
names = [k for k, _ in m.named_parameters()]
fwd = {"low": oracle.light_forward, "medium": oracle.medium_forward, "high": oracle.complex_forward}[name]
with oracle.train_mode():
    ro = fwd(sd, x)
rg = dict(zip(names, torch.autograd.grad((ro - tgt).abs().mean(), [sd[k] for k in names], allow_unused=True)))
so = sim(name, sd, x)
sg = dict(zip(names, torch.autograd.grad((so - tgt).abs().mean(), [sd[k] for k in names], allow_unused=True)))
out = m(x)
loss, _ = DehazingLoss(1.0, 0.0, 0.0)(out, tgt)
loss.backward()
print("out err ours", (out - ro).abs().max().item(), "sim", (so - ro).abs().max().item())
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-20)).item()
for k, p in m.named_parameters():
    print(f"{k:45s} ours/ref {rel(p.grad, rg[k]):.4f}  sim/ref {rel(sg[k], rg[k]):.4f}  ours/sim {rel(p.grad, sg[k]):.4f}  |ref| {rg[k].norm().item():.3e}")
