import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, "tests")
import torch
from helpers import make_branch, make_classifier
from adam_dehaze_b200 import ops
cnt = {"n": 0}
orig = ops.fold_bn
def counting(*a, **k):
    cnt["n"] += 1
    return orig(*a, **k)
ops.fold_bn = counting
torch.set_grad_enabled(False)
m = make_branch("medium").cuda()
x = torch.rand(2, 3, 64, 64, device="cuda")
for i in range(3):
    eng = m._branch_engine()
    sig0 = eng._ver._sig
    m(x)
    sig1 = eng._ver._sig
    changed = None
    if sig0 is not None and sig0 != sig1:
        changed = [(a, b) for a, b in zip(sig0, sig1) if a != b][:3]
    print(i, "fold_bn calls so far", cnt["n"], "changed:", changed)
