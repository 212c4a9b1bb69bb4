import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from helpers import make_classifier
from adam_dehaze_b200 import _lib, ops
import adam_dehaze_b200.training.autograd as ag
clf = make_classifier("densenet121").cuda().train()
x = torch.rand(16, 3, 512, 512, device="cuda")
for _ in range(2):
    lg, _ = clf(x); lg.sum().backward()
torch.cuda.synchronize()
inner = _lib.call
calls = []
def timed(name, *a):
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(); r = inner(name, *a); eb.record()
    calls.append((name, ea, eb, a))
    return r
_lib.call = timed; ops._lib.call = timed; ag._lib.call = timed
lg, _ = clf(x); lg.sum().backward()
torch.cuda.synchronize()
by = {}
rows = []
for nm, ea, eb, a in calls:
    ms = ea.elapsed_time(eb)
    by[nm] = by.get(nm, 0) + ms
    if nm == "adb_bn_train_stats":
        rows.append((ms, a[1], a[2], a[3]))
print({k: round(v, 2) for k, v in sorted(by.items(), key=lambda kv: -kv[1])})
rows.sort(reverse=True)
print(rows[:12]); print(len(rows), sum(r[0] for r in rows))
