import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from adam_dehaze_b200 import ops, _lib
torch.cuda.init()
x = torch.empty(1, device="cuda")
def bench(fn, n=300):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = (time.perf_counter() - t0) / n * 1e6
    torch.cuda.synchronize()
    return dt
print("torch.empty small      %.1f us" % bench(lambda: torch.empty(1024, device="cuda")))
print("torch.empty 512MB      %.1f us" % bench(lambda: torch.empty((16, 512, 512, 64), dtype=torch.bfloat16, device="cuda")))
keep = []
print("torch.empty 512MB kept %.1f us" % bench(lambda: keep.append(torch.empty((4, 512, 512, 64), dtype=torch.bfloat16, device="cuda")), 100))
keep.clear()
a = torch.randn(16, 128, 128, 128, device="cuda").to(torch.bfloat16)
b = torch.randn(16, 128, 128, 128, device="cuda").to(torch.bfloat16)
out = torch.empty(128, 128, 3, 3, device="cuda")
print("ops.wgrad host         %.1f us" % bench(lambda: ops.wgrad(a, b, out=out), 100))
from adam_dehaze_b200.ops import ConvSpec
w = torch.randn(128, 128, 3, 3, device="cuda")
spec = ConvSpec.from_conv(w, pad=1)
dst = torch.empty_like(a)
print("ops.conv2d host        %.1f us" % bench(lambda: ops.conv2d(spec, a, dst=dst), 100))
print("current_stream         %.1f us" % bench(lambda: _lib.current_stream(), 1000))
st = _lib.current_stream()
print("adb_add_bf16 call      %.1f us" % bench(lambda: _lib.call("adb_add_bf16", _lib.ptr(a), 128, _lib.ptr(b), 128, 1024, 128, st), 300))
