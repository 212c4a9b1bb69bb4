"""Time adb_bn_train_stats / adb_affine_act / adb_bn_bwd / adb_bn_relu_bwd on the shapes of the training step (CUDA events)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from adam_dehaze_b200 import _lib  # noqa: E402

P = _lib.ptr
for (px, c, pitch) in [(4194304, 64, 64), (4194304, 96, 96), (1048576, 192, 192), (262144, 384, 384), (262144, 64, 256), (262144, 224, 256),
                       (65536, 480, 512), (16384, 512, 1024), (16384, 992, 1024), (4096, 992, 1024), (4096, 128, 128)]:
    z = torch.randn(px, pitch, device="cuda").to(torch.bfloat16)
    g, b = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    scratch = torch.empty(int(_lib.load().adb_bn_scratch_floats(px, c)), device="cuda")
    st4 = [torch.empty(c, device="cuda") for _ in range(4)]
    y = torch.empty(px, c, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(px, c, device="cuda").to(torch.bfloat16)
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")

    def stats():
        _lib.call("adb_bn_train_stats", P(z), px, c, pitch, P(g), P(b), 1e-5, 0.1, None, None, None, P(scratch), P(st4[0]), P(st4[1]), P(st4[2]),
                  P(st4[3]), _lib.current_stream())

    def act():
        _lib.call("adb_affine_act", P(z), pitch, px, c, P(st4[2]), P(st4[3]), None, 0, 1, P(y), c, _lib.current_stream())

    def bwd():
        _lib.call("adb_bn_bwd", P(dy), c, P(y), c, P(z), pitch, px, c, 1, P(g), P(st4[0]), P(st4[1]), P(scratch), P(dy), c, P(dy), c, P(dg), P(db), 0,
                  _lib.current_stream())

    def bwd_relu():
        _lib.call("adb_bn_relu_bwd", P(dy), c, P(z), pitch, px, c, P(st4[2]), P(st4[3]), P(g), P(st4[0]), P(st4[1]), P(scratch), P(dy), c, 0,
                  P(dg), P(db), 0, _lib.current_stream())
    out = []
    for fn, passes in ((stats, 1), (act, 2), (bwd, 7), (bwd_relu, 5)):
        fn()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 5
        out.append(f"{fn.__name__} {ms:7.3f} ms {passes * px * c * 2 / ms / 1e9:6.2f} TB/s")
    print(f"px={px:8d} c={c:4d} pitch={pitch:4d}  " + "   ".join(out), flush=True)
