"""Dump the in-kernel clock64 timeline of CTA 0 for one conv shape (developer aid)."""
import ctypes as C
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from adam_dehaze_b200 import ops, _lib
import prof_conv

def detail(names):
    """Epilogue sub-step stamps (tune flag 64): tag:delta sequence of warp 4 / lane 0 of CTA 0.
    tags: 1 before tfull wait, 2 after, 3 slab staging free, 4 slab computed (TMEM+residual -> smem), 5 fenced,
    6 TMA store issued, 7 DOT tmem ready, 8 DOT stored, 9 tile done; residual path: 13 residual bounced through the staging
    buffer, 14 first half of the next item's residual loads issued, 15 accumulator in registers (tcgen05.wait::ld); 16 TMA store
    instruction issued (6 then follows its commit_group)."""
    for shape in prof_conv.SHAPES:
        if shape[0] not in names:
            continue
        prof_conv.run(shape, None, 2)
        ms, tf = prof_conv.run(shape, {"flags": 64}, 1)
        torch.cuda.synchronize()
        buf = (C.c_int64 * (6 * 256))()
        _lib.call("adb_debug_timeline", buf, 6 * 256)
        ev = [(int(v) >> 56, int(v) & ((1 << 56) - 1)) for v in buf if v != 0]
        print(f"== {shape[0]}  {ms:.3f} ms {tf:.0f} TF/s (epilogue detail)")
        out, prev = [], None
        for tag, clk in ev[:400]:
            out.append(f"{tag}:{0 if prev is None else clk - prev}")
            prev = clk
        print(" ".join(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--detail":
        return detail(sys.argv[2:])
    names = sys.argv[1:] or ["med_64_3x3", "med_256_3x3"]
    for shape in prof_conv.SHAPES:
        if shape[0] not in names:
            continue
        prof_conv.run(shape, None, 2)
        ms, tf = prof_conv.run(shape, {"flags": 4}, 1)
        torch.cuda.synchronize()
        buf = (C.c_int64 * (6 * 256))()
        _lib.call("adb_debug_timeline", buf, 6 * 256)
        t = [list(buf[r * 256:(r + 1) * 256]) for r in range(6)]
        t0 = min(v for r in t for v in r if v > 0)
        print(f"== {shape[0]}  {ms:.3f} ms {tf:.0f} TF/s")
        for r, nm in enumerate(["A-prod wait done", "B-prod wait done", "MMA ready", "MMA issued", "epi start", "epi end"]):
            vals = [v - t0 for v in t[r] if v > 0][:40]
            print(f"{nm:18s}", " ".join(str(v) for v in vals))

main()
