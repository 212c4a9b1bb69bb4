"""Wait-cycle statistics of one CTA of the rolling-row conv kernel (csrc/conv_roll.cu, tune flag 4) — who stalls on whom.

    python tools/roll_stats.py [shape names from tools/prof_conv.py ...]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from adam_dehaze_b200 import _lib  # noqa: E402
import prof_conv  # noqa: E402


def main():
    names = sys.argv[1:] or ["light_32_3x3", "med_64_3x3", "dense_3x3_128_32"]
    for shape in prof_conv.SHAPES:
        if shape[0] not in names:
            continue
        prof_conv.run(shape, None, 2)
        for cta in (0, 77):
            ms, tf = prof_conv.run(shape, {"flags": 4, "acc": cta}, 1)
            torch.cuda.synchronize()
            buf = (C.c_int64 * (6 * 256))()
            _lib.call("adb_debug_timeline", buf, 6 * 256)
            b = list(buf)
            rows = max(1, b[19])
            print(f"== {shape[0]} CTA {cta}: {ms:.3f} ms {tf:.0f} TF/s; input rows {b[19]}, A boxes {b[1]}, output rows (warp 4) {b[36]}")
            print(f"   A producer : total {b[2]} cyc, waiting for a free slot {b[0]} ({b[0] / max(1, b[2]):.0%})")
            print(f"   MMA issuer : total {b[20]} cyc = {b[20] / rows:.0f}/row; ring wait {b[16]} ({b[16] / max(1, b[20]):.0%}), "
                  f"operand wait {b[17]} ({b[17] / max(1, b[20]):.0%}), issue {b[18]} ({b[18] / max(1, b[20]):.0%})")
            er = max(1, b[36])
            print(f"   epilogue w4: total {b[37]} cyc = {b[37] / er:.0f}/row; accumulator wait {b[32]} ({b[32] / max(1, b[37]):.0%}), "
                  f"staging wait {b[33] / er:.0f}/row, slab+store {b[34] / er:.0f}/row, zero+return {b[35] / er:.0f}/row")


main()
