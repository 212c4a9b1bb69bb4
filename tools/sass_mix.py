"""Instruction-mix table of the tensor-core kernels in libadb200.so (`cuobjdump -sass`): proves the tcgen05 / TMA / TMEM path
(UTCHMMA, UTMALDG, UTMASTG, LDTM, STTM) and the packed-pair epilogue arithmetic (FFMA2, FADD2, F2FP.RELU), and that no legacy
HMMA (mma.sync) is present.      python tools/sass_mix.py [path/to/libadb200.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "FFMA2", "FADD2", "F2FP.RELU", "HMMA"]


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "adam_dehaze_b200", "libadb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name).split("(")[0]
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        for op in OPS:
            if re.search(r"\b" + re.escape(op) + r"\b", line) or (op == "F2FP.RELU" and "F2FP.RELU" in line):
                counts[name][op] += 1
    print("| kernel | " + " | ".join(OPS) + " |")
    print("|---|" + "---:|" * len(OPS))
    for k, c in counts.items():
        if c["UTCHMMA"] or c["HMMA"]:
            print(f"| `{k}` | " + " | ".join(str(c[o]) for o in OPS) + " |")
    print(f"\nkernels in the library: {len(counts)}; with HMMA (legacy mma.sync): {sum(1 for c in counts.values() if c['HMMA'])}")


if __name__ == "__main__":
    main()
