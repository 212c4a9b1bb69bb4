"""Build libadb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m adam_dehaze_b200.build [--force]

The library is written next to this file so that it travels with a repo snapshot to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadb200.so")
STAMP = os.path.join(HERE, ".libadb200.stamp")
SOURCES = ["adb_host.cu", "conv_igemm.cu", "conv_roll.cu", "conv_wgrad.cu", "pointwise.cu", "train.cu", "metrics.cu", "route.cu", "guard_fp32.cu", "input_pipe.cu"]
HEADERS = ["adb_ptx.cuh", "adb_host.h", "conv_common.cuh", os.path.join("..", "..", "include", "adb200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-shared", "-cudart", "static",
]


def _digest():
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _compile_one(args):
    src, obj, flags, verbose = args
    cmd = [nvcc_path()] + flags + ["-c", src, "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    """Compile every translation unit (in parallel, objects cached per source digest under csrc/.obj) and link."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(CSRC, ".obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    hdr = hashlib.sha256()
    for f in HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            hdr.update(fh.read())
    hdr.update(" ".join(NVCC_FLAGS).encode())
    jobs, objs = [], []
    for s in SOURCES:
        h = hdr.copy()
        with open(os.path.join(CSRC, s), "rb") as fh:
            h.update(fh.read())
        obj = os.path.join(objdir, s.replace(".cu", "") + "." + h.hexdigest()[:16] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            for old in os.listdir(objdir):
                if old.startswith(s.replace(".cu", "") + "."):
                    os.remove(os.path.join(objdir, old))
            jobs.append((os.path.join(CSRC, s), obj, cflags, verbose))
    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for src, rc, out in ex.map(_compile_one, jobs):
            if verbose:
                print(out)
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError(f"nvcc failed compiling {src}")
    cmd = [nvcc_path()] + NVCC_FLAGS + objs + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libadb200.so")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
