"""adam_dehaze_b200 — a B200-native (sm_100a) implementation of the ADAM-Dehaze hot path.

Layout (only what the path needs):
  csrc/        hand-written CUDA: tcgen05/TMA implicit-GEMM conv, HBM-bound fused kernels, routing; C-ABI in
               include/adb200.h, built in-tree as libadb200.so (python -m adam_dehaze_b200.build)
  _lib.py      ctypes binding (fails loudly: no CPU / torch fallback)
  ops.py       weight packing + one Python call per fused kernel
  engine.py    per-module execution plans (Light / Medium / Complex branches, ResNet and DenseNet HDEN)
  models/      drop-in mirrors of the reference's models/classifier.py, models/routing.py, models/dehazing/*
  training/    drop-in mirror of training/loss.py
"""
__version__ = "0.1.0"
