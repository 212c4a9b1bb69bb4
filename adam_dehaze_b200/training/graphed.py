"""One whole optimisation step as a CUDA graph.

A training step of the joint model (train_joint.py:129-150) is ~1.7 k libadb200 launches issued from Python; at ~85 us of host
work per launch the step is launch-bound on the host (147 ms of queueing for ~125 ms of kernels).  `GraphedStep` runs the
step function a few times eagerly (every lazily built packing, index map, workspace and kernel attribute then exists), captures
one more execution — forward, loss, backward through torch.autograd, the flat gradient bucket's NCCL all-reduce and the fused
Adam launch — into a `torch.cuda.CUDAGraph`, and replays it: one host call per step.

What makes the step capturable: nothing on the path reads a device value on the host (bucket counts, losses and per-parameter
Adam step counts live on the device; `LossMeter` reads the loss one step late from pinned memory), every buffer is a torch
allocation (taken from the graph's private pool during capture) and every kernel is launched on the current stream.
Limits: tensor shapes are frozen (feed batches of the captured shape through `copy_inputs`; run a ragged last batch through
the eager step), and host-side scalars baked into launches — the learning rate — need a re-capture when they change.
"""
import torch


class GraphedStep:
    def __init__(self, step_fn, static_inputs=(), warmup=3):
        """step_fn(): one step over `static_inputs` (tensors it closes over); returns a tensor or tuple of tensors (e.g. the loss)."""
        self.step_fn, self.static_inputs = step_fn, tuple(static_inputs)
        # Warm up and capture on ONE side stream: autograd pins every AccumulateGrad node to the stream it was created on, and a
        # node of another stream (e.g. from eager steps on the default stream whose loss tensor is still referenced somewhere —
        # drop such references first) would make the capture wait on that stream, which invalidates it.
        import gc
        gc.collect()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                out = step_fn()
            del out
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=side):
            self.outputs = step_fn()

    def copy_inputs(self, *tensors):
        for dst, src in zip(self.static_inputs, tensors):
            dst.copy_(src, non_blocking=True)

    def __call__(self, *tensors):
        if tensors:
            self.copy_inputs(*tensors)
        self.graph.replay()
        return self.outputs
