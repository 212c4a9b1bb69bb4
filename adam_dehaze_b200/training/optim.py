"""Optimizer step and gradient exchange of the training drivers (train_dehazing.py:33-37,91-92; train_joint.py:84-88,149-150).

`FlatAdam` keeps every parameter as a view of ONE fp32 buffer, so a step is: gather the .grad tensors into the flat
gradient buffer, one NCCL all-reduce over NVLink when torch.distributed is initialised (replicas + one
gradient all-reduce per step, mean), and ONE adb_adam_step_segments launch with torch.optim.Adam's arithmetic
(weight_decay = L2 added to the gradient).  There is no torch fallback: the step runs on the CUDA device only.

DistributedDataParallel / torch.optim.Adam behaviours kept:
  * replicas start equal: the flat parameter buffer is broadcast from rank 0 at construction;
  * every rank joins every all-reduce, whatever it computed: a parameter without a gradient on this rank contributes
    zeros, and a per-parameter "has a gradient" flag rides at the tail of the same bucket;
  * a parameter whose .grad is None on EVERY rank is skipped entirely (no moment decay, no weight decay, no step count),
    as torch.optim.Adam does — a branch that saw no sample under HardRouter joint training does not drift.
"""
import torch

from .. import _lib


_SETTER_FORM = [None]      # which call form of torch._C._autograd._unsafe_set_version_counter this torch accepts


def _bump_versions(params):
    """The step writes parameter memory from a libadb200 kernel; tell torch (and the weight-packing caches keyed on
    `_version`) that the tensors changed.  torch >= 2.5 takes (sequence of tensors, sequence of versions), older builds
    (tensor, version); the form is probed once — a failed pybind call formats its arguments into the exception text, which for
    a CUDA tensor is a device synchronisation plus a tensor print."""
    params = list(params)
    if not params:
        return
    setter = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
    if setter is not None and _SETTER_FORM[0] != "none":
        if _SETTER_FORM[0] in (None, "seq"):
            try:
                setter(tuple(params), tuple(p._version + 1 for p in params))
                _SETTER_FORM[0] = "seq"
                return
            except (TypeError, RuntimeError):
                if _SETTER_FORM[0] == "seq":
                    raise
        try:
            probe = torch.zeros(1)
            setter(probe, probe._version + 1)          # probe on a CPU scalar: cheap to format if it fails
            for p in params:
                setter(p, p._version + 1)
            _SETTER_FORM[0] = "one"
            return
        except (TypeError, RuntimeError):
            _SETTER_FORM[0] = "none"
    with torch.no_grad():
        for p in params:
            p.add_(0)          # in-place no-op: bumps the version for floating-point and integer tensors alike


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("FlatAdam: empty parameter list")
        dev = self.params[0].device
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        self.step_count = 0
        sizes = [p.numel() for p in self.params]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + (s + 3) // 4 * 4)     # 16-byte aligned views
        total = self.offsets[-1]
        nseg = len(self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # gradient bucket + one "has a gradient" flag per parameter at its tail: one collective carries both
        self.bucket = torch.zeros(total + (nseg + 3) // 4 * 4, dtype=torch.float32, device=dev)
        self.grad = self.bucket[:total]
        self.live = self.bucket[total:total + nseg]
        self.seg_off = torch.tensor(self.offsets[:-1], dtype=torch.int64, device=dev)
        self.seg_step = torch.zeros(nseg, dtype=torch.int32, device=dev)          # per-parameter step counts (torch.optim.Adam 'step')
        self._seg_bc = torch.zeros((2, nseg), dtype=torch.float32, device=dev)
        self._live_host = torch.zeros(nseg, dtype=torch.float32).pin_memory() if dev.type == "cuda" else torch.zeros(nseg)
        self.allreduce_events = None       # a list here collects (start, end) CUDA events around each all-reduce (bench.py)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad_views = []
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                if p.dtype != torch.float32:
                    raise TypeError("FlatAdam: fp32 master parameters expected")
                view = self.flat[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self.grad_views.append(self.grad[o:o + p.numel()].view(p.shape))
        if self.world_size() > 1:          # replicas start from rank 0's parameters, as DistributedDataParallel does
            import torch.distributed as dist
            dist.broadcast(self.flat, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)

    @property
    def param_groups(self):
        return [{"params": self.params, "lr": self.lr}]

    def state_dict(self):
        """torch.optim.Adam's layout ({'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}), so the
        'optimizer_state_dict' entry of a checkpoint (train_dehazing.py:196-203) is interchangeable with the reference's."""
        state = {}
        steps = self.seg_step.tolist() if self.step_count > 0 else []
        if self.step_count > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                if steps[i] == 0:          # never stepped: torch.optim.Adam holds no state for it
                    continue
                n = p.numel()
                state[i] = {"step": torch.tensor(float(steps[i])),
                            "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FlatAdam.load_state_dict: expected one parameter group over the same parameters")
        g = groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        steps = set()
        seg_steps = [0] * len(self.params)
        with torch.no_grad():
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                st = sd["state"].get(i, sd["state"].get(str(i)))
                if st is None:
                    continue
                seg_steps[i] = int(float(st["step"]))
                if tuple(st["exp_avg"].shape) != tuple(p.shape):
                    raise ValueError(f"FlatAdam.load_state_dict: state {i} has shape {tuple(st['exp_avg'].shape)}, parameter {tuple(p.shape)}")
                n = p.numel()
                self.exp_avg[o:o + n].view(p.shape).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + n].view(p.shape).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            self.seg_step.copy_(torch.tensor(seg_steps, dtype=torch.int32))
        self.step_count = max(steps) if steps else 0

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None

    def world_size(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce_gradients(self):
        """Gather the .grad tensors into the flat bucket and sum it over the ranks (one collective per step).
        Returns the world size (the Adam kernel folds the 1/world mean into its gradient read)."""
        with torch.no_grad():
            have = [(v, p.grad) for v, p in zip(self.grad_views, self.params) if p.grad is not None]
            none = [v for v, p in zip(self.grad_views, self.params) if p.grad is None]
            if none:       # parameters that saw no sample on this rank still take part in the collective (zeros)
                torch._foreach_zero_(none)
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            flags = [0.0 if p.grad is None else 1.0 for p in self.params]
            if flags != getattr(self, "_last_flags", None):       # (the pinned staging buffer is rewritten only when the set changes)
                torch.cuda.current_stream().synchronize() if self.flat.is_cuda and getattr(self, "_last_flags", None) is not None else None
                self._live_host.copy_(torch.tensor(flags))
                self._last_flags = flags
            self.live.copy_(self._live_host, non_blocking=True)
        world = self.world_size()
        if world > 1:
            import torch.distributed as dist
            ev = None
            if self.allreduce_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            dist.all_reduce(self.bucket, group=self.group)
            if ev is not None:
                ev[1].record()
                self.allreduce_events.append(ev)
        return world

    def step(self):
        if self.flat.device.type != "cuda":
            raise RuntimeError("FlatAdam.step: parameters must live on the CUDA device (no CPU path)")
        self.step_count += 1
        world = self.reduce_gradients()
        b1, b2 = self.betas
        _lib.call("adb_adam_step_segments", _lib.ptr(self.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                  self.flat.numel(), float(self.lr), float(b1), float(b2), float(self.eps), float(self.weight_decay),
                  1.0 / world, _lib.ptr(self.seg_off), len(self.params), _lib.ptr(self.live), _lib.ptr(self.seg_step),
                  _lib.ptr(self._seg_bc[0]), _lib.ptr(self._seg_bc[1]), _lib.current_stream())
        _bump_versions(self.params)
