"""Drop-in for training/loss.py (reference loss.py:110-241): DehazingLoss, JointLoss and their factories.

Built on the B200 path so far: the L1 reconstruction term and the cross-entropy term, forward AND backward, as fused
warp-shuffle reductions (adb_l1_mse_fwd / adb_l1_bwd / adb_ce_fwd_bwd) wrapped in torch.autograd.Function so the
reference's `loss.backward()` call keeps working.  The VGG16 content term (loss.py:47-84) and the LPIPS term
(loss.py:86-108) need the conv dgrad kernels and are not built yet: with their lambdas non-zero the forward raises
NotImplementedError instead of silently dropping a term.  Pass lambda_content=0, lambda_perceptual=0 to train on
L1 (+CE) only.
"""
import torch
import torch.nn as nn

from .. import ops
from .. import _lib


class _L1Mean(torch.autograd.Function):
    """mean |pred - target| (nn.L1Loss, loss.py:121) — one pass forward, one pass backward."""

    @staticmethod
    def forward(ctx, pred, target):
        pred, target = pred.contiguous().float(), target.contiguous().float()
        ctx.save_for_backward(pred, target)
        return ops.l1_mse(pred, target)[0]

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        grad = torch.empty_like(pred)
        # dL/dpred = g * sign(pred - target) / numel; g is a device scalar, folded in afterwards without a host sync
        _lib.call("adb_l1_bwd", _lib.ptr(pred), _lib.ptr(target), pred.numel(), 1.0, _lib.ptr(grad), _lib.current_stream())
        return grad * g, None


class _CrossEntropy(torch.autograd.Function):
    """nn.CrossEntropyLoss (mean reduction, loss.py:177) with the gradient produced in the same launch."""

    @staticmethod
    def forward(ctx, logits, labels):
        loss, grad = ops.cross_entropy(logits.contiguous().float(), labels.contiguous(), 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected CUDA tensors — this package runs on B200 (sm_100a) only and has no CPU path")


class DehazingLoss(nn.Module):
    def __init__(self, lambda_l1=1.0, lambda_content=0.1, lambda_perceptual=0.1):
        super().__init__()
        self.lambda_l1, self.lambda_content, self.lambda_perceptual = lambda_l1, lambda_content, lambda_perceptual

    def forward(self, pred, target):
        """Returns (total, {'l1','content','perceptual','total'}) like loss.py:125-162."""
        _need_cuda(pred, "DehazingLoss")
        if self.lambda_content != 0 or self.lambda_perceptual != 0:
            raise NotImplementedError(
                "DehazingLoss: the VGG16 content term and the LPIPS term are not built on the B200 path yet "
                "(they need the conv backward kernels); construct with lambda_content=0, lambda_perceptual=0")
        l1 = _L1Mean.apply(pred, target)
        zero = torch.zeros((), device=pred.device)
        total = self.lambda_l1 * l1
        return total, {"l1": l1, "content": zero, "perceptual": zero, "total": total}


class JointLoss(nn.Module):
    def __init__(self, lambda_dehazing=1.0, lambda_classification=0.2, lambda_detection=0.5, config=None,
                 dehazing_loss=None):
        super().__init__()
        self.lambda_dehazing = lambda_dehazing
        self.lambda_classification = lambda_classification
        self.lambda_detection = lambda_detection
        self.dehazing_loss = dehazing_loss if dehazing_loss is not None else DehazingLoss()

    def forward(self, pred, target_clear, pred_intensity=None, target_intensity=None, detection_loss=None):
        """Returns (total, {'dehazing','classification','detection','total','dehazing_components'}), loss.py:179-224."""
        dehazing, parts = self.dehazing_loss(pred, target_clear)
        if pred_intensity is not None and target_intensity is not None:
            ce = _CrossEntropy.apply(pred_intensity, target_intensity)
        else:
            ce = torch.tensor(0.0, device=pred.device)
        det = detection_loss if detection_loss is not None else torch.tensor(0.0, device=pred.device)
        total = self.lambda_dehazing * dehazing + self.lambda_classification * ce + self.lambda_detection * det
        return total, {"dehazing": dehazing, "classification": ce, "detection": det, "total": total,
                       "dehazing_components": parts}


def get_dehazing_loss(config):
    """Factory (loss.py:226-232): fixed lambdas 1.0 / 0.1 / 0.1 as in the reference."""
    return DehazingLoss(lambda_l1=1.0, lambda_content=0.1, lambda_perceptual=0.1)


def get_joint_loss(config):
    """Factory (loss.py:234-241)."""
    jt = config["joint_training"]
    return JointLoss(lambda_dehazing=jt["lambda_dehazing"], lambda_classification=jt["lambda_classification"],
                     lambda_detection=jt["lambda_detection"], config=config)
