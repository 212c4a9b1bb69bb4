"""Drop-in for the reference's models/dehazing/base_model.py building blocks.

Same classes, constructor signatures, sub-module names (hence state_dict keys and random-init order) as
/root/reference/models/dehazing/base_model.py:4-96; the arithmetic runs in libadb200 kernels (see engine.py).
The modules are parameter containers: nn.Conv2d / nn.BatchNorm2d objects are never called.
"""
import torch
import torch.nn as nn

from ... import engine as _engine


class ConvBlock(nn.Module):
    """Conv2d (bias only without BN) -> BatchNorm2d -> activation (reference base_model.py:4-24)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, use_bn=True, activation=nn.ReLU()):
        super().__init__()
        seq = [nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=not use_bn)]
        if use_bn:
            seq.append(nn.BatchNorm2d(out_channels))
        if activation is not None:
            if not isinstance(activation, nn.ReLU):
                raise ValueError("ConvBlock on the B200 path fuses ReLU or no activation only")
            seq.append(activation)
        self.block = nn.Sequential(*seq)

    def forward(self, x):
        _engine.require_cuda_any(x, "ConvBlock")
        return _engine.block_engine(self).conv_block(x)


class ResidualBlock(nn.Module):
    """Two ConvBlocks with an identity skip and a trailing ReLU (reference base_model.py:26-41)."""

    def __init__(self, channels, kernel_size=3):
        super().__init__()
        self.conv1 = ConvBlock(channels, channels, kernel_size, padding=kernel_size // 2)
        self.conv2 = ConvBlock(channels, channels, kernel_size, padding=kernel_size // 2, activation=None)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        _engine.require_cuda_any(x, "ResidualBlock")
        return _engine.block_engine(self).residual_block(x)


class AttentionBlock(nn.Module):
    """Channel gate (avg+max pooled MLP) then 7x7 spatial gate (reference base_model.py:43-78)."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.max_pool = nn.AdaptiveMaxPool2d(1)
        self.fc = nn.Sequential(
            nn.Conv2d(channels, channels // reduction, 1, bias=False),
            nn.ReLU(inplace=True),
            nn.Conv2d(channels // reduction, channels, 1, bias=False),
        )
        self.sigmoid = nn.Sigmoid()
        self.conv_spatial = nn.Conv2d(2, 1, kernel_size=7, padding=3, bias=False)

    def forward(self, x):
        _engine.require_cuda_any(x, "AttentionBlock")
        return _engine.block_engine(self).attention_block(x)


class BaseDehazeModel(nn.Module):
    """Common base of the branch models (reference base_model.py:80-96)."""

    def __init__(self):
        super().__init__()

    def forward(self, x):
        raise NotImplementedError

    def get_info(self):
        total = sum(p.numel() for p in self.parameters())
        return {
            "model_type": self.__class__.__name__,
            "params": total,
            "trainable_params": sum(p.numel() for p in self.parameters() if p.requires_grad),
        }

    # -- shared by the three default branches
    _engine_kind = None

    def _branch_engine(self):
        eng = self.__dict__.get("_adb_engine")
        if eng is None:
            eng = _engine.BranchEngine(self, self._engine_kind)
            self.__dict__["_adb_engine"] = eng
        return eng

    def forward_bucket(self, x, out, index, n_dev, count=None):
        """Routed forward: dehaze rows index[0:*n_dev] of `x` into the same rows of `out` (device-side count)."""
        return self._branch_engine().forward(x, out=out, index=index, n_dev=n_dev, count=count)
