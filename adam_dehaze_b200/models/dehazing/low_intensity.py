"""Drop-in for models/dehazing/low_intensity.py — the Light branch (reference low_intensity.py:5-54, 127-140)."""
import torch
import torch.nn as nn

from .base_model import BaseDehazeModel, ConvBlock, ResidualBlock


class LightweightDehazeModel(BaseDehazeModel):
    """3 -> C stem, n residual blocks, C -> C -> 3 head with sigmoid, alpha-blended with the input (no clamp)."""
    _engine_kind = "light"

    def __init__(self, in_channels=3, base_channels=32, n_blocks=3):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        self.init_conv = ConvBlock(in_channels, base_channels, kernel_size=3, padding=1)
        self.residual_blocks = nn.Sequential(*[ResidualBlock(base_channels) for _ in range(n_blocks)])
        self.output_conv = nn.Sequential(
            ConvBlock(base_channels, base_channels, kernel_size=3, padding=1),
            nn.Conv2d(base_channels, in_channels, kernel_size=3, padding=1),
            nn.Sigmoid(),
        )
        self.skip_alpha = nn.Parameter(torch.tensor(0.1))

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="LightweightDehazeModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info


def create_low_intensity_model(config):
    """Factory with the reference's config keys (low_intensity.py:127-140)."""
    cfg = config["dehazing"]["low"]
    if cfg["model_type"] != "lightweight":
        raise NotImplementedError(
            "LowIntensityDehazeModel (model_type != 'lightweight', low_intensity.py:56-125) is a non-default variant "
            "not built on the B200 path yet (SURVEY.md §8f rank 4)")
    return LightweightDehazeModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
