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


class LowIntensityDehazeModel(BaseDehazeModel):
    """Non-default Light variant (reference low_intensity.py:56-125): 3x3 stem, one stride-2 encoder stage, n_blocks
    residual blocks at half resolution, one transposed-conv decoder with a concat skip, a three-conv sigmoid head;
    output clamp(x + (out - 0.5) * 2, 0, 1)."""
    _engine_kind = "low_unet"

    def __init__(self, in_channels=3, base_channels=32, n_blocks=3):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        c, c2 = base_channels, base_channels * 2
        self.init_conv = ConvBlock(in_channels, c, kernel_size=3, padding=1)
        self.down1 = nn.Sequential(ConvBlock(c, c2, kernel_size=4, stride=2, padding=1), ResidualBlock(c2))
        self.bottleneck = nn.Sequential(*[ResidualBlock(c2) for _ in range(n_blocks - 1)])
        self.up1 = nn.Sequential(nn.ConvTranspose2d(c2, c, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(c),
                                 nn.ReLU(inplace=True))
        self.output_conv = nn.Sequential(
            ConvBlock(c2, c, kernel_size=3, padding=1),
            ConvBlock(c, c, kernel_size=3, padding=1),
            nn.Conv2d(c, in_channels, kernel_size=3, padding=1),
            nn.Sigmoid(),
        )

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="LowIntensityDehazeModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info


def create_low_intensity_model(config):
    """Factory with the reference's config keys (low_intensity.py:127-140)."""
    cfg = config["dehazing"]["low"]
    if cfg["model_type"] != "lightweight":      # the reference sends every other value here (low_intensity.py:135-140)
        return LowIntensityDehazeModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
    return LightweightDehazeModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
