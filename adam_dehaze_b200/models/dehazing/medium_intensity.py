"""Drop-in for models/dehazing/medium_intensity.py — the Medium branch (reference medium_intensity.py:5-126, 201-215)."""
import torch.nn as nn

from .base_model import BaseDehazeModel, ConvBlock, ResidualBlock


def _up(cin, cout):
    """ConvTranspose2d(4, 2, 1) -> BN -> ReLU -> ResidualBlock, the decoder stage of the reference (medium:50-58)."""
    return [nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True), ResidualBlock(cout)]


class MediumIntensityDehazeModel(BaseDehazeModel):
    """U-Net C/2C/4C: 7x7 stem, two stride-2 encoders with two residual blocks each, a two-block bottleneck, two
    transposed-conv decoders with concat skips, a three-conv tanh head; output clamp(x + residual, 0, 1)."""
    _engine_kind = "unet"

    def __init__(self, in_channels=3, base_channels=64, n_blocks=6):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        c1, c2, c4 = base_channels, base_channels * 2, base_channels * 4
        self.init_conv = ConvBlock(in_channels, c1, kernel_size=7, padding=3)
        self.encoder = nn.ModuleList([
            nn.Sequential(ConvBlock(c1, c2, kernel_size=4, stride=2, padding=1), ResidualBlock(c2), ResidualBlock(c2)),
            nn.Sequential(ConvBlock(c2, c4, kernel_size=4, stride=2, padding=1), ResidualBlock(c4), ResidualBlock(c4)),
        ])
        self.bottleneck = nn.Sequential(ResidualBlock(c4), ResidualBlock(c4))
        self.decoder = nn.ModuleList([nn.Sequential(*_up(c4, c2)), nn.Sequential(*_up(c2 * 2, c1))])
        self.output_conv = nn.Sequential(
            ConvBlock(c1 * 2, c1, kernel_size=3, padding=1),
            ConvBlock(c1, c1 // 2, kernel_size=3, padding=1),
            nn.Conv2d(c1 // 2, in_channels, kernel_size=3, padding=1),
            nn.Tanh(),
        )

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="MediumIntensityDehazeModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info


class COrunInspiredModel(BaseDehazeModel):
    """Non-default Medium variant (reference medium_intensity.py:128-199): 7x7 stem, three scales (full, max-pooled /2 and
    /4, bilinearly upsampled back with align_corners=True) fused by a 1x1 conv, n_blocks residual blocks at full resolution,
    a two-conv tanh head; output clamp(x + residual, 0, 1)."""
    _engine_kind = "corun"

    def __init__(self, in_channels=3, base_channels=64, n_blocks=6):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        c = base_channels
        self.init_conv = ConvBlock(in_channels, c, kernel_size=7, padding=3)
        self.scale1_conv = ConvBlock(c, c, kernel_size=3, padding=1)
        self.scale2_conv = nn.Sequential(nn.MaxPool2d(kernel_size=2, stride=2), ConvBlock(c, c * 2, kernel_size=3, padding=1),
                                         nn.UpsamplingBilinear2d(scale_factor=2))
        self.scale3_conv = nn.Sequential(nn.MaxPool2d(kernel_size=4, stride=4), ConvBlock(c, c * 4, kernel_size=3, padding=1),
                                         nn.UpsamplingBilinear2d(scale_factor=4))
        self.fusion_conv = ConvBlock(c + c * 2 + c * 4, c * 2, kernel_size=1, padding=0)
        self.residual_blocks = nn.Sequential(*[ResidualBlock(c * 2) for _ in range(n_blocks)])
        self.output_conv = nn.Sequential(ConvBlock(c * 2, c, kernel_size=3, padding=1),
                                         nn.Conv2d(c, in_channels, kernel_size=3, padding=1), nn.Tanh())

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="COrunInspiredModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info


def create_medium_intensity_model(config):
    """Factory with the reference's config keys (medium_intensity.py:201-215)."""
    cfg = config["dehazing"]["medium"]
    if cfg["model_type"] == "corun":
        return COrunInspiredModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
    return MediumIntensityDehazeModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
