"""Drop-in for models/dehazing/high_intensity.py — the Complex branch (reference high_intensity.py:6-147, 225-239)."""
import torch.nn as nn

from .base_model import AttentionBlock, BaseDehazeModel, ConvBlock, ResidualBlock


def _down(cin, cout):
    return nn.Sequential(ConvBlock(cin, cout, kernel_size=4, stride=2, padding=1), ResidualBlock(cout),
                         ResidualBlock(cout), AttentionBlock(cout))


def _up(cin, cout):
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1), nn.BatchNorm2d(cout),
                         nn.ReLU(inplace=True), ResidualBlock(cout), AttentionBlock(cout))


class HighIntensityDehazeModel(BaseDehazeModel):
    """The Medium U-Net at C/2C/4C = 96/192/384 with an AttentionBlock after every stage, plus a 3->16->16->1 sigmoid
    guidance branch that scales the tanh residual: clamp(x + residual * guidance, 0, 1)."""
    _engine_kind = "unet_attn"

    def __init__(self, in_channels=3, base_channels=96, n_blocks=9):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        c1, c2, c4 = base_channels, base_channels * 2, base_channels * 4
        self.init_conv = ConvBlock(in_channels, c1, kernel_size=7, padding=3)
        self.encoder = nn.ModuleList([_down(c1, c2), _down(c2, c4)])
        self.bottleneck = nn.Sequential(ResidualBlock(c4), AttentionBlock(c4), ResidualBlock(c4), AttentionBlock(c4))
        self.decoder = nn.ModuleList([_up(c4, c2), _up(c2 * 2, c1)])
        self.output_conv = nn.Sequential(
            ConvBlock(c1 * 2, c1, kernel_size=3, padding=1),
            ConvBlock(c1, c1 // 2, kernel_size=3, padding=1),
            nn.Conv2d(c1 // 2, in_channels, kernel_size=3, padding=1),
            nn.Tanh(),
        )
        self.detail_branch = nn.Sequential(
            ConvBlock(in_channels, 16, kernel_size=3, padding=1),
            ConvBlock(16, 16, kernel_size=3, padding=1),
            nn.Conv2d(16, 1, kernel_size=1, padding=0),
            nn.Sigmoid(),
        )

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="HighIntensityDehazeModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info

class DualBranchAttentionModel(BaseDehazeModel):
    """Non-default Complex variant (reference high_intensity.py:149-223): a global branch (7x7 stem, two max-pooled
    residual+attention stages, bilinear upsampling back), a full-resolution local branch, a sigmoid "transmission" head and
    a tanh fusion head over their concat; output clamp(x + (1 - transmission) * residual, 0, 1)."""
    _engine_kind = "dual"

    def __init__(self, in_channels=3, base_channels=96, n_blocks=9):
        super().__init__()
        self.in_channels, self.base_channels, self.n_blocks = in_channels, base_channels, n_blocks
        c, ch, cq = base_channels, base_channels // 2, base_channels // 4
        self.global_branch = nn.Sequential(
            ConvBlock(in_channels, c, kernel_size=7, padding=3),
            nn.MaxPool2d(kernel_size=2, stride=2), ResidualBlock(c), AttentionBlock(c),
            nn.MaxPool2d(kernel_size=2, stride=2), ResidualBlock(c), AttentionBlock(c),
            ResidualBlock(c), nn.UpsamplingBilinear2d(scale_factor=2),
            ResidualBlock(c), nn.UpsamplingBilinear2d(scale_factor=2),
            ConvBlock(c, ch, kernel_size=3, padding=1),
        )
        self.local_branch = nn.Sequential(ConvBlock(in_channels, ch, kernel_size=3, padding=1), ResidualBlock(ch),
                                          ResidualBlock(ch), ConvBlock(ch, ch, kernel_size=3, padding=1))
        self.transmission_branch = nn.Sequential(ConvBlock(c, ch, kernel_size=3, padding=1), ConvBlock(ch, cq, kernel_size=3, padding=1),
                                                 nn.Conv2d(cq, 1, kernel_size=1, padding=0), nn.Sigmoid())
        self.fusion_conv = nn.Sequential(ConvBlock(c, ch, kernel_size=3, padding=1),
                                         nn.Conv2d(ch, in_channels, kernel_size=3, padding=1), nn.Tanh())

    def forward(self, x):
        return self._branch_engine().forward(x)

    def get_info(self):
        info = super().get_info()
        info.update(model_type="DualBranchAttentionModel", base_channels=self.base_channels, n_blocks=self.n_blocks)
        return info


def create_high_intensity_model(config):
    """Factory with the reference's config keys (high_intensity.py:225-239)."""
    cfg = config["dehazing"]["high"]
    if cfg["model_type"] == "dual_branch":
        return DualBranchAttentionModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
    return HighIntensityDehazeModel(base_channels=cfg["channels"], n_blocks=cfg["blocks"])
