"""ORACLE — test infrastructure only.  Never imported by the product package (adam_dehaze_b200/); only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.

A plain PyTorch fp32 restatement of the ADAM-Dehaze hot path, written functionally over a `state_dict` (the reference's
own parameter names) so that it runs with any weights on any device and travels to the GPU box, where
/root/reference does not exist.  Every function cites the reference lines it restates.

Pinning: tests/test_oracle_golden.py checks each function against tests/golden/*.pt, which were produced by importing the
UNMODIFIED reference modules from /root/reference (script: oracle/make_golden.py).  Exceptions, "parity unpinned":
  * perceptual_lpips(): the `lpips` package is not installed anywhere offline, so the LPIPS-alex arithmetic follows the
    published definition (see the docstring) and no reference output pins it;
  * classifier with model_name='densenet121': the reference has no DenseNet arm (classifier.py:22-69 raises ValueError);
    the oracle is torchvision.models.densenet121 + the reference's own head.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default


# --------------------------------------------------------------------------- building blocks (base_model.py)
_TRAIN_BN = False


class train_mode:
    """`with train_mode():` — BatchNorm layers use batch statistics, as under model.train() (train_dehazing.py:66).
    The functional state_dict is not updated (running statistics are a side effect the caller can compute itself)."""

    def __enter__(self):
        global _TRAIN_BN
        self._old, _TRAIN_BN = _TRAIN_BN, True

    def __exit__(self, *a):
        global _TRAIN_BN
        _TRAIN_BN = self._old


def _bn(sd, p, x):
    """nn.BatchNorm2d, base_model.py:15-16: running statistics in eval mode, batch statistics inside train_mode()."""
    if _TRAIN_BN:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, EPS)


def conv_block(sd, p, x, stride=1, padding=1, bn=True, act="relu"):
    """ConvBlock: Conv2d(bias iff no BN) -> BN -> activation, base_model.py:4-24."""
    y = F.conv2d(x, sd[p + ".block.0.weight"], sd.get(p + ".block.0.bias"), stride=stride, padding=padding)
    if bn:
        y = _bn(sd, p + ".block.1", y)
    if act == "relu":
        y = F.relu(y)
    return y


def residual_block(sd, p, x):
    """ResidualBlock: ConvBlock(ReLU) -> ConvBlock(no act) -> += x -> ReLU, base_model.py:26-41."""
    y = conv_block(sd, p + ".conv1", x)
    y = conv_block(sd, p + ".conv2", y, act=None)
    return F.relu(y + x)


def attention_block(sd, p, x):
    """AttentionBlock (CBAM-style channel gate then 7x7 spatial gate), base_model.py:43-78."""
    def fc(v):
        return F.conv2d(F.relu(F.conv2d(v, sd[p + ".fc.0.weight"])), sd[p + ".fc.2.weight"])
    gate = torch.sigmoid(fc(F.adaptive_avg_pool2d(x, 1)) + fc(F.adaptive_max_pool2d(x, 1)))
    x = x * gate
    stats = torch.cat([x.mean(dim=1, keepdim=True), x.max(dim=1, keepdim=True)[0]], dim=1)
    return x * torch.sigmoid(F.conv2d(stats, sd[p + ".conv_spatial.weight"], padding=3))


def _up_block(sd, p, x):
    """ConvTranspose2d(4,2,1) -> BN -> ReLU (decoder heads, medium_intensity.py:53-55; high_intensity.py:57-59)."""
    y = F.conv_transpose2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=2, padding=1)
    return F.relu(_bn(sd, p + ".1", y))


def _match(x, ref):
    """Bilinear fallback when sizes differ (only if H or W is not a multiple of 4), medium:92-99; high:110-116."""
    if x.shape[2:] != ref.shape[2:]:
        x = F.interpolate(x, size=ref.shape[2:], mode="bilinear", align_corners=False)
    return x


# --------------------------------------------------------------------------- branches
def light_forward(sd, x, n_blocks=3):
    """LightweightDehazeModel.forward, low_intensity.py:33-45.  No clamp."""
    f = conv_block(sd, "init_conv", x)
    for i in range(n_blocks):
        f = residual_block(sd, f"residual_blocks.{i}", f)
    f = conv_block(sd, "output_conv.0", f)
    out = torch.sigmoid(F.conv2d(f, sd["output_conv.1.weight"], sd["output_conv.1.bias"], padding=1))
    a = sd["skip_alpha"]
    return (1 - a) * x + a * out


def _unet_forward(sd, x, attention):
    """Shared body of MediumIntensityDehazeModel.forward (medium_intensity.py:78-114) and
    HighIntensityDehazeModel.forward (high_intensity.py:96-132) up to the tanh residual."""
    feats = [conv_block(sd, "init_conv", x, padding=3)]
    for e in range(2):
        f = conv_block(sd, f"encoder.{e}.0", feats[-1], stride=2, padding=1)
        f = residual_block(sd, f"encoder.{e}.1", f)
        f = residual_block(sd, f"encoder.{e}.2", f)
        if attention:
            f = attention_block(sd, f"encoder.{e}.3", f)
        feats.append(f)
    b = feats[-1]
    if attention:   # Res, Attn, Res, Attn  (high_intensity.py:44-49)
        b = attention_block(sd, "bottleneck.1", residual_block(sd, "bottleneck.0", b))
        b = attention_block(sd, "bottleneck.3", residual_block(sd, "bottleneck.2", b))
    else:           # Res, Res            (medium_intensity.py:42-45)
        b = residual_block(sd, "bottleneck.1", residual_block(sd, "bottleneck.0", b))
    x1 = residual_block(sd, "decoder.0.3", _up_block(sd, "decoder.0", b))
    if attention:
        x1 = attention_block(sd, "decoder.0.4", x1)
    x1 = torch.cat([_match(x1, feats[-2]), feats[-2]], dim=1)
    x2 = residual_block(sd, "decoder.1.3", _up_block(sd, "decoder.1", x1))
    if attention:
        x2 = attention_block(sd, "decoder.1.4", x2)
    x2 = torch.cat([_match(x2, feats[0]), feats[0]], dim=1)
    r = conv_block(sd, "output_conv.0", x2)
    r = conv_block(sd, "output_conv.1", r)
    return torch.tanh(F.conv2d(r, sd["output_conv.2.weight"], sd["output_conv.2.bias"], padding=1))


def medium_forward(sd, x):
    """MediumIntensityDehazeModel.forward, medium_intensity.py:78-117."""
    return torch.clamp(x + _unet_forward(sd, x, attention=False), 0, 1)


def complex_forward(sd, x):
    """HighIntensityDehazeModel.forward, high_intensity.py:92-138."""
    g = conv_block(sd, "detail_branch.0", x)
    g = conv_block(sd, "detail_branch.1", g)
    g = torch.sigmoid(F.conv2d(g, sd["detail_branch.2.weight"], sd["detail_branch.2.bias"]))
    return torch.clamp(x + _unet_forward(sd, x, attention=True) * g, 0, 1)


# --------------------------------------------------------------------------- non-default variants
def low_unet_forward(sd, x):
    """LowIntensityDehazeModel.forward, low_intensity.py:96-115 (n_blocks - 1 bottleneck blocks, read from the keys)."""
    f0 = conv_block(sd, "init_conv", x)
    f = residual_block(sd, "down1.1", conv_block(sd, "down1.0", f0, stride=2, padding=1))
    i = 0
    while f"bottleneck.{i}.conv1.block.0.weight" in sd:
        f = residual_block(sd, f"bottleneck.{i}", f)
        i += 1
    up = _up_block(sd, "up1", f)
    r = conv_block(sd, "output_conv.0", torch.cat([up, f0], dim=1))
    r = conv_block(sd, "output_conv.1", r)
    out = torch.sigmoid(F.conv2d(r, sd["output_conv.2.weight"], sd["output_conv.2.bias"], padding=1))
    return torch.clamp(x + (out - 0.5) * 2, 0, 1)


def corun_forward(sd, x):
    """COrunInspiredModel.forward, medium_intensity.py:170-190.  nn.UpsamplingBilinear2d == bilinear, align_corners=True."""
    f0 = conv_block(sd, "init_conv", x, padding=3)
    s1 = conv_block(sd, "scale1_conv", f0)
    s2 = F.interpolate(conv_block(sd, "scale2_conv.1", F.max_pool2d(f0, 2, 2)), scale_factor=2, mode="bilinear", align_corners=True)
    s3 = F.interpolate(conv_block(sd, "scale3_conv.1", F.max_pool2d(f0, 4, 4)), scale_factor=4, mode="bilinear", align_corners=True)
    f = conv_block(sd, "fusion_conv", torch.cat([s1, s2, s3], dim=1), padding=0)
    i = 0
    while f"residual_blocks.{i}.conv1.block.0.weight" in sd:
        f = residual_block(sd, f"residual_blocks.{i}", f)
        i += 1
    r = conv_block(sd, "output_conv.0", f)
    r = torch.tanh(F.conv2d(r, sd["output_conv.1.weight"], sd["output_conv.1.bias"], padding=1))
    return torch.clamp(x + r, 0, 1)


def dual_branch_forward(sd, x):
    """DualBranchAttentionModel.forward, high_intensity.py:203-223."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)   # noqa: E731
    g = conv_block(sd, "global_branch.0", x, padding=3)
    g = attention_block(sd, "global_branch.3", residual_block(sd, "global_branch.2", F.max_pool2d(g, 2, 2)))
    g = attention_block(sd, "global_branch.6", residual_block(sd, "global_branch.5", F.max_pool2d(g, 2, 2)))
    g = up(residual_block(sd, "global_branch.7", g))
    g = up(residual_block(sd, "global_branch.9", g))
    g = conv_block(sd, "global_branch.11", g)
    l = conv_block(sd, "local_branch.0", x)
    l = residual_block(sd, "local_branch.2", residual_block(sd, "local_branch.1", l))
    l = conv_block(sd, "local_branch.3", l)
    cat = torch.cat([g, l], dim=1)
    t = conv_block(sd, "transmission_branch.1", conv_block(sd, "transmission_branch.0", cat))
    t = torch.sigmoid(F.conv2d(t, sd["transmission_branch.2.weight"], sd["transmission_branch.2.bias"]))
    r = conv_block(sd, "fusion_conv.0", cat)
    r = torch.tanh(F.conv2d(r, sd["fusion_conv.1.weight"], sd["fusion_conv.1.bias"], padding=1))
    return torch.clamp(x + (1 - t) * r, 0, 1)


BRANCH_FORWARD = {"low": light_forward, "medium": medium_forward, "high": complex_forward,
                  "low_unet": low_unet_forward, "corun": corun_forward, "dual_branch": dual_branch_forward}


# --------------------------------------------------------------------------- classifier (HDEN)
def _basic_block(sd, p, x, stride):
    """torchvision BasicBlock as called through models/classifier.py:24-36."""
    y = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1)))
    y = _bn(sd, p + ".bn2", F.conv2d(y, sd[p + ".conv2.weight"], padding=1))
    if p + ".downsample.0.weight" in sd:
        x = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(y + x)


def resnet_features(sd, x, p="backbone", blocks=(2, 2, 2, 2)):
    """torchvision resnet18/34 with fc = Identity (classifier.py:24-36): features [B, 512]."""
    y = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=2, padding=3)))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, nb in enumerate(blocks):
        for bi in range(nb):
            y = _basic_block(sd, f"{p}.layer{li + 1}.{bi}", y, stride=2 if (li > 0 and bi == 0) else 1)
    return torch.flatten(F.adaptive_avg_pool2d(y, 1), 1)


def densenet121_features(sd, x, p="backbone"):
    """torchvision densenet121 with classifier = Identity (north_star HDEN; not in the reference — parity unpinned)."""
    y = F.conv2d(x, sd[p + ".features.conv0.weight"], stride=2, padding=3)
    y = F.max_pool2d(F.relu(_bn(sd, p + ".features.norm0", y)), 3, 2, 1)
    for bi, layers in enumerate((6, 12, 24, 16)):
        for li in range(layers):
            q = f"{p}.features.denseblock{bi + 1}.denselayer{li + 1}"
            t = F.conv2d(F.relu(_bn(sd, q + ".norm1", y)), sd[q + ".conv1.weight"])
            t = F.conv2d(F.relu(_bn(sd, q + ".norm2", t)), sd[q + ".conv2.weight"], padding=1)
            y = torch.cat([y, t], 1)
        if bi < 3:
            q = f"{p}.features.transition{bi + 1}"
            y = F.avg_pool2d(F.conv2d(F.relu(_bn(sd, q + ".norm", y)), sd[q + ".conv.weight"]), 2, 2)
    y = F.relu(_bn(sd, p + ".features.norm5", y))
    return torch.flatten(F.adaptive_avg_pool2d(y, 1), 1)


def classifier_forward(sd, x, model_name="resnet18"):
    """FogIntensityClassifier.forward in eval mode (dropout = identity): returns (logits, features), classifier.py:80-97."""
    if model_name == "resnet18":
        feats = resnet_features(sd, x, blocks=(2, 2, 2, 2))
    elif model_name == "resnet34":
        feats = resnet_features(sd, x, blocks=(3, 4, 6, 3))
    elif model_name == "densenet121":
        feats = densenet121_features(sd, x)
    else:
        raise ValueError(f"Unsupported model: {model_name}")
    h = F.relu(F.linear(feats, sd["classifier.1.weight"], sd["classifier.1.bias"]))
    return F.linear(h, sd["classifier.4.weight"], sd["classifier.4.bias"]), feats


# --------------------------------------------------------------------------- routing (routing.py)
def route_indices(logits=None, intensity=None):
    """argmax + the three ascending bucket lists, routing.py:40-57 (x[mask] keeps ascending batch order)."""
    if intensity is None:
        intensity = torch.argmax(logits, dim=1)
    return intensity, [torch.nonzero(intensity == k).flatten() for k in range(3)]


def hard_route(branch_sds, x, logits=None, intensity=None):
    """HardRouter.forward, routing.py:23-68.  branch_sds: {'low','medium','high'} -> state_dict."""
    intensity, buckets = route_indices(logits, intensity)
    out = torch.zeros_like(x)
    for k, name in enumerate(("low", "medium", "high")):
        if buckets[k].numel():
            out[buckets[k]] = BRANCH_FORWARD[name](branch_sds[name], x[buckets[k]])
    return out, intensity, buckets


def soft_route(branch_sds, x, logits, temperature):
    """SoftRouter.forward, routing.py:90-132."""
    w = F.softmax(logits / temperature, dim=1)
    outs = {name: BRANCH_FORWARD[name](branch_sds[name], x) for name in ("low", "medium", "high")}
    blend = torch.zeros_like(x)
    for i, name in enumerate(("low", "medium", "high")):
        blend += w[:, i].view(-1, 1, 1, 1) * outs[name]
    return blend, w, outs


def gate_weights(sd, feats, p="gate_network"):
    """GatedRouter gate MLP in eval mode, routing.py:155-163."""
    h = F.relu(F.linear(feats, sd[p + ".0.weight"], sd[p + ".0.bias"]))
    h = F.relu(F.linear(h, sd[p + ".3.weight"], sd[p + ".3.bias"]))
    return F.softmax(F.linear(h, sd[p + ".5.weight"], sd[p + ".5.bias"]), dim=1)


# --------------------------------------------------------------------------- losses (training/loss.py)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
VGG16_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")


def vgg16_prefix(vgg_sd, x, last_index):
    """torchvision vgg16().features[:last_index+1] (conv3x3+ReLU / MaxPool2d(2)), loss.py:73-74."""
    idx = 0
    for v in VGG16_CFG:
        if idx > last_index:
            break
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            x = F.conv2d(x, vgg_sd[f"{idx}.weight"], vgg_sd[f"{idx}.bias"], padding=1)
            idx += 1
            if idx > last_index:
                break
            x = F.relu(x)
            idx += 1
    return x


def content_loss(vgg_sd, pred, target):
    """ContentLoss.forward, loss.py:47-84.  The layer map sends 'relu2_2/3_3/4_3' to indices 9/16/23, which are the
    MaxPool layers pool2/pool3/pool4 — restated as written, not as named."""
    mean = torch.tensor(IMAGENET_MEAN, device=pred.device).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD, device=pred.device).view(1, 3, 1, 1)
    p, t = (pred - mean) / std, (target - mean) / std
    loss = 0.0
    for idx in (9, 16, 23):
        loss = loss + F.mse_loss(vgg16_prefix(vgg_sd, p, idx), vgg16_prefix(vgg_sd, t, idx))
    return loss / 3


def perceptual_lpips(alex_sd, lin_ws, pred, target):
    """PerceptualLoss.forward, loss.py:86-108, with LPIPS(net='alex') restated from its published definition
    (Zhang et al. 2018; lpips==0.1.4 lpips/lpips.py): inputs mapped to [-1,1] (loss.py:104-105), LPIPS' ScalingLayer
    (x - shift)/scale, five AlexNet ReLU taps, channel-wise unit normalisation, squared difference, a non-negative
    1x1 'lin' layer per tap, spatial mean, sum over taps -> [B,1,1,1].  PARITY UNPINNED (package absent offline)."""
    shift = torch.tensor([-.030, -.088, -.188], device=pred.device).view(1, 3, 1, 1)
    scale = torch.tensor([.458, .448, .450], device=pred.device).view(1, 3, 1, 1)

    def taps(x):
        x = (2 * x - 1 - shift) / scale
        outs = []
        x = F.relu(F.conv2d(x, alex_sd["0.weight"], alex_sd["0.bias"], stride=4, padding=2)); outs.append(x)
        x = F.max_pool2d(x, 3, 2)
        x = F.relu(F.conv2d(x, alex_sd["3.weight"], alex_sd["3.bias"], padding=2)); outs.append(x)
        x = F.max_pool2d(x, 3, 2)
        x = F.relu(F.conv2d(x, alex_sd["6.weight"], alex_sd["6.bias"], padding=1)); outs.append(x)
        x = F.relu(F.conv2d(x, alex_sd["8.weight"], alex_sd["8.bias"], padding=1)); outs.append(x)
        x = F.relu(F.conv2d(x, alex_sd["10.weight"], alex_sd["10.bias"], padding=1)); outs.append(x)
        return outs

    def unit(f):
        return f / (torch.sqrt(torch.sum(f ** 2, dim=1, keepdim=True)) + 1e-10)

    total = 0
    for fp, ft, w in zip(taps(pred), taps(target), lin_ws):
        d = (unit(fp) - unit(ft)) ** 2
        total = total + F.conv2d(d, w).mean(dim=(2, 3), keepdim=True)
    return total


def dehazing_loss(pred, target, content, perceptual, lambdas=(1.0, 0.1, 0.1)):
    """DehazingLoss.forward combination, loss.py:125-162 (perceptual reduced with .mean())."""
    l1 = F.l1_loss(pred, target)
    perceptual = perceptual.mean() if perceptual.dim() > 0 else perceptual
    total = lambdas[0] * l1 + lambdas[1] * content + lambdas[2] * perceptual
    return total, {"l1": l1, "content": content, "perceptual": perceptual, "total": total}


def joint_loss(dehazing, logits=None, labels=None, detection=None, lambdas=(1.0, 0.2, 0.5)):
    """JointLoss.forward combination, loss.py:179-224."""
    ce = F.cross_entropy(logits, labels) if (logits is not None and labels is not None) else torch.tensor(0.0, device=dehazing.device)
    det = detection if detection is not None else torch.tensor(0.0, device=dehazing.device)
    total = lambdas[0] * dehazing + lambdas[1] * ce + lambdas[2] * det
    return total, {"dehazing": dehazing, "classification": ce, "detection": det, "total": total}


# --------------------------------------------------------------------------- synthetic hazy inputs
def synth_hazy(n, h, w, betas=(0.03, 0.06, 0.09), seed=42, device="cpu"):
    """Clear J ~ U[0,1), hazy I = clip(J*t + A(1-t)), t = exp(-beta*d), depth from utils/helpers.py:241-249 scaled to
    metres (x100), A = 0.8 (helpers.py:250-255).  beta cycles round-robin over the batch.  Returns (hazy, clear, labels)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    clear = torch.rand(n, 3, h, w, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    d = (0.3 + 0.7 * torch.sqrt((xx - 0.5) ** 2 + (yy - 0.2) ** 2)) * 100.0
    labels = torch.arange(n) % len(betas)
    beta = torch.tensor(betas)[labels].view(n, 1, 1, 1)
    t = torch.exp(-beta * d.view(1, 1, h, w))
    hazy = torch.clamp(clear * t + 0.8 * (1 - t), 0, 1)
    return hazy.to(device), clear.to(device), labels.to(device)


# --------------------------------------------------------------------------- detection hand-off (models/detection.py)
def detection_normalize(dehazed):
    """IntegratedDetectionSystem.forward, models/detection.py:109-121: per image, (img - mean[c]) / std[c] with the ImageNet
    constants; returns the list of [3,H,W] tensors the detector receives.  Pinned by tests/golden/detection_handoff.pt
    (oracle/make_golden_detection.py runs the unmodified reference class)."""
    mean = torch.tensor([0.485, 0.456, 0.406], device=dehazed.device).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=dehazed.device).view(3, 1, 1)
    return [img.clone().sub(mean).div(std) for img in dehazed]


# --------------------------------------------------------------------------- image-quality metrics (evaluation/metrics.py)
def image_metrics(pred, target):
    """calculate_image_metrics, evaluation/metrics.py:13-36, for ONE image pair ([3,H,W] tensors in [0,1]).

    scikit-image is not installed offline, so PSNR/SSIM follow its published algorithms (skimage 0.19+
    `peak_signal_noise_ratio(data_range=1)`; `structural_similarity(gray_t, gray_p, data_range=1)` with the defaults:
    7x7 uniform_filter, use_sample_covariance=True, K1=0.01, K2=0.03, mean over the map cropped by 3 pixels) restated
    with scipy.ndimage.uniform_filter in float64.  PARITY UNPINNED against skimage itself."""
    import numpy as np
    from scipy.ndimage import uniform_filter
    p = pred.detach().cpu().double().numpy().transpose(1, 2, 0)
    t = target.detach().cpu().double().numpy().transpose(1, 2, 0)
    mse = np.mean((t - p) ** 2)
    psnr = 10.0 * np.log10(1.0 / mse)
    x, y = t.mean(axis=2), p.mean(axis=2)
    win, npix = 7, 49
    cov_norm = npix / (npix - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return {"psnr": float(psnr), "ssim": float(s[pad:-pad, pad:-pad].mean())}
