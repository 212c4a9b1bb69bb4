"""GPU parity of the HBM-bound kernels (attention passes, pooling, head, routing, blend, losses) against torch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from adam_dehaze_b200 import ops
    return ops


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.cuda.synchronize()


@pytest.mark.parametrize("c,h,w,n", [(96, 32, 64, 2), (192, 16, 32, 2), (384, 8, 16, 1), (96, 37, 21, 1)])
def test_attention_block(c, h, w, n):
    """AttentionBlock (base_model.py:64-78) on a bf16 map; tolerance 1e-2 relative (bf16 output rounding)."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, c, h, w, generator=g).relu().to(torch.bfloat16).float().cuda()
    fc0 = (torch.randn(c // 16, c, 1, 1, generator=g) / c ** 0.5).cuda()
    fc2 = (torch.randn(c, c // 16, 1, 1, generator=g) / (c // 16) ** 0.5).cuda()
    wsp = (torch.randn(1, 2, 7, 7, generator=g) / 98 ** 0.5).cuda()
    y = ops.attention(ops.nchw_to_nhwc(x), ops.AttnParams(fc0, fc2, wsp))
    fc = lambda t: F.conv2d(F.relu(F.conv2d(t, fc0)), fc2)
    ch = torch.sigmoid(fc(F.adaptive_avg_pool2d(x, 1)) + fc(F.adaptive_max_pool2d(x, 1)))
    xr = x * ch
    sp = torch.sigmoid(F.conv2d(torch.cat([xr.mean(1, keepdim=True), xr.max(1, keepdim=True)[0]], 1), wsp, padding=3))
    ref = xr * sp
    out = ops.nhwc_to_nchw(y)
    err = (out - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err


def test_maxpool_avgpool_head():
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 33, 48, generator=g).to(torch.bfloat16).float().cuda()
    xh = ops.nchw_to_nhwc(x)
    mp = ops.nhwc_to_nchw(ops.maxpool3x3s2(xh))
    assert torch.equal(mp, F.max_pool2d(x, 3, 2, 1))
    ap = ops.global_avgpool(xh)
    assert torch.allclose(ap, x.mean((2, 3)), rtol=1e-4, atol=1e-5)
    feat = torch.randn(4, 512, generator=g).cuda()
    w1 = torch.randn(256, 512, generator=g).cuda() / 22
    b1 = torch.randn(256, generator=g).cuda()
    w2 = torch.randn(3, 256, generator=g).cuda() / 16
    b2 = torch.randn(3, generator=g).cuda()
    logits = ops.head_mlp(feat, w1, b1, w2, b2)
    ref = F.linear(F.relu(F.linear(feat, w1, b1)), w2, b2)
    assert torch.allclose(logits, ref, rtol=1e-4, atol=1e-4)


def _route_ref(logits):
    inten = torch.argmax(logits, dim=1)
    return inten, [torch.nonzero(inten == k).flatten() for k in range(3)]


@pytest.mark.parametrize("b", [1, 7, 32, 256, 1500, 4097])
def test_route_bit_exact(b):
    """argmax + bucket lists must equal torch.argmax / nonzero exactly, incl. exact ties, 1-ulp gaps and NaNs."""
    ops = _ops()
    g = torch.Generator().manual_seed(b)
    logits = torch.randn(b, 3, generator=g)
    if b >= 7:
        logits[1] = torch.tensor([0.5, 0.5, 0.5])                       # 3-way tie -> class 0
        logits[2] = torch.tensor([0.1, 0.7, 0.7])                       # tie -> first max (1)
        v = torch.tensor(0.3)
        logits[3] = torch.stack([v, torch.nextafter(v, torch.tensor(1.0)), v])  # 1-ulp gap -> 1
        logits[4] = torch.tensor([0.2, float("nan"), 5.0])              # NaN counts as max
        logits[5] = torch.tensor([float("-inf"), float("-inf"), float("-inf")])
        logits[6] = torch.tensor([-0.0, 0.0, -0.0])                     # signed zeros tie -> 0
    logits = logits.cuda()
    inten, masks, bidx, bcnt = ops.route(logits)
    ref_int, ref_lists = _route_ref(logits)
    assert torch.equal(inten, ref_int)
    cnt = bcnt.cpu().tolist()
    for k in range(3):
        assert torch.equal(masks[k], ref_int == k)
        assert cnt[k] == ref_lists[k].numel()
        assert torch.equal(bidx[k, :cnt[k]].long(), ref_lists[k])
    assert sum(cnt) == b


def test_route_given_intensity():
    ops = _ops()
    inten_in = torch.tensor([2, 0, 1, 1, 0, 2, 2, 5], device="cuda")
    inten, masks, bidx, bcnt = ops.route(intensity=inten_in)
    assert torch.equal(inten, inten_in)
    assert bcnt.cpu().tolist() == [2, 2, 3]     # class 5 is routed nowhere, as in routing.py:46-50
    assert bidx[0, :2].tolist() == [1, 4] and bidx[1, :2].tolist() == [2, 3] and bidx[2, :3].tolist() == [0, 5, 6]


def test_blend3():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    ys = [torch.rand(4, 3, 16, 32, generator=g).cuda() for _ in range(3)]
    logits = torch.randn(4, 3, generator=g).cuda()
    out, wts = ops.blend3(*ys, logits, 0.5)
    w = F.softmax(logits / 0.5, dim=1)
    ref = torch.zeros_like(ys[0])
    for i in range(3):
        ref += w[:, i].view(4, 1, 1, 1) * ys[i]
    assert torch.allclose(wts, w, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_losses():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    p = torch.rand(3, 3, 33, 47, generator=g).cuda()
    t = torch.rand(3, 3, 33, 47, generator=g).cuda()
    out = ops.l1_mse(p, t)
    assert torch.allclose(out[0], F.l1_loss(p, t), rtol=1e-4)
    assert torch.allclose(out[1], F.mse_loss(p, t), rtol=1e-4)
    logits = torch.randn(16, 3, generator=g).cuda().requires_grad_(True)
    labels = torch.randint(0, 3, (16,), generator=g).cuda()
    loss, grad = ops.cross_entropy(logits.detach(), labels, grad_scale=0.2)
    ref = F.cross_entropy(logits, labels)
    (0.2 * ref).backward()
    assert torch.allclose(loss[0], ref.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(grad, logits.grad, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("kh,kw,pad,stride,kp,h,w", [(1, 3, 1, 1, 16, 9, 300), (1, 7, 3, 1, 32, 5, 257), (7, 7, 3, 2, 160, 64, 160),
                                                     (1, 5, 2, 1, 16, 8, 40)])   # last one: the generic (untiled) kernel
def test_stem_pack_bit_exact(kh, kw, pad, stride, kp, h, w):
    """adb_stem_pack against an unfold-based restatement (bf16 rounding of the same fp32 pixels -> bit-exact), with a
    routed index list and a live count below n."""
    import torch.nn.functional as F
    from adam_dehaze_b200 import ops
    g = torch.Generator().manual_seed(5)
    xb = torch.rand(5, 3, h, w, generator=g).cuda()
    index = torch.tensor([3, 0, 4, 1], dtype=torch.int32, device="cuda")
    n_dev = torch.tensor([3], dtype=torch.int32, device="cuda")
    wo = (w + 2 * pad - kw) // stride + 1
    ho = (h + 2 * pad - kh) // stride + 1 if kh > 1 else h
    out = torch.full((4, ho, wo, kp), -1.0, dtype=torch.bfloat16, device="cuda")
    ops.stem_pack(xb, kw, pad, kp, stride=stride, kh=kh, index=index, n_dev=n_dev, n=4, out=out)
    xs = xb[index[:3].long()]
    if kh == 1:
        cols = F.unfold(xs, (1, kw), padding=(0, pad), stride=(1, stride))          # [n, 3*kw, h*wo], channel-major (c, s)
        cols = cols.view(3, 3, 1, kw, ho, wo)
    else:
        cols = F.unfold(xs, (kh, kw), padding=pad, stride=stride).view(3, 3, kh, kw, ho, wo)
    ref = cols.permute(0, 4, 5, 2, 3, 1).reshape(3, ho, wo, kh * kw * 3)             # j = (r*kw + s)*3 + c
    assert torch.equal(out[:3, ..., :kh * kw * 3].float(), ref.to(torch.bfloat16).float())
    assert out[:3, ..., kh * kw * 3:].float().abs().max().item() == 0.0
    assert (out[3].float() == -1.0).all()


@pytest.mark.parametrize("k", [2, 4])
def test_maxpool_kxk_bit_exact(k):
    import torch.nn.functional as F
    from adam_dehaze_b200 import ops
    x = torch.randn(3, 24, 16, 40, generator=torch.Generator().manual_seed(8)).to(torch.bfloat16).float().cuda()
    y = ops.maxpool_kxk(ops.nchw_to_nhwc(x), k)
    assert torch.equal(ops.nhwc_to_nchw(y), F.max_pool2d(x, k, k))


@pytest.mark.parametrize("scale", [2, 4])
def test_upsample_bilinear_align_corners(scale):
    """nn.UpsamplingBilinear2d(scale_factor) into a channel slice of a wider buffer; fp32 interpolation of bf16 inputs,
    bf16 output: |err| <= 1 bf16 ulp of the largest magnitude (2^-8 relative)."""
    from adam_dehaze_b200 import ops
    x = torch.randn(2, 16, 9, 13, generator=torch.Generator().manual_seed(9)).to(torch.bfloat16).float().cuda()
    out = torch.full((2, 9 * scale, 13 * scale, 40), 3.0, dtype=torch.bfloat16, device="cuda")
    ops.upsample_bilinear(ops.nchw_to_nhwc(x), scale, out=out, c_off=8)
    ref = torch.nn.UpsamplingBilinear2d(scale_factor=scale)(x)
    got = ops.nhwc_to_nchw(out)[:, 8:24]
    assert (got - ref).abs().max().item() <= ref.abs().max().item() * 2 ** -8
    assert (out[..., :8].float() == 3.0).all() and (out[..., 24:].float() == 3.0).all()


def test_image_metrics_on_device():
    """PSNR / SSIM on the device vs the oracle's restatement of skimage's algorithms.
    Tolerance: fp32 window sums vs float64 — |dPSNR| <= 1e-3 dB, |dSSIM| <= 1e-4."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import adam_oracle as oracle
    from adam_dehaze_b200.evaluation.metrics import ImageQualityMetrics, calculate_image_metrics, image_metrics
    g = torch.Generator().manual_seed(3)
    base = torch.nn.functional.interpolate(torch.rand(3, 3, 12, 20, generator=g), size=(45, 77), mode="bilinear")
    tgt = base.clamp(0, 1).cuda()
    pred = (base + 0.05 * torch.randn(3, 3, 45, 77, generator=g)).clamp(0, 1).cuda()
    psnr, ssim = image_metrics(pred, tgt)
    for i in range(3):
        ref = oracle.image_metrics(pred[i], tgt[i])
        assert abs(psnr[i].item() - ref["psnr"]) <= 1e-3, (psnr[i].item(), ref["psnr"])
        assert abs(ssim[i].item() - ref["ssim"]) <= 1e-4, (ssim[i].item(), ref["ssim"])
    one = calculate_image_metrics(pred[0], tgt[0])
    assert abs(one["psnr"] - psnr[0].item()) < 1e-6 and set(one) == {"psnr", "ssim"}
    m = ImageQualityMetrics(with_lpips=False)
    m.add_sample(pred[0], tgt[0], category="low")
    m.add_sample(pred[1:], tgt[1:])
    avg = m.compute_averages()
    assert avg["low"]["samples"] == 1 and avg["all"]["samples"] == 2
    assert abs(avg["all"]["psnr"] - psnr[1:].mean().item()) < 1e-4
    p_same, s_same = image_metrics(tgt, tgt)
    assert torch.isinf(p_same).all() and (s_same - 1).abs().max().item() < 1e-5


def test_image_metrics_closed_form_cases():
    """PSNR / SSIM pins that need no skimage (evaluation/metrics.py:13-36; skimage is absent offline, so the oracle's restatement
    is otherwise unpinned): identical images -> SSIM exactly 1 and PSNR = +inf (or the 1e-12 floor's 120 dB); a constant offset d
    -> PSNR = -20 log10(d) analytically and SSIM = (2 mu1 mu2 + C1) / (mu1^2 + mu2^2 + C1) on a constant image (zero variance);
    SSIM is symmetric in its arguments and decreases as noise grows.  Values from the definitions in Wang et al. 2004 with
    skimage's defaults (K1 = 0.01, K2 = 0.03, 7x7 uniform window, data_range = 1)."""
    from adam_dehaze_b200.evaluation.metrics import image_metrics
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand((2, 3, 48, 64), generator=g, device="cuda")
    psnr, ssim = image_metrics(x, x.clone())
    assert (ssim - 1.0).abs().max().item() <= 1e-6 and (psnr >= 100).all()
    d = 0.1
    psnr, _ = image_metrics((x * 0.8), (x * 0.8 + d))
    assert (psnr - 20.0).abs().max().item() <= 1e-3                     # -20 log10(0.1) = 20 dB
    a, b = torch.full((1, 3, 32, 32), 0.4, device="cuda"), torch.full((1, 3, 32, 32), 0.6, device="cuda")
    _, ssim = image_metrics(a, b)
    c1 = 0.01 ** 2
    want = (2 * 0.4 * 0.6 + c1) / (0.4 ** 2 + 0.6 ** 2 + c1)
    # (fp32 windows: u_xx - u_x^2 cancels to ~1e-7 against C2 = 9e-4 — skimage on float32 input has the same noise)
    assert abs(ssim.item() - want) <= 1e-3, (ssim.item(), want)
    n1, n2 = x + 0.05 * torch.randn_like(x), x + 0.2 * torch.randn_like(x)
    s_ab = image_metrics(x, n1.clamp(0, 1))[1]
    s_ba = image_metrics(n1.clamp(0, 1), x)[1]
    assert (s_ab - s_ba).abs().max().item() <= 1e-6
    assert (image_metrics(x, n2.clamp(0, 1))[1] < s_ab).all()


def test_lpips_structure_without_the_package():
    """LPIPS pins that hold for ANY weights (lpips is absent offline, loss.py:86-108 -> parity unpinned): d(x, x) = 0,
    d is symmetric, non-negative (lin layers are non-negative, features are unit-normalised), scale-consistent with the
    published definition (a sum over five taps of spatial means), and the fallback init is the same in every process."""
    from adam_dehaze_b200.training.loss import PerceptualLoss
    torch.manual_seed(1)
    a = PerceptualLoss().cuda()
    torch.manual_seed(2)
    b = PerceptualLoss().cuda()
    assert all(torch.equal(p, q) for p, q in zip(a.state_dict().values(), b.state_dict().values()))   # rank-independent fallback
    assert a.pretrained is False
    g = torch.Generator(device="cuda").manual_seed(4)
    x, y = torch.rand((2, 3, 96, 96), generator=g, device="cuda"), torch.rand((2, 3, 96, 96), generator=g, device="cuda")
    with torch.no_grad():
        assert a(x, x.clone()).abs().max().item() <= 1e-12
        dxy, dyx = a(x, y), a(y, x)
    assert dxy.shape == (2, 1, 1, 1) and (dxy > 0).all()
    assert (dxy - dyx).abs().max().item() <= 1e-3 * dxy.abs().max().item()
    with pytest.raises(FileNotFoundError):
        PerceptualLoss(allow_random_weights=False)


def test_cross_entropy_ignore_index_and_bad_labels():
    """nn.CrossEntropyLoss semantics (loss.py:177): rows labelled -100 are ignored (mean over the others, zero gradient); a
    label outside [0, classes) raises through the device error flag instead of reading out of bounds."""
    from adam_dehaze_b200 import _lib, ops
    logits = torch.randn(6, 3, device="cuda")
    labels = torch.tensor([0, -100, 2, 1, -100, 2], device="cuda")
    loss, grad = ops.cross_entropy(logits, labels)
    keep = labels != -100
    ref_logits = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(ref_logits, labels, ignore_index=-100)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-6 and (grad - ref_logits.grad).abs().max().item() <= 1e-6
    assert grad[~keep].abs().max().item() == 0.0
    ops.cross_entropy(logits, torch.tensor([0, 1, 7, 1, 0, 2], device="cuda"))
    torch.cuda.synchronize()
    with pytest.raises(_lib.AdbError):
        _lib.call("adb_kernel_error_flag")
    _lib.call("adb_kernel_error_flag")            # reading the flag clears it


def test_detection_normalisation_is_differentiable():
    from adam_dehaze_b200.models.detection import IMAGENET_STD, normalize_for_detection
    x = torch.rand(2, 3, 16, 24, device="cuda", requires_grad=True)
    y = normalize_for_detection(x)
    g = torch.randn_like(y)
    y.backward(g)
    want = g / torch.tensor(IMAGENET_STD, device="cuda").view(1, 3, 1, 1)
    assert (x.grad - want).abs().max().item() <= 1e-6
