"""GPU parity of the tcgen05 implicit-GEMM conv (adb_conv2d) against torch fp32 convolutions on the same bf16-rounded
operands.  Tolerance: inputs/weights are exactly representable, so the only differences are fp32 accumulation order
and the final bf16 rounding of the output: |err| <= 1e-2 * max|ref| + 1e-3 (stated per test)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from adam_dehaze_b200 import ops
    return ops


def _rand_fm(n, c, h, w, seed, relu=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, c, h, w, generator=g)
    if relu:
        x = x.relu()
    return x.to(torch.bfloat16).float().cuda()


def _rand_w(shape, seed, fan_in):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) / fan_in ** 0.5).to(torch.bfloat16).float().cuda()


def _bn(c, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(c, generator=g).cuda() + 0.5, torch.randn(c, generator=g).cuda() * 0.1,
            torch.randn(c, generator=g).cuda() * 0.1, torch.rand(c, generator=g).cuda() + 0.5, 1e-5)


def _bn_ref(y, bn):
    g, b, m, v, eps = bn
    return (y - m.view(1, -1, 1, 1)) / torch.sqrt(v.view(1, -1, 1, 1) + eps) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def _close(out, ref, rel=1e-2, abs_=1e-3):
    err = (out - ref).abs().max().item()
    bound = rel * ref.abs().max().item() + abs_
    assert err <= bound, f"max err {err:.4g} > {bound:.4g} (ref max {ref.abs().max().item():.4g})"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    from adam_dehaze_b200 import _lib
    torch.cuda.synchronize()
    _lib.call("adb_kernel_error_flag")


def test_layout_roundtrip():
    ops = _ops()
    x = _rand_fm(2, 20, 9, 13, 0)
    y = ops.nchw_to_nhwc(x, 24)
    assert y.shape == (2, 9, 13, 24)
    assert torch.equal(y[..., :20].permute(0, 3, 1, 2).float(), x)
    assert torch.equal(y[..., 20:].float(), torch.zeros_like(y[..., 20:].float()))
    z = ops.nhwc_to_nchw(y, 20)
    assert torch.equal(z, x)


CONV_CASES = [
    # (cin, cout, k, n, h, w, tune)
    (64, 64, 3, 1, 32, 32, None),
    (64, 64, 1, 1, 16, 16, None),
    (32, 32, 3, 2, 32, 64, None),
    (16, 16, 3, 1, 32, 32, None),
    (96, 96, 3, 1, 64, 64, None),
    (192, 192, 3, 1, 32, 32, None),
    (384, 384, 3, 1, 16, 32, None),
    (256, 256, 3, 1, 16, 16, None),
    (128, 128, 3, 1, 24, 40, None),          # ragged tiles (w not a power of two)
    (64, 48, 3, 1, 32, 32, None),
    (64, 64, 3, 4, 128, 256, None),          # many tiles per CTA: pipeline/accumulator phase wrap
    (64, 64, 3, 2, 64, 128, {"mt": 2}),
    (256, 256, 3, 2, 64, 128, {"flags": 16}),            # CTA pair: 256 channels as two 128-channel N tiles (the default rule)
    (256, 256, 3, 2, 64, 128, {"flags": 16 + 2048}),     # ... and with the split forbidden: one 256-column tile
    (192, 192, 3, 2, 32, 128, {"flags": 16 + 1024}),     # forced split: two 96-channel N tiles
    (384, 384, 3, 1, 32, 64, {"flags": 16 + 1024}),      # forced split of both 192-channel tiles: four N tiles
    (96, 96, 3, 2, 64, 128, {"mt": 2}),
    # single accumulator stage with two sub-tiles per CTA (N = 160 .. 192: the epilogue is exposed, accumulator released early)
    (192, 192, 3, 2, 32, 256, {"flags": 16}),            # N = 192 CTA pair, many tiles per CTA
    (192, 192, 3, 1, 24, 200, {"flags": 32}),            # 1-CTA, ragged
    (384, 384, 3, 1, 16, 128, {"flags": 16}),            # two N tiles of 192, six channel chunks
    (128, 128, 3, 2, 64, 128, {"flags": 16, "mt": 2, "acc": 1}),   # forced single stage on a shape that would double-buffer
    (96, 96, 3, 4, 64, 128, {"mt": 2, "acc": 1}),        # 32-channel slabs, ragged last K chunk
    (160, 160, 3, 2, 32, 128, {"flags": 16}),            # N = 160: five 32-channel slabs per sub-tile
    (64, 64, 3, 2, 64, 128, {"stages": 2, "acc": 1}),
    (64, 64, 3, 2, 64, 256, None),           # halo re-use path: tile = one image row (TW = 128), 128-byte swizzle
    (96, 96, 3, 1, 32, 128, None),           # halo path with 64-byte swizzle rows (Ck = 32)
    (48, 48, 3, 1, 32, 128, None),           # halo path with 32-byte swizzle rows (Ck = 16)
    (64, 64, 3, 2, 64, 256, {"flags": 1}),   # one TMA box per tap (no halo), same shape
    (192, 192, 3, 1, 16, 256, None),
    (256, 256, 3, 1, 8, 128, {"mt": 2}),
    # CTA-pair mode (cluster of 2, tcgen05 cta_group::2) forced on: flags bit 4
    (64, 64, 3, 2, 64, 128, {"flags": 16}),
    (64, 64, 3, 4, 128, 256, {"flags": 16}),          # many pair-tiles per cluster: ring / accumulator phase wrap
    (96, 96, 3, 2, 64, 128, {"flags": 16, "mt": 2}),
    (128, 128, 3, 1, 24, 40, {"flags": 16}),          # ragged: the peer's sub-tile partly / wholly outside the image
    (192, 192, 3, 1, 16, 256, {"flags": 16}),
    (256, 256, 3, 2, 16, 128, {"flags": 16}),
    (384, 384, 3, 1, 16, 32, {"flags": 16}),          # two N tiles of 192
    (64, 48, 3, 1, 32, 32, {"flags": 16}),            # N = 48: 24 weight rows per CTA
    (64, 64, 3, 1, 6, 128, {"flags": 16, "mt": 1}),   # 3 pair-rows: odd row count
    (512, 128, 1, 2, 16, 32, {"flags": 16}),          # 1x1 (DenseNet bottleneck shape)
    # 2-D tiles (16 rows x 8 pixels, all nine taps share one halo box; descriptor SBO = one halo row) forced on: flags bit 8
    (32, 32, 3, 2, 32, 64, {"flags": 256}),
    (64, 64, 3, 2, 64, 128, {"flags": 256, "mt": 1}),
    (64, 64, 3, 2, 64, 128, {"flags": 256, "mt": 2}),
    (128, 32, 3, 2, 40, 72, {"flags": 256}),           # DenseNet 3x3 shape, ragged both ways
    (96, 96, 3, 1, 48, 24, {"flags": 256}),            # ragged K chunks + 2-D tiles
    (64, 64, 3, 4, 128, 256, {"flags": 256 | 16}),     # 2-D tiles in CTA-pair mode
]


@pytest.mark.parametrize("cin,cout,k,n,h,w,tune", CONV_CASES)
def test_conv_s1(cin, cout, k, n, h, w, tune):
    ops = _ops()
    x = _rand_fm(n, cin, h, w, 1)
    wt = _rand_w((cout, cin, k, k), 2, cin * k * k)
    bn = _bn(cout, 3)
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(x), tune=tune)
    out = ops.nhwc_to_nchw(y, cout)
    ref = F.relu(_bn_ref(F.conv2d(x, wt, padding=k // 2), bn))
    _close(out, ref)
    if spec.cout_pad > cout:
        assert y[..., cout:].float().abs().max().item() == 0.0


@pytest.mark.parametrize("tune", [None, {"flags": 16}, {"flags": 256}], ids=["auto", "pair", "tile2d"])
def test_conv_bias_residual(tune):
    ops = _ops()
    x = _rand_fm(2, 64, 32, 48, 4)
    res = _rand_fm(2, 64, 32, 48, 5)
    wt = _rand_w((64, 64, 3, 3), 6, 576)
    bias = torch.randn(64, device="cuda") * 0.1
    spec = ops.ConvSpec.from_conv(wt, bias=bias, act=ops.ACT_RELU)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(x), residual=ops.nchw_to_nhwc(res), tune=tune)
    ref = F.relu(F.conv2d(x, wt, bias, padding=1) + res)
    _close(ops.nhwc_to_nchw(y), ref)


RES_CASES = [
    # (c, n, h, w, tune): the residual epilogue's addressing forms and slab widths
    (192, 2, 16, 128, {"flags": 16}),                    # CTA pair, two sub-tiles, three 64-channel slabs per warp pair (Complex)
    (192, 1, 9, 200, {"flags": 16}),                     # ... ragged in both directions (rows outside the image load nothing)
    (96, 2, 16, 128, None),                              # 32-channel slabs
    (48, 1, 12, 64, None),                               # 16-channel slabs
    (128, 2, 24, 24, None),                              # 16-pixel-wide tiles: a warp's rows span several image rows
    (384, 1, 16, 16, {"flags": 16}),                     # ... with two N tiles
    (256, 2, 32, 128, {"flags": 16}),                    # 256 channels as two 128-channel N tiles, accumulators double-buffered
]


@pytest.mark.parametrize("c,n,h,w,tune", RES_CASES)
@pytest.mark.parametrize("in_place", [False, True], ids=["res", "res_is_dst"])
def test_conv_residual_forms(c, n, h, w, tune, in_place):
    """relu(bn(conv(x)) + r) with r a separate map and with r == dst (ResidualBlock.conv2 in the engine), 1e-2 relative."""
    ops = _ops()
    x = _rand_fm(n, c, h, w, 90)
    res = _rand_fm(n, c, h, w, 91)
    wt = _rand_w((c, c, 3, 3), 92, 9 * c)
    bn = _bn(c, 93)
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU)
    spec.w_fold = None                                   # the tap-by-tap kernel is the one under test
    r = ops.nchw_to_nhwc(res)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(x), residual=r, dst=r if in_place else None, tune=tune)
    ref = F.relu(_bn_ref(F.conv2d(x, wt, padding=1), bn) + res)
    _close(ops.nhwc_to_nchw(y), ref)


@pytest.mark.parametrize("act", ["tanh", "sigmoid", "none"])
def test_conv_activations(act):
    ops = _ops()
    x = _rand_fm(1, 32, 16, 32, 7)
    wt = _rand_w((32, 32, 3, 3), 8, 288)
    code = {"tanh": ops.ACT_TANH, "sigmoid": ops.ACT_SIGMOID, "none": ops.ACT_NONE}[act]
    fn = {"tanh": torch.tanh, "sigmoid": torch.sigmoid, "none": lambda t: t}[act]
    y = ops.conv2d(ops.ConvSpec.from_conv(wt, act=code), ops.nchw_to_nhwc(x))
    _close(ops.nhwc_to_nchw(y), fn(F.conv2d(x, wt, padding=1)))


@pytest.mark.parametrize("tune", [None, {"flags": 16}], ids=["auto", "pair"])
@pytest.mark.parametrize("cin,cout,k,pad,n,h,w", [(64, 128, 4, 1, 1, 32, 32), (96, 192, 4, 1, 2, 64, 64),
                                                  (64, 128, 3, 1, 1, 32, 64), (64, 128, 1, 0, 1, 32, 32)])
def test_conv_s2(cin, cout, k, pad, n, h, w, tune):
    ops = _ops()
    x = _rand_fm(n, cin, h, w, 9)
    wt = _rand_w((cout, cin, k, k), 10, cin * k * k)
    bn = _bn(cout, 11)
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU, stride=2, pad=pad)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(x), tune=tune)
    ref = F.relu(_bn_ref(F.conv2d(x, wt, stride=2, padding=pad), bn))
    assert y.shape[1:3] == ref.shape[2:]
    _close(ops.nhwc_to_nchw(y, cout), ref)


@pytest.mark.parametrize("tune", [None, {"flags": 16}], ids=["auto", "pair"])
@pytest.mark.parametrize("cin,cout,n,h,w", [(128, 64, 1, 16, 16), (384, 192, 1, 16, 32), (256, 64, 2, 32, 32)])
def test_conv_transpose(cin, cout, n, h, w, tune):
    ops = _ops()
    x = _rand_fm(n, cin, h, w, 12)
    wt = _rand_w((cin, cout, 4, 4), 13, cin * 4)
    bias = torch.randn(cout, device="cuda") * 0.1
    bn = _bn(cout, 14)
    spec = ops.ConvSpec.from_convT(wt, bias=bias, bn=bn, act=ops.ACT_RELU)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(x), tune=tune)
    ref = F.relu(_bn_ref(F.conv_transpose2d(x, wt, bias, stride=2, padding=1), bn))
    assert tuple(y.shape[1:3]) == (2 * h, 2 * w)
    _close(ops.nhwc_to_nchw(y, cout), ref)


@pytest.mark.parametrize("tune", [None, {"flags": 16}], ids=["auto", "pair"])
@pytest.mark.parametrize("c0,c1,cout", [(64, 64, 64), (96, 96, 96), (192, 192, 96)])
def test_conv_concat_sources(c0, c1, cout, tune):
    ops = _ops()
    a = _rand_fm(1, c0, 32, 32, 15)
    b = _rand_fm(1, c1, 32, 32, 16)
    wt = _rand_w((cout, c0 + c1, 3, 3), 17, (c0 + c1) * 9)
    spec = ops.ConvSpec.from_conv(wt, act=ops.ACT_RELU)
    y = ops.conv2d(spec, ops.nchw_to_nhwc(a), ops.nchw_to_nhwc(b), tune=tune)
    ref = F.relu(F.conv2d(torch.cat([a, b], 1), wt, padding=1))
    _close(ops.nhwc_to_nchw(y, cout), ref)


def test_conv_into_channel_slice():
    ops = _ops()
    x = _rand_fm(1, 64, 16, 16, 18)
    wt = _rand_w((32, 64, 3, 3), 19, 576)
    dst = torch.zeros((1, 16, 16, 96), dtype=torch.bfloat16, device="cuda")
    ops.conv2d(ops.ConvSpec.from_conv(wt), ops.nchw_to_nhwc(x), dst=dst, dst_c_off=64)
    ref = F.conv2d(x, wt, padding=1)
    _close(ops.nhwc_to_nchw(dst)[:, 64:], ref)
    assert dst[..., :64].float().abs().max().item() == 0.0


@pytest.mark.parametrize("k,cout,kp", [(7, 64, 32), (3, 32, 16), (7, 96, 32), (3, 16, 16)])
def test_stem(k, cout, kp):
    ops = _ops()
    g = torch.Generator().manual_seed(20)
    x = torch.rand(2, 3, 32, 64, generator=g).cuda()
    wt = _rand_w((cout, 3, k, k), 21, 3 * k * k)
    bn = _bn(cout, 22)
    spec = ops.ConvSpec.from_stem(wt, kp, bn=bn, act=ops.ACT_RELU)
    packed = ops.stem_pack(x, k, k // 2, kp)
    y = ops.conv2d(spec, packed)
    xr = x.to(torch.bfloat16).float()
    ref = F.relu(_bn_ref(F.conv2d(xr, wt, padding=k // 2), bn))
    _close(ops.nhwc_to_nchw(y, cout), ref)


def test_dot_epilogue():
    ops = _ops()
    x = _rand_fm(2, 16, 32, 64, 23)
    wt = _rand_w((16, 16, 3, 3), 24, 144)
    bn = _bn(16, 25)
    dw = torch.randn(16, device="cuda")
    out = torch.empty((2, 32, 64), dtype=torch.float32, device="cuda")
    ops.conv2d(ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU), ops.nchw_to_nhwc(x), epi=ops.EPI_DOT, dot=(dw, 0.25, out))
    feat = F.relu(_bn_ref(F.conv2d(x, wt, padding=1), bn))
    ref = torch.sigmoid((feat * dw.view(1, -1, 1, 1)).sum(1) + 0.25)
    _close(out, ref, rel=2e-3, abs_=1e-4)


@pytest.mark.parametrize("mode", ["blend", "residual", "guided"])
def test_image_epilogue(mode):
    ops = _ops()
    g = torch.Generator().manual_seed(26)
    xb = torch.rand(5, 3, 32, 64, generator=g).cuda()
    index = torch.tensor([4, 1, 3], dtype=torch.int32, device="cuda")
    feat = _rand_fm(3, 32, 32, 64, 27)
    wt = _rand_w((3, 32, 3, 3), 28, 288)
    bias = torch.randn(3, device="cuda") * 0.1
    out = torch.zeros_like(xb)
    guid = torch.rand(3, 32, 64, device="cuda")
    alpha = torch.tensor(0.1, device="cuda")
    if mode == "blend":
        spec = ops.ConvSpec.from_conv(wt, bias=bias, act=ops.ACT_SIGMOID)
        img = dict(mode=ops.IMG_BLEND, x=xb, out=out, index=index, alpha=alpha)
    elif mode == "residual":
        spec = ops.ConvSpec.from_conv(wt, bias=bias, act=ops.ACT_TANH)
        img = dict(mode=ops.IMG_RESIDUAL, x=xb, out=out, index=index)
    else:
        spec = ops.ConvSpec.from_conv(wt, bias=bias, act=ops.ACT_TANH)
        img = dict(mode=ops.IMG_GUIDED, x=xb, out=out, index=index, guidance=guid)
    ops.conv2d(spec, ops.nchw_to_nhwc(feat), epi=ops.EPI_IMAGE, image=img)
    v = F.conv2d(feat, wt, bias, padding=1)
    xs = xb[index.long()]
    if mode == "blend":
        ref = 0.9 * xs + 0.1 * torch.sigmoid(v)
    elif mode == "residual":
        ref = torch.clamp(xs + torch.tanh(v), 0, 1)
    else:
        ref = torch.clamp(xs + torch.tanh(v) * guid.unsqueeze(1), 0, 1)
    _close(out[index.long()], ref, rel=2e-3, abs_=2e-4)
    untouched = [i for i in range(5) if i not in (4, 1, 3)]
    assert out[untouched].abs().max().item() == 0.0


@pytest.mark.parametrize("tune", [None, {"flags": 16}], ids=["auto", "pair"])
def test_dynamic_bucket_count(tune):
    """n_dev < n: images beyond the live count are not touched (routed bucket without a host round-trip)."""
    ops = _ops()
    x = _rand_fm(4, 64, 16, 32, 29)
    wt = _rand_w((64, 64, 3, 3), 30, 576)
    n_dev = torch.tensor([5], dtype=torch.int32, device="cuda")   # bucket holds 5 images, this launch covers [3, 7)
    dst = torch.full((4, 16, 32, 64), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.conv2d(ops.ConvSpec.from_conv(wt), ops.nchw_to_nhwc(x), dst=dst, n_dev=n_dev, n_start=3, tune=tune)
    ref = F.conv2d(x, wt, padding=1)
    _close(ops.nhwc_to_nchw(dst)[:2], ref[:2])
    assert (dst[2:].float() == 7.0).all()


@pytest.mark.parametrize("cin,pitch,cout,h,w,n", [(64, 64, 128, 16, 32, 2), (96, 256, 128, 24, 40, 2), (224, 256, 128, 16, 16, 3),
                                                   (512, 1024, 256, 8, 16, 2)])
def test_conv1x1_fused_pre_activation(cin, pitch, cout, h, w, n):
    """relu(x*s + b) fused into the A operand of a 1x1 conv (DenseNet norm1/relu1 -> conv1), reading a channel prefix of a
    wider buffer.  Reference: torch fp32 on the bf16-rounded operands; the fused path rounds the activated input to bf16
    exactly like the unfused adb_affine_relu pass did."""
    ops = _ops()
    x = _rand_fm(n, pitch, h, w, 31)
    wt = _rand_w((cout, cin, 1, 1), 32, cin)
    g = torch.Generator().manual_seed(33)
    s = (torch.rand(cin, generator=g) + 0.5).cuda()
    b = (torch.randn(cin, generator=g) * 0.3).cuda()
    bn = _bn(cout, 34)
    act_in = F.relu(x[:, :cin] * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)).to(torch.bfloat16).float()
    ref = F.relu(_bn_ref(F.conv2d(act_in, wt), bn))
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU, pad=0)
    got = ops.conv2d(spec, ops.nchw_to_nhwc(x, pitch), c0=cin, pre=(s.contiguous(), b.contiguous()))
    _close(ops.nhwc_to_nchw(got, cout), ref)


# ---------------------------------------------------------------- rolling-row kernel (csrc/conv_roll.cu)
ROLL_CASES = [
    # (c0, c1, cout, n, h, w, residual, tune)
    (32, 0, 32, 2, 40, 128, False, None),             # Light 32->32: 64-byte operand rows, ring of 16 wraps twice
    (32, 0, 32, 1, 70, 256, True, None),              # + residual, two strips, two segments (SH = 64 then 6 rows)
    (64, 0, 64, 2, 33, 128, True, None),              # Medium 64->64: ring of 8, odd height
    (128, 0, 32, 2, 24, 128, False, None),            # DenseNet conv2: two 64-channel chunks
    (64, 64, 64, 1, 20, 256, False, None),            # channel concat of two sources
    (96, 0, 32, 1, 18, 192, False, None),             # ragged K chunk (96 = 64 + 32) and a ragged last strip (192 = 128 + 64)
    (32, 0, 32, 3, 130, 136, True, {"mt": 2}),        # 16-row segments: many segment boundaries, 8-pixel last strip
    (16, 0, 32, 1, 9, 128, False, None),              # 32-byte operand rows
    (32, 0, 32, 1, 1, 128, False, None),              # a single image row
    (64, 0, 64, 1, 2, 512, True, {"stages": 2}),
    (32, 0, 32, 24, 64, 128, True, None),             # several segments per CTA: slot phases carry across segments
    (32, 0, 32, 2, 40, 128, False, {"acc": 1}),       # one input row per operand box (default: 4 for a ring of 16)
    (32, 0, 32, 2, 43, 256, True, {"acc": 2}),        # two rows per box, odd row count: a short last group
    (64, 0, 64, 2, 33, 128, True, {"acc": 1}),
    (128, 0, 32, 1, 67, 128, False, {"mt": 1}),       # 8-row segments with 4-row boxes: groups cut by segment ends
]


@pytest.mark.parametrize("c0,c1,cout,n,h,w,residual,tune", ROLL_CASES)
def test_conv_roll(c0, c1, cout, n, h, w, residual, tune):
    """3x3 convs with 3*cout <= 256 and W >= 128 take the rolling-row kernel; same tolerance as the tap-by-tap kernel, and
    the two kernels agree with each other to bf16 output rounding."""
    ops = _ops()
    a = _rand_fm(n, c0, h, w, 40)
    b = _rand_fm(n, c1, h, w, 41) if c1 else None
    wt = _rand_w((cout, c0 + c1, 3, 3), 42, (c0 + c1) * 9)
    bn = _bn(cout, 43)
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU)
    assert spec.w_fold is not None
    res = _rand_fm(n, cout, h, w, 44) if residual else None
    kw = dict(residual=ops.nchw_to_nhwc(res) if residual else None)
    an, bn_ = ops.nchw_to_nhwc(a), (ops.nchw_to_nhwc(b) if c1 else None)
    y = ops.conv2d(spec, an, bn_, tune=tune, **kw)
    x = torch.cat([a, b], 1) if c1 else a
    ref = _bn_ref(F.conv2d(x, wt, padding=1), bn)
    ref = F.relu(ref + res if residual else ref)
    _close(ops.nhwc_to_nchw(y, cout), ref)
    y_old = ops.conv2d(spec, an, bn_, tune={"flags": 512}, **kw)
    _close(ops.nhwc_to_nchw(y, cout), ops.nhwc_to_nchw(y_old, cout), rel=8e-3, abs_=1e-3)


def test_conv_roll_channel_slice_in_place_and_bucket_count():
    """DenseNet use: the 3x3 conv stores its 32 channels into a slice of the block buffer; ResidualBlock use: dst == residual
    (in place); routed buckets: images beyond the live count stay untouched."""
    ops = _ops()
    x = _rand_fm(3, 128, 16, 128, 45)
    wt = _rand_w((32, 128, 3, 3), 46, 128 * 9)
    dst = torch.zeros((3, 16, 128, 96), dtype=torch.bfloat16, device="cuda")
    n_dev = torch.tensor([6], dtype=torch.int32, device="cuda")     # bucket of 6, this launch covers [4, 7): 2 live images
    ops.conv2d(ops.ConvSpec.from_conv(wt), ops.nchw_to_nhwc(x), dst=dst, dst_c_off=32, n_dev=n_dev, n_start=4)
    ref = F.conv2d(x, wt, padding=1)
    _close(ops.nhwc_to_nchw(dst)[:2, 32:64], ref[:2])
    assert dst[..., :32].float().abs().max().item() == 0.0 and dst[..., 64:].float().abs().max().item() == 0.0
    assert dst[2].float().abs().max().item() == 0.0
    f = _rand_fm(2, 32, 48, 256, 47)
    w2 = _rand_w((32, 32, 3, 3), 48, 288)
    fn = ops.nchw_to_nhwc(f)
    ops.conv2d(ops.ConvSpec.from_conv(w2, act=ops.ACT_RELU), fn.clone(), dst=fn, residual=fn)
    _close(ops.nhwc_to_nchw(fn), F.relu(F.conv2d(f, w2, padding=1) + f))


@pytest.mark.parametrize("mode", ["blend", "residual", "guided"])
@pytest.mark.parametrize("cin,h,w", [(32, 24, 160), (48, 9, 128)])
def test_conv_roll_image_head(mode, cin, h, w):
    """The 3-channel output heads on the rolling-row schedule (N = 3 x 16): same arithmetic and tolerance as test_image_epilogue,
    and the same result as the tap-by-tap kernel."""
    ops = _ops()
    g = torch.Generator().manual_seed(50)
    xb = torch.rand(5, 3, h, w, generator=g).cuda()
    index = torch.tensor([4, 1, 3], dtype=torch.int32, device="cuda")
    feat = _rand_fm(3, cin, h, w, 51)
    wt = _rand_w((3, cin, 3, 3), 52, cin * 9)
    bias = torch.randn(3, device="cuda") * 0.1
    guid = torch.rand(3, h, w, device="cuda")
    alpha = torch.tensor(0.1, device="cuda")
    act = ops.ACT_SIGMOID if mode == "blend" else ops.ACT_TANH
    spec = ops.ConvSpec.from_conv(wt, bias=bias, act=act)
    assert spec.w_fold is not None

    def run(tune):
        out = torch.zeros_like(xb)
        img = {"blend": dict(mode=ops.IMG_BLEND, x=xb, out=out, index=index, alpha=alpha),
               "residual": dict(mode=ops.IMG_RESIDUAL, x=xb, out=out, index=index),
               "guided": dict(mode=ops.IMG_GUIDED, x=xb, out=out, index=index, guidance=guid)}[mode]
        ops.conv2d(spec, ops.nchw_to_nhwc(feat), epi=ops.EPI_IMAGE, image=img, tune=tune)
        return out
    out = run(None)
    v = F.conv2d(feat, wt, bias, padding=1)
    xs = xb[index.long()]
    ref = {"blend": 0.9 * xs + 0.1 * torch.sigmoid(v), "residual": torch.clamp(xs + torch.tanh(v), 0, 1),
           "guided": torch.clamp(xs + torch.tanh(v) * guid.unsqueeze(1), 0, 1)}[mode]
    _close(out[index.long()], ref, rel=2e-3, abs_=2e-4)
    assert out[[0, 2]].abs().max().item() == 0.0
    _close(out, run({"flags": 512}), rel=1e-3, abs_=1e-4)


@pytest.mark.parametrize("c", [16, 32])
def test_conv_roll_dot_head(c):
    ops = _ops()
    x = _rand_fm(2, c, 20, 256, 53)
    wt = _rand_w((c, c, 3, 3), 54, c * 9)
    bn = _bn(c, 55)
    dw = torch.randn(c, device="cuda")
    out = torch.empty((2, 20, 256), dtype=torch.float32, device="cuda")
    spec = ops.ConvSpec.from_conv(wt, bn=bn, act=ops.ACT_RELU)
    ops.conv2d(spec, ops.nchw_to_nhwc(x), epi=ops.EPI_DOT, dot=(dw, 0.25, out))
    feat = F.relu(_bn_ref(F.conv2d(x, wt, padding=1), bn))
    ref = torch.sigmoid((feat * dw.view(1, -1, 1, 1)).sum(1) + 0.25)
    _close(out, ref, rel=2e-3, abs_=1e-4)
    old = torch.empty_like(out)
    ops.conv2d(spec, ops.nchw_to_nhwc(x), epi=ops.EPI_DOT, dot=(dw, 0.25, old), tune={"flags": 512})
    _close(out, old, rel=1e-3, abs_=1e-4)


@pytest.mark.parametrize("cout,h,w,n", [(64, 64, 256, 2), (64, 32, 96, 1)])
def test_stem_7x7_stride2_space_to_depth(cout, h, w, n):
    """HDEN stem (torchvision conv1 / conv0 + BatchNorm + ReLU): space-to-depth operand (adb_stem_pack kh = kw = 2, stride 2) +
    the 4x4-tap conv kind ADB_CONV_K4_S2D against F.conv2d(7x7, stride 2, pad 3) on the bf16-rounded image."""
    ops = _ops()
    g = torch.Generator().manual_seed(60)
    x = torch.rand(n, 3, h, w, generator=g).cuda()
    wt = _rand_w((cout, 3, 7, 7), 61, 147)
    bn = _bn(cout, 62)
    spec = ops.ConvSpec.from_stem_s2d(wt, bn=bn, act=ops.ACT_RELU)
    cols = ops.stem_pack(x, 2, 0, 16, stride=2, kh=2)
    assert cols.shape == (n, h // 2, w // 2, 16)
    y = ops.conv2d(spec, cols)
    ref = F.relu(_bn_ref(F.conv2d(x.to(torch.bfloat16).float(), wt, stride=2, padding=3), bn))
    _close(ops.nhwc_to_nchw(y, cout), ref)


# ----------------------------------------------------------------------------- epilogue channel statistics (adb_conv_desc.stat_out)
STAT_CASES = [
    # (c0, c1, cout, k, stride, n, h, w, tune)
    (128, 0, 128, 3, 1, 2, 32, 64, None),
    (128, 0, 128, 3, 1, 2, 32, 64, {"flags": 16}),       # CTA pair: both CTAs own slots
    (128, 0, 128, 3, 1, 1, 24, 40, None),                # ragged tiles: rows outside the image count nothing
    (96, 96, 96, 3, 1, 2, 16, 48, None),                 # two sources, 32-channel slabs
    (64, 0, 128, 4, 2, 2, 64, 64, None),                 # stride-2 encoder conv
    (256, 0, 256, 3, 1, 1, 16, 16, None),
    (192, 0, 48, 1, 1, 2, 16, 32, None),                 # 1x1, 16-channel slabs, cout 48 of 48
    (64, 0, 64, 3, 1, 2, 64, 96, {"mt": 2}),             # several sub-tiles per CTA tile (w < 128: not the rolling-row kernel)
    (512, 0, 512, 3, 1, 2, 8, 16, None),                 # two N tiles share the pixel tile's slots
]


@pytest.mark.parametrize("c0,c1,cout,k,stride,n,h,w,tune", STAT_CASES)
def test_conv_epilogue_statistics(c0, c1, cout, k, stride, n, h, w, tune):
    """(sum, sum of squares) per channel folded from the epilogue partials == the same sums over the stored bf16 output
    (fp64 on the host), to fp32 rounding of the per-slot partials: 1e-5 relative."""
    ops = _ops()
    a = _rand_fm(n, c0, h, w, 70)
    b = _rand_fm(n, c1, h, w, 71) if c1 else None
    wt = _rand_w((cout, c0 + c1, k, k), 72, (c0 + c1) * k * k)
    bias = torch.randn(cout, device="cuda") * 0.1
    spec = ops.ConvSpec.from_conv(wt, bias=bias, stride=stride, pad=1 if k > 1 else 0)
    y, part = ops.conv2d(spec, ops.nchw_to_nhwc(a), None if b is None else ops.nchw_to_nhwc(b), tune=tune, stats=True)
    assert part is not None and part.shape[0] == n and part.shape[2:] == (2, spec.cout_pad)
    y_plain = ops.conv2d(spec, ops.nchw_to_nhwc(a), None if b is None else ops.nchw_to_nhwc(b), tune=tune)
    assert torch.equal(y, y_plain)                       # the statistics do not disturb the output
    yd = y.double().reshape(-1, spec.cout_pad)
    s = part.double().sum((0, 1))
    ref0, ref1 = yd.sum(0), (yd * yd).sum(0)
    assert (s[0] - ref0).abs().max().item() <= 1e-5 * yd.abs().sum(0).max().item() + 1e-6
    assert (s[1] - ref1).abs().max().item() <= 1e-5 * ref1.max().item() + 1e-6


def test_conv_epilogue_statistics_not_offered_by_rolling_row_launch():
    ops = _ops()
    x = _rand_fm(1, 64, 16, 256, 73)
    wt = _rand_w((64, 64, 3, 3), 74, 576)
    spec = ops.ConvSpec.from_conv(wt)
    assert spec.w_fold is not None
    y, part = ops.conv2d(spec, ops.nchw_to_nhwc(x), stats=True)
    assert part is None
    _close(ops.nhwc_to_nchw(y), F.conv2d(x, wt, padding=1))


@pytest.mark.parametrize("c,n,h,w", [(128, 4, 32, 32), (256, 2, 16, 24), (48, 2, 40, 40)])
def test_bn_finalize_from_epilogue_statistics(c, n, h, w):
    """adb_bn_finalize_stats(partials) == adb_bn_train_stats(z): mean / rstd / scale / shift and the running statistics
    (both fp64 folds of fp32 partials of the same bf16 values; 2e-5 relative)."""
    import ctypes as C
    from adam_dehaze_b200 import _lib
    ops = _ops()
    x = _rand_fm(n, 64, h, w, 75)
    wt = _rand_w((c, 64, 3, 3), 76, 576) * 3.0
    bias = torch.randn(c, device="cuda")
    spec = ops.ConvSpec.from_conv(wt, bias=bias)
    spec.w_fold = None
    z, part = ops.conv2d(spec, ops.nchw_to_nhwc(x), stats=True)
    assert part is not None
    cp = spec.cout_pad
    px = n * h * w
    g = torch.rand(cp, device="cuda") + 0.5
    bt = torch.randn(cp, device="cuda")
    outs = []
    for use_part in (False, True):
        rm, rv = torch.zeros(cp, device="cuda"), torch.ones(cp, device="cuda")
        nbt = torch.zeros((), dtype=torch.int64, device="cuda")
        stats = torch.empty(4 * cp, device="cuda")
        ptrs = [_lib.ptr(stats[i * cp:(i + 1) * cp]) for i in range(4)]
        if use_part:
            slots = part.shape[0] * part.shape[1]
            scratch = torch.empty(int(_lib.load().adb_bn_stat_scratch_floats(slots, cp)), device="cuda")
            _lib.call("adb_bn_finalize_stats", _lib.ptr(part), slots, cp, px, cp, _lib.ptr(g), _lib.ptr(bt), 1e-5, 0.1,
                      _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(nbt), _lib.ptr(scratch), *ptrs, _lib.current_stream())
        else:
            scratch = torch.empty(int(_lib.load().adb_bn_scratch_floats(px, cp)), device="cuda")
            _lib.call("adb_bn_train_stats", _lib.ptr(z), px, cp, cp, _lib.ptr(g), _lib.ptr(bt), 1e-5, 0.1,
                      _lib.ptr(rm), _lib.ptr(rv), _lib.ptr(nbt), _lib.ptr(scratch), *ptrs, _lib.current_stream())
        outs.append((stats.clone(), rm, rv, int(nbt.item())))
    (s0, rm0, rv0, k0), (s1, rm1, rv1, k1) = outs
    assert k0 == k1 == 1
    for u, v in ((s0, s1), (rm0, rm1), (rv0, rv1)):
        assert (u - v).abs().max().item() <= 2e-5 * u.abs().max().item() + 1e-6
    # and against torch's batch statistics of the stored z
    zf = z.float().reshape(-1, cp)
    assert (s1[:cp] - zf.mean(0)).abs().max().item() <= 1e-4 * zf.abs().max().item()
    assert (s1[cp:2 * cp] - 1.0 / torch.sqrt(zf.var(0, unbiased=False) + 1e-5)).abs().max().item() <= 1e-3 * s1[cp:2 * cp].abs().max().item()


@pytest.mark.parametrize("c,n,live,h,w", [(96, 3, 3, 64, 128), (192, 4, 2, 32, 64), (384, 2, 2, 16, 32)])
def test_attention_pool_from_epilogue_partials(c, n, live, h, w):
    """AttentionBlock's (avg, max) pool folded from the producing conv's (sum, max) partials == adb_attn_pool over the stored
    map: max bit-exact, sums to fp32 reassociation (1e-5 relative); and the attention output built on it stays within the
    one bf16 ulp of the one built on the pool pass.  `live` < n: the device-side image count of a routed bucket."""
    from adam_dehaze_b200 import _lib
    ops = _ops()
    x = _rand_fm(n, c, h, w, 80)
    res = _rand_fm(n, c, h, w, 81)
    wt = _rand_w((c, c, 3, 3), 82, 9 * c)
    spec = ops.ConvSpec.from_conv(wt, bn=_bn(c, 83), act=ops.ACT_RELU)
    n_dev = torch.tensor([live], dtype=torch.int32, device="cuda")
    f = ops.nchw_to_nhwc(res).clone()
    f, part = ops.conv2d(spec, ops.nchw_to_nhwc(x), dst=f, residual=f, stats="pool", n_dev=n_dev)
    assert part is not None and part.shape[0] == n
    pool_ref = torch.empty(ops.pool_scratch_floats(n, h, w, c), device="cuda")
    _lib.call("adb_attn_pool", _lib.ptr(f), n, h, w, c, _lib.ptr(n_dev), 0, _lib.ptr(pool_ref), _lib.current_stream())
    pool = torch.full((n * 2 * c,), float("nan"), device="cuda")
    scratch = torch.empty(int(_lib.load().adb_attn_pool_stat_scratch_floats(n, part.shape[1], c)), device="cuda")
    _lib.call("adb_attn_pool_from_stats", _lib.ptr(part), n, part.shape[1], part.shape[3], c, _lib.ptr(n_dev), 0, _lib.ptr(scratch),
              _lib.ptr(pool), _lib.current_stream())
    a = pool.view(n, 2, c)[:live]
    b = pool_ref[:n * 2 * c].view(n, 2, c)[:live]
    assert torch.equal(a[:, 1], b[:, 1])
    assert (a[:, 0] - b[:, 0]).abs().max().item() <= 1e-5 * f[:live].float().abs().sum((1, 2)).max().item()
    g = torch.Generator(device="cpu").manual_seed(84)
    ap = ops.AttnParams((torch.randn(c // 16, c, 1, 1, generator=g) * 0.1).cuda(), (torch.randn(c, c // 16, 1, 1, generator=g) * 0.1).cuda(),
                        (torch.randn(1, 2, 7, 7, generator=g) * 0.1).cuda())
    y0 = ops.attention(f, ap, n=n, n_dev=n_dev)
    y1 = ops.attention(f, ap, n=n, n_dev=n_dev, pool_partials=part)
    _close(y1[:live].float(), y0[:live].float(), rel=1e-2, abs_=1e-4)     # one bf16 ulp of the output
