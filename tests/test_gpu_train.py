"""GPU parity of the training-step kernels: conv weight gradients (adb_wgrad) and data gradients
(adb_conv2d with transformed weights) against torch autograd in fp32 on the same bf16-rounded operands.

Tolerance (stated per test): operands are exactly representable in bf16 and products accumulate in fp32, so the only
differences are summation order: |err| <= 2e-3 * max|ref| + 1e-4 for weight gradients."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from adam_dehaze_b200 import ops
    return ops


def _fm(n, c, h, w, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).float().cuda()


def _close(out, ref, rel, abs_):
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


def _wgrad_ref(x, dz, k, stride, pad):
    w = torch.zeros(dz.shape[1], x.shape[1], k, k, device=x.device, requires_grad=True)
    y = F.conv2d(x, w, stride=stride, padding=pad)
    (g,) = torch.autograd.grad(y, w, dz)
    return g


@pytest.mark.parametrize("cin,cout,k,h,w,n", [
    (64, 64, 3, 32, 32, 2),       # one 64-channel group each side
    (16, 16, 3, 16, 48, 1),       # boxes wider than the channel count (TMA zero fill), ragged width
    (32, 48, 3, 24, 40, 3),       # ragged tiles
    (96, 96, 3, 16, 128, 1),      # two groups, second partly out of range
    (192, 128, 3, 16, 32, 2),     # several channel chunks
    (128, 272, 3, 16, 32, 1),     # three M tiles
    (64, 32, 1, 16, 32, 2),       # 1x1
])
def test_wgrad_s1(cin, cout, k, h, w, n):
    ops = _ops()
    x, dz = _fm(n, cin, h, w, 1), _fm(n, cout, h, w, 2)
    ref = _wgrad_ref(x, dz, k, 1, k // 2)
    got = ops.wgrad(ops.nchw_to_nhwc(dz, cout), ops.nchw_to_nhwc(x, cin), kh=k, kw=k, pad=k // 2)
    _close(got, ref, 2e-3, 1e-4)


def test_wgrad_concat_sources_and_true_rows():
    ops = _ops()
    n, h, w = 2, 16, 64
    xa, xb, dz = _fm(n, 64, h, w, 3), _fm(n, 32, h, w, 4), _fm(n, 3, h, w, 5)
    ref = _wgrad_ref(torch.cat([xa, xb], 1), dz, 3, 1, 1)
    got = ops.wgrad(ops.nchw_to_nhwc(dz, 16), ops.nchw_to_nhwc(xa, 64), ops.nchw_to_nhwc(xb, 32), cs_true=3)
    assert got.shape == (3, 96, 3, 3)
    _close(got, ref, 2e-3, 1e-4)


@pytest.mark.parametrize("cin,cout,h,w,n", [(64, 128, 32, 64, 2), (96, 192, 32, 32, 1)])
def test_wgrad_s2_4x4(cin, cout, h, w, n):
    ops = _ops()
    x, dz = _fm(n, cin, h, w, 6), _fm(n, cout, h // 2, w // 2, 7)
    ref = _wgrad_ref(x, dz, 4, 2, 1)
    got = ops.wgrad(ops.nchw_to_nhwc(dz, cout), ops.nchw_to_nhwc(x, cin), kind=ops.CONV_S2, kh=4, kw=4, pad=1)
    _close(got, ref, 2e-3, 1e-4)


def test_wgrad_conv_transpose():
    """ConvTranspose2d(4,2,1) weight gradient = the stride-2 form with the maps swapped."""
    ops = _ops()
    n, ci, co, h, w = 2, 128, 64, 16, 32
    x, dy = _fm(n, ci, h, w, 8), _fm(n, co, 2 * h, 2 * w, 9)
    wt = torch.zeros(ci, co, 4, 4, device="cuda", requires_grad=True)
    y = F.conv_transpose2d(x, wt, stride=2, padding=1)
    (ref,) = torch.autograd.grad(y, wt, dy)
    got = ops.wgrad(ops.nchw_to_nhwc(x, ci), ops.nchw_to_nhwc(dy, co), kind=ops.CONV_S2, kh=4, kw=4, pad=1)
    _close(got, ref, 2e-3, 1e-4)


@pytest.mark.parametrize("k,kp", [(3, 16), (7, 32)])
def test_wgrad_stem(k, kp):
    ops = _ops()
    n, h, w, co = 2, 16, 64, 32
    g = torch.Generator().manual_seed(10)
    x = torch.rand(n, 3, h, w, generator=g).cuda()
    dz = _fm(n, co, h, w, 11)
    cols = ops.stem_pack(x, k, k // 2, kp)                       # bf16-rounds x: compare against the rounded image
    ref = _wgrad_ref(x.to(torch.bfloat16).float(), dz, k, 1, k // 2)
    got = ops.wgrad(ops.nchw_to_nhwc(dz, co), cols, kh=k, kw=1, pad=k // 2, layout=ops.WG_STEM, stem_kw=k)
    assert got.shape == (co, 3, k, k)
    _close(got, ref, 2e-3, 1e-4)


def test_wgrad_accumulate():
    ops = _ops()
    x, dz = _fm(1, 64, 16, 32, 12), _fm(1, 64, 16, 32, 13)
    ref = _wgrad_ref(x, dz, 3, 1, 1)
    xs, dzs = ops.nchw_to_nhwc(x, 64), ops.nchw_to_nhwc(dz, 64)
    out = ops.wgrad(dzs, xs)
    ops.wgrad(dzs, xs, out=out, accumulate=True)
    _close(out, 2 * ref, 2e-3, 1e-4)
