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
    torch.set_grad_enabled(True)      # (tests/test_oracle_golden.py switches autograd off at import time)
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


@pytest.mark.parametrize("mode", [1, 2], ids=["channel_major", "tap_packed"])
@pytest.mark.parametrize("cin,cout,k,h,w,n,kind", [
    (64, 64, 3, 32, 64, 2, 0), (32, 32, 3, 20, 48, 2, 0), (128, 64, 3, 16, 40, 1, 0), (96, 48, 3, 24, 32, 2, 0),
    (64, 16, 1, 16, 32, 2, 0), (64, 64, 4, 32, 64, 2, 1), (160, 32, 3, 8, 16, 3, 0), (64, 64, 3, 4, 16, 2, 0),
    (96, 96, 3, 16, 48, 2, 0), (64, 128, 4, 32, 64, 1, 1), (96, 192, 4, 16, 32, 2, 1), (128, 256, 1, 16, 32, 1, 0), (48, 80, 3, 16, 32, 1, 0),
    (64, 192, 3, 16, 32, 2, 0), (96, 160, 3, 16, 32, 1, 0), (32, 224, 3, 8, 16, 1, 0), (192, 192, 3, 8, 32, 2, 0),
])
def test_wgrad_both_kernels(cin, cout, k, h, w, n, kind, mode):
    """The channel-major kernel (M = dZ channels) and the tap-packed kernel (M = (tap, X channel) pairs, chosen when the
    gradient has <= 64 channels) give the same weight gradient."""
    ops = _ops()
    stride = 2 if kind == 1 else 1
    pad = 1 if k > 1 else 0
    x, dz = _fm(n, cin, h, w, 14), _fm(n, cout, h // stride, w // stride, 15)
    ref = _wgrad_ref(x, dz, k, stride, pad)
    got = ops.wgrad(ops.nchw_to_nhwc(dz, cout), ops.nchw_to_nhwc(x, cin), kind=ops.CONV_S2 if kind == 1 else ops.CONV_S1,
                    kh=k, kw=k, pad=pad, mode=mode)
    _close(got, ref, 2e-3, 1e-4)


@pytest.mark.parametrize("mode", [1, 2], ids=["channel_major", "tap_packed"])
def test_wgrad_gradient_is_trailing_slice_of_block_buffer(mode):
    """DenseNet conv2: dZ is the LAST 32 channels of the block-buffer gradient (autograd.conv_plain).  The 64-channel TMA
    box must stop at the slice (zero fill), neither reading the neighbour's values into the result nor running past the
    end of the allocation; the buffer sits at the very end of its own cudaMalloc block here."""
    ops = _ops()
    n, h, w, pitch, co, ci = 2, 8, 16, 256, 32, 128
    raw = torch.empty(n * h * w * pitch, dtype=torch.bfloat16, device="cuda")
    buf = raw.view(n, h, w, pitch)
    buf.copy_(torch.randn(n, h, w, pitch, generator=torch.Generator().manual_seed(31)).to(torch.bfloat16))
    x = _fm(n, ci, h, w, 32)
    dz = buf[..., pitch - co:]
    ref = _wgrad_ref(x, dz.permute(0, 3, 1, 2).float(), 3, 1, 1)
    got = ops.wgrad(dz, ops.nchw_to_nhwc(x, ci), kh=3, kw=3, pad=1, cs=co, mode=mode)
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


# --------------------------------------------------------------------------------------------------------------------
# Branch models in train() mode: forward (batch-statistics BatchNorm) + backward through an L1 loss, against the fp32
# oracle under torch autograd on the same weights and inputs.
#
# Tolerances.  north_star: loss values and gradients within 1e-2 relative.
#   * loss value: |loss - loss_ref| <= 1e-2 |loss_ref|                                              (asserted)
#   * every backward KERNEL is within 1e-2 of torch autograd on identical operands                   (unit tests below)
#   * END-TO-END gradients of a deep ReLU network cannot agree with fp32 to 1e-2 once activations are STORED in bf16
#     (north_star: bf16 with fp32 accumulate): a forward perturbation of relative size e flips a fraction ~e of the
#     ReLU / clamp / |.| masks, which moves a white-noise gradient by ~sqrt(2e) per layer (12 % at e = 1 %).  The
#     fp32 oracle itself, run with bf16-rounded conv/BN outputs (helpers.oracle_bf16_storage), deviates from its own
#     fp32 result by 16 % (Light) to 60 % (Complex) on this random-init / random-target case.  The end-to-end assertion
#     is therefore relative to that floor: ||g - g_ref|| <= 1.25 * ||g_sim - g_ref|| + 0.02 ||g_ref|| over the whole
#     gradient, the same per parameter tensor with slack 2.0x + 0.05, and cos(g, g_ref) > 0 for every tensor whose
#     bf16-storage oracle gradient is itself within 70 % of the fp32 one.
def _train_case(name, n, h, w, seed=5):
    from helpers import make_branch, oracle_bf16_storage, rand_image   # (puts oracle/ on sys.path)
    import adam_oracle as oracle
    from adam_dehaze_b200.training.loss import DehazingLoss
    m = make_branch(name).cuda().train()
    x = rand_image(n, h, w, seed).cuda()
    tgt = rand_image(n, h, w, seed + 1).cuda()
    sd = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
    fwd = {"low": oracle.light_forward, "medium": oracle.medium_forward, "high": oracle.complex_forward,
           "low_unet": oracle.low_unet_forward, "corun": oracle.corun_forward, "dual_branch": oracle.dual_branch_forward}[name]
    names = [k for k, _ in m.named_parameters()]

    def grads_of(out):
        loss = (out - tgt).abs().mean()
        return loss.detach(), dict(zip(names, torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)))
    with oracle.train_mode():
        ref_out = fwd(sd, x)
        sim_out = oracle_bf16_storage(fwd, sd, x)
    ref_loss, ref_grads = grads_of(ref_out)
    _, sim_grads = grads_of(sim_out)

    crit = DehazingLoss(lambda_l1=1.0, lambda_content=0.0, lambda_perceptual=0.0)
    rm_before = {k: v.clone() for k, v in m.state_dict().items() if k.endswith("running_mean")}
    out = m(x)
    loss, parts = crit(out, tgt)
    loss.backward()
    torch.cuda.synchronize()
    return m, out, loss, ref_out.detach(), sim_out.detach(), ref_loss, ref_grads, sim_grads, rm_before


@pytest.mark.parametrize("name,n,h,w", [("low", 2, 32, 48), ("medium", 2, 64, 64), ("high", 2, 64, 64), ("low_unet", 2, 48, 64),
                                        ("corun", 2, 64, 96), ("dual_branch", 2, 64, 64)])
def test_branch_train_step_matches_oracle(name, n, h, w):
    from helpers import psnr
    m, out, loss, ref_out, sim_out, ref_loss, ref_grads, sim_grads, rm_before = _train_case(name, n, h, w)
    # forward: batch statistics amplify bf16 rounding at 16x16 bottlenecks; bound by the bf16-storage oracle's own error
    floor = (sim_out - ref_out).abs().max().item()
    assert (out - ref_out).abs().max().item() <= max(2e-2, 1.5 * floor)
    assert psnr(out.detach(), ref_out) >= min(45.0, psnr(sim_out, ref_out) - 2.0)   # 45 dB, or the bf16-storage floor
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    tot_err = tot_sim = tot_ref = 0.0
    bad = []
    for k, p in m.named_parameters():
        assert p.grad is not None, f"{k}: no gradient"
        g, r, s = p.grad.float(), ref_grads[k], sim_grads[k]
        assert g.shape == r.shape
        rn = r.norm().item()
        if rn < 1e-7:      # ConvTranspose biases ahead of a batch-statistics BatchNorm: true gradient 0, oracle holds noise
            assert g.abs().max().item() <= 1e-6
            continue
        e, es = (g - r).norm().item(), (s - r).norm().item()
        tot_err += e * e; tot_sim += es * es; tot_ref += rn * rn
        cos = (g * r).sum().item() / (g.norm().item() * rn + 1e-30)
        # (a tensor whose bf16-storage oracle is itself off by > 70 % is noise-dominated: its direction carries no information;
        #  one whose bf16-storage oracle is off by > 25 % — the attention MLPs at random init — gets 3x that floor instead of 2x:
        #  its error moves by a third with the fp32 reassociation of any upstream reduction)
        if e > (2.0 if es <= 0.25 * rn else 3.0) * es + 0.05 * rn or (cos <= 0 and es < 0.7 * rn):
            bad.append((k, e / rn, es / rn, cos))
    assert not bad, f"(name, ours/ref, bf16-oracle/ref, cos): {bad[:8]}"
    assert tot_err ** 0.5 <= 1.25 * tot_sim ** 0.5 + 0.02 * tot_ref ** 0.5, (tot_err ** 0.5 / tot_ref ** 0.5, tot_sim ** 0.5 / tot_ref ** 0.5)
    # running statistics were updated with momentum 0.1 (model.train() side effect)
    sd = m.state_dict()
    assert any((sd[k] - v).abs().max().item() > 0 for k, v in rm_before.items())
    assert int(sd[[k for k in sd if k.endswith("num_batches_tracked")][0]].item()) == 1


# --------------------------------------------------------------------------------------------------------------------
# Unit parity of the HBM-bound training kernels against torch autograd (fp32) on bf16-representable inputs.
def _nhwc(x, pitch=None):
    return _ops().nchw_to_nhwc(x, pitch or x.shape[1])


def _nchw(t, c):
    return _ops().nhwc_to_nchw(t, c)


def _ptr(t):
    from adam_dehaze_b200 import _lib
    return _lib.ptr(t)


def _call(name, *a):
    from adam_dehaze_b200 import _lib
    _lib.call(name, *a, _lib.current_stream())


@pytest.mark.parametrize("c,act,with_res", [(32, 1, False), (96, 1, True), (64, 0, False), (384, 1, True)])
def test_bn_train_forward_backward(c, act, with_res):
    from adam_dehaze_b200 import _lib
    n, h, w = 2, 12, 20
    z = (_fm(n, c, h, w, 20) * 1.5 + 0.3).to(torch.bfloat16).float()
    res = _fm(n, c, h, w, 21) if with_res else None
    dy = _fm(n, c, h, w, 22)
    g = torch.Generator().manual_seed(23)
    gamma = (torch.rand(c, generator=g) + 0.5).cuda().requires_grad_(True)
    beta = (torch.randn(c, generator=g) * 0.1).cuda().requires_grad_(True)
    rm, rv = torch.zeros(c).cuda(), torch.ones(c).cuda()
    zr = z.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if with_res else None
    yref = F.batch_norm(zr, rm.clone(), rv.clone(), gamma, beta, True, 0.1, 1e-5)
    if with_res:
        yref = yref + rr
    if act:
        yref = F.relu(yref)
    grads = torch.autograd.grad(yref, [zr, gamma, beta] + ([rr] if with_res else []), dy)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    F.batch_norm(z, rm_ref, rv_ref, None, None, True, 0.1, 1e-5)

    px = n * h * w
    zt = _nhwc(z)
    scratch = torch.empty(int(_lib.load().adb_bn_scratch_floats(px, c)), device="cuda")
    mean, rstd, scale, shift = (torch.empty(c, device="cuda") for _ in range(4))
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    _call("adb_bn_train_stats", _ptr(zt), px, c, c, _ptr(gamma.detach()), _ptr(beta.detach()), 1e-5, 0.1, _ptr(rm), _ptr(rv), _ptr(nbt),
          _ptr(scratch), _ptr(mean), _ptr(rstd), _ptr(scale), _ptr(shift))
    rt = _nhwc(res) if with_res else None
    yt = torch.empty_like(zt)
    _call("adb_affine_act", _ptr(zt), c, px, c, _ptr(scale), _ptr(shift), _ptr(rt), c, act, _ptr(yt), c)
    _close(_nchw(yt, c), yref.detach(), 1e-2, 1e-3)
    _close(rm, rm_ref, 1e-4, 1e-6)
    _close(rv, rv_ref, 1e-4, 1e-6)
    assert int(nbt.item()) == 1
    # backward from the exact fp32 y (rounded to bf16 like the tape keeps it)
    yk = _nhwc(yref.detach())
    dyt = _nhwc(dy)
    gt, dzt = torch.empty_like(dyt), torch.empty_like(dyt)
    dgamma, dbeta = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    _call("adb_bn_bwd", _ptr(dyt), c, _ptr(yk) if act else None, c, _ptr(zt), c, px, c, act, _ptr(gamma.detach()), _ptr(mean), _ptr(rstd),
          _ptr(scratch), _ptr(gt), c, _ptr(dzt), c, _ptr(dgamma), _ptr(dbeta), 0)
    _close(_nchw(dzt, c), grads[0], 1e-2, 1e-3)
    _close(dgamma, grads[1], 5e-3, 1e-3)
    _close(dbeta, grads[2], 5e-3, 1e-3)
    if with_res:
        _close(_nchw(gt, c), grads[3], 1e-2, 1e-3)
    if act == 1 and not with_res:
        # the y-free path: ReLU mask recomputed from z with the forward's (scale, shift); in place over dy, then accumulated
        # into a wider gradient buffer's prefix (DenseNet block buffer)
        dg2, db2 = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        dy2 = dyt.clone()
        _call("adb_bn_relu_bwd", _ptr(dy2), c, _ptr(zt), c, px, c, _ptr(scale), _ptr(shift), _ptr(gamma.detach()), _ptr(mean), _ptr(rstd),
              _ptr(scratch), _ptr(dy2), c, 0, _ptr(dg2), _ptr(db2), 0)
        _close(_nchw(dy2, c), grads[0], 1e-2, 1e-3)
        _close(dg2, grads[1], 5e-3, 1e-3)
        _close(db2, grads[2], 5e-3, 1e-3)
        _close(dg2, dgamma, 1e-5, 1e-5)        # same sums as the y-based kernel up to fp32 summation order
        _close(db2, dbeta, 1e-5, 1e-5)
        pitch = c + 16
        base = torch.randn(n, h, w, pitch, generator=torch.Generator().manual_seed(5)).to(torch.bfloat16).cuda()
        acc = base.clone()
        _call("adb_bn_relu_bwd", _ptr(dyt), c, _ptr(zt), c, px, c, _ptr(scale), _ptr(shift), _ptr(gamma.detach()), _ptr(mean), _ptr(rstd),
              _ptr(scratch), _ptr(acc), pitch, 1, _ptr(dg2), _ptr(db2), 0)
        assert torch.equal(acc[..., :c], (base[..., :c].float() + dy2.float()).to(torch.bfloat16))
        assert torch.equal(acc[..., c:], base[..., c:])


@pytest.mark.parametrize("mode,act", [(0, 3), (1, 2), (2, 2)])
def test_img_head_forward_backward(mode, act):
    n, h, w = 2, 20, 36
    g = torch.Generator().manual_seed(30 + mode)
    z = (torch.randn(n, 3, h, w, generator=g)).to(torch.bfloat16).float().cuda()
    x = torch.rand(n, 3, h, w, generator=g).cuda()
    gd = torch.rand(n, 1, h, w, generator=g).cuda()
    alpha = torch.tensor(0.1).cuda()
    dout = torch.randn(n, 3, h, w, generator=g).cuda()
    zr, gr, ar = z.clone().requires_grad_(True), gd.clone().requires_grad_(True), alpha.clone().requires_grad_(True)
    v = torch.sigmoid(zr) if act == 3 else torch.tanh(zr)
    if mode == 0:
        ref = (1 - ar) * x + ar * v
    elif mode == 1:
        ref = torch.clamp(x + v, 0, 1)
    else:
        ref = torch.clamp(x + v * gr, 0, 1)
    ins = [zr] + ([ar] if mode == 0 else []) + ([gr] if mode == 2 else [])
    grads = torch.autograd.grad(ref, ins, dout)
    zt = _nhwc(z, 16)
    out = torch.empty_like(x)
    gflat = gd.reshape(n, h, w).contiguous()
    _call("adb_img_head_fwd", _ptr(zt), 16, _ptr(x), _ptr(gflat) if mode == 2 else None, _ptr(alpha.reshape(1)) if mode == 0 else None,
          mode, act, n, h, w, _ptr(out))
    _close(out, ref.detach(), 1e-5, 1e-5)
    dz = torch.full_like(zt, 7.0)
    dgd = torch.empty_like(gflat)
    red = torch.empty(4, device="cuda")
    _call("adb_img_head_bwd", _ptr(dout), _ptr(zt), 16, _ptr(x), _ptr(gflat) if mode == 2 else None,
          _ptr(alpha.reshape(1)) if mode == 0 else None, mode, act, n, h, w, _ptr(dz), _ptr(dgd) if mode == 2 else None, _ptr(red))
    _close(_nchw(dz, 3), grads[0], 1e-2, 1e-4)
    assert dz[..., 3:].abs().max().item() == 0
    _close(red[:3], grads[0].sum(dim=(0, 2, 3)), 1e-2, 1e-3)
    if mode == 0:
        _close(red[3:4], grads[1].reshape(1), 1e-3, 1e-3)
    if mode == 2:
        _close(dgd, grads[1].reshape(n, h, w), 1e-4, 1e-5)


def test_dot_head_forward_backward():
    n, h, w, c = 2, 16, 24, 16
    y = _fm(n, c, h, w, 40)
    g = torch.Generator().manual_seed(41)
    wv = (torch.randn(c, generator=g) * 0.3).cuda().requires_grad_(True)
    bv = torch.tensor([0.2]).cuda().requires_grad_(True)
    dg = torch.randn(n, h, w, generator=g).cuda()
    yr = y.clone().requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(yr, wv.view(1, c, 1, 1), bv)).reshape(n, h, w)
    grads = torch.autograd.grad(ref, [yr, wv, bv], dg)
    yt = _nhwc(y)
    out = torch.empty(n, h, w, device="cuda")
    _call("adb_dot_head_fwd", _ptr(yt), c, c, _ptr(wv.detach()), _ptr(bv.detach()), n * h * w, _ptr(out))
    _close(out, ref.detach(), 1e-4, 1e-5)
    dy = torch.empty_like(yt)
    red = torch.empty(c + 1, device="cuda")
    _call("adb_dot_head_bwd", _ptr(dg), _ptr(out), _ptr(yt), c, c, _ptr(wv.detach()), n * h * w, _ptr(dy), c, _ptr(red))
    _close(_nchw(dy, c), grads[0], 1e-2, 1e-4)
    _close(red[:c], grads[1], 1e-3, 1e-3)
    _close(red[c:], grads[2], 1e-3, 1e-3)


@pytest.mark.parametrize("k,stride,pad,scale", [(2, 2, 0, 2), (4, 4, 0, 4), (3, 2, 1, 2), (3, 2, 0, 4)])
def test_maxpool_and_bilinear_backward(k, stride, pad, scale):
    """adb_maxpool_bwd (first-maximum tie rule) and adb_upsample_bilinear_bwd (align_corners=True) vs torch autograd."""
    n, c, h, w = 2, 32, 24, 40
    g = torch.Generator().manual_seed(80)
    x = torch.randn(n, c, h, w, generator=g).to(torch.bfloat16).float().cuda()
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, k, stride, pad)
    dy = _fm(n, c, yr.shape[2], yr.shape[3], 81)
    (ref,) = torch.autograd.grad(yr, xr, dy)
    xt = _nhwc(x)
    yt = torch.empty((n, yr.shape[2], yr.shape[3], c), dtype=torch.bfloat16, device="cuda")
    _call("adb_maxpool_fwd", _ptr(xt), n, h, w, c, k, stride, pad, _ptr(yt))
    assert torch.equal(_nchw(yt, c), yr.detach())
    dx = torch.empty_like(xt)
    _call("adb_maxpool_bwd", _ptr(_nhwc(dy)), _ptr(xt), _ptr(yt), n, h, w, c, k, stride, pad, _ptr(dx))
    _close(_nchw(dx, c), ref, 1e-2, 1e-3)
    # bilinear
    ur = F.interpolate(xr, scale_factor=scale, mode="bilinear", align_corners=True)
    du = _fm(n, c, h * scale, w * scale, 82)
    (ref_u,) = torch.autograd.grad(ur, xr, du)
    dxu = torch.empty_like(xt)
    _call("adb_upsample_bilinear_bwd", _ptr(_nhwc(du)), c, 0, n, h, w, c, scale, _ptr(dxu))
    _close(_nchw(dxu, c), ref_u, 1e-2, 1e-3)


@pytest.mark.parametrize("c,h,w", [(96, 24, 40), (192, 16, 16), (384, 8, 24)])
def test_attention_backward(c, h, w):
    import adam_dehaze_b200.ops as ops
    from adam_dehaze_b200 import _lib
    n, cr = 2, c // 16
    x = _fm(n, c, h, w, 50).relu()
    dy = _fm(n, c, h, w, 51)
    g = torch.Generator().manual_seed(52)
    w1 = (torch.randn(cr, c, 1, 1, generator=g) / c ** 0.5).cuda().requires_grad_(True)
    w2 = (torch.randn(c, cr, 1, 1, generator=g) / cr ** 0.5).cuda().requires_grad_(True)
    ws = (torch.randn(1, 2, 7, 7, generator=g) * 0.1).cuda().requires_grad_(True)
    xr = x.clone().requires_grad_(True)

    def fc(v):
        return F.conv2d(F.relu(F.conv2d(v, w1)), w2)
    gate = torch.sigmoid(fc(F.adaptive_avg_pool2d(xr, 1)) + fc(F.adaptive_max_pool2d(xr, 1)))
    u = xr * gate
    stats = torch.cat([u.mean(1, keepdim=True), u.max(1, keepdim=True)[0]], 1)
    ref = u * torch.sigmoid(F.conv2d(stats, ws, padding=3))
    grads = torch.autograd.grad(ref, [xr, w1, w2, ws], dy)

    ap = ops.AttnParams(w1.detach(), w2.detach(), ws.detach())
    keep = {}
    xt = _nhwc(x)
    yt = ops.attention(xt, ap, scratch=keep)
    _close(_nchw(yt, c), ref.detach(), 1e-2, 1e-3)
    scratch = torch.empty(int(_lib.load().adb_attn_bwd_scratch_floats(n, h, w, c)), device="cuda")
    dx = torch.empty_like(xt)
    dw1, dw2, dws = torch.empty(cr * c, device="cuda"), torch.empty(c * cr, device="cuda"), torch.empty(98, device="cuda")
    _call("adb_attn_bwd", _ptr(_nhwc(dy)), _ptr(xt), n, h, w, c, _ptr(keep[("pool", n, h, w, c)]), _ptr(keep[("gate", n, c)]),
          _ptr(keep[("stats", n, h, w)]), _ptr(keep[("spatial", n, h, w)]), _ptr(ap.w1), _ptr(ap.w2), cr, _ptr(ap.wsp), _ptr(scratch),
          _ptr(dx), _ptr(dw1), _ptr(dw2), _ptr(dws))
    _close(_nchw(dx, c), grads[0], 1e-2, 1e-3)
    _close(dw1.view_as(w1), grads[1], 1e-2, 1e-4)
    _close(dw2.view_as(w2), grads[2], 1e-2, 1e-4)
    _close(dws.view_as(ws), grads[3], 1e-2, 1e-4)


@pytest.mark.parametrize("kind", ["s1", "s1_head", "s2", "convT"])
def test_dgrad_through_forward_kernel(kind):
    """Data gradients = adb_conv2d with transformed weights (training/autograd.py)."""
    import adam_dehaze_b200.ops as ops
    from adam_dehaze_b200.training import autograd as ag
    n, h, w = 2, 16, 32
    g = torch.Generator().manual_seed(60)
    if kind in ("s1", "s1_head"):
        ci, co = (64, 96) if kind == "s1" else (32, 3)
        wt = (torch.randn(co, ci, 3, 3, generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16).float().cuda()
        x = _fm(n, ci, h, w, 61).requires_grad_(True)
        y = F.conv2d(x, wt, padding=1)
        dz = _fm(n, co, h, w, 62)
        spec = ag._dgrad_spec_s1(wt)
        got = ops.conv2d(spec, _nhwc(dz, ops.pad16(co)))
    elif kind == "s2":
        ci, co = 64, 128
        wt = (torch.randn(co, ci, 4, 4, generator=g) / (16 * ci) ** 0.5).to(torch.bfloat16).float().cuda()
        x = _fm(n, ci, h, w, 61).requires_grad_(True)
        y = F.conv2d(x, wt, stride=2, padding=1)
        dz = _fm(n, co, h // 2, w // 2, 62)
        got = ops.conv2d(ops.ConvSpec.from_convT(wt), _nhwc(dz))
    else:
        ci, co = 128, 64
        wt = (torch.randn(ci, co, 4, 4, generator=g) / (16 * ci) ** 0.5).to(torch.bfloat16).float().cuda()
        x = _fm(n, ci, h, w, 61).requires_grad_(True)
        y = F.conv_transpose2d(x, wt, stride=2, padding=1)
        dz = _fm(n, co, 2 * h, 2 * w, 62)
        got = ops.conv2d(ops.ConvSpec.from_conv(wt, stride=2, pad=1), _nhwc(dz))
    (ref,) = torch.autograd.grad(y, x, dz)
    _close(_nchw(got, ci), ref, 1e-2, 1e-3)


def test_flat_adam_matches_torch_adam():
    from adam_dehaze_b200.training.optim import FlatAdam
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.randn(17, 5, device="cuda")), torch.nn.Parameter(torch.randn(33, device="cuda")),
          torch.nn.Parameter(torch.randn((), device="cuda"))]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(qs, lr=1e-2, weight_decay=1e-4)
    opt = FlatAdam(ps, lr=1e-2, weight_decay=1e-4)
    v0 = [p._version for p in ps]
    for step in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        opt.step()
        ref.step()
    for p, q in zip(ps, qs):
        _close(p.detach(), q.detach(), 1e-5, 1e-6)
    assert all(p._version > v for p, v in zip(ps, v0))     # packing caches keyed on _version see the update


def test_soft_router_train_step():
    """SoftRouter under train_joint.py:141-150 semantics: blend of the three branches in train() mode, gradients to every
    branch and to the logits; loss decreases over a few FlatAdam steps on a fixed batch."""
    from helpers import CONFIG, make_branch, rand_image
    from adam_dehaze_b200.models.routing import SoftRouter
    from adam_dehaze_b200.training.loss import DehazingLoss
    from adam_dehaze_b200.training.optim import FlatAdam
    models = {k: make_branch(k).cuda() for k in ("low", "medium", "high")}
    router = SoftRouter(models, classifier=None, temperature=0.5).cuda().train()
    x, tgt = rand_image(2, 64, 64, 9).cuda(), rand_image(2, 64, 64, 10).cuda()
    logits = torch.tensor([[0.3, -0.2, 0.1], [0.0, 0.5, -0.4]], device="cuda", requires_grad=True)
    crit = DehazingLoss(1.0, 0.0, 0.0)
    opt = FlatAdam(router.parameters(), lr=1e-3, weight_decay=1e-4)
    losses = []
    for it in range(4):
        opt.zero_grad()
        out, info = router(x, logits)
        loss, _ = crit(out, tgt)
        loss.backward()
        if it == 0:
            assert logits.grad is not None and logits.grad.abs().max().item() > 0
            # blend backward against autograd on the individual outputs
            ys = [info["individual_outputs"][k].detach() for k in ("low", "medium", "high")]
            lg = logits.detach().clone().requires_grad_(True)
            wts = torch.softmax(lg / 0.5, 1)
            ref = sum(wts[:, k].view(-1, 1, 1, 1) * ys[k] for k in range(3))
            ((ref - tgt).abs().mean()).backward()
            _close(logits.grad, lg.grad, 1e-3, 1e-6)
            assert all(p.grad is not None for p in router.parameters())
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses


# --------------------------------------------------------------------------------------------------------------------
# Perceptual loss terms against the oracle (fp32, torch autograd) with the SAME random-init trunks.
# Tolerances: values 1e-2 relative (north_star).  d/d(pred) flows through 10-13 bf16 conv layers with ReLU / max-pool
# masks, so — as for the branch gradients above — it is bounded by the fp32 oracle's own error under bf16 storage:
# ||g - g_ref|| <= 1.5 ||g_sim - g_ref|| + 0.05 ||g_ref||, and cos(g, g_ref) >= 0.9.
def _smooth_pair(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    base = F.interpolate(torch.rand(n, 3, h // 8, w // 8, generator=g), size=(h, w), mode="bilinear", align_corners=False)
    tgt = base.clamp(0, 1).cuda()
    pred = (base * 0.8 + 0.15 + 0.05 * torch.rand(n, 3, h, w, generator=g)).clamp(0, 1).cuda()
    return pred, tgt


def _grad_vs_oracle(fn_ref, fn_ours, pred):
    from helpers import oracle_bf16_storage
    p1 = pred.clone().requires_grad_(True)
    v_ref = fn_ref(p1)
    (g_ref,) = torch.autograd.grad(v_ref.sum(), p1)
    p2 = pred.clone().requires_grad_(True)
    v_sim = oracle_bf16_storage(fn_ref, p2)
    (g_sim,) = torch.autograd.grad(v_sim.sum(), p2)
    p3 = pred.clone().requires_grad_(True)
    v = fn_ours(p3)
    v.sum().backward()
    torch.cuda.synchronize()
    return v.detach(), v_ref.detach(), p3.grad, g_ref, g_sim


def test_content_loss_matches_oracle():
    from helpers import make_branch  # noqa: F401  (sys.path)
    import adam_oracle as oracle
    from adam_dehaze_b200.training.loss import ContentLoss
    torch.manual_seed(11)
    crit = ContentLoss().cuda()
    vsd = {k: v.detach().float() for k, v in crit.model.state_dict().items()}
    pred, tgt = _smooth_pair(2, 64, 96, 70)
    v, v_ref, g, g_ref, g_sim = _grad_vs_oracle(lambda p: oracle.content_loss(vsd, p, tgt), lambda p: crit(p, tgt), pred)
    assert abs(v.item() - v_ref.item()) <= 1e-2 * abs(v_ref.item()), (v.item(), v_ref.item())
    e, es, rn = (g - g_ref).norm().item(), (g_sim - g_ref).norm().item(), g_ref.norm().item()
    cos = (g * g_ref).sum().item() / (g.norm().item() * rn)
    assert e <= 1.5 * es + 0.05 * rn and cos >= 0.9, (e / rn, es / rn, cos)
    with torch.no_grad():
        assert abs(crit(pred, tgt).item() - v_ref.item()) <= 1e-2 * abs(v_ref.item())


def test_lpips_matches_oracle():
    from helpers import make_branch  # noqa: F401
    import adam_oracle as oracle
    from adam_dehaze_b200.training.loss import PerceptualLoss
    torch.manual_seed(12)
    crit = PerceptualLoss().cuda()
    asd = {}
    for i, idx in enumerate((0, 3, 6, 8, 10)):
        conv = crit.loss_fn.convs()[i]
        asd[f"{idx}.weight"], asd[f"{idx}.bias"] = conv.weight.detach().float(), conv.bias.detach().float()
    lins = [w.detach().float() for w in crit.loss_fn.lin_weights()]
    pred, tgt = _smooth_pair(2, 96, 128, 71)
    v, v_ref, g, g_ref, g_sim = _grad_vs_oracle(lambda p: oracle.perceptual_lpips(asd, lins, p, tgt), lambda p: crit(p, tgt), pred)
    assert v.shape == (2, 1, 1, 1)
    assert (v - v_ref).abs().max().item() <= 1e-2 * v_ref.abs().max().item(), (v.flatten().tolist(), v_ref.flatten().tolist())
    e, es, rn = (g - g_ref).norm().item(), (g_sim - g_ref).norm().item(), g_ref.norm().item()
    cos = (g * g_ref).sum().item() / (g.norm().item() * rn)
    assert e <= 1.5 * es + 0.05 * rn and cos >= 0.9, (e / rn, es / rn, cos)


def test_dehazing_loss_full_and_joint_loss():
    """DehazingLoss with the reference's lambdas (1.0, 0.1, 0.1) and JointLoss on top: components, total, gradient flow."""
    from helpers import CONFIG
    from adam_dehaze_b200.training.loss import get_dehazing_loss, get_joint_loss
    torch.manual_seed(13)
    crit = get_joint_loss(CONFIG).cuda()
    pred, tgt = _smooth_pair(2, 64, 64, 72)
    pred.requires_grad_(True)
    logits = torch.tensor([[0.2, 0.1, -0.3], [1.0, -1.0, 0.0]], device="cuda", requires_grad=True)
    labels = torch.tensor([0, 2], device="cuda")
    total, parts = crit(pred, tgt, logits, labels)
    dc = parts["dehazing_components"]
    want = 1.0 * (1.0 * dc["l1"] + 0.1 * dc["content"] + 0.1 * dc["perceptual"]) + 0.2 * parts["classification"]
    assert abs(total.item() - want.item()) <= 1e-5 * abs(want.item())
    assert abs(parts["classification"].item() - F.cross_entropy(logits.detach(), labels).item()) <= 1e-5
    total.backward()
    assert pred.grad is not None and pred.grad.abs().max().item() > 0 and logits.grad is not None
    assert set(parts) == {"dehazing", "classification", "detection", "total", "dehazing_components"}
    assert set(dc) == {"l1", "content", "perceptual", "total"}


@pytest.mark.parametrize("arch,shape", [("resnet18", (4, 128, 160)), ("densenet121", (3, 128, 128))])
def test_classifier_train_step_matches_oracle(arch, shape):
    """HDEN (resnet18, and the north_star's densenet121) in train() mode — train_joint.py:117-150 trains it through the router: batch-statistics BN through the
    BasicBlocks (3x3 stride-2 and 1x1 stride-2 downsample convs, max-pool, global average pool) and the head MLP, forward
    and backward vs the fp32 oracle under autograd.  Dropout is disabled (p = 0) for the comparison; same end-to-end
    gradient criterion as the branch models (bf16-storage floor)."""
    from helpers import make_classifier, oracle_bf16_storage, rand_image
    import adam_oracle as oracle
    clf = make_classifier(arch).cuda().train()
    clf.classifier[0].p = 0.0
    clf.classifier[3].p = 0.0
    x = rand_image(*shape, 21).cuda()
    labels = torch.tensor([0, 2, 1, 1], device="cuda")[:shape[0]]
    sd = {k: v.detach().clone().float().requires_grad_(v.dtype.is_floating_point) for k, v in clf.state_dict().items()}
    names = [k for k, _ in clf.named_parameters()]

    def run(fn):
        with oracle.train_mode():
            lg, ft = fn()
        loss = F.cross_entropy(lg, labels)
        return lg.detach(), ft.detach(), loss.detach(), dict(zip(names, torch.autograd.grad(loss, [sd[k] for k in names])))
    ref_lg, ref_ft, ref_loss, ref_g = run(lambda: oracle.classifier_forward(sd, x, arch))
    sim_lg, sim_ft, _, sim_g = run(lambda: oracle_bf16_storage(oracle.classifier_forward, sd, x, arch))
    floor_ft = (sim_ft - ref_ft).abs().max().item()       # batch statistics over 4x5x4 samples at layer4 amplify bf16 rounding
    floor_lg = (sim_lg - ref_lg).abs().max().item()
    from adam_dehaze_b200.training.loss import _CrossEntropy
    logits, feats = clf(x)
    loss = _CrossEntropy.apply(logits, labels)
    loss.backward()
    torch.cuda.synchronize()
    assert (logits - ref_lg).abs().max().item() <= max(3e-2 * max(1.0, ref_lg.abs().max().item()), 1.5 * floor_lg)
    assert (feats - ref_ft).abs().max().item() <= max(3e-2 * ref_ft.abs().max().item(), 1.5 * floor_ft)
    assert abs(loss.item() - ref_loss.item()) <= 1e-2 * abs(ref_loss.item())
    tot_e = tot_s = tot_r = 0.0
    for k, p in clf.named_parameters():
        assert p.grad is not None and p.grad.shape == ref_g[k].shape, k
        e, es, rn = (p.grad - ref_g[k]).norm().item(), (sim_g[k] - ref_g[k]).norm().item(), ref_g[k].norm().item()
        tot_e += e * e; tot_s += es * es; tot_r += rn * rn
        assert (p.grad * ref_g[k]).sum().item() > 0 or rn < 1e-7, k
    assert tot_e ** 0.5 <= 1.25 * tot_s ** 0.5 + 0.02 * tot_r ** 0.5, (tot_e ** 0.5 / tot_r ** 0.5, tot_s ** 0.5 / tot_r ** 0.5)
    # dropout active: a second forward differs, and the head still back-propagates
    clf.classifier[0].p, clf.classifier[3].p = 0.3, 0.2
    a, _ = clf(x)
    b, _ = clf(x)
    assert (a - b).abs().max().item() > 0


def test_joint_training_step_with_classifier():
    """train_joint.py:129-150: classifier(x) -> SoftRouter(x, logits) -> JointLoss -> backward -> optimizer over
    router.parameters() (which include the classifier and the three branches)."""
    from helpers import CONFIG, make_branch, make_classifier, rand_image
    from adam_dehaze_b200.models.routing import create_router
    from adam_dehaze_b200.training.loss import JointLoss, DehazingLoss
    from adam_dehaze_b200.training.optim import FlatAdam
    branches = {k: make_branch(k) for k in ("low", "medium", "high")}
    clf = make_classifier("resnet18")
    router = create_router(branches, clf, dict(CONFIG, routing={"type": "soft", "temperature": 0.5})).cuda().train()
    crit = JointLoss(1.0, 0.2, 0.5, dehazing_loss=DehazingLoss(1.0, 0.0, 0.0)).cuda()
    params = list(router.parameters())
    assert any(p is q for p in params for q in clf.parameters())        # the classifier trains with the router
    opt = FlatAdam(params, lr=5e-5, weight_decay=1e-4)
    x, tgt = rand_image(3, 64, 64, 31).cuda(), rand_image(3, 64, 64, 32).cuda()
    labels = torch.tensor([0, 1, 2], device="cuda")
    losses = []
    for it in range(3):
        opt.zero_grad()
        logits, _ = clf(x)
        out, info = router(x, logits)
        loss, parts = crit(out, tgt, logits, labels)
        loss.backward()
        if it == 0:
            missing = [k for k, p in router.named_parameters() if p.grad is None]
            assert not missing, missing[:5]
        opt.step()
        losses.append(loss.item())
    assert all(l == l for l in losses) and losses[-1] < losses[0] + 0.05, losses


def test_gated_router_train_step():
    """GatedRouter (routing.py:134-226) in train() mode: the gate MLP on the classifier features is trained through the
    blend; gate-MLP gradients (dropout off) against torch autograd on the same features / branch outputs."""
    from helpers import CONFIG, make_branch, make_classifier, rand_image
    from adam_dehaze_b200.models.routing import create_router
    branches = {k: make_branch(k) for k in ("low", "medium", "high")}
    clf = make_classifier("resnet18")
    torch.manual_seed(3)
    router = create_router(branches, clf, dict(CONFIG, routing={"type": "gated", "temperature": 0.5})).cuda().train()
    router.gate_network[2].p = 0.0
    clf.classifier[0].p = clf.classifier[3].p = 0.0
    x, tgt = rand_image(2, 64, 64, 41).cuda(), rand_image(2, 64, 64, 42).cuda()
    out, info = router(x)
    loss = (out - tgt).abs().mean()
    loss.backward()
    g = router.gate_network
    assert all(p.grad is not None for p in router.parameters())
    # reference: same features and branch outputs, gate MLP + softmax + blend in torch
    with torch.no_grad():
        _, feats = clf(x)       # train-mode forward again: identical batch statistics, dropout disabled
    ys = [info["individual_outputs"][k].detach() for k in ("low", "medium", "high")]
    ws = [p.detach().clone().requires_grad_(True) for p in (g[0].weight, g[0].bias, g[3].weight, g[3].bias, g[5].weight, g[5].bias)]
    h = F.relu(F.linear(F.relu(F.linear(feats, ws[0], ws[1])), ws[2], ws[3]))
    wts = torch.softmax(F.linear(h, ws[4], ws[5]), 1)
    ref = sum(wts[:, k].view(-1, 1, 1, 1) * ys[k] for k in range(3))
    (ref - tgt).abs().mean().backward()
    assert torch.allclose(info["gate_weights"], wts.detach(), atol=2e-3)
    for p, r in zip((g[0].weight, g[0].bias, g[3].weight, g[3].bias, g[5].weight, g[5].bias), ws):
        _close(p.grad, r.grad, 5e-2, 1e-7)


def test_hard_router_train_mode():
    """HardRouter under model.train(): every image goes through its own branch's sub-batch (batch-statistics BN over the
    bucket), gradients reach exactly the branches that received images."""
    from helpers import CONFIG, make_branch, rand_image
    from adam_dehaze_b200.models.routing import create_router
    branches = {k: make_branch(k) for k in ("low", "medium", "high")}
    router = create_router(branches, None, CONFIG).cuda().train()
    x, tgt = rand_image(5, 64, 64, 51).cuda(), rand_image(5, 64, 64, 52).cuda()
    labels = torch.tensor([1, 0, 1, 1, 0], device="cuda")
    out, info = router(x, intensity=labels)
    assert torch.equal(info["intensity"], labels)
    (out - tgt).abs().mean().backward()
    assert all(p.grad is not None for p in branches["low"].parameters())
    assert all(p.grad is not None for p in branches["medium"].parameters())
    assert all(p.grad is None for p in branches["high"].parameters())          # no sample routed there
    with torch.no_grad():
        alone = branches["medium"](x[[0, 2, 3]].contiguous())                   # same sub-batch -> same batch statistics
    assert (out[[0, 2, 3]] - alone).abs().max().item() <= 1e-6


def test_loss_meter_reads_every_step_one_step_late():
    """training.train_dehazing.LossMeter: step i's loss reaches the host when step i+1 is pushed (or at flush), through
    pinned memory — the numbers the reference accumulates with loss.item() (train_dehazing.py:95), without a per-step sync."""
    from adam_dehaze_b200.training.train_dehazing import LossMeter
    m = LossMeter()
    vals = [0.5, 1.25, 2.0, 4.0]
    seen = []
    for v in vals:
        seen.append(m.push(torch.tensor(v, device="cuda") * 1.0))
    assert seen == [None, 0.5, 1.25, 2.0]
    assert m.flush() == 4.0 and m.count == 4 and abs(m.total - sum(vals)) < 1e-6
    assert m.flush() == 4.0 and m.count == 4        # idempotent


def test_train_driver_resume_continues_the_same_run(tmp_path):
    """train_dehazing_model(..., resume=True) (main.py --resume, which the reference parses and ignores): two epochs run in one go
    and one epoch + a resumed second epoch end with the same parameters, Adam step count and checkpoint keys
    (train_dehazing.py:196-203)."""
    import copy
    import os
    from helpers import CONFIG, make_branch
    from adam_dehaze_b200.training.loss import DehazingLoss
    from adam_dehaze_b200.training.train_dehazing import synthetic_loader, train_dehazing_model

    def run(ck_dir, schedule):
        cfg = copy.deepcopy(CONFIG)
        cfg["device"] = "cuda:0"
        cfg["dehazing"]["low"]["learning_rate"] = 1e-3
        cfg["dehazing"]["checkpoint_dir"] = str(ck_dir)
        train = synthetic_loader(2, 6, 32, 32, "cuda:0", seed=5)
        val = synthetic_loader(1, 6, 32, 32, "cuda:0", seed=6)
        model = None
        for epochs, resume in schedule:
            model = make_branch("low", seed=11)          # a fresh process would start from fresh weights too
            model = train_dehazing_model(model, "low", cfg, train_loader=train, val_loader=val, epochs=epochs,
                                         criterion=DehazingLoss(1.0, 0.0, 0.0), resume=resume)
        return torch.load(os.path.join(str(ck_dir), "low", "last_checkpoint.pth"), map_location="cpu")

    a = run(tmp_path / "straight", [(2, False)])
    b = run(tmp_path / "resumed", [(1, False), (2, True)])
    assert a["epoch"] == b["epoch"] == 1
    assert set(a) >= {"epoch", "model_state_dict", "optimizer_state_dict", "val_psnr", "val_ssim", "val_loss"}
    for k in a["model_state_dict"]:
        # (a resume that lost the Adam moments or the step count would move every weight by ~lr = 1e-3 on its first step)
        assert torch.allclose(a["model_state_dict"][k].float(), b["model_state_dict"][k].float(), rtol=1e-3, atol=2e-5), k
    sa, sb = a["optimizer_state_dict"]["state"], b["optimizer_state_dict"]["state"]
    assert float(sa[0]["step"]) == float(sb[0]["step"]) > 0
    assert torch.allclose(sa[0]["exp_avg"], sb[0]["exp_avg"], rtol=1e-3, atol=1e-7)


def test_weight_cache_refreshes_every_packing_in_one_launch():
    """After an optimizer step the bf16 packings of ALL parameters (forward, row-folded, data-gradient layouts) are refreshed
    by one adb_gather_cast_multi launch; each must equal a fresh packing of the updated weight."""
    from helpers import make_branch, rand_image
    from adam_dehaze_b200 import _lib
    from adam_dehaze_b200.training import autograd as ag
    from adam_dehaze_b200.training.loss import DehazingLoss
    from adam_dehaze_b200.training.optim import FlatAdam
    m = make_branch("medium").cuda().train()
    opt = FlatAdam(m.parameters(), lr=1e-2)
    crit = DehazingLoss(1.0, 0.0, 0.0)
    x, tgt = rand_image(2, 64, 128, 3).cuda(), rand_image(2, 64, 128, 4).cuda()
    calls = []
    inner = _lib.call

    def spy(name, *a):
        calls.append(name)
        return inner(name, *a)
    for it in range(3):
        opt.zero_grad()
        if it == 2:
            _lib.call = ag._lib.call = spy
        try:
            loss, _ = crit(m(x), tgt)
        finally:
            _lib.call = ag._lib.call = inner
        loss.backward()
        opt.step()
    assert calls.count("adb_gather_cast_multi") == 1 and calls.count("adb_gather_cast") == 0
    # one more forward refreshes against the latest weights; compare with fresh packings
    opt.zero_grad()
    crit(m(x), tgt)[0].backward()
    cache = m.__dict__["_adb_engine"].train_cache if "_adb_engine" in m.__dict__ else m.engine.train_cache
    checked = 0
    for key, e in cache._c.items():
        if e["maps"] is None or key[0] != "f":
            continue
        w = e["weight"].detach()
        if w.dim() != 4 or w.shape[2] != 3 or w.shape[1] % 16:
            continue
        spec = e["val"]
        fresh = ag.ConvSpec.from_conv(w, stride=1, pad=1) if spec.kind == 0 else None
        if fresh is None or fresh.w_packed.shape != spec.w_packed.shape:
            continue
        assert torch.equal(spec.w_packed, fresh.w_packed)
        if spec.w_fold is not None:
            assert torch.equal(spec.w_fold, fresh.w_fold)
        checked += 1
    assert checked >= 8


def test_graphed_step_matches_eager_steps():
    """training/graphed.GraphedStep: the captured step (forward, loss, backward, FlatAdam) replayed N times leaves the same
    parameters, Adam moments / step counts and BatchNorm statistics as N eager steps (capture records, it does not execute)."""
    from helpers import make_branch, rand_image
    from adam_dehaze_b200.training.graphed import GraphedStep
    from adam_dehaze_b200.training.loss import DehazingLoss
    from adam_dehaze_b200.training.optim import FlatAdam
    x, tgt = rand_image(2, 64, 128, 7).cuda(), rand_image(2, 64, 128, 8).cuda()
    crit = DehazingLoss(1.0, 0.0, 0.0)

    def make():
        m = make_branch("low").cuda().train()
        opt = FlatAdam(m.parameters(), lr=1e-3, weight_decay=1e-4)

        def step():
            opt.zero_grad()
            loss, _ = crit(m(x), tgt)
            loss.backward()
            opt.step()
            return loss
        return m, opt, step
    m_e, opt_e, step_e = make()
    for _ in range(4):
        loss_e = step_e()
    m_g, opt_g, step_g = make()
    gs = GraphedStep(step_g, warmup=2)          # 2 eager steps + 1 capture (recorded, not executed)
    for _ in range(2):
        loss_g = gs()
    torch.cuda.synchronize()
    assert abs(loss_g.item() - loss_e.item()) <= 1e-5 * abs(loss_e.item())
    for (k, a), (_, b) in zip(m_e.state_dict().items(), m_g.state_dict().items()):
        assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-7), k
    assert torch.equal(opt_e.seg_step, opt_g.seg_step) and int(opt_g.seg_step[0]) == 4
    assert torch.allclose(opt_e.exp_avg, opt_g.exp_avg, rtol=1e-5, atol=1e-9)
