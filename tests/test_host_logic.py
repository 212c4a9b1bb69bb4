"""CPU: host-side logic — weight packing / BN folding against a Python restatement of the kernel's K walk, factories,
error behaviour, the C-ABI export list, and sharding (incl. a world_size-2 gloo run)."""
import os
import re
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from helpers import CONFIG, ROOT, make_branch, make_classifier

from adam_dehaze_b200 import _lib, ops, sharding

torch.set_grad_enabled(False)


# ----------------------------------------------------------------------------- K-walk emulation (mirrors conv_igemm.cu build())
def _gather(src, dh, dw, stride=1, ph=0, pw=0):
    """src NHWC float; returns A[n, ho, wo, c] = src[n, stride*(ho+dh)+ph, stride*(wo+dw)+pw, c] with zero fill."""
    n, h, w, c = src.shape
    ho, wo = h // stride, w // stride
    out = torch.zeros(n, ho, wo, c)
    for y in range(ho):
        yy = stride * (y + dh) + ph
        if not (0 <= yy < h):
            continue
        for x in range(wo):
            xx = stride * (x + dw) + pw
            if 0 <= xx < w:
                out[:, y, x] = src[:, yy, xx]
    return out


def _fl2(u):
    return u // 2  # python floor division == the kernel's fl2()


def emulate(spec, srcs):
    srcs = [s.float() for s in srcs]
    wp = spec.w_packed.float()
    cols = []
    if spec.kind == ops.CONV_S1:
        ph_, pw_ = (spec.kh - 1) // 2, (spec.kw - 1) // 2
        for r in range(spec.kh):
            for s in range(spec.kw):
                cols += [_gather(t, r - ph_, s - pw_) for t in srcs]
        a = torch.cat(cols, dim=3)
        y = a @ wp.t()
    elif spec.kind == ops.CONV_S2:
        for r in range(spec.kh):
            for s in range(spec.kw):
                u, v = r - spec.pad, s - spec.pad
                cols += [_gather(t, _fl2(u), _fl2(v), 2, u - 2 * _fl2(u), v - 2 * _fl2(v)) for t in srcs]
        y = torch.cat(cols, dim=3) @ wp.t()
    else:
        n, h, w, _ = srcs[0].shape
        y = torch.zeros(n, 2 * h, 2 * w, wp.shape[1])
        for a_ in range(2):
            for b_ in range(2):
                cols = []
                for i in range(2):
                    for j in range(2):
                        cols += [_gather(t, (1 - i) if a_ else -i, (1 - j) if b_ else -j) for t in srcs]
                y[:, a_::2, b_::2] = torch.cat(cols, dim=3) @ wp[a_ * 2 + b_].t()
    y = y * spec.scale + spec.shift
    return F.relu(y) if spec.act == ops.ACT_RELU else y


def _bn(c, g):
    return (torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1, torch.randn(c, generator=g) * 0.1,
            torch.rand(c, generator=g) + 0.5, 1e-5)


def _bn_ref(y, bn):
    return F.batch_norm(y, bn[2], bn[3], bn[0], bn[1], False, 0.0, bn[4])


def _bf(t):
    return t.to(torch.bfloat16).float()


def test_pack_conv_s1_and_concat():
    g = torch.Generator().manual_seed(0)
    a, b = _bf(torch.randn(1, 16, 6, 7, generator=g)), _bf(torch.randn(1, 32, 6, 7, generator=g))
    w = _bf(torch.randn(20, 48, 3, 3, generator=g) * 0.1)
    bn = _bn(20, g)
    spec = ops.ConvSpec.from_conv(w, bn=bn, act=ops.ACT_RELU)
    assert spec.cout_pad == 32 and spec.w_packed.shape == (32, 9 * 48)
    y = emulate(spec, [a.permute(0, 2, 3, 1), b.permute(0, 2, 3, 1)])
    ref = F.relu(_bn_ref(F.conv2d(torch.cat([a, b], 1), w, padding=1), bn))
    assert torch.allclose(y[..., :20].permute(0, 3, 1, 2), ref, atol=1e-4)
    assert y[..., 20:].abs().max() == 0


def test_pack_conv_fold_rows_accumulate_into_adjacent_output_rows():
    """The rolling-row kernel's algebra (csrc/conv_roll.cu): E_j[w,(blk,co)] = sum_{s,ci} X[j,w+s-1,ci] Wf[s][blk*cp+co][ci],
    out[h] = E_{h-1}[blk 2] + E_h[blk 1] + E_{h+1}[blk 0], must reproduce F.conv2d(padding=1) with the packing of
    ops.pack_conv_weight_fold."""
    import torch.nn.functional as F
    from adam_dehaze_b200 import ops
    torch.manual_seed(0)
    co, ci, h, w = 24, 16, 7, 10
    wt = torch.randn(co, ci, 3, 3)
    x = torch.randn(1, ci, h, w)
    with ops.pack_as(torch.float32):
        wf = ops.pack_conv_weight_fold(wt)            # [3][3*cp][ci]
    cp = ops.pad16(co)
    assert wf.shape == (3, 3 * cp, ci) and cp == 32
    # 3*cout_pad <= 256 -> cout_pad 16 (image / guidance heads), 32, 64; 96 channels do not fit one MMA's N
    assert ops.fold_eligible(co, 3, 3, 1, 1) and ops.fold_eligible(16, 3, 3, 1, 1) and ops.fold_eligible(3, 3, 3, 1, 1)
    assert not ops.fold_eligible(96, 3, 3, 1, 1) and not ops.fold_eligible(32, 3, 3, 2, 1) and not ops.fold_eligible(32, 1, 1, 1, 0)
    xp = F.pad(x, (1, 1, 0, 0))[0].permute(1, 2, 0)    # [h][w+2][ci]
    out = torch.zeros(h, w, cp)
    for j in range(h):
        e = sum(xp[j, s:s + w] @ wf[s].t() for s in range(3))       # [w][3*cp]
        for blk in range(3):
            hh = j - 1 + blk
            if 0 <= hh < h:
                out[hh] += e[:, blk * cp:(blk + 1) * cp]
    ref = F.conv2d(x, wt, padding=1)[0].permute(1, 2, 0)
    assert torch.allclose(out[..., :co], ref, atol=1e-4)
    assert out[..., co:].abs().max() == 0


def test_pack_stem_space_to_depth_form_equals_the_7x7_stride2_conv():
    """ops.pack_stem_s2d_weight: torchvision's 7x7 / stride 2 / pad 3 stem (classifier.py:24-36) == a 4x4-tap conv at offsets
    -2..+1 over the space-to-depth image (ADB_CONV_K4_S2D, include/adb200.h)."""
    import torch.nn.functional as F
    from adam_dehaze_b200 import ops
    torch.manual_seed(1)
    w = torch.randn(24, 3, 7, 7)
    x = torch.randn(2, 3, 20, 28)
    with ops.pack_as(torch.float32):
        wp = ops.pack_stem_s2d_weight(w)                                    # [32, 256]
    assert wp.shape == (32, 256)
    xs = torch.zeros(2, 16, 10, 14)
    for py in range(2):
        for px in range(2):
            q = (py * 2 + px) * 3
            xs[:, q:q + 3] = x[:, :, py::2, px::2]
    w4 = wp.view(32, 4, 4, 16).permute(0, 3, 1, 2)                           # [co, ch, R, S]
    got = F.conv2d(F.pad(xs, (2, 1, 2, 1)), w4)                              # taps at offsets -2..+1
    ref = F.conv2d(x, w, stride=2, padding=3)
    assert torch.allclose(got[:, :24], ref, atol=1e-4) and got[:, 24:].abs().max() == 0


@pytest.mark.parametrize("k,pad", [(4, 1), (3, 1), (1, 0)])
def test_pack_conv_s2(k, pad):
    g = torch.Generator().manual_seed(1)
    x = _bf(torch.randn(2, 16, 8, 12, generator=g))
    w = _bf(torch.randn(16, 16, k, k, generator=g) * 0.1)
    bias = torch.randn(16, generator=g)
    spec = ops.ConvSpec.from_conv(w, bias=bias, stride=2, pad=pad)
    y = emulate(spec, [x.permute(0, 2, 3, 1)])
    assert torch.allclose(y.permute(0, 3, 1, 2), F.conv2d(x, w, bias, stride=2, padding=pad), atol=1e-4)


def test_pack_conv_transpose_phases():
    g = torch.Generator().manual_seed(2)
    x = _bf(torch.randn(1, 16, 5, 6, generator=g))
    wt = _bf(torch.randn(16, 24, 4, 4, generator=g) * 0.1)
    bias = torch.randn(24, generator=g)
    bn = _bn(24, g)
    spec = ops.ConvSpec.from_convT(wt, bias=bias, bn=bn, act=ops.ACT_RELU)
    assert spec.w_packed.shape == (4, 32, 64)
    y = emulate(spec, [x.permute(0, 2, 3, 1)])
    ref = F.relu(_bn_ref(F.conv_transpose2d(x, wt, bias, stride=2, padding=1), bn))
    assert torch.allclose(y[..., :24].permute(0, 3, 1, 2), ref, atol=1e-4)


@pytest.mark.parametrize("k,kp", [(3, 16), (7, 32)])
def test_pack_stem(k, kp):
    """stem_pack layout (out[..., s*3+c] = x[c, h, w+s-pad]) + kh x 1 conv == the kxk stem conv."""
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 9, 11, generator=g)
    w = _bf(torch.randn(8, 3, k, k, generator=g) * 0.1)
    spec = ops.ConvSpec.from_stem(w, kp)
    assert (spec.kh, spec.kw) == (k, 1) and spec.w_packed.shape == (16, k * kp)
    xp = F.pad(x, (k // 2, k // 2, 0, 0))
    packed = torch.zeros(1, 9, 11, kp)
    for s in range(k):
        for c in range(3):
            packed[0, :, :, s * 3 + c] = xp[0, c, :, s:s + 11]
    y = emulate(spec, [packed])
    assert torch.allclose(y[..., :8].permute(0, 3, 1, 2), F.conv2d(x, w, padding=k // 2), atol=1e-4)


def test_fold_bn_matches_batchnorm_eval():
    g = torch.Generator().manual_seed(4)
    bn = _bn(5, g)
    bias = torch.randn(5, generator=g)
    s, b = ops.fold_bn(5, bias, bn)
    acc = torch.randn(3, 5, 4, 4, generator=g)
    ref = _bn_ref(acc + bias.view(1, -1, 1, 1), bn)
    assert torch.allclose(acc * s[:5].view(1, -1, 1, 1) + b[:5].view(1, -1, 1, 1), ref, atol=1e-5)
    assert s.numel() == 16 and s[5:].abs().sum() == 0 and b[5:].abs().sum() == 0


# ----------------------------------------------------------------------------- boundary behaviour
def test_factories_and_errors():
    from adam_dehaze_b200.models.classifier import create_classifier
    from adam_dehaze_b200.models.routing import GatedRouter, HardRouter, SoftRouter, create_router
    from adam_dehaze_b200.training.loss import get_dehazing_loss, get_joint_loss
    branches = {n: make_branch(n) for n in ("low", "medium", "high")}
    clf = make_classifier()
    for kind, cls in (("hard", HardRouter), ("soft", SoftRouter), ("gated", GatedRouter)):
        cfg = dict(CONFIG, routing={"type": kind, "temperature": 0.5})
        r = create_router(branches, clf, cfg)
        assert isinstance(r, cls)
    assert create_router(branches, clf, dict(CONFIG, routing={"type": "soft", "temperature": 0.5})).temperature == 0.5
    with pytest.raises(ValueError, match="Unsupported routing type"):
        create_router(branches, clf, dict(CONFIG, routing={"type": "nope", "temperature": 1}))
    with pytest.raises(ValueError, match="Unsupported model"):
        create_classifier(dict(CONFIG, classifier={"model": "vgg11", "num_classes": 3, "pretrained": False}))
    with pytest.raises(ValueError, match="Unsupported ResNet variant"):
        create_classifier(dict(CONFIG, classifier={"model": "resnet101", "num_classes": 3, "pretrained": False}))
    assert get_dehazing_loss(CONFIG).lambda_content == 0.1
    jl = get_joint_loss(CONFIG)
    assert (jl.lambda_dehazing, jl.lambda_classification, jl.lambda_detection) == (1.0, 0.2, 0.5)
    from adam_dehaze_b200.models.dehazing.base_model import BaseDehazeModel
    with pytest.raises(NotImplementedError):
        BaseDehazeModel()(torch.zeros(1))
    info = branches["high"].get_info()
    assert info["model_type"] == "HighIntensityDehazeModel" and info["params"] == 16320576 and info["base_channels"] == 96
    assert branches["medium"].get_info()["params"] == 7228835 and branches["low"].get_info()["params"] == 66756


def test_no_cpu_fallback_and_train_mode_refused():
    m = make_branch("low")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.rand(1, 3, 16, 16))
    clf = make_classifier()
    with pytest.raises(RuntimeError, match="no CPU path"):
        clf(torch.rand(1, 3, 32, 32))
    # train() mode has no CPU path either (the training kernels are CUDA only)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.train()(torch.rand(1, 3, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU path"):
        clf.train()(torch.rand(1, 3, 32, 32))
    from adam_dehaze_b200 import engine
    v = make_branch("corun").train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        v(torch.rand(1, 3, 64, 64))


def test_wgrad_desc_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "adb200.h")).read()
    body = hdr[hdr.index("typedef struct adb_wgrad_desc {"):hdr.index("} adb_wgrad_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"[\s\*](\w+)\s*[;,]", body)
    assert names == [f[0] for f in _lib.WgradDesc._fields_]


def test_dgrad_weight_transforms_match_autograd():
    """The data-gradient launches are ordinary convs with transformed weights (training/autograd.py): check the
    transforms against torch autograd on CPU in fp32 (3x3 stride 1; 4x4 stride 2 <-> ConvTranspose2d(4,2,1))."""
    import torch.nn.functional as F
    with torch.enable_grad():
        _check_dgrad_transforms(F)


def _check_dgrad_transforms(F):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, 8, 10, generator=g, requires_grad=True)
    w = torch.randn(6, 4, 3, 3, generator=g)
    dz = torch.randn(1, 6, 8, 10, generator=g)
    (ref,) = torch.autograd.grad(F.conv2d(x, w, padding=1), x, dz)
    wd = w.flip(2, 3).permute(1, 0, 2, 3)                      # _dgrad_spec_s1's weight
    assert torch.allclose(F.conv2d(dz, wd, padding=1), ref, atol=1e-5)
    w4 = torch.randn(6, 4, 4, 4, generator=g)
    dz2 = torch.randn(1, 6, 4, 5, generator=g)
    (ref2,) = torch.autograd.grad(F.conv2d(x, w4, stride=2, padding=1), x, dz2)
    assert torch.allclose(F.conv_transpose2d(dz2, w4, stride=2, padding=1), ref2, atol=1e-5)   # weight as is
    wt = torch.randn(4, 6, 4, 4, generator=g)
    dy = torch.randn(1, 6, 16, 20, generator=g)
    (ref3,) = torch.autograd.grad(F.conv_transpose2d(x, wt, stride=2, padding=1), x, dy)
    assert torch.allclose(F.conv2d(dy, wt, stride=2, padding=1), ref3, atol=1e-4)              # weight as is


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "adam_dehaze_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+(oracle|adam_oracle)", src, re.M), f


# ----------------------------------------------------------------------------- C ABI
def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "adb200.h")).read()
    declared = sorted(set(re.findall(r"ADB_API[^;]*?\b(adb_\w+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in adb200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, set(_lib.EXPORTED_SYMBOLS) ^ set(declared)
    assert lib.adb_version() >= 100
    # no compute without a GPU: the device check must report, not crash
    if not torch.cuda.is_available():
        assert lib.adb_device_check() != 0 and "device" in _lib.last_error().lower()


def test_conv_desc_struct_matches_header_field_order():
    hdr = open(os.path.join(ROOT, "include", "adb200.h")).read()
    body = hdr[hdr.index("typedef struct adb_conv_desc {"):hdr.index("} adb_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"[\s\*](\w+)\s*[;,]", body)
    assert names == [f[0] for f in _lib.ConvDesc._fields_]


# ----------------------------------------------------------------------------- sharding
def test_shard_bounds_partition():
    for total in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(total, world)
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ADB_ROOT"])
from adam_dehaze_b200 import sharding
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["ADB_PORT"], rank=int(os.environ["RANK"]), world_size=2)
total = 11
lo, hi = sharding.shard_bounds(total, dist.get_rank(), 2)
mine = torch.zeros(total, dtype=torch.int64); mine[lo:hi] = 1
dist.all_reduce(mine)                       # every image owned by exactly one rank
t = torch.tensor([float(hi - lo) * (dist.get_rank() + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max-over-ranks timing rule
assert mine.tolist() == [1] * total, mine
assert t.item() == 10.0, t
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


def test_sharding_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER)
    port = str(29000 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), ADB_ROOT=ROOT, ADB_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_rebalance_plan_equalises_every_class():
    """sharding.rebalance_plan: every rank ends with its equal share (+-1) of every class, images move only off ranks that
    hold more than their share, nothing is lost or duplicated, and unrouted class ids stay put."""
    import random
    rng = random.Random(0)
    for world in (2, 3, 8):
        for trial in range(5):
            labels = [[rng.choice([0, 0, 0, 1, 2, 2, 7]) if (r + trial) % 2 else rng.choice([0, 1, 1, 1, 2]) for _ in range(rng.randrange(0, 40))]
                      for r in range(world)]
            send = sharding.unrouted_stay(labels, sharding.rebalance_plan(labels))
            for r in range(world):
                sent = sorted(i for d in range(world) for i in send[r][d])
                assert sent == list(range(len(labels[r])))                       # every image goes to exactly one place
            for k in range(3):
                held = [sum(1 for s in range(world) for i in send[s][d] if labels[s][i] == k) for d in range(world)]
                assert max(held) - min(held) <= 1, (k, held)
                for r in range(world):
                    had = sum(1 for v in labels[r] if v == k)
                    kept = sum(1 for i in send[r][r] if labels[r][i] == k)
                    assert kept == min(had, held[r])                             # no image leaves a rank that is not over its share
            assert all(labels[r][i] != 7 or i in send[r][r] for r in range(world) for i in range(len(labels[r])))


_GLOO_A2A_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ADB_ROOT"])
from adam_dehaze_b200 import sharding
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["ADB_PORT"], rank=rank, world_size=2)
labels = [[2, 2, 2, 0, 2, 1, 2], [0, 0, 1, 0, 0, 0, 5]]
ex = sharding.Exchange(labels, rank)
x = torch.arange(7, dtype=torch.float32).view(7, 1, 1, 1) + 100 * rank
got, lab = ex.forward(x)
assert [int((lab == k).sum()) for k in range(3)] == ([3, 1, 3] if rank == 0 else [3, 1, 2]), lab
src = [int(v) // 100 for v in got.flatten().tolist()]
assert all(labels[s][int(v) % 100] == int(l) for s, v, l in zip(src, got.flatten().tolist(), lab.tolist()))   # labels travel with their images
back = ex.backward(got * 2)
assert torch.equal(back, x * 2)                      # results return to their owner in the original order
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


def test_rebalance_exchange_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_A2A_WORKER)
    port = str(31000 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), ADB_ROOT=ROOT, ADB_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


_GLOO_GRAD_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["ADB_ROOT"])
from adam_dehaze_b200.training.optim import FlatAdam
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + os.environ["ADB_PORT"], rank=rank, world_size=2)
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(()))]
before = [p.detach().clone() for p in params]
opt = FlatAdam(params, lr=1e-3, weight_decay=1e-4)
assert all(torch.equal(p.detach(), b) for p, b in zip(params, before))          # flattening keeps the values
assert all(p.data_ptr() == opt.flat.data_ptr() + 4 * o for p, o in zip(params, opt.offsets))   # views of ONE buffer
params[0].grad = torch.full((3, 5), float(rank + 1))
params[2].grad = torch.tensor(10.0 * (rank + 1))
# params[1] saw no sample on this rank: it must still take part in the collective with zeros
world = opt.reduce_gradients()
assert world == 2
assert torch.equal(opt.grad_views[0], torch.full((3, 5), 3.0)), opt.grad_views[0]
assert torch.equal(opt.grad_views[1], torch.zeros(7))
assert opt.grad_views[2].item() == 30.0
try:
    opt.step()
    raise SystemExit("FlatAdam.step ran on CPU tensors")
except RuntimeError as e:
    assert "no CPU path" in str(e)
dist.barrier(); dist.destroy_process_group()
print("ok")
"""


def test_gradient_bucket_allreduce_world2_gloo(tmp_path):
    """Training multi-GPU path: one flat gradient bucket, one sum all-reduce per step; ranks whose shard
    missed a branch contribute zeros so collectives stay matched."""
    script = tmp_path / "g.py"
    script.write_text(_GLOO_GRAD_WORKER)
    port = str(31000 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), ADB_ROOT=ROOT, ADB_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_weight_repack_index_maps():
    """After an optimizer step every packing is refreshed by one gather through an index map derived once
    (training/autograd.derive_index_maps).  The map must reproduce the direct packing for every layout in use:
    conv, stride-1 dgrad (rotated / transposed / row-padded / channel-sliced), ConvTranspose phases, embedded 3x3 stride-2
    dgrad, horizontally packed stems."""
    from adam_dehaze_b200.training import autograd as ag
    from adam_dehaze_b200.ops import ConvSpec
    g = torch.Generator().manual_seed(4)
    cases = [
        (torch.randn(48, 32, 3, 3, generator=g), lambda wt: ConvSpec.from_conv(wt, pad=1)),
        (torch.randn(32, 32, 3, 3, generator=g), lambda wt: ConvSpec.from_conv(wt, pad=1)),        # carries a w_fold packing too
        (torch.randn(3, 32, 3, 3, generator=g), lambda wt: ag._dgrad_spec_s1(wt)),
        (torch.randn(64, 96, 3, 3, generator=g), lambda wt: [(o, ag._dgrad_spec_s1(wt[:, o:o + 32])) for o in (0, 32, 64)]),
        (torch.randn(64, 32, 4, 4, generator=g), lambda wt: ConvSpec.from_convT(wt)),
        (torch.randn(32, 16, 3, 3, generator=g), lambda wt: ConvSpec.from_convT(ag._embed4x4(ag._pad_rows16(wt), 1))),
        (torch.randn(32, 16, 1, 1, generator=g), lambda wt: ConvSpec.from_convT(ag._embed4x4(ag._pad_rows16(wt), 0))),
        (torch.randn(32, 3, 7, 7, generator=g), lambda wt: ConvSpec.from_stem(wt, 32)),
        (torch.randn(128, 64, 4, 4, generator=g), lambda wt: ConvSpec.from_conv(wt[32:96], stride=2, pad=1)),
    ]
    for w, build in cases:
        val, maps = ag.derive_index_maps(build, w)
        w2 = torch.randn(w.shape, generator=g)
        fresh = ag._flat_specs(build(w2))
        # one map per w_packed, then one per row-folded packing (rolling-row kernel) of the specs that carry one
        targets = [sp.w_packed for sp in fresh] + [sp.w_fold for sp in fresh if sp.w_fold is not None]
        assert len(maps) == len(targets) >= 1
        src = torch.cat([torch.zeros(1), w2.reshape(-1)])
        for (packed, idx), tgt in zip(maps, targets):
            got = src[(idx.long() + 1)].to(torch.bfloat16).view(tgt.shape)
            assert packed.shape == tgt.shape and torch.equal(got, tgt)


def test_roofline_traffic_json_follows_from_the_committed_ncu_capture():
    """bench.py's roofline.traffic is read from profiles/r2_conv_traffic.json; that file must be exactly what
    tools/conv_traffic.py computes from the committed ncu csv of the roofline pass (no hand-edited number)."""
    import json
    import subprocess
    csv_path = os.path.join(ROOT, "profiles", "r2", "ncu_conv_traffic.csv")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "conv_traffic.py"), csv_path], capture_output=True, text=True, check=True,
                         cwd=ROOT)
    got = json.loads(out.stdout)
    want = json.load(open(os.path.join(ROOT, "profiles", "r2_conv_traffic.json")))
    assert got["launches"] == want["launches"] == 179
    assert abs(got["dram_bytes_per_launch_avg"] - want["dram_bytes_per_launch_avg"]) <= 1e-6 * want["dram_bytes_per_launch_avg"]
    assert set(got["per_model"]) == {"low", "medium", "high", "densenet121"}
    assert got["per_model"]["densenet121"]["launches"] == 120 and got["per_model"]["high"]["launches"] == 26


def test_tape_sweep_leaves_no_reference_cycle():
    """After the reverse sweep the tape holds neither closures nor gradients: everything a step allocated is released by
    reference count, in a fixed order (a leftover tape<->closure cycle made the caching allocator's state differ from step
    to step and stalled steps in cudaMalloc — profiles/r1_train_summary.md)."""
    import gc
    import weakref
    from adam_dehaze_b200.training.autograd import Tape

    class Probe:
        pass
    order = []
    t = Tape(None)
    probes = [Probe() for _ in range(3)]
    refs = [weakref.ref(p) for p in probes]
    alive_at_run = []
    for i, p in enumerate(probes):
        def fn(i=i, p=p, t=t):
            order.append(i)
            alive_at_run.append([r() is not None for r in refs])
            t.pg[i] = i
        t.back.append(fn)
    t.head_backward = lambda dout, t=t: order.append("head")
    del probes, p, fn
    gc.disable()
    try:
        pg = t.backward(None)
        assert order == ["head", 2, 1, 0] and pg == {0: 0, 1: 1, 2: 2}
        # a closure (and what it captured) is gone as soon as it has run: when closure 0 runs, probes 2 and 1 are already dead
        assert alive_at_run[-1] == [True, False, False]
        assert all(r() is None for r in refs)
        assert t.back == [] and t.head_backward is None and t.pg == {}
        wt = weakref.ref(t)
        del t
        assert wt() is None          # freed by reference count alone (the cycle collector is off)
    finally:
        gc.enable()


def test_detection_handoff_factories_and_errors():
    """models/detection.py drop-in surface (reference detection.py:7-139): unknown detector names raise ValueError, the
    integrated system freezes the detector's parameters, and the hand-off has no CPU path."""
    import torch.nn as nn
    from adam_dehaze_b200.models import detection as det
    with pytest.raises(ValueError, match="Unsupported detection model"):
        det.DetectionModel(model_name="yolo")
    with pytest.raises(ValueError):
        det.create_detection_model({"detection": {"model": "nope", "pretrained": False}})
    detector = nn.Linear(2, 2)
    system = det.create_integrated_system(nn.Identity(), detector)
    assert isinstance(system, det.IntegratedDetectionSystem)
    assert all(not p.requires_grad for p in system.detection_model.parameters())
    with pytest.raises(RuntimeError, match="no CPU path"):
        det.normalize_for_detection(torch.rand(1, 3, 8, 8))


def test_flat_adam_state_dict_is_interchangeable_with_torch_adam():
    """FlatAdam.state_dict()/load_state_dict() use torch.optim.Adam's layout, so 'optimizer_state_dict' in a checkpoint
    written by either side loads into the other (train_dehazing.py:196-203)."""
    from adam_dehaze_b200.training.optim import FlatAdam
    g = torch.Generator().manual_seed(9)
    shapes = [(4, 3, 3, 3), (4,), (7, 5)]
    ref_params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    ours_params = [torch.nn.Parameter(p.detach().clone()) for p in ref_params]
    adam = torch.optim.Adam(ref_params, lr=3e-4, weight_decay=1e-4)
    for _ in range(2):
        for p in ref_params:
            p.grad = torch.randn(p.shape, generator=g)
        adam.step()
    sd = adam.state_dict()
    flat = FlatAdam(ours_params, lr=1.0)
    assert flat.state_dict()["state"] == {}
    flat.load_state_dict(sd)
    assert flat.step_count == 2 and flat.lr == 3e-4 and flat.weight_decay == 1e-4 and tuple(flat.betas) == (0.9, 0.999)
    for i, (p, o) in enumerate(zip(ours_params, flat.offsets)):
        assert torch.equal(flat.exp_avg[o:o + p.numel()].view(p.shape), sd["state"][i]["exp_avg"])
        assert torch.equal(flat.exp_avg_sq[o:o + p.numel()].view(p.shape), sd["state"][i]["exp_avg_sq"])
    back = flat.state_dict()
    adam2 = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ref_params], lr=1.0)
    adam2.load_state_dict(back)
    st2 = adam2.state_dict()
    assert st2["param_groups"][0]["lr"] == 3e-4
    for i in range(len(shapes)):
        assert torch.equal(st2["state"][i]["exp_avg"], sd["state"][i]["exp_avg"]) and float(st2["state"][i]["step"]) == 2.0
    with pytest.raises(ValueError):
        FlatAdam([torch.nn.Parameter(torch.zeros(3))]).load_state_dict(sd)


def test_integration_doc_names_every_exported_entry_point():
    """INTEGRATION.md's entry-point table must account for every ADB_API symbol of include/adb200.h (names may be grouped as
    `adb_x_fwd/bwd`)."""
    hdr = open(os.path.join(ROOT, "include", "adb200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    declared = sorted(set(re.findall(r"ADB_API[^;]*?\b(adb_\w+)\s*\(", hdr)))
    missing = [s for s in declared
               if s not in doc and s.replace("_fwd", "_fwd/bwd") not in doc and not (s.endswith("_bwd") and s[:-4] + "_fwd/bwd" in doc)]
    assert not missing, missing


def test_bench_keeps_stdout_to_one_json_line_under_nccl(monkeypatch):
    """bench.quiet_nccl_stdout: NCCL honours NCCL_DEBUG_FILE only above the VERSION level, and prints its version line to
    stdout at VERSION — so VERSION is raised to WARN and the log file defaults to stderr; explicit settings are kept."""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setenv("NCCL_DEBUG", "VERSION")
    monkeypatch.delenv("NCCL_DEBUG_FILE", raising=False)
    bench.quiet_nccl_stdout()
    assert os.environ["NCCL_DEBUG"] == "WARN" and os.environ["NCCL_DEBUG_FILE"] == "/dev/stderr"
    monkeypatch.setenv("NCCL_DEBUG", "INFO")
    monkeypatch.setenv("NCCL_DEBUG_FILE", "/tmp/nccl.%h.%p.log")
    bench.quiet_nccl_stdout()
    assert os.environ["NCCL_DEBUG"] == "INFO" and os.environ["NCCL_DEBUG_FILE"] == "/tmp/nccl.%h.%p.log"
    monkeypatch.delenv("NCCL_DEBUG", raising=False)
    monkeypatch.delenv("NCCL_DEBUG_FILE", raising=False)
    bench.quiet_nccl_stdout()
    assert "NCCL_DEBUG" not in os.environ and os.environ["NCCL_DEBUG_FILE"] == "/dev/stderr"


def test_bench_clock_sampler_parses_nvidia_smi_lines():
    """The `clocks` object of the bench line: median SM clock under load, max clock, throttle reasons, peak power."""
    sys.path.insert(0, ROOT)
    import bench

    class Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    s = bench.ClockSampler(0)
    s.proc = Proc()
    s.lines = ["1650, 1965, 980.5, Not Active, Not Active, Not Active, Active",
               "1665, 1965, 990.1, Not Active, Not Active, Not Active, Active",
               "1950, 1965, 400.0, Not Active, Not Active, Not Active, Not Active",
               "garbage line", "N/A, 1965, 1.0, x, x, x, x"]
    out = s.stop()
    assert out["sm_mhz"] == 1665.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["power_w_max"] == 990.1 and out["samples"] == 3
