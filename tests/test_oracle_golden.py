"""CPU: the oracle (oracle/adam_oracle.py) is pinned against the golden vectors produced from the UNMODIFIED reference
modules (oracle/make_golden.py), and the drop-in modules are shown to initialise bit-identically to the reference
(state_dict sha256 fingerprints), so `oracle(state_dict(drop-in), x)` is the reference's output for any x."""
import pytest
import torch

from helpers import CONFIG, SEED, fingerprint, golden, make_branch, make_classifier, rand_image

import adam_oracle as oracle

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name", ["low", "medium", "high", "low_unet", "corun", "dual_branch"])
def test_branch_oracle_matches_reference(name):
    g = golden(f"branch_{name}.pt")
    m = make_branch(name)
    sd = m.state_dict()
    assert list(sd.keys()) == g["keys"]                      # checkpoint layout
    assert fingerprint(sd) == g["fingerprint"]               # same seed -> same weights as the reference
    assert m.get_info() == g["info"]
    for case in g["cases"]:
        n, h, w = case["shape"]
        out = oracle.BRANCH_FORWARD[name](sd, rand_image(n, h, w, case["seed"]))
        assert torch.allclose(out, case["out"], rtol=0, atol=2e-6), (out - case["out"]).abs().max()
        if name == "low":
            assert 0 < out.min() and out.max() < 1           # Light output in (0,1) without a clamp
        else:
            assert 0 <= out.min() and out.max() <= 1


def test_blocks_oracle_matches_reference():
    from adam_dehaze_b200.models.dehazing.base_model import AttentionBlock, ResidualBlock
    g = golden("blocks.pt")
    torch.manual_seed(SEED)
    rb, ab = ResidualBlock(32).eval(), AttentionBlock(96).eval()
    assert fingerprint(rb.state_dict()) == g["res_fp"] and fingerprint(ab.state_dict()) == g["attn_fp"]
    xr = torch.randn(1, 32, 16, 24, generator=torch.Generator().manual_seed(5))
    xa = torch.randn(1, 96, 16, 24, generator=torch.Generator().manual_seed(6)).relu()
    sd_r = {"rb." + k: v for k, v in rb.state_dict().items()}
    sd_a = {"ab." + k: v for k, v in ab.state_dict().items()}
    assert torch.allclose(oracle.residual_block(sd_r, "rb", xr), g["res_out"], atol=2e-6)
    assert torch.allclose(oracle.attention_block(sd_a, "ab", xa), g["attn_out"], atol=2e-6)


def test_classifier_oracle_matches_reference():
    g = golden("classifier_resnet18.pt")
    clf = make_classifier("resnet18")
    sd = clf.state_dict()
    assert list(sd.keys()) == g["keys"] and fingerprint(sd) == g["fingerprint"]
    assert clf.feature_dim == g["feature_dim"] == 512 and clf.num_classes == 3 and clf.model_name == "resnet18"
    n, h, w = g["shape"]
    logits, feats = oracle.classifier_forward(sd, rand_image(n, h, w, g["seed"]))
    assert torch.allclose(logits, g["logits"], atol=2e-6) and torch.allclose(feats, g["features"], atol=2e-5)


def test_densenet121_oracle_matches_torchvision():
    """north_star HDEN arm: no reference code exists (parity unpinned); the oracle is pinned on torchvision itself."""
    clf = make_classifier("densenet121")
    assert clf.feature_dim == 1024
    x = rand_image(1, 64, 64, 3)
    ref_feats = clf.backbone(x)                               # torchvision forward, classifier = Identity
    logits, feats = oracle.classifier_forward(clf.state_dict(), x, "densenet121")
    assert torch.allclose(feats, ref_feats, atol=1e-5)
    assert logits.shape == (1, 3)


def _branch_sds():
    return {n: make_branch(n).state_dict() for n in ("low", "medium", "high")}


def test_routing_oracle_matches_reference():
    g = golden("routing.pt")
    sds = _branch_sds()
    clf_sd = make_classifier("resnet18").state_dict()
    n, h, w = g["shape"]
    x = rand_image(n, h, w, g["seed"])
    # natural run: random-init HDEN sends everything to one class
    logits, _ = oracle.classifier_forward(clf_sd, x)
    out, inten, _ = oracle.hard_route(sds, x, logits=logits)
    assert torch.equal(inten, g["natural_intensity"])
    assert torch.allclose(out, g["natural_out"], atol=2e-6)
    # crafted logits with exact ties: first max wins
    inten_c, buckets = oracle.route_indices(logits=g["crafted_logits"])
    assert torch.equal(inten_c, g["crafted_intensity"])
    for k in range(3):
        assert torch.equal(buckets[k], torch.nonzero(g["crafted_masks"][k]).flatten())
    out_c, _, _ = oracle.hard_route(sds, x, intensity=inten_c)
    assert torch.allclose(out_c, g["crafted_out"], atol=2e-6)
    blend, wts, _ = oracle.soft_route(sds, x, g["crafted_logits"], 0.5)
    assert torch.allclose(wts, g["soft_weights"], atol=1e-7) and torch.allclose(blend, g["soft_out"], atol=2e-6)
    # router checkpoint layout: classifier.* and models.{low,medium,high}.* prefixes
    from adam_dehaze_b200.models.routing import create_router
    router = create_router({n_: make_branch(n_) for n_ in ("low", "medium", "high")}, make_classifier(), CONFIG)
    assert list(router.state_dict().keys()) == g["router_keys"]


def test_losses_oracle_matches_reference():
    import torchvision.models as tvm
    g = golden("losses.pt")
    torch.manual_seed(SEED)
    vgg_sd = tvm.vgg16(weights=None).features.state_dict()
    assert fingerprint(vgg_sd) == g["vgg_fingerprint"]
    n, h, w = g["shape"]
    pred, tgt = rand_image(n, h, w, g["pred_seed"]), rand_image(n, h, w, g["target_seed"])
    content = oracle.content_loss(vgg_sd, pred, tgt)
    assert torch.allclose(content, g["content"], rtol=1e-5)
    # the golden run stubbed LPIPS with mean((x-t)^2) on the [-1,1] images (oracle/make_golden.py)
    stub = (((2 * pred - 1) - (2 * tgt - 1)) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    total, parts = oracle.dehazing_loss(pred, tgt, content, stub)
    assert torch.allclose(total, g["dehazing_total"], rtol=1e-5)
    for k in ("l1", "content", "perceptual"):
        assert torch.allclose(parts[k], g["dehazing_parts"][k], rtol=1e-5)
    jt, jp = oracle.joint_loss(total, g["joint_logits"], g["joint_labels"])
    assert torch.allclose(jt, g["joint_total"], rtol=1e-5) and torch.allclose(jp["classification"], g["joint_ce"], rtol=1e-6)


def test_lpips_restatement_shape_and_properties():
    """LPIPS is unpinned (package absent): check the restated definition's invariants only."""
    import torchvision.models as tvm
    torch.manual_seed(0)
    alex = tvm.alexnet(weights=None).features.state_dict()
    lins = [torch.rand(1, c, 1, 1) for c in (64, 192, 384, 256, 256)]
    a, b = rand_image(2, 64, 64, 1), rand_image(2, 64, 64, 2)
    d_ab = oracle.perceptual_lpips(alex, lins, a, b)
    assert d_ab.shape == (2, 1, 1, 1) and (d_ab > 0).all()
    assert torch.allclose(oracle.perceptual_lpips(alex, lins, a, a), torch.zeros(2, 1, 1, 1))
    assert torch.allclose(d_ab, oracle.perceptual_lpips(alex, lins, b, a), rtol=1e-5)


def test_synth_hazy_contract():
    hazy, clear, labels = oracle.synth_hazy(6, 32, 64)
    assert hazy.shape == clear.shape == (6, 3, 32, 64) and labels.tolist() == [0, 1, 2, 0, 1, 2]
    assert 0 <= hazy.min() and hazy.max() <= 1
    # heavier beta -> closer to the airlight 0.8 everywhere
    assert (hazy[2] - 0.8).abs().mean() < (hazy[0] - 0.8).abs().mean()


@pytest.mark.parametrize("name", ["low", "medium", "high"])
def test_train_mode_oracle_matches_reference(name):
    """oracle.train_mode() (batch-statistics BatchNorm) + torch autograd reproduces one training step of the unmodified
    reference: output, L1 loss and every parameter gradient (fixtures: oracle/make_golden_train.py)."""
    g = golden(f"train_{name}.pt")
    m = make_branch(name)
    assert fingerprint(m.state_dict()) == g["fingerprint"]
    n, h, w = g["shape"]
    x, tgt = rand_image(n, h, w, g["x_seed"]), rand_image(n, h, w, g["target_seed"])
    fwd = {"low": oracle.light_forward, "medium": oracle.medium_forward, "high": oracle.complex_forward}[name]
    with torch.enable_grad():
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
        with oracle.train_mode():
            out = fwd(sd, x)
        loss = (out - tgt).abs().mean()
        names = [k for k, _ in m.named_parameters()]
        grads = dict(zip(names, torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)))
    assert torch.allclose(out, g["out"], rtol=0, atol=5e-6), (out - g["out"]).abs().max()
    assert abs(loss.item() - g["loss"].item()) <= 1e-6
    for k in names:
        ref = g["grads"][k]
        got = grads[k]
        if isinstance(ref, dict):
            assert tuple(got.shape) == ref["shape"]
            scale = max(ref["norm"], 1e-9)
            assert abs(got.norm().item() - ref["norm"]) <= 2e-3 * scale + 1e-9, k
            assert (got.flatten()[:8] - ref["head"]).abs().max().item() <= 2e-3 * max(ref["head"].abs().max().item(), scale / got.numel() ** 0.5) + 1e-9, k
        else:
            assert (got - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-9, k


def test_detection_handoff_oracle_matches_reference():
    """oracle.detection_normalize == what the unmodified IntegratedDetectionSystem hands its detector (detection.py:109-121)."""
    g = golden("detection_handoff.pt")
    n, h, w = g["shape"]
    x = rand_image(n, h, w, g["seed"])
    dehazed = x * 0.5 + 0.25
    assert torch.equal(dehazed, g["dehazed"])
    got = torch.stack(oracle.detection_normalize(dehazed))
    assert torch.equal(got, g["normalized"])



def test_input_oracle_matches_cv2_golden_vectors():
    """oracle/input_oracle.py (restatement of cv2.cvtColor + cv2.resize(INTER_LINEAR, uint8) + ToTensor, dataset.py:76-99) against
    tests/golden/input_pipeline.pt, which oracle/make_golden_input.py produced with cv2 + torchvision themselves: bit-exact."""
    import numpy as np
    import input_oracle
    g = golden("input_pipeline.pt")
    assert len(g["cases"]) >= 8
    for case in g["cases"]:
        bgr = case["bgr"].numpy()
        size = (case["size"], case["size"])
        for key, flip in (("tensor", 0), ("hflip", 1), ("vflip", 2)):
            got = input_oracle.load_transform(bgr, size, flip)
            assert got.dtype == np.float32 and np.array_equal(got, case[key].numpy()), (bgr.shape, size, key)


def test_input_oracle_matches_live_cv2_when_present():
    """Wider sweep against the installed cv2 (skipped where cv2 is absent, e.g. a slim box): up/down-scales, the exact-2x
    INTER_AREA shortcut, odd sizes, single rows/columns."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    import input_oracle
    rng = np.random.default_rng(7)
    for (sh, sw, dh, dw) in [(480, 640, 256, 256), (100, 130, 256, 256), (512, 512, 256, 256), (255, 257, 256, 256), (64, 64, 256, 256),
                             (333, 777, 512, 512), (1, 50, 32, 32), (50, 1, 32, 32), (31, 33, 64, 16), (1024, 2048, 512, 512)]:
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(input_oracle.resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh))), (sh, sw, dh, dw)
