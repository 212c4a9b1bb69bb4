"""GPU: adb_image_u8_to_f32 (csrc/input_pipe.cu) against the cv2/torchvision golden vectors and the numpy oracle — bit-exact
(integer resize arithmetic, one IEEE division).  Reference: data/dataset.py:73-99."""
import numpy as np
import pytest
import torch

from helpers import golden

import input_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _sync():
    yield
    from adam_dehaze_b200 import _lib
    torch.cuda.synchronize()
    _lib.call("adb_kernel_error_flag")


def test_input_kernel_matches_cv2_golden_vectors():
    from adam_dehaze_b200.data.pipeline import u8_to_tensor
    g = golden("input_pipeline.pt")
    for case in g["cases"]:
        src = case["bgr"].cuda().unsqueeze(0).contiguous()
        size = (case["size"], case["size"])
        out = u8_to_tensor(src, size=size)
        assert torch.equal(out[0].cpu(), case["tensor"]), tuple(case["bgr"].shape)
        for key, bit in (("hflip", 1), ("vflip", 2)):
            fl = torch.tensor([bit], dtype=torch.uint8, device="cuda")
            assert torch.equal(u8_to_tensor(src, size=size, flips=fl)[0].cpu(), case[key])


@pytest.mark.parametrize("sh,sw,dh,dw", [(480, 640, 256, 256), (100, 130, 256, 256), (512, 512, 256, 256), (255, 257, 256, 256),
                                         (333, 777, 512, 512), (1, 50, 32, 32), (50, 1, 32, 32), (31, 33, 64, 16), (720, 1280, 512, 512)])
def test_input_kernel_matches_oracle_batch(sh, sw, dh, dw):
    from adam_dehaze_b200.data.pipeline import u8_to_tensor
    rng = np.random.default_rng(sh * 7 + sw)
    imgs = rng.integers(0, 256, (3, sh, sw, 3), dtype=np.uint8)
    flips = [0, 3, 1]
    out = u8_to_tensor(torch.from_numpy(imgs).cuda(), size=(dh, dw), flips=torch.tensor(flips, dtype=torch.uint8, device="cuda"))
    for i in range(3):
        want = input_oracle.load_transform(imgs[i], (dh, dw), flips[i])
        assert np.array_equal(out[i].cpu().numpy(), want), i


def test_input_kernel_full_resolution_properties():
    """BASELINE size (1024 x 2048), no oracle needed: same-size conversion is exactly uint8 / 255 with the channels swapped;
    a double flip is the identity; a 2048 x 4096 source halves by the (a+b+c+d+2)>>2 rule."""
    from adam_dehaze_b200.data.pipeline import u8_to_tensor
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.randint(0, 256, (2, 1024, 2048, 3), generator=g, device="cuda", dtype=torch.uint8)
    out = u8_to_tensor(src)
    # (reference values on the CPU: torch's CUDA division by a scalar multiplies by the reciprocal, ToTensor on the host divides)
    want = src.cpu().flip(3).permute(0, 3, 1, 2).float() / 255.0
    assert torch.equal(out.cpu(), want)
    both = torch.full((2,), 3, dtype=torch.uint8, device="cuda")
    assert torch.equal(u8_to_tensor(src, flips=both).flip(2, 3), out)
    big = torch.randint(0, 256, (1, 2048, 4096, 3), generator=g, device="cuda", dtype=torch.uint8)
    half = u8_to_tensor(big, size=(1024, 2048), bgr=False)
    b = big.int()
    area = ((b[:, 0::2, 0::2] + b[:, 0::2, 1::2] + b[:, 1::2, 0::2] + b[:, 1::2, 1::2] + 2) >> 2).permute(0, 3, 1, 2).cpu().float() / 255.0
    assert torch.equal(half.cpu(), area)


def test_device_loader_reads_a_dataset_directory(tmp_path):
    """HazyImageFolder + DeviceLoader on a directory laid out like the reference's dataset (dataset.py:21-56): files are
    decoded by cv2 on the host, uploaded as uint8 and transformed on the device; batches equal the reference transform."""
    cv2 = pytest.importorskip("cv2")
    from adam_dehaze_b200.data.pipeline import DeviceLoader, HazyImageFolder
    rng = np.random.default_rng(1)
    truth = {}
    for level, shape in (("low", (40, 56)), ("medium", (32, 32)), ("high", (64, 64))):
        for kind in ("hazy", "clear", "dehazed"):
            d = tmp_path / "val" / level / kind
            d.mkdir(parents=True)
            for i in range(2):
                img = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
                cv2.imwrite(str(d / f"{level}{i}.png"), img)
                truth[(level, kind, i)] = img
    folder = HazyImageFolder(str(tmp_path), split="val", img_size=32)
    assert len(folder) == 6
    loader = DeviceLoader(folder, batch_size=4, augment=False, workers=2)
    seen = 0
    for batch in loader:
        assert batch["hazy"].is_cuda and batch["hazy"].shape[1:] == (3, 32, 32) and batch["intensity"].dtype == torch.long
        for j, name in enumerate(batch["name"]):
            level, i = name[:-5], int(name[-5])
            assert batch["intensity"][j].item() == {"low": 0, "medium": 1, "high": 2}[level]
            for kind in ("hazy", "clear", "dehazed"):
                want = input_oracle.load_transform(truth[(level, kind, i)], (32, 32))
                assert np.array_equal(batch[kind][j].cpu().numpy(), want), (name, kind)
            seen += 1
    assert seen == 6
