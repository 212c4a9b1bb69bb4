"""ORACLE — test infrastructure only (see oracle/adam_oracle.py).  CPU restatement (numpy) of the per-sample input transform of
the reference loader, /root/reference/data/dataset.py:73-99: cv2.cvtColor(BGR2RGB) -> cv2.resize(img_size) ->
transforms.ToTensor(), plus the tensor flips of the training split (dataset.py:58-63).

cv2.resize is a third-party dependency (opencv-python, requirements.txt: opencv-python>=4.5; 4.13.0 in this image).  Its
INTER_LINEAR for 8-bit images is restated here from the published algorithm (opencv/modules/imgproc/src/resize.cpp) and
PINNED against cv2 itself: oracle/make_golden_input.py writes tests/golden/input_pipeline.pt from cv2 + torchvision, and
tests/test_oracle_golden.py::test_input_oracle_* compare (also live against cv2 when it is importable).
"""
import numpy as np


def resize_linear_u8(src, dh, dw):
    """cv2.resize(src, (dw, dh)) for uint8 HWC images, default interpolation (INTER_LINEAR), bit for bit.

    * exact 2x downscale in both axes -> OpenCV switches to the INTER_AREA fast path: (a+b+c+d+2) >> 2;
    * otherwise per axis f = float32((d+0.5)*scale - 0.5), s = floor(f), frac = f - s; shorts rint((1-frac)*2048),
      rint(frac*2048); columns collapse onto the edge pixel (frac = 0) when s is out of range, rows only clamp indices;
      horizontal pass in int32, vertical pass ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2."""
    sh, sw = src.shape[:2]
    if (sh, sw) == (dh, dw):
        return src.copy()
    s = src.astype(np.int32)
    if sh == 2 * dh and sw == 2 * dw:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def coeffs(dn, sn, vertical):
        scale = np.float64(sn) / np.float64(dn)
        f = ((np.arange(dn, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s0 = np.floor(f).astype(np.int64)
        fr = (f - s0.astype(np.float32)).astype(np.float32)
        if not vertical:
            lo = s0 < 0
            s0[lo] = 0
            fr[lo] = 0
            hi = s0 >= sn - 1
            s0[hi] = sn - 1
            fr[hi] = 0
        a0 = np.rint((np.float32(1.0) - fr) * np.float32(2048.0)).astype(np.int32)
        a1 = np.rint(fr * np.float32(2048.0)).astype(np.int32)
        return np.clip(s0, 0, sn - 1), np.clip(s0 + 1, 0, sn - 1), a0, a1
    x0, x1, ax0, ax1 = coeffs(dw, sw, False)
    y0, y1, by0, by1 = coeffs(dh, sh, True)
    rows = s[:, x0] * ax0[None, :, None] + s[:, x1] * ax1[None, :, None]
    r0, r1 = rows[y0], rows[y1]
    out = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def load_transform(img_bgr_u8, size=None, flip=0):
    """dataset.py:76-99 for one decoded image: BGR2RGB, resize to (size, size) if needed, ToTensor -> float32 [3, H, W];
    flip bit 0 = horizontal, bit 1 = vertical (RandomHorizontalFlip / RandomVerticalFlip on the tensor)."""
    rgb = img_bgr_u8[:, :, ::-1]
    if size is not None and (rgb.shape[0] != size[0] or rgb.shape[1] != size[1]):
        rgb = resize_linear_u8(np.ascontiguousarray(rgb), size[0], size[1])
    t = np.ascontiguousarray(rgb.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
    if flip & 1:
        t = t[:, :, ::-1]
    if flip & 2:
        t = t[:, ::-1, :]
    return np.ascontiguousarray(t)
