"""CPU restatement (numpy) of the reference's per-sample input pipeline, /root/reference/patchgan/io.py:38-58 -- TEST
INFRASTRUCTURE ONLY: imported by tests/, never by the product (patchgan_b200/io.py runs pg_prep_batch_u8 on the device).

    img    = read_image(jpg, RGB) / 255.                  io.py:42   uint8 (3,H,W) -> float32
    labels = read_image(png, GRAY) + 1                    io.py:43   uint8 arithmetic: 255 + 1 wraps to 0
    stacked = cat((img, labels)) -> Resize((S,S), antialias=None) [-> RandomHorizontalFlip, RandomVerticalFlip]   :46-49
    mask[i] = (labels == label_i)                         io.py:54-56  exact float equality on the INTERPOLATED labels

The resize is torchvision's Resize on a float tensor = aten::upsample_bilinear2d(align_corners=False, antialias=False);
its arithmetic lives in PyTorch (torch>=1.13, setup.py:35; 2.11 here) and is restated below: source index
max(scale * (dst + 0.5) - 0.5, 0) with scale = in / out in float32, lambda1 = src - floor(src), lambda0 = 1 - lambda1,
value = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11), each operation rounded to float32.

Pinned by tests/test_oracle_golden.py::test_input_pipeline_matches_reference against tests/golden/io.npz, which
tests/golden/make_golden.py records from the live reference's COCOStuffDataset on synthetic jpg / png files: masks
bit-exact, image within 1 ulp (ATen's vectorised kernel orders the image arithmetic differently in the last bit).
"""
import numpy as np

F32 = np.float32


def _index_weights(in_size, out_size):
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    real = np.maximum((scale * (dst + F32(0.5))).astype(F32) - F32(0.5), F32(0)).astype(F32)
    i0 = np.minimum(np.floor(real).astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.minimum((real - i0.astype(F32)).astype(F32), F32(1))
    l0 = (F32(1) - l1).astype(F32)
    return i0, i1, l0, l1


def resize_bilinear(x, out_h, out_w):
    """x: float32 (C, H, W) -> (C, out_h, out_w); identity when the size does not change (scale 1: lambda1 = 0)."""
    C, H, W = x.shape
    y0, y1, ly0, ly1 = _index_weights(H, out_h)
    x0, x1, lx0, lx1 = _index_weights(W, out_w)
    a, b = x[:, y0], x[:, y1]
    LX0, LX1 = lx0[None, None, :], lx1[None, None, :]
    LY0, LY1 = ly0[None, :, None], ly1[None, :, None]
    t0 = ((a[:, :, x0] * LX0).astype(F32) + (a[:, :, x1] * LX1).astype(F32)).astype(F32)
    t1 = ((b[:, :, x0] * LX0).astype(F32) + (b[:, :, x1] * LX1).astype(F32)).astype(F32)
    return ((t0 * LY0).astype(F32) + (t1 * LY1).astype(F32)).astype(F32)


def prepare_sample(img_u8, lab_u8, labels, size, flip=0):
    """img_u8: uint8 (3,H,W); lab_u8: uint8 (H,W); labels: requested label ids (sorted like io.py:17);
    size: (S, S) target; flip: bit 0 horizontal, bit 1 vertical.  Returns (img float32 (3,S,S), mask float32 (L,S,S))."""
    img = (img_u8.astype(F32) / F32(255.)).astype(F32)
    lab = (lab_u8.astype(np.uint8) + np.uint8(1)).astype(np.uint8).astype(F32)
    stacked = np.concatenate([img, lab[None]], axis=0)
    if stacked.shape[1:] != tuple(size):
        stacked = resize_bilinear(stacked, size[0], size[1])
    if flip & 1:
        stacked = stacked[:, :, ::-1]
    if flip & 2:
        stacked = stacked[:, ::-1, :]
    img, lab = stacked[:3], stacked[3]
    mask = np.stack([(lab == F32(v)) for v in np.sort(labels)]).astype(F32)
    return np.ascontiguousarray(img), mask
