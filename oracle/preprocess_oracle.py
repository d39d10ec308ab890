"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input stage.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(shapy_b200.preprocess -> shapy_preprocess_forward) never does.

Reference chain for one detected person (regressor/human_shape):
    read_img                utils/img_utils.py:57-61      uint8 HxWx3 -> float32 / 255.0, clip [0, 1]
    Crop.__call__           data/transforms/transforms.py:522-541  -> crop(np_image, center, scale, [S, S])
    crop                    utils/transf_utils.py:51-96   window [ul, br) from transform(..., invert=1) (integer
                                                          truncation), zero padding outside the image,
                                                          cv2.resize(..., INTER_LINEAR) to S x S
    transform/get_transform utils/transf_utils.py:9-48    h = 200 * scale; 3x3 float32 matrix, np.linalg.inv
    ToTensor                HWC -> CHW
    Normalize.__call__      data/transforms/transforms.py:710-733  clamp [0, 1]; (x - mean) / std

Pinned: tests/golden/preprocess.npz holds outputs of the reference's own crop() + the Normalize arithmetic run in the
build container (cv2 4.13) on seeded images, written by tools/make_golden.py; tests/test_oracle_pins.py checks
this file against them (<= 1e-6 absolute before normalisation: cv2's SIMD paths differ from scalar float by an ulp).
"""
import numpy as np


def get_transform(center, scale, res):
    """transf_utils.py:9-36 with rot = 0."""
    h = 200 * scale
    t = np.zeros((3, 3), dtype=np.float32)
    t[0, 0] = float(res[1]) / h
    t[1, 1] = float(res[0]) / h
    t[0, 2] = res[1] * (-float(center[0]) / h + .5)
    t[1, 2] = res[0] * (-float(center[1]) / h + .5)
    t[2, 2] = 1
    return t.astype(np.float32)


def transform_inv(pt, center, scale, res):
    """transf_utils.py:41-48 with invert = 1."""
    t = np.linalg.inv(get_transform(center, scale, res))
    new_pt = np.array([pt[0] - 1, pt[1] - 1, 1.], dtype=np.float32).T
    new_pt = np.dot(t, new_pt)
    return new_pt[:2].astype(int) + 1


def crop_window(center, scale, res):
    """Upper-left (inclusive) and bottom-right (exclusive) corner of the source window, transf_utils.py:52-56."""
    ul = np.array(transform_inv([1, 1], center, scale, res)) - 1
    br = np.array(transform_inv([res[0] + 1, res[1] + 1], center, scale, res)) - 1
    return ul, br


def _axis(dst, src):
    """cv::resize INTER_LINEAR source index / weight per destination index (imgproc/resize.cpp, resizeGeneric_)."""
    inv_scale = float(dst) / float(src)          # double
    scale = 1.0 / inv_scale
    s0 = np.zeros(dst, np.int64)
    w1 = np.zeros(dst, np.float32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= src - 1:
            s, f = src - 1, np.float32(0)
        s0[d], w1[d] = s, f
    return s0, w1


def crop_resize(img_u8, ul, br, size):
    """float32 (size, size, 3): window [ul, br) of img_u8 / 255 with zero padding, bilinearly resized like cv2."""
    H, W = img_u8.shape[:2]
    src_w, src_h = int(br[0] - ul[0]), int(br[1] - ul[1])
    if src_w <= 0 or src_h <= 0:
        raise ValueError('empty crop window')
    win = np.zeros((src_h, src_w, 3), np.float32)
    x0, x1 = max(0, -ul[0]), min(br[0], W) - ul[0]
    y0, y1 = max(0, -ul[1]), min(br[1], H) - ul[1]
    if x1 > x0 and y1 > y0:
        src = img_u8[max(0, ul[1]):min(H, br[1]), max(0, ul[0]):min(W, br[0])].astype(np.float32) / np.float32(255.0)
        win[y0:y1, x0:x1] = np.clip(src, 0, 1)
    sx, fx = _axis(size, src_w)
    sy, fy = _axis(size, src_h)
    sx1, sy1 = np.minimum(sx + 1, src_w - 1), np.minimum(sy + 1, src_h - 1)
    a1 = fx[None, :, None]
    a0 = np.float32(1) - a1
    b1 = fy[:, None, None]
    b0 = np.float32(1) - b1
    r0 = win[sy][:, sx] * a0 + win[sy][:, sx1] * a1
    r1 = win[sy1][:, sx] * a0 + win[sy1][:, sx1] * a1
    return (r0 * b0 + r1 * b1).astype(np.float32)


def preprocess(img_u8, center, scale, size, mean, std):
    """(3, size, size) float32: the tensor the reference feeds to the network for one person."""
    ul, br = crop_window(center, scale, [size, size])
    c = crop_resize(img_u8, ul, br, size)
    c = np.clip(c, 0, 1).transpose(2, 0, 1)
    m = np.asarray(mean, np.float32)[:, None, None]
    s = np.asarray(std, np.float32)[:, None, None]
    return ((c - m) / s).astype(np.float32)
