"""On-GPU input stage: uint8 full images + (center, scale) per person -> the normalised fp32 crops of the network.

Host mirror of the reference's per-person CPU chain:

    read_img     regressor/human_shape/utils/img_utils.py:57-61       uint8 -> float32 / 255
    Crop         regressor/human_shape/data/transforms/transforms.py:522-541 -> crop(image, center, scale, [S, S])
    crop         regressor/human_shape/utils/transf_utils.py:51-96    integer window, zero padding, cv2.resize bilinear
    Normalize    regressor/human_shape/data/transforms/transforms.py:710-733

The window corners are host logic and follow transf_utils.py:9-56 operation by operation (float32 3x3 matrix,
np.linalg.inv, integer truncation) so they are the same integers the reference computes; the per-pixel work runs in
`shapy_preprocess_forward` (csrc/preprocess.cu).  There is no CPU fallback for the pixels.
"""
import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # regressor/human_shape/config/datasets_defaults.py:37-38
IMAGENET_STD = (0.229, 0.224, 0.225)


def get_transform(center, scale, res) -> np.ndarray:
    """transf_utils.py:9-36 (rot = 0): crop-space <- image-space affine map, float32."""
    h = 200 * scale
    t = np.zeros((3, 3), dtype=np.float32)
    t[0, 0] = float(res[1]) / h
    t[1, 1] = float(res[0]) / h
    t[0, 2] = res[1] * (-float(center[0]) / h + .5)
    t[1, 2] = res[0] * (-float(center[1]) / h + .5)
    t[2, 2] = 1
    return t.astype(np.float32)


def _transform_inv(pt, center, scale, res):
    """transf_utils.py:41-48 with invert=1: 1-based crop pixel -> 1-based image pixel (truncated)."""
    t = np.linalg.inv(get_transform(center, scale, res))
    new_pt = np.array([pt[0] - 1, pt[1] - 1, 1.], dtype=np.float32).T
    new_pt = np.dot(t, new_pt)
    return new_pt[:2].astype(int) + 1


def crop_window(center, scale, size: int):
    """(ul, br): upper-left (inclusive) and bottom-right (exclusive) image pixel of the window, transf_utils.py:52-56."""
    if not (np.isfinite(scale) and scale > 0):
        raise ValueError(f'crop scale must be positive and finite, got {scale}')
    res = [size, size]
    ul = np.array(_transform_inv([1, 1], center, scale, res)) - 1
    br = np.array(_transform_inv([res[0] + 1, res[1] + 1], center, scale, res)) - 1
    if br[0] <= ul[0] or br[1] <= ul[1]:
        raise ValueError(f'empty crop window for center {center}, scale {scale}')   # the reference fails inside cv2.resize
    return ul, br


class InputStage:
    """Packs a batch of uint8 images and crop requests into pinned staging buffers and runs the device kernel.

    images : sequence of HxWx3 uint8 arrays (numpy or torch, RGB as read_img returns them)
    persons: sequence of (image_index, center(x, y), scale) -- `center` / `scale` as produced by the reference's
             bbox_to_center_scale (scale = bbox size / 200)
    """

    def __init__(self, device, size: int = 224, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('shapy_b200.preprocess.InputStage needs a CUDA device (there is no CPU path)')
        self.size = int(size)
        self.mean = (C.c_float * 3)(*[float(m) for m in mean])
        self.std = (C.c_float * 3)(*[float(s) for s in std])
        self._host = None      # pinned uint8 staging buffer
        self._hdesc = None     # pinned descriptor table
        self._copied = None    # event: the previous batch's async H2D copies have read the staging buffers

    def pack(self, images, persons):
        """Returns (pinned image bytes, pinned descriptor table as uint8) for one batch."""
        if self._copied is not None:
            self._copied.synchronize()         # the staging buffers are about to be overwritten
        offs, total = [], 0
        arrs = []
        for im in images:
            a = im.numpy() if torch.is_tensor(im) else np.asarray(im)
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError('InputStage: images must be HxWx3 uint8')
            arrs.append(np.ascontiguousarray(a))
            offs.append(total)
            total += a.size
        if self._host is None or self._host.numel() < total:
            self._host = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        hb = self._host.numpy()
        for a, o in zip(arrs, offs):
            hb[o:o + a.size] = a.reshape(-1)
        n = len(persons)
        nbytes = n * C.sizeof(_lib.ImageDesc)
        if self._hdesc is None or self._hdesc.numel() < nbytes:
            self._hdesc = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        table = (_lib.ImageDesc * n).from_buffer(self._hdesc.numpy())
        for i, (idx, center, scale) in enumerate(persons):
            ul, br = crop_window(center, scale, self.size)
            H, W = arrs[idx].shape[:2]
            table[i].offset, table[i].height, table[i].width = offs[idx], H, W
            table[i].ul_x, table[i].ul_y, table[i].br_x, table[i].br_y = int(ul[0]), int(ul[1]), int(br[0]), int(br[1])
        return self._host[:total], self._hdesc[:nbytes]

    def uniform_table(self, n: int, height: int, width: int, center=None, scale=None) -> torch.Tensor:
        """Pinned descriptor table (as uint8) for `n` images of one size stored back to back, one person per image
        (default: the window of a detection that covers the whole image) -- the layout a serving loop or the
        rank-0 scatter of `shapy_b200.dist` ships."""
        center = (width / 2.0, height / 2.0) if center is None else center
        scale = max(height, width) / 200.0 if scale is None else scale
        ul, br = crop_window(center, scale, self.size)
        nbytes = n * C.sizeof(_lib.ImageDesc)
        buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        table = (_lib.ImageDesc * n).from_buffer(buf.numpy())
        for i in range(n):
            table[i].offset, table[i].height, table[i].width = i * height * width * 3, height, width
            table[i].ul_x, table[i].ul_y, table[i].br_x, table[i].br_y = int(ul[0]), int(ul[1]), int(br[0]), int(br[1])
        return buf

    def run_device(self, dimg: torch.Tensor, ddesc: torch.Tensor, n: int, out: torch.Tensor = None) -> torch.Tensor:
        """Device-resident entry: `dimg` uint8 image bytes, `ddesc` the descriptor table (uint8 view), both already on
        the device; launches the kernel on the current stream."""
        if not (dimg.is_cuda and ddesc.is_cuda and dimg.dtype == torch.uint8):
            raise RuntimeError('InputStage.run_device: device uint8 tensors expected (there is no CPU path)')
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(n, 3, self.size, self.size, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().shapy_preprocess_forward(_lib.ptr(dimg), _lib.ptr(ddesc), n, self.size, self.mean, self.std,
                                                           _lib.ptr(out), _lib.stream_ptr()), 'preprocess_forward')
        return out

    def __call__(self, images, persons) -> torch.Tensor:
        """(len(persons), 3, size, size) fp32 on the device, on the current stream."""
        if len(persons) == 0 or len(images) == 0:
            raise ValueError('InputStage: empty batch')
        hb, hd = self.pack(images, persons)
        n = len(persons)
        with torch.cuda.device(self.device):
            dimg = hb.to(self.device, non_blocking=True)
            ddesc = hd.to(self.device, non_blocking=True)
            self._copied = torch.cuda.Event()
            self._copied.record()
            out = torch.empty(n, 3, self.size, self.size, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().shapy_preprocess_forward(_lib.ptr(dimg), _lib.ptr(ddesc), n, self.size, self.mean, self.std,
                                                           _lib.ptr(out), _lib.stream_ptr()), 'preprocess_forward')
        return out
