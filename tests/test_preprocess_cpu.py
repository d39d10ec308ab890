"""Input stage, CPU side: the numpy restatement, the host mirror of the window arithmetic and the
host-compiled copy of the kernel's per-pixel function against crops produced by the REFERENCE's own
transf_utils.crop + ToTensor + Normalize (tests/golden/preprocess.npz, written by tools/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import build_oracle, preprocess_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preprocess.npz'))
N = int(G['n'])
TOL = 1e-6      # absolute, on normalised values of magnitude <= 2.7: one float32 ulp of slack for summation order


def case(i):
    return G[f'img{i}'], G[f'center{i}'], float(G[f'scale{i}']), int(G[f'size{i}'])


@pytest.mark.parametrize('i', range(N))
def test_oracle_matches_reference_crops(i):
    img, center, scale, size = case(i)
    out = po.preprocess(img, center, scale, size, G['mean'], G['std'])
    assert out.shape == (3, size, size) and out.dtype == np.float32
    assert np.abs(out - G[f'out{i}']).max() <= TOL
    # the container's IPP-accelerated cv2 build deviates from OpenCV's portable code by this much on the same crop
    assert float(G[f'ipp_dev{i}']) < 5e-5


@pytest.mark.parametrize('i', range(N))
def test_host_mirror_window_is_the_reference_window(i):
    from shapy_b200 import preprocess
    img, center, scale, size = case(i)
    ul, br = preprocess.crop_window(center, scale, size)
    assert (ul == G[f'ul{i}']).all() and (br == G[f'br{i}']).all()


def test_host_mirror_rejects_empty_window_and_cpu():
    from shapy_b200 import preprocess
    with pytest.raises(ValueError):
        preprocess.crop_window(np.float32([10, 10]), 0.0, 224)
    with pytest.raises(RuntimeError):
        preprocess.InputStage('cpu')


@pytest.mark.parametrize('i', range(N))
def test_kernel_pixel_function_compiled_for_host(i):
    """shapy_b200/csrc/preprocess.cuh is the code the GPU runs; oracle/preprocess_host.cpp compiles it for the CPU."""
    lib = C.CDLL(build_oracle.build_preprocess_host())
    lib.preprocess_host.argtypes = [C.c_void_p] + [C.c_int] * 7 + [C.c_void_p] * 3
    img, center, scale, size = case(i)
    img = np.ascontiguousarray(img)
    ul, br = G[f'ul{i}'], G[f'br{i}']
    mean, std = np.ascontiguousarray(G['mean']), np.ascontiguousarray(G['std'])
    out = np.empty((3, size, size), np.float32)
    lib.preprocess_host(img.ctypes.data, img.shape[0], img.shape[1], int(ul[0]), int(ul[1]), int(br[0]), int(br[1]), size,
                        mean.ctypes.data, std.ctypes.data, out.ctypes.data)
    assert np.abs(out - G[f'out{i}']).max() <= TOL
    assert np.array_equal(out, po.preprocess(img, center, scale, size, mean, std))
