"""Input stage kernel (shapy_preprocess_forward through shapy_b200.preprocess.InputStage) against the reference's own
crops (tests/golden/preprocess.npz) and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_oracle as po

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'preprocess.npz'))
N = int(G['n'])
TOL = 2e-6     # the device compiler contracts a*b + c*d into FMAs: one more ulp than the CPU bar


def test_kernel_matches_reference_crops():
    from shapy_b200.preprocess import InputStage
    for size in sorted({int(G[f'size{i}']) for i in range(N)}):
        idx = [i for i in range(N) if int(G[f'size{i}']) == size]
        stage = InputStage('cuda', size=size, mean=G['mean'].tolist(), std=G['std'].tolist())
        images = [G[f'img{i}'] for i in idx]
        persons = [(k, G[f'center{i}'], float(G[f'scale{i}'])) for k, i in enumerate(idx)]
        out = stage(images, persons).cpu().numpy()
        for k, i in enumerate(idx):
            assert np.abs(out[k] - G[f'out{i}']).max() <= TOL, i


def test_several_people_in_one_image_at_network_resolution():
    from shapy_b200.preprocess import InputStage, IMAGENET_MEAN, IMAGENET_STD
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    other = rng.integers(0, 256, (333, 217, 3), dtype=np.uint8)
    persons = [(0, np.float32([320, 240]), 1.9), (0, np.float32([20, 30]), 0.8), (1, np.float32([100, 300]), 2.4),
               (0, np.float32([630, 470]), 1.1)]
    stage = InputStage('cuda', size=224)
    out = stage([img, other], persons)
    assert out.shape == (4, 3, 224, 224) and out.dtype == torch.float32
    for k, (j, c, s) in enumerate(persons):
        ref = po.preprocess([img, other][j], c, s, 224, IMAGENET_MEAN, IMAGENET_STD)
        assert np.abs(out[k].cpu().numpy() - ref).max() <= TOL, k
    # a second batch through the same stage reuses the staging buffers
    out2 = stage([other], [(0, np.float32([100, 300]), 2.4)])
    assert np.abs(out2[0].cpu().numpy() - po.preprocess(other, np.float32([100, 300]), 2.4, 224, IMAGENET_MEAN, IMAGENET_STD)).max() <= TOL
