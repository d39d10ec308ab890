"""HRNet-W48 through the C ABI against the reference module's own outputs on the seeded synthetic
checkpoint (tests/golden/ref_hrnet.npz) and against the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import net_oracle
from shapy_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope='module')
def backbone():
    return synth.build_synthetic_regressor().backbone


def _inputs():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 224, 224, generator=g)
    x64 = torch.randn(1, 3, 64, 96, generator=g)
    return x, x64


@pytest.mark.parametrize('engine', [1, 0])
def test_hrnet_vs_reference_golden(backbone, golden_dir, engine):
    g = np.load(os.path.join(golden_dir, 'ref_hrnet.npz'))
    x, x64 = _inputs()
    bb = backbone.cuda().eval()
    bb.engine, bb.precision_mode = engine, 1
    bb.invalidate()
    out = bb(x.cuda())
    assert rel(out['concat'], g['concat']) < 1e-4, rel(out['concat'], g['concat'])
    assert rel(out['layer1'][:, ::7, ::5, ::5], g['layer1_sub']) < 1e-4
    assert rel(out['layer4'][:, ::16], g['layer4']) < 1e-4
    out64 = bb(x64.cuda())                     # non-square, non-224 input: 16x24 ... 2x3 feature maps
    assert rel(out64['concat'], g['concat64']) < 1e-4


def test_hrnet_fp16_mode_close(backbone, golden_dir):
    """BASELINE config 2 numerics: single-pass fp16 operands (not the parity mode)."""
    g = np.load(os.path.join(golden_dir, 'ref_hrnet.npz'))
    x, _ = _inputs()
    bb = backbone.cuda().eval()
    bb.engine, bb.precision_mode = 0, 0
    bb.invalidate()
    out = bb(x.cuda())
    assert rel(out['concat'], g['concat']) < 3e-2
    bb.precision_mode = 1
    bb.invalidate()


def test_hrnet_batch_independence(backbone):
    """Images are independent: a permutation of the batch permutes the outputs bit for bit (same launch
    configuration), and a sub-batch matches to fp32 rounding (the tile configuration depends on the batch
    size, which changes the summation order)."""
    bb = backbone.cuda().eval()
    bb.engine, bb.precision_mode = 0, 1
    bb.invalidate()
    x = torch.randn(9, 3, 224, 224, generator=torch.Generator().manual_seed(5)).cuda()
    full = bb(x)['concat']
    perm = torch.tensor([4, 0, 8, 2, 6, 1, 7, 3, 5]).cuda()
    assert torch.equal(bb(x[perm].contiguous())['concat'], full[perm])
    part = bb(x[3:5].contiguous())['concat']
    assert rel(part, full[3:5]) < 1e-5
