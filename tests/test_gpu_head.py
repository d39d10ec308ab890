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


def test_head_vs_reference_golden(golden_dir):
    from shapy_b200 import ops
    g = np.load(os.path.join(golden_dir, 'ref_head.npz'))
    sd = {k: v.cuda() for k, v in synth.make_head_state_dict().items()}
    p = 'regressor.module.'
    out = ops.head_forward(torch.from_numpy(g['feats']).cuda(), sd[p + 'layer_000.0.weight'], sd[p + 'layer_000.0.bias'],
                           sd[p + 'layer_001.0.weight'], sd[p + 'layer_001.0.bias'], sd[p + 'output_layer.weight'],
                           sd[p + 'output_layer.bias'], sd['regressor.mean_param'].reshape(-1), 3)
    for k in range(3):
        assert rel(out[k], g['params'][k]) < 1e-5
        # the regressed update itself (params - mean), not just mean + small delta
        d = out[k].cpu() - synth.mean_params()
        dr = torch.from_numpy(g['params'][k]) - synth.mean_params()
        assert rel(d, dr) < 1e-4


@pytest.mark.parametrize('B', [1, 64, 97])
def test_head_vs_oracle(B):
    from shapy_b200 import ops
    sd = synth.make_head_state_dict(seed=4)
    feats = torch.randn(B, 2048, generator=torch.Generator().manual_seed(B)).abs()
    ref = net_oracle.head_forward(sd, feats)
    c = {k: v.cuda() for k, v in sd.items()}
    p = 'regressor.module.'
    out = ops.head_forward(feats.cuda(), c[p + 'layer_000.0.weight'], c[p + 'layer_000.0.bias'],
                           c[p + 'layer_001.0.weight'], c[p + 'layer_001.0.bias'], c[p + 'output_layer.weight'],
                           c[p + 'output_layer.bias'], c['regressor.mean_param'].reshape(-1), 3)
    for k in range(3):
        assert rel(out[k], ref[k]) < 1e-5


def test_collapsed_head_matches_reference_golden(golden_dir):
    """The affine-collapsed evaluation (default in the module) against the reference's own outputs."""
    from shapy_b200 import ops
    g = np.load(os.path.join(golden_dir, 'ref_head.npz'))
    sd = synth.make_head_state_dict()
    p = 'regressor.module.'
    MfT, Mp, c = ops.collapse_head(sd[p + 'layer_000.0.weight'], sd[p + 'layer_000.0.bias'], sd[p + 'layer_001.0.weight'],
                                   sd[p + 'layer_001.0.bias'], sd[p + 'output_layer.weight'], sd[p + 'output_layer.bias'], 2048)
    out = ops.head_forward_collapsed(torch.from_numpy(g['feats']).cuda(), MfT.cuda(), Mp.cuda(), c.cuda(),
                                     sd['regressor.mean_param'].reshape(-1).cuda(), 3)
    for k in range(3):
        assert rel(out[k], g['params'][k]) < 1e-5
        d = out[k].cpu() - synth.mean_params()
        dr = torch.from_numpy(g['params'][k]) - synth.mean_params()
        assert rel(d, dr) < 1e-4
