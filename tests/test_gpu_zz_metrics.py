"""P2P evaluation metric kernels (shapy_p2p_error through shapy_b200.metrics.v2vhdError) against the reference's outputs."""
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'p2p.npz'))


def regressor_files(tmp_path):
    paths = []
    for name, V in (('in', G['v_in'].shape[1]), ('tg', G['v_tg'].shape[1])):
        rp, col, val = G[f'{name}_row_ptr'], G[f'{name}_col'], G[f'{name}_val']
        m = sp.csr_matrix((val.astype(np.float64), col, rp), shape=(len(rp) - 1, V))
        paths.append(str(tmp_path / f'{name}.pkl'))
        with open(paths[-1], 'wb') as f:
            pickle.dump(m, f)
    return paths


@pytest.mark.parametrize('align,tag', [(True, 'aligned'), (False, 'raw')])
def test_metric_matches_reference(tmp_path, align, tag):
    from shapy_b200 import metrics
    m = metrics.v2vhdError(*regressor_files(tmp_path), align=align)
    mean, err = m(torch.from_numpy(G['v_in']).cuda(), torch.from_numpy(G['v_tg']).cuda())
    mean, err = mean.cpu().numpy(), err.cpu().numpy()
    assert err.shape == G[f'error_{tag}'].shape
    # the reference works in float64; fp32 points of magnitude ~1 give errors to ~1e-6 relative
    assert np.abs(mean - G[f'mean_{tag}']).max() / G[f'mean_{tag}'].max() < 1e-5
    assert np.abs(err - G[f'error_{tag}']).max() / G[f'error_{tag}'].max() < 1e-5


@pytest.mark.parametrize('name,tag', [('translation', 'aligned'), ('none', 'raw')])
def test_v2v_point_error_matches_reference(name, tag):
    """shapy_v2v_error through the PointError mirror against the reference's PointError outputs (`v2v_t` / `v2v`)."""
    from shapy_b200 import metrics
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'v2v.npz'))
    pe = metrics.PointError(metrics.build_alignment(name))
    err = pe(torch.from_numpy(g['est']).cuda(), torch.from_numpy(g['gt']).cuda())
    ref = g[f'error_{tag}']
    assert err.shape == ref.shape
    assert np.abs(err.cpu().numpy() - ref).max() / ref.max() < 2e-5      # both sides fp32: eps * |coordinate| / error
    assert np.abs(pe.last_mean.cpu().numpy() - ref.mean(1)).max() / ref.mean(1).max() < 2e-5
    with pytest.raises(NotImplementedError):
        metrics.build_alignment('procrustes')
