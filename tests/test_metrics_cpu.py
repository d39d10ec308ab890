"""Point-to-point evaluation metric, CPU side: the numpy restatement, the module mirror and the
host-compiled copy of the kernels' per-point functions against outputs of the REFERENCE's own v2vhdError
(tests/golden/p2p.npz, written by tools/make_golden.py)."""
import ctypes as C
import os
import pickle

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import build_oracle, metrics_oracle as mo

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'p2p.npz'))
IN = (G['in_row_ptr'], G['in_col'], G['in_val'])
TG = (G['tg_row_ptr'], G['tg_col'], G['tg_val'])


@pytest.mark.parametrize('align,tag', [(True, 'aligned'), (False, 'raw')])
def test_oracle_matches_reference_metric(align, tag):
    mean, err = mo.p2p_error(IN, TG, G['v_in'], G['v_tg'], align)
    assert np.abs(mean - G[f'mean_{tag}']).max() / G[f'mean_{tag}'].max() < 1e-9
    assert np.abs(err - G[f'error_{tag}']).max() / G[f'error_{tag}'].max() < 1e-7


@pytest.mark.parametrize('align,tag', [(1, 'aligned'), (0, 'raw')])
def test_kernel_point_functions_compiled_for_host(align, tag):
    lib = C.CDLL(build_oracle.build_metrics_host())
    lib.p2p_host.argtypes = [C.c_void_p] * 8 + [C.c_int] * 5 + [C.c_void_p] * 2
    arrs = [np.ascontiguousarray(G[k]) for k in ('in_row_ptr', 'in_col', 'in_val', 'tg_row_ptr', 'tg_col', 'tg_val', 'v_in', 'v_tg')]
    B, V1 = G['v_in'].shape[:2]
    V2, P = G['v_tg'].shape[1], len(G['in_row_ptr']) - 1
    err, mean = np.empty((B, P), np.float32), np.empty(B, np.float32)
    lib.p2p_host(*[a.ctypes.data for a in arrs], B, P, V1, V2, align, err.ctypes.data, mean.ctypes.data)
    assert np.abs(mean - G[f'mean_{tag}']).max() / G[f'mean_{tag}'].max() < 1e-6
    assert np.abs(err - G[f'error_{tag}']).max() / G[f'error_{tag}'].max() < 1e-6


def regressor_files(tmp_path):
    paths = []
    for name, (rp, col, val), V in (('in', IN, G['v_in'].shape[1]), ('tg', TG, G['v_tg'].shape[1])):
        m = sp.csr_matrix((val.astype(np.float64), col, rp), shape=(len(rp) - 1, V))
        paths.append(str(tmp_path / f'{name}.pkl'))
        with open(paths[-1], 'wb') as f:
            pickle.dump(m, f)
    return paths


def test_module_mirror_buffers_and_cpu_refusal(tmp_path):
    from shapy_b200 import metrics
    m = metrics.v2vhdError(*regressor_files(tmp_path), align=True)
    assert sorted(m.state_dict()) == ['input_point_regressor', 'target_point_regressor']       # metrics.py:392-396
    assert m.input_point_regressor.is_sparse and tuple(m.input_point_regressor.shape) == (len(IN[0]) - 1, G['v_in'].shape[1])
    with pytest.raises(RuntimeError):
        m(torch.from_numpy(G['v_in']), torch.from_numpy(G['v_tg']))


@pytest.mark.parametrize('align,tag', [(True, 'aligned'), (False, 'raw')])
def test_v2v_oracle_matches_reference_point_error(align, tag):
    """metrics_oracle.v2v_error against the reference's own PointError(TranslationAlignment | NoAlignment) outputs
    (tests/golden/v2v.npz, tools/make_golden.py v2v)."""
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'v2v.npz'))
    err = mo.v2v_error(g['est'], g['gt'], align)
    assert err.shape == g[f'error_{tag}'].shape
    assert np.abs(err - g[f'error_{tag}']).max() / g[f'error_{tag}'].max() < 5e-6     # the reference works in float32 here: eps * |coordinate| / error
