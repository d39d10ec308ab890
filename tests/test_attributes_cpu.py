"""B2A attribute head, CPU side: the numpy restatement, the host mirror (module names, checkpoint
loading, gender routing codes) and the host-compiled copy of the kernel's output function against outputs of the
REFERENCE's own Polynomial module (tests/golden/b2a.npz, written by tools/make_golden.py)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import attributes_oracle as ao, build_oracle

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'b2a.npz'))
GENDERS = [None if x == '<none>' else str(x) for x in G['genders']]
TOL = 1e-6      # relative to the largest rating: fp32 GEMM summation order


class Target:
    """Minimal stand-in for the reference's target structures (has_field / get_field)."""

    def __init__(self, gender):
        self.gender = gender

    def has_field(self, name):
        return name == 'gender' and self.gender is not None

    def get_field(self, name):
        return self.gender


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def test_oracle_matches_reference_module():
    out = ao.b2a_by_gender(G['betas'], GENDERS, (G['w_male'], G['b_male']), (G['w_female'], G['b_female']))
    assert rel(out, G['attributes']) < TOL
    assert (out[[3, 4, 6]] == 0).all()            # '', None, 'neutral' get a row of zeros
    ref_idx = [tuple(x) for x in G['indices_000'].tolist()] + [tuple(x) for x in G['indices_001'].tolist()]
    assert ref_idx == ao.feature_indices(10)


def test_kernel_output_function_compiled_for_host():
    from shapy_b200 import attributes
    lib = C.CDLL(build_oracle.build_attributes_host())
    lib.b2a_host.argtypes = [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p]
    codes = attributes.gender_codes([Target(g) for g in GENDERS], len(GENDERS))
    assert codes.tolist() == [0, 1, 1, 2, 2, 0, 2, 1, 0]
    arrs = [np.ascontiguousarray(G['betas'], np.float32), codes] + [np.ascontiguousarray(G[k], np.float32)
                                                                    for k in ('w_male', 'b_male', 'w_female', 'b_female')]
    out = np.empty((len(GENDERS), 15), np.float32)
    lib.b2a_host(*[a.ctypes.data for a in arrs], len(GENDERS), 10, 15, out.ctypes.data)
    assert rel(out, G['attributes']) < TOL


def test_mirror_names_and_lightning_checkpoint_loading(tmp_path):
    from shapy_b200 import attributes
    m = attributes.B2A(10, 15)
    assert sorted(m.state_dict()) == ['b2a.indices_000', 'b2a.indices_001', 'b2a.linear.bias', 'b2a.linear.weight']
    assert torch.equal(m.b2a.indices_001, torch.from_numpy(G['indices_001']))
    sd = {'b2a.linear.weight': torch.from_numpy(G['w_male']), 'b2a.linear.bias': torch.from_numpy(G['b_male']),
          'b2a.indices_000': torch.from_numpy(G['indices_000']), 'b2a.indices_001': torch.from_numpy(G['indices_001'])}
    path = str(tmp_path / 'males.ckpt')
    torch.save({'state_dict': sd, 'hyper_parameters': {'cfg': {'num_shape_comps': 10}}}, path)   # Lightning layout
    loaded = attributes.B2A.load_from_checkpoint(path)
    assert torch.equal(loaded.b2a.linear.weight, sd['b2a.linear.weight']) and loaded.b2a.input_dim == 10
    with pytest.raises(RuntimeError):
        loaded(torch.zeros(2, 10))               # CPU tensors: no fallback
    assert attributes.gender_codes(None, 3).tolist() == [2, 2, 2]


def test_a2b_host_logic(tmp_path):
    """A2B mirror, host side: attribute selection from a checkpoint's config (config.py:373-413), feature vector
    (a2b.py:569-592) against the oracle restatement, checkpoint round trip with the reference's `a2b.*` names."""
    import torch
    from oracle import attributes_oracle as ao2
    from shapy_b200 import attributes
    cfg = {'ds_gender': 'female', 'female_attributes': {'big': True, 'pear_shaped': True, 'tall': True, 'petite': False},
           'measurements': {'height_gt': True, 'weight_gt': False, 'weight_bg': True}, 'bodytalk_meas_preprocess': True,
           'network': {'type': 'polynomial'}}
    m = attributes.A2B(cfg)
    assert m.selected_attr_idx == [0, 8, 14] and m.selected_mmts == ['height_gt', 'weight_bg'] and m.input_feature_size == 5
    assert m.a2b.linear.weight.shape == (10, 5 + 15)
    rating = torch.rand(4, 15)
    mm = {'height_gt': torch.tensor([1.6, 1.7, 1.8, 1.9]), 'weight_bg': torch.tensor([50., 60., 70., 80.])}
    fv, noise = m.create_input_feature_vec({'rating': rating, **mm})
    ref = ao2.a2b_features(rating.numpy(), [0, 8, 14], {k: v.numpy() for k, v in mm.items()}, ['height_gt', 'weight_bg'], True)
    assert np.abs(fv.numpy() - ref).max() < 1e-5 and float(noise.abs().max()) == 0.0
    path = str(tmp_path / 'a2b.ckpt')
    torch.save({'state_dict': {f'a2b.{k}': v for k, v in m.a2b.state_dict().items()}, 'hyper_parameters': {'cfg': cfg}}, path)
    m2 = attributes.A2B.load_from_checkpoint(path)
    assert torch.equal(m2.a2b.linear.weight, m.a2b.linear.weight) and m2.selected_mmts == m.selected_mmts
    with pytest.raises(ValueError):
        attributes.A2B({**cfg, 'female_attributes': {'no_such_attribute': True}})
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 5))                 # CPU tensors: there is no CPU path for the regression itself
