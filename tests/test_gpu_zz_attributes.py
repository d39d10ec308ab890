"""B2A kernel (shapy_b2a_forward) against the reference Polynomial outputs, stand-alone and inside the regressor."""
import os

import numpy as np
import pytest
import torch

from oracle import attributes_oracle as ao

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'b2a.npz'))
GENDERS = [None if x == '<none>' else str(x) for x in G['genders']]


class Target:
    def __init__(self, gender):
        self.gender = gender

    def has_field(self, name):
        return name == 'gender' and self.gender is not None

    def get_field(self, name):
        return self.gender


def heads():
    from shapy_b200 import attributes
    out = []
    for g in ('male', 'female'):
        m = attributes.B2A(10, 15)
        with torch.no_grad():
            m.b2a.linear.weight.copy_(torch.from_numpy(G[f'w_{g}']))
            m.b2a.linear.bias.copy_(torch.from_numpy(G[f'b_{g}']))
        out.append(m.cuda().eval())
    return out


def test_kernel_matches_reference_module():
    from shapy_b200 import attributes
    males, females = heads()
    codes = torch.from_numpy(attributes.gender_codes([Target(g) for g in GENDERS], len(GENDERS)))
    out = attributes.b2a_forward(torch.from_numpy(G['betas']).cuda(), codes, males, females).cpu().numpy()
    assert np.abs(out - G['attributes']).max() / np.abs(G['attributes']).max() < 1e-6
    assert (out[[3, 4, 6]] == 0).all()
    single = males(torch.from_numpy(G['betas']).cuda()).cpu().numpy()          # B2A.forward: one regressor for every row
    ref = ao.polynomial_forward(G['betas'], G['w_male'], G['b_male'])
    assert np.abs(single - ref).max() / np.abs(ref).max() < 1e-6


def test_regressor_reports_attributes(tmp_path):
    from shapy_b200 import synth
    from shapy_b200.human_shape.models import build_model
    paths = {}
    for g in ('male', 'female'):
        sd = {'b2a.linear.weight': torch.from_numpy(G[f'w_{g}']), 'b2a.linear.bias': torch.from_numpy(G[f'b_{g}']),
              'b2a.indices_000': torch.from_numpy(G['indices_000']), 'b2a.indices_001': torch.from_numpy(G['indices_001'])}
        paths[g] = str(tmp_path / f'{g}.ckpt')
        torch.save({'state_dict': sd, 'hyper_parameters': {'cfg': {}}}, paths[g])
    plain = synth.build_synthetic_regressor()
    cfg = synth.make_exp_cfg()
    cfg['network']['smplx']['use_b2a'] = True
    cfg['network']['smplx']['b2a_males_checkpoint'] = paths['male']
    cfg['network']['smplx']['b2a_females_checkpoint'] = paths['female']
    model = build_model(cfg)['network']
    assert model.use_b2a
    model.load_state_dict(plain.state_dict(), strict=False)     # same synthetic backbone / head / body model
    model = model.cuda().eval()
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
    targets = [Target('male'), Target('female'), Target(None)]
    with torch.no_grad():
        out = model(x, targets)
    betas = out['stage_02']['betas'].cpu().numpy()
    ref = ao.b2a_by_gender(betas, ['male', 'female', None], (G['w_male'], G['b_male']), (G['w_female'], G['b_female']))
    got = out['attributes'].cpu().numpy()
    assert got.shape == (3, 15) and (got[2] == 0).all()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
