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


# ------------------------------------------------------------------------------------------------ A2B
class FullTarget(Target):
    def __init__(self, gender, attributes=None, height=None, weight=None):
        super().__init__(gender)
        self.fields = {'attributes': attributes, 'height': height, 'weight': weight}

    def has_field(self, name):
        return (name == 'gender' and self.gender is not None) or self.fields.get(name) is not None

    def get_field(self, name, default=None):
        if name == 'gender':
            return self.gender
        v = self.fields.get(name)
        return default if v is None else v


def _a2b_cfg(gender, ntype='polynomial'):
    from shapy_b200.attributes import ATTRIBUTE_NAMES
    names = [n.lower().replace(' ', '_') for n in ATTRIBUTE_NAMES[gender]]
    sel = {n: (i % 3 != 1) for i, n in enumerate(names)}                     # 10 of the 15 attributes
    return {'ds_gender': gender, f'{gender}_attributes': sel, 'measurements': {'height_gt': True, 'weight_gt': True,
            'height_bg': False, 'weight_bg': True}, 'num_shape_comps': 10, 'bodytalk_meas_preprocess': True,
            'network': {'type': ntype}}


@pytest.mark.parametrize('ntype', ['polynomial', 'linear'])
def test_a2b_kernel_vs_oracle(ntype):
    from shapy_b200 import attributes
    g = torch.Generator().manual_seed(9)
    mods = {}
    for gender in ('male', 'female'):
        m = attributes.A2B(_a2b_cfg(gender, ntype))
        lin = m.a2b.linear if ntype == 'polynomial' else m.a2b
        with torch.no_grad():
            lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) * 0.05)
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=g) * 0.1)
        mods[gender] = m.cuda().eval()
    B = 9
    genders = ['male', 'female', None, 'Female', 'm', '', 'male', 'female', 'female']
    rating = torch.rand(B, 15, generator=g) * 4 + 1
    mm = {'height_gt': torch.rand(B, generator=g) * 0.4 + 1.5, 'weight_gt': torch.rand(B, generator=g) * 40 + 50,
          'height_bg': torch.rand(B, generator=g) * 0.4 + 1.5, 'weight_bg': torch.rand(B, generator=g) * 40 + 50}
    batch = {'rating': rating.cuda(), **{k: v.cuda() for k, v in mm.items()}}
    fm, _ = mods['male'].create_input_feature_vec(batch)
    ff, _ = mods['female'].create_input_feature_vec(batch)
    assert fm.shape == (B, 13)
    ref_fm = ao.a2b_features(rating.numpy(), mods['male'].selected_attr_idx, {k: v.numpy() for k, v in mm.items()},
                             mods['male'].selected_mmts, True)
    assert np.abs(fm.cpu().numpy() - ref_fm).max() < 1e-5
    codes = attributes.gender_codes([Target(x) for x in genders], B)
    out = attributes.a2b_forward(fm, codes, mods['male'], mods['female'], ff).cpu().numpy()

    def wb(m):
        lin = m.a2b.linear if ntype == 'polynomial' else m.a2b
        return lin.weight.detach().cpu().numpy(), lin.bias.detach().cpu().numpy()
    ref = ao.a2b_by_gender(fm.cpu().numpy(), ff.cpu().numpy(), genders, wb(mods['male']), wb(mods['female']),
                           linear=(ntype == 'linear'))
    assert np.abs(out - ref).max() / np.abs(ref).max() < 2e-6
    assert (out[[2, 5]] == 0).all()


def test_regressor_a2b_refined_betas(tmp_path):
    """use_a2b through the regressor: betas_ref / v_shaped_ref of the last stage (iterative_regressor.py:778-852)."""
    from shapy_b200 import attributes, synth
    from shapy_b200.human_shape.models import build_model
    from oracle import smplx_oracle
    g = torch.Generator().manual_seed(10)
    paths, mods = {}, {}
    for gender in ('male', 'female'):
        m = attributes.A2B(_a2b_cfg(gender))
        with torch.no_grad():
            m.a2b.linear.weight.copy_(torch.randn(m.a2b.linear.weight.shape, generator=g) * 0.01)
            m.a2b.linear.bias.copy_(torch.randn(10, generator=g) * 0.1)
        paths[gender] = str(tmp_path / f'a2b_{gender}.ckpt')
        torch.save({'state_dict': {f'a2b.{k}': v for k, v in m.a2b.state_dict().items()},
                    'hyper_parameters': {'cfg': _a2b_cfg(gender)}}, paths[gender])
        mods[gender] = m
    plain = synth.build_synthetic_regressor()
    cfg = synth.make_exp_cfg()
    cfg['network']['smplx'].update(use_a2b=True, a2b_males_checkpoint=paths['male'], a2b_females_checkpoint=paths['female'],
                                   num_attributes=15)
    model = build_model(cfg)['network']
    assert model.use_a2b
    model.load_state_dict(plain.state_dict(), strict=False)
    model = model.cuda().eval()
    x = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(4)).cuda()
    rating = torch.rand(3, 15, generator=g) * 4 + 1
    targets = [FullTarget('male', rating[0].tolist(), 1.80, 82.0), FullTarget('female', rating[1].tolist()), FullTarget(None)]
    with torch.no_grad():
        out = model(x, targets)
    st = out['stage_02']
    meas = {k: v.cpu().numpy() for k, v in out['measurements'].items()}
    feats = {}
    for gender, h0, w0 in (('male', 1.71, 71.0), ('female', 1.59, 62.0)):
        hg = np.float32([1.80, h0, h0]) if gender == 'male' else np.float32([1.80, h0, h0])
        wg = np.float32([82.0, w0, w0])
        r = np.stack([rating[0].numpy(), rating[1].numpy(), np.zeros(15, np.float32)])
        feats[gender] = ao.a2b_features(r, mods[gender].selected_attr_idx,
                                        {'height_gt': hg, 'weight_gt': wg, 'height_bg': meas['height'], 'weight_bg': meas['mass']},
                                        mods[gender].selected_mmts, True)

    def wb(m):
        return m.a2b.linear.weight.detach().numpy(), m.a2b.linear.bias.detach().numpy()
    ref = ao.a2b_by_gender(feats['male'], feats['female'], ['male', 'female', None], wb(mods['male']), wb(mods['female']))
    got = st['betas_ref'].cpu().numpy()
    assert got.shape == (3, 10) and (got[2] == 0).all()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-5
    vs = smplx_oracle.forward_shape(synth.make_smplx(), torch.from_numpy(ref))
    assert float((st['v_shaped_ref'].cpu() - vs).abs().max()) < 1e-5
