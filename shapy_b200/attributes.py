"""B2A attribute head: SMPL-X betas -> linguistic attribute ratings.

Host mirror of
    attributes/attributes/attributes_betas/polynomial.py:21-140   Polynomial (degree-2 features + Linear)
    attributes/attributes/attributes_betas/b2a.py:24-52,117-118   B2A (Lightning module wrapping it as `.b2a`)
    regressor/human_shape/models/common/iterative_regressor.py:146-171,761-776   loading + per-gender routing
Same state-dict names (`b2a.linear.weight`, `b2a.linear.bias`, `b2a.indices_000`, `b2a.indices_001`), so the
reference's Lightning checkpoints (`ckpt['state_dict']`) load unchanged.  The arithmetic is one launch of
`shapy_b2a_forward` for both genders (csrc/attributes.cu); there is no CPU fallback.
"""
from itertools import combinations_with_replacement

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class Polynomial(nn.Module):
    """polynomial.py:21-52: parameters `linear.{weight,bias}`, buffers `indices_000`, `indices_001`."""

    def __init__(self, input_dim: int, output_dim: int, degree: int = 2, alpha: float = 0.0):
        super().__init__()
        if degree != 2:
            raise NotImplementedError('shapy_b200 Polynomial: only degree 2 (the released B2A models) is built')
        self.input_dim, self.output_dim, self.degree, self.alpha = input_dim, output_dim, degree, alpha
        combos = [c for d in range(1, degree + 1) for c in combinations_with_replacement(range(input_dim), d)]
        self.coeff_size = len(combos)
        self.linear = nn.Linear(self.coeff_size, output_dim)
        for ii in range(degree):
            idx = torch.tensor([c for c in combos if len(c) == ii + 1], dtype=torch.long)
            self.register_buffer(f'indices_{ii:03d}', idx)

    def forward(self, x):
        g = torch.zeros(x.shape[0], dtype=torch.int32, device=x.device)
        return b2a_forward(x, g, self, self)


class B2A(nn.Module):
    """b2a.py:24-52: holds the network as `.b2a`; forward(x) = self.b2a(x) (b2a.py:117-118)."""

    def __init__(self, input_dim: int = 10, output_dim: int = 15, degree: int = 2):
        super().__init__()
        self.b2a = Polynomial(input_dim, output_dim, degree)

    @staticmethod
    def load_from_checkpoint(checkpoint_path, cfg=None, map_location='cpu'):
        """Reads a Lightning checkpoint of the reference's B2A (`state_dict` with `b2a.*` keys)."""
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
        w = sd['b2a.linear.weight']
        n_feat, n = w.shape[1], 1
        while n + n * (n + 1) // 2 < n_feat:
            n += 1
        if n + n * (n + 1) // 2 != n_feat:
            raise ValueError(f'B2A checkpoint: {n_feat} input features are not a degree-2 polynomial basis')
        obj = B2A(n, w.shape[0], 2)
        obj.load_state_dict({k: v for k, v in sd.items() if k.startswith('b2a.')}, strict=True)
        return obj.eval()

    def forward(self, x):
        return self.b2a(x)


# attribute vocabulary of the rating vectors (attributes/attributes/utils/constants.py:38-73): pure data, needed to turn
# the `{gender}_attributes: {name: bool}` switches of an A2B checkpoint's config into indices of the rating vector
ATTRIBUTE_NAMES = {
    'female': ['Big', 'Broad Shoulders', 'Feminine', 'Large Breasts', 'Long Legs', 'Long Neck', 'Long Torso', 'Muscular',
               'Pear Shaped', 'Petite', 'Short', 'Short Arms', 'Skinny Legs', 'Slim Waist', 'Tall'],
    'male': ['Average', 'Big', 'Broad Shoulders', 'Delicate Build', 'Long Legs', 'Long Neck', 'Long Torso', 'Masculine',
             'Muscular', 'Rectangular', 'Short', 'Short Arms', 'Skinny Arms', 'Soft Body', 'Tall'],
}


def features_from_config(cfg):
    """attributes/attributes/utils/config.py:373-413 (non-synthetic datasets): (attribute names, their indices in the
    rating vector, measurement feature names) switched on in an A2B config."""
    ds_gender = cfg.get('ds_gender', 'female')
    names = ATTRIBUTE_NAMES[ds_gender]
    attributes = []
    if cfg.get('use_attributes', True):
        attributes = [k for k, v in (cfg.get(f'{ds_gender}_attributes') or {}).items() if v]
    idx = [i for i, v in enumerate(names) if v.lower().replace(' ', '_') in attributes]
    if len(idx) != len(attributes):
        raise ValueError('Some selected attributes are not annotated.')
    mmts = []
    if cfg.get('use_measurements', True):
        mmts = [k for k, v in (cfg.get('measurements') or {}).items() if v]
    return attributes, idx, mmts


class A2B(nn.Module):
    """Attributes (+ height / weight) -> betas: host mirror of attributes/attributes/attributes_betas/a2b.py:91-279
    (module), 569-602 (create_input_feature_vec) for the network types the released regressors use -- `polynomial`
    (degree 2, polynomial.py) and `linear`.  The network is held as `.a2b` with the reference's parameter names, so a
    Lightning checkpoint's `state_dict` loads unchanged; forward(x) = self.a2b(x) (a2b.py:278-279) runs on the device
    through `shapy_b2a_forward` (the same degree-2 kernel as B2A, csrc/attributes.cu)."""

    def __init__(self, cfg=None, input_dim=None, output_dim=None, network_type=None):
        super().__init__()
        cfg = cfg or {}
        self.cfg = cfg
        self.betas_size = int(output_dim or cfg.get('num_shape_comps', 10))
        self.ds_gender = cfg.get('ds_gender', 'female')
        self.bodytalk_meas_preprocess = bool(cfg.get('bodytalk_meas_preprocess', False))
        if cfg.get('use_attr_noise', False) or cfg.get('use_srb_noise', False):
            raise ValueError('shapy_b200 A2B: input-noise options are training-time features and are not supported')
        if input_dim is None:
            self.selected_attr, self.selected_attr_idx, self.selected_mmts = features_from_config(cfg)
            input_dim = len(self.selected_attr) + len(self.selected_mmts)
        else:
            self.selected_attr, self.selected_attr_idx, self.selected_mmts = [], list(range(input_dim)), []
        self.input_feature_size = int(input_dim)
        ntype = network_type or (cfg.get('network') or {}).get('type', 'polynomial')
        if ntype == 'polynomial':
            degree = ((cfg.get('network') or {}).get('polynomial') or {}).get('degree', 2)
            self.a2b = Polynomial(self.input_feature_size, self.betas_size, degree)
        elif ntype == 'linear':
            self.a2b = nn.Linear(self.input_feature_size, self.betas_size)
        else:
            raise NotImplementedError(f'shapy_b200 A2B: network type `{ntype}` is not on the released-model path '
                                      '(polynomial / linear are)')
        self.network_type = ntype

    @staticmethod
    def load_from_checkpoint(checkpoint_path, cfg=None, map_location='cpu'):
        """Reads a Lightning checkpoint of the reference's A2B: `hyper_parameters['cfg']` + `state_dict` (`a2b.*`)."""
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if cfg is None:
            cfg = (ckpt.get('hyper_parameters') or {}).get('cfg', {})
        obj = A2B(cfg)
        sd = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
        obj.load_state_dict({k: v for k, v in sd.items() if k.startswith('a2b.')}, strict=True)
        return obj.eval()

    def create_input_feature_vec(self, batch):
        """a2b.py:569-602 without the training-time noise: [rating[:, selected] | one column per selected measurement]."""
        fv = batch['rating'][:, self.selected_attr_idx]
        if not torch.is_tensor(fv):
            fv = torch.from_numpy(np.asarray(fv)).to(dtype=torch.float32)
        noise = torch.zeros_like(fv)
        for name in self.selected_mmts:
            meas = batch[name].reshape(-1, 1)
            if not torch.is_tensor(meas):
                meas = torch.from_numpy(np.asarray(meas)).to(dtype=torch.float32)
            meas = meas.to(device=fv.device, dtype=fv.dtype)
            if self.bodytalk_meas_preprocess:
                if 'height' in name:
                    meas = meas * 100
                if 'mass' in name or 'weight' in name:
                    meas = meas.pow(1.0 / 3.0)
            fv = torch.hstack((fv, meas))
        return fv, noise

    def forward(self, x):
        g = torch.zeros(x.shape[0], dtype=torch.int32, device=x.device)
        return a2b_forward(x, g, self, self)


def _a2b_wb(module):
    net = module.a2b if isinstance(module, A2B) else module
    if isinstance(net, Polynomial):
        return net.linear.weight, net.linear.bias, net.input_dim, True
    return net.weight, net.bias, net.in_features, False


def a2b_forward(features_m: torch.Tensor, gender, males, females, features_f: torch.Tensor = None):
    """(B, num_betas): rows of gender 0 go through `males` with `features_m`, rows of gender 1 through `females` with
    `features_f` (the reference builds one feature vector per gender, iterative_regressor.py:808-836); other rows are
    zero (843-850).  One launch of `shapy_a2b_forward`."""
    features_f = features_m if features_f is None else features_f
    if not (torch.is_tensor(features_m) and features_m.is_cuda and features_f.is_cuda):
        raise RuntimeError('shapy_b200.attributes: features must be CUDA tensors (there is no CPU fallback)')
    dev = features_m.device
    wm, bm, n, poly = _a2b_wb(males)
    wf, bf, nf, polyf = _a2b_wb(females)
    if n != nf or wm.shape != wf.shape or poly != polyf:
        raise ValueError('male and female A2B regressors must have the same architecture')
    if features_m.shape[1] != n or features_f.shape != features_m.shape:
        raise ValueError(f'A2B expects {n} input features per gender, got {tuple(features_m.shape)} / {tuple(features_f.shape)}')
    gender = torch.as_tensor(gender).to(device=dev, dtype=torch.int32).contiguous()
    ws = [t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (wm, bm, wf, bf)]
    fm, ff = features_m.contiguous().float(), features_f.contiguous().float()
    out = torch.empty(fm.shape[0], wm.shape[0], dtype=torch.float32, device=dev)
    if fm.shape[0] == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().shapy_a2b_forward(_lib.ptr(fm), _lib.ptr(ff), _lib.ptr(gender), _lib.ptr(ws[0]), _lib.ptr(ws[1]),
                                                _lib.ptr(ws[2]), _lib.ptr(ws[3]), fm.shape[0], n, wm.shape[0], 0 if poly else 1,
                                                _lib.ptr(out), _lib.stream_ptr()), 'a2b_forward')
    return out


def gender_codes(targets, batch_size: int) -> np.ndarray:
    """iterative_regressor.py:762-767: first letter of the target's `gender` field -> 0 male, 1 female, 2 none."""
    genders = []
    for x in (targets if targets is not None else []):
        genders.append(x.get_field('gender') if (x is not None and x.has_field('gender')) else None)
    genders = (genders + [None] * batch_size)[:batch_size]      # no target / no field -> 'n' (a row of zeros)
    letters = [x.lower()[0] if (x is not None and x != '') else 'n' for x in genders]
    return np.array([0 if c == 'm' else (1 if c == 'f' else 2) for c in letters], dtype=np.int32)


def _wb(module):
    p = module.b2a if isinstance(module, B2A) else module
    return p.linear.weight, p.linear.bias, p.input_dim


def b2a_forward(betas: torch.Tensor, gender: torch.Tensor, males, females) -> torch.Tensor:
    """(B, num_outputs): rows with gender 0 / 1 through the male / female regressor, other rows zero."""
    if not (torch.is_tensor(betas) and betas.is_cuda):
        raise RuntimeError('shapy_b200.attributes: betas must be a CUDA tensor (there is no CPU fallback)')
    wm, bm, n = _wb(males)
    wf, bf, nf = _wb(females)
    if n != nf or wm.shape != wf.shape:
        raise ValueError('male and female B2A regressors must have the same shape')
    if betas.shape[1] != n:
        raise ValueError(f'B2A expects {n} betas, got {betas.shape[1]}')
    dev = betas.device
    betas = betas.contiguous().float()
    gender = gender.to(device=dev, dtype=torch.int32).contiguous()
    ws = [t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (wm, bm, wf, bf)]
    out = torch.empty(betas.shape[0], wm.shape[0], dtype=torch.float32, device=dev)
    if betas.shape[0] == 0:
        return out
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().shapy_b2a_forward(_lib.ptr(betas), _lib.ptr(gender), _lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]),
                                                _lib.ptr(ws[3]), betas.shape[0], n, wm.shape[0], _lib.ptr(out), _lib.stream_ptr()),
                   'b2a_forward')
    return out
