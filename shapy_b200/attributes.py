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
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().shapy_b2a_forward(_lib.ptr(betas), _lib.ptr(gender), _lib.ptr(ws[0]), _lib.ptr(ws[1]), _lib.ptr(ws[2]),
                                                _lib.ptr(ws[3]), betas.shape[0], n, wm.shape[0], _lib.ptr(out), _lib.stream_ptr()),
                   'b2a_forward')
    return out
