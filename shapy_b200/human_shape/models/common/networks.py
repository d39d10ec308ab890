"""Regression head, host mirror of regressor/human_shape/models/common/networks.py (MLP 308-400,
IterativeRegression 403-592, build_regressor 727-762).  Same module / parameter names
(`module.layer_000.0.weight`, `module.output_layer.bias`, `mean_param`), so reference checkpoints load.
The nn.Linear objects are parameter containers; the forward runs shapy_head_forward (csrc/head.cu).
SHAPY_A's head has no activation and no normalisation (configs/b2a_expose_hrnet_demo.yaml:200-207);
other variants are rejected instead of silently computed differently."""

import torch
import torch.nn as nn

from .... import ops as _ops


def _get(cfg, key, default=None):
    if cfg is None:
        return default
    v = cfg.get(key, default) if hasattr(cfg, 'get') else getattr(cfg, key, default)
    return default if v is None else v


class MLP(nn.Module):
    def __init__(self, input_dim, output_dim, layers=None, activation=None, normalization=None, dropout=0.0,
                 gain=0.01, preactivated=False, flatten=True, **kwargs):
        super().__init__()
        layers = list(layers) if layers is not None else []
        act = _get(activation, 'type', 'none') if activation else 'none'
        norm = _get(normalization, 'type', 'none') if normalization else 'none'
        if str(act).lower() != 'none' or str(norm).lower() != 'none':
            raise ValueError('shapy_b200 MLP: only activation/normalization type "none" (SHAPY_A) is implemented')
        if len(layers) != 2:
            raise ValueError('shapy_b200 MLP: exactly two hidden layers (SHAPY_A: [1024, 1024]) are implemented')
        self.flatten, self.input_dim, self.output_dim = flatten, input_dim, output_dim
        self.num_layers = len(layers)
        cur = input_dim
        self.blocks = []
        for i, dim in enumerate(layers):
            lin = nn.Linear(cur, dim, bias=True)
            cur = dim
            mods = [lin] + ([nn.Dropout(dropout)] if dropout > 0.0 else [])
            block = nn.Sequential(*mods)
            self.add_module('layer_{:03d}'.format(i), block)
            self.blocks.append(block)
        self.output_layer = nn.Linear(cur, output_dim)
        # init_weights(..., gain=gain, init_type='xavier', distr='uniform')  (networks.py:378-382)
        nn.init.xavier_uniform_(self.output_layer.weight, gain=gain)

    def extra_repr(self):
        return f'Input ({self.input_dim}) -> Output ({self.output_dim})\nFlatten: {self.flatten}'


class IterativeRegression(nn.Module):
    def __init__(self, module, mean_param, num_stages=1, append_params=True, learn_mean=False, detach_mean=False,
                 dim=1, **kwargs):
        super().__init__()
        if not append_params:
            raise ValueError('shapy_b200 IterativeRegression: append_params=False is not implemented')
        self.module = module
        self._num_stages = num_stages
        self.dim, self.append_params, self.detach_mean, self.learn_mean = dim, append_params, detach_mean, learn_mean
        self._collapsed = None
        if learn_mean:
            self.register_parameter('mean_param', nn.Parameter(mean_param, requires_grad=True))
        else:
            self.register_buffer('mean_param', mean_param)

    def get_mean(self):
        return self.mean_param.clone()

    @property
    def num_stages(self):
        return self._num_stages

    # The SHAPY_A MLP has no activation, so every stage is an affine map; `collapse = True` contracts the three
    # Linear layers once per weight load (fp64, host) and runs 2 tiny launches instead of 10 skinny GEMMs.
    # Differences to the layer-by-layer evaluation are fp32 rounding (~1e-6 relative).
    collapse = True

    def invalidate(self):
        self._collapsed = None

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_collapsed' else copy.deepcopy(v, memo)
        return new

    def forward(self, features, cond=None, **kwargs):
        """Returns (parameters, deltas): lists with one (B, P) tensor per stage."""
        if self.training:
            raise RuntimeError('shapy_b200 IterativeRegression is inference-only: call .eval() first')
        if cond is not None:
            raise NotImplementedError('shapy_b200 IterativeRegression: explicit `cond` is not implemented')
        m = self.module
        mean = self.mean_param.reshape(-1)
        if self.collapse:
            key = str(features.device)
            if getattr(self, '_collapsed', None) is None or self._collapsed[0] != key:
                mats = _ops.collapse_head(m.layer_000[0].weight, m.layer_000[0].bias, m.layer_001[0].weight,
                                          m.layer_001[0].bias, m.output_layer.weight, m.output_layer.bias,
                                          features.shape[1])
                self._collapsed = (key, [t.to(features.device) for t in mats])
            out = _ops.head_forward_collapsed(features, *self._collapsed[1], mean, self._num_stages)
        else:
            out = _ops.head_forward(features, m.layer_000[0].weight, m.layer_000[0].bias, m.layer_001[0].weight,
                                    m.layer_001[0].bias, m.output_layer.weight, m.output_layer.bias, mean,
                                    self._num_stages)
        parameters = [out[i] for i in range(self._num_stages)]
        deltas = [parameters[0] - self.mean_param.reshape(1, -1)]
        return parameters, deltas


def build_regressor(network_cfg, input_dim, output_dim, param_mean):
    regressor_type = _get(network_cfg, 'type', 'mlp')
    if regressor_type != 'iterative-mlp':
        raise ValueError(f'shapy_b200 implements the iterative-mlp regressor only, got: {regressor_type}')
    mlp_cfg = dict(_get(network_cfg, 'mlp', {}) or {})
    append_params = _get(network_cfg, 'append_params', True)
    regressor = MLP(input_dim + append_params * param_mean.numel(), output_dim, **mlp_cfg)
    kw = {k: network_cfg.get(k) for k in ('num_stages', 'append_params', 'learn_mean', 'detach_mean')
          if hasattr(network_cfg, 'get') and network_cfg.get(k) is not None}
    it = IterativeRegression(regressor, param_mean, **kw)
    return it, it.num_stages
