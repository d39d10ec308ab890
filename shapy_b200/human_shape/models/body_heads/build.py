"""Mirror of regressor/human_shape/models/body_heads/build.py:5-32."""
from .registry import BODY_HEAD_REGISTRY


def build(exp_cfg):
    network_cfg = exp_cfg.get('network', {})
    body_cfg = exp_cfg.get('body_model', {})
    network_type = network_cfg.get('type', 'smplx')
    if network_type != 'SMPLXRegressor':
        raise ValueError(f'shapy_b200 implements SMPLXRegressor only, got: {network_type}')
    loss_cfg = (exp_cfg.get('losses', {}) or {}).get('body', {})
    return BODY_HEAD_REGISTRY.get(network_type)(body_cfg, network_cfg=network_cfg.get('smplx', {}),
                                                loss_cfg=loss_cfg)
