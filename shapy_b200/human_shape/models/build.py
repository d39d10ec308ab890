"""build_model, mirror of regressor/human_shape/models/build.py:14-36."""
from .body_heads import BODY_HEAD_REGISTRY, build_body_head


def build_model(exp_cfg):
    network_cfg = exp_cfg.get('network', {})
    net_type = network_cfg.get('type', 'expose')
    if net_type not in BODY_HEAD_REGISTRY:
        raise ValueError(f'Unknown network type: {net_type}')
    if exp_cfg.get('use_adv_training', False):
        raise NotImplementedError
    return {'network': build_body_head(exp_cfg), 'discriminator': None, 'discriminator_loss': None}
