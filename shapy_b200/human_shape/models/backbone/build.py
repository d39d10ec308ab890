"""Mirror of regressor/human_shape/models/backbone/build.py: only the `hrnet` backbone is on the
SHAPY_A hot path (configs/b2a_expose_hrnet_demo.yaml:196-199)."""
from .hrnet import build as build_hrnet


def build_backbone(backbone_cfg):
    bb_type = backbone_cfg.get('type', 'hrnet')
    if bb_type != 'hrnet':
        raise ValueError(f'shapy_b200 implements the hrnet backbone only, got: {bb_type}')
    pretrained = backbone_cfg.get('pretrained', True)
    model = build_hrnet(backbone_cfg, pretrained=pretrained)
    return model, model.get_output_dim()
