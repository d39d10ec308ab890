"""Weak-perspective camera, host mirror of regressor/human_shape/models/camera/camera_projection.py
(build_cam_proj 44-89, WeakPerspectiveCamera 168-213).  The projection of the 123 model joints is fused
into the SMPL-X joints kernel (csrc/smplx.cu: smplx_joints_kernel); this module carries the config
and exposes the same objects to callers."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class CameraParams(object):
    attributes = ['translation', 'rotation', 'scale', 'focal_length']
    KEYS = ['translation', 'rotation', 'scale', 'focal_length']

    def __init__(self, translation=None, rotation=None, scale=None, focal_length=None, scale_first=False):
        self.translation, self.rotation, self.scale = translation, rotation, scale
        self.focal_length, self.scale_first = focal_length, scale_first

    def keys(self):
        return [k for k in self.KEYS if getattr(self, k) is not None]

    def get(self, key, default=None):
        return getattr(self, key, default)

    def __getitem__(self, key):
        return getattr(self, key)


class WeakPerspectiveCamera(nn.Module):
    def __init__(self, scale_first=False, **kwargs):
        super().__init__()
        self.scale_first = scale_first


def build_cam_proj(camera_cfg, dtype=torch.float32):
    get = (lambda k, d=None: camera_cfg.get(k, d)) if camera_cfg is not None else (lambda k, d=None: d)
    camera_type = get('type', 'weak-persp') or 'weak-persp'
    pos = get('pos_func', 'softplus')
    if camera_type.lower() != 'weak-persp' or pos != 'softplus':
        raise ValueError('shapy_b200 implements the weak-persp camera with pos_func=softplus only')
    wp = get('weak_persp', {}) or {}
    if wp.get('scale_first', False):
        raise ValueError('shapy_b200: weak_persp.scale_first=True is not implemented')
    mean_scale = wp.get('mean_scale', 0.9)
    mean_scale = math.log(math.exp(mean_scale) - 1)
    return {'camera': WeakPerspectiveCamera(scale_first=False), 'mean': torch.tensor([mean_scale, 0.0, 0.0]),
            'scale_func': F.softplus, 'dim': 3}
