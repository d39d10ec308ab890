"""SMPLXRegressor, mirror of regressor/human_shape/models/body_heads/body_heads.py:211-283 (+ the
SMPLRegressor pose / shape spaces it inherits, 84-131)."""
import os.path as osp
import pickle

import numpy as np
import torch

from ..body_models import build_body_model
from ..common.iterative_regressor import HMRLikeRegressor
from ..common.pose_utils import build_pose_parameterization
from .registry import BODY_HEAD_REGISTRY


@BODY_HEAD_REGISTRY.register()
class SMPLXRegressor(HMRLikeRegressor):
    def __init__(self, body_model_cfg, network_cfg, loss_cfg=None, dtype=torch.float32):
        self.predict_face = network_cfg.get('predict_face', True)
        if self.predict_face or network_cfg.get('predict_hands', False):
            raise ValueError('shapy_b200 SMPLXRegressor: predict_face / predict_hands must be False (SHAPY_A, '
                             'configs/b2a_expose_hrnet_demo.yaml:181-182)')
        super().__init__(body_model_cfg, network_cfg, loss_cfg, dtype=dtype)

    def _build_model(self, body_model_cfg):
        self.body_model_cfg = body_model_cfg
        model = build_body_model(body_model_cfg)
        self.model_type = model.name
        self.curr_model_cfg = body_model_cfg.get(self.model_type, {}) or {}
        return model

    def _build_pose_space(self, body_model_cfg):
        mean_pose_path = osp.expandvars(self.curr_model_cfg.get('mean_pose_path', '') or '')
        self.mean_poses_dict = {}
        if osp.exists(mean_pose_path):
            with open(mean_pose_path, 'rb') as f:
                self.mean_poses_dict = pickle.load(f)
        gcfg = dict(self.curr_model_cfg.get('global_rot', {}) or {'type': 'cont_rot_repr'})
        bcfg = dict(self.curr_model_cfg.get('body_pose', {}) or {'type': 'cont_rot_repr'})
        global_rot_desc = build_pose_parameterization(1, **gcfg)
        self.global_rot_decoder = global_rot_desc.decoder
        body_pose_desc = build_pose_parameterization(num_angles=self.model.num_body_joints,
                                                     mean=self.mean_poses_dict.get('body_pose', None), **bcfg)
        self.body_pose_decoder = body_pose_desc.decoder
        # rotate the model 180 degrees around the x axis (body_heads.py:101-108)
        global_rot_desc.mean[3] = -1
        return {'global_rot': global_rot_desc, 'body_pose': body_pose_desc}

    def _build_blendshape_space(self, body_model_cfg, dtype=torch.float32):
        num_betas = self.model.num_betas
        shape_mean_path = osp.expandvars(body_model_cfg.get('shape_mean_path', '') or '')
        if osp.exists(shape_mean_path):
            shape_mean = torch.from_numpy(np.load(shape_mean_path, allow_pickle=True)).to(dtype=dtype).reshape(
                1, -1)[:, :num_betas].reshape(-1)
        else:
            shape_mean = torch.zeros([num_betas], dtype=dtype)
        return {'betas': dict(dim=num_betas, mean=shape_mean)}
