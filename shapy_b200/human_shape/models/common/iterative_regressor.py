"""HMRLikeRegressor, host mirror of
regressor/human_shape/models/common/iterative_regressor.py (ctor 40-209, forward 623-870).

Same constructor, registered buffers (`{name}_idxs`, `{name}_mean`, `param_mean`), sub-module names
(backbone, regressor, model, global_rot_decoder, body_pose_decoder, body_measurements) and output dict.
Inference only: the compute is five C-ABI calls (HRNet, head, 6D decode, SMPL-X, measurements), plus one for the
B2A attribute head when its two checkpoints are configured (iterative_regressor.py:146-171, 761-776).  The A2B head
(attributes -> betas, a second body-model pass) is not built.
"""
import os.path as osp
from collections import defaultdict

import torch
import torch.nn as nn

from .... import ops as _ops
from ....attributes import B2A, b2a_forward, gender_codes
from ....body_measurements import BodyMeasurements
from ..backbone.build import build_backbone
from ..body_models.utils import KeypointTensor
from ..camera.camera_projection import CameraParams, build_cam_proj
from .networks import build_regressor


class HMRLikeRegressor(nn.Module):
    def __init__(self, body_model_cfg, network_cfg, loss_cfg=None, dtype=torch.float32):
        super().__init__()
        self.pose_last_stage = network_cfg.get('pose_last_stage', True)
        camera_data = build_cam_proj(network_cfg.get('camera', {}), dtype=dtype)
        self.projection = camera_data['camera']
        self.camera_scale_func = camera_data['scale_func']
        camera_space = {'dim': camera_data['dim'], 'mean': camera_data['mean']}
        self.model = self._build_model(body_model_cfg)
        pose_space = self._build_pose_space(body_model_cfg)
        blendshape_space = self._build_blendshape_space(body_model_cfg)
        self.pose_space, self.blendshape_space, self.appearance_space = pose_space, blendshape_space, {}
        param_dict = {}
        for name, desc in pose_space.items():
            param_dict[name] = {'dim': desc.dim, 'mean': desc.mean}
        param_dict.update(blendshape_space)
        param_dict['camera'] = camera_space
        mean_lst, start = [], 0
        for name, desc in param_dict.items():
            self.register_buffer(f'{name}_idxs', torch.arange(start, start + desc['dim'], dtype=torch.long))
            mean_lst.append(desc['mean'].view(-1))
            start += desc['dim']
            self.register_buffer(f'{name}_mean', desc['mean'])
        self.param_names = list(param_dict.keys())
        self._slices = {}
        s = 0
        for name, desc in param_dict.items():
            self._slices[name] = (s, s + desc['dim'])
            s += desc['dim']
        param_mean = torch.cat(mean_lst).view(1, -1)
        self._param_dim = param_mean.numel()
        self.register_buffer('param_mean', param_mean)
        self.backbone, feat_dims = build_backbone(network_cfg.get('backbone', {}) or {})
        self.feature_key = network_cfg.get('feature_key', 'avg_pooling')
        self._feat_dim = feat_dims[self.feature_key]
        self.regressor, self._num_stages = build_regressor(network_cfg, self._feat_dim, self._param_dim,
                                                           param_mean=param_mean)
        meas_def = osp.expandvars(network_cfg.get('meas_definition_path', '') or '')
        meas_vert = osp.expandvars(network_cfg.get('meas_vertices_path', '') or '')
        lm = network_cfg.get('meas_landmarks', None)     # shapy_b200 extension: landmarks given inline
        self.compute_measurements = bool(network_cfg.get('compute_measurements', False)) and (
            lm is not None or (osp.exists(meas_def) and osp.exists(meas_vert)))
        if self.compute_measurements:
            cfg = {'meas_definition_path': meas_def, 'meas_vertices_path': meas_vert}
            if lm is not None:
                cfg['landmarks'] = lm
            self.body_measurements = BodyMeasurements(cfg)
        # betas -> attributes regressors, one per gender (iterative_regressor.py:146-171)
        b2a_m = osp.expandvars(network_cfg.get('b2a_males_checkpoint', '') or '')
        b2a_f = osp.expandvars(network_cfg.get('b2a_females_checkpoint', '') or '')
        self.use_b2a = bool(network_cfg.get('use_b2a', False)) and osp.exists(b2a_m) and osp.exists(b2a_f)
        if self.use_b2a:
            self.b2a_males = B2A.load_from_checkpoint(b2a_m)
            self.b2a_females = B2A.load_from_checkpoint(b2a_f)
            for mod in (self.b2a_males, self.b2a_females):
                for p in mod.parameters():
                    p.requires_grad = False
        if network_cfg.get('use_a2b', False) and osp.exists(osp.expandvars(network_cfg.get('a2b_males_checkpoint', '') or '')):
            raise NotImplementedError('shapy_b200: the A2B head (iterative_regressor.py:173-204, 778-852) is not built')
        self.use_a2b = False

    # properties of the reference class
    param_dim = property(lambda self: self._param_dim)
    feat_dim = property(lambda self: self._feat_dim)
    num_stages = property(lambda self: self._num_stages)
    num_betas = property(lambda self: self.model.num_betas)

    def get_mean(self):
        return self.param_mean

    def flat_params_to_dict(self, param_tensor):
        return {name: torch.index_select(param_tensor, 1, getattr(self, f'{name}_idxs')) for name in self.param_names}

    def compute_features(self, images, extra_features=None):
        return self.backbone(images)[self.feature_key]

    def forward(self, images, targets=None, compute_losses=True, cond=None, extra_features=None, **kwargs):
        if self.training:
            raise RuntimeError('shapy_b200 regressor is inference-only: call .eval() first')
        batch_size = len(images)
        features = self.compute_features(images)
        parameters, _ = self.regressor(features)
        num_stages = len(parameters)
        param_dicts = []
        for params in parameters:
            d = {}
            for name in self.param_names:
                a, b = self._slices[name]
                val = params[:, a:b].contiguous()
                if hasattr(self, f'{name}_decoder'):
                    d[name] = getattr(self, f'{name}_decoder')(val)
                    d[f'raw_{name}'] = val
                else:
                    d[name] = val
            param_dicts.append(d)
        if not self.pose_last_stage:
            raise NotImplementedError('shapy_b200: pose_last_stage=False is not implemented (SHAPY_A uses True)')
        merged = param_dicts[-1]
        camera_params = merged['camera']
        model_output = self.model(get_skin=True, return_shaped=True, global_rot=merged['global_rot'],
                                  body_pose=merged['body_pose'], betas=merged['betas'], camera=camera_params)
        out_params = defaultdict(lambda: dict())
        last = f'stage_{num_stages - 1:02d}'
        for key in ('vertices', 'v_shaped', 'joints'):
            out_params[last][key] = model_output[key]
        scale = self.camera_scale_func(camera_params[:, 0].view(-1, 1))
        translation = camera_params[:, 1:3]
        proj_joints = KeypointTensor.from_obj(model_output['proj_joints'], model_output['joints'])
        out_params['proj_joints'] = proj_joints
        out_params['num_stages'] = num_stages
        out_params['features'] = features
        out_params['camera_parameters'] = CameraParams(translation=translation, scale=scale,
                                                       scale_first=getattr(self.projection, 'scale_first', False))
        stage_keys = []
        for n in range(num_stages):
            key = f'stage_{n:02d}'
            stage_keys.append(key)
            out_params[key]['faces'] = model_output['faces']
            out_params[key].update(param_dicts[n])
        if self.compute_measurements:
            meas = self.body_measurements.forward_vertices(out_params[last]['v_shaped'], self.model.faces_i32)
            meas_dict = {name: d['tensor'] for name, d in meas['measurements'].items()}
            out_params[last].update(measurements=meas_dict)
            out_params.update(measurements=meas_dict)
        out_params['stage_keys'] = stage_keys
        out_params[stage_keys[-1]]['proj_joints'] = proj_joints
        if self.use_b2a:
            codes = torch.from_numpy(gender_codes(targets, batch_size))
            out_params['attributes'] = b2a_forward(merged['betas'], codes, self.b2a_males, self.b2a_females)
        out_params['losses'] = {}
        return out_params
