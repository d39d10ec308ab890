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

from ....attributes import A2B, B2A, a2b_forward, b2a_forward, gender_codes
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
        # attributes (+ height / weight) -> betas regressors, one per gender (iterative_regressor.py:173-204)
        a2b_m = osp.expandvars(network_cfg.get('a2b_males_checkpoint', '') or '')
        a2b_f = osp.expandvars(network_cfg.get('a2b_females_checkpoint', '') or '')
        self.num_attributes = network_cfg.get('num_attributes', False)
        self.use_a2b = bool(network_cfg.get('use_a2b', False)) and osp.exists(a2b_m) and osp.exists(a2b_f)
        if self.use_a2b:
            if not self.compute_measurements:
                raise ValueError('shapy_b200: use_a2b needs compute_measurements (height_bg / weight_bg features)')
            self.a2b_males = A2B.load_from_checkpoint(a2b_m)
            self.a2b_females = A2B.load_from_checkpoint(a2b_f)
            for mod in (self.a2b_males, self.a2b_females):
                for p in mod.parameters():
                    p.requires_grad = False

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
        if self.use_a2b:
            self._a2b(out_params, targets, batch_size, last, images.device)
        out_params['losses'] = {}
        return out_params

    def _a2b(self, out_params, targets, batch_size, last, device):
        """iterative_regressor.py:778-852: refined betas from the attribute ratings + height / weight of the targets
        (missing values default to 1.71 m / 71 kg for the male and 1.59 m / 62 kg for the female regressor) and the
        measured height / mass of the predicted body; then only the shape blend (`forward_shape`)."""
        targets = list(targets) if targets is not None else []
        targets = (targets + [None] * batch_size)[:batch_size]

        def field(t, name, default):
            if t is not None and hasattr(t, 'get_field'):
                try:
                    return t.get_field(name, default)
                except TypeError:
                    return t.get_field(name) if t.has_field(name) else default
            return default
        n_attr = int(self.num_attributes or len(self.a2b_males.selected_attr_idx) or 15)
        attr = torch.stack([torch.as_tensor(t.get_field('attributes'), dtype=torch.float32)
                            if (t is not None and t.has_field('attributes')) else torch.zeros(n_attr) for t in targets]).to(device)
        meas = out_params['measurements']
        vecs = {}
        for g, mod, h0, w0 in (('m', self.a2b_males, 1.71, 71.0), ('f', self.a2b_females, 1.59, 62.0)):
            batch = {'rating': attr,
                     'height_gt': torch.tensor([field(t, 'height', h0) for t in targets], dtype=torch.float32, device=device),
                     'weight_gt': torch.tensor([field(t, 'weight', w0) for t in targets], dtype=torch.float32, device=device),
                     'height_bg': meas['height'], 'weight_bg': meas['mass']}
            vecs[g], _ = mod.create_input_feature_vec(batch)
        codes = torch.from_numpy(gender_codes(targets, batch_size))
        betas_ref = a2b_forward(vecs['m'], codes, self.a2b_males, self.a2b_females, vecs['f'])
        out_params[last]['betas_ref'] = betas_ref
        out_params[last]['v_shaped_ref'] = self.model.forward_shape(betas_ref)['v_shaped']
