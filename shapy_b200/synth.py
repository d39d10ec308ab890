"""Seeded synthetic stand-ins for the licensed assets the reference needs.

The SMPL-X model file, the SHAPY_A checkpoint, ``all_means.pkl`` and
``SMPLX_to_J14.pkl`` are licensed downloads that are not in the reference repo
(documentation/INSTALL.md:39-121), so every parity test / benchmark runs on
synthetic weights of the real shapes:

* ``make_smplx``       -- SMPL-X sized body model on the REAL 20 908-face topology
                          and a real T-pose body (tests/golden/img00_body.npz,
                          extracted from the reference's only golden sample).
* ``make_state_dict``  -- order-independent, per-key seeded initialisation of any
                          ``state_dict`` (He-normal convs, randomised BN stats).
* ``mean_params``      -- the 145-d mean vector of the SHAPY_A parameter space.
"""
import hashlib
import math
import os
import re

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(_HERE), 'tests', 'golden')

# Public SMPL-X kinematic tree (55 joints); consistent with HEAD_IDX = 15 ->
# chain [15, 12, 9, 6, 3, 0] (reference body_models.py:531-532, 587).
SMPLX_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 15, 15, 15,
                 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35, 20, 37, 38,
                 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50, 21, 52, 53]

NUM_VERTS = 10475
NUM_FACES = 20908
NUM_JOINTS = 55


def load_body_fixture(path=None):
    path = path or os.path.join(GOLDEN_DIR, 'img00_body.npz')
    d = np.load(path)
    return {k: d[k] for k in d.files}


def make_smplx(seed: int = 2, fixture=None, dtype=torch.float32) -> dict:
    """Synthetic SMPL-X (V=10475, J=55) with the buffer names of the reference module
    (body_models.py:112-166, 563-597).  Dense tensors, as the reference stores them."""
    fx = fixture or load_body_fixture()
    g = torch.Generator().manual_seed(seed)
    V, J = NUM_VERTS, NUM_JOINTS
    faces = torch.from_numpy(fx['faces'].astype(np.int64))
    v_template = torch.from_numpy(fx['v_shaped'].astype(np.float32))
    shapedirs_all = torch.randn(V, 3, 20, generator=g) * 5e-3
    posedirs = torch.randn((J - 1) * 9, V * 3, generator=g) * 1e-3

    # Spatially coherent joint regressor and skinning weights (like the real SMPL-X, whose weights are smooth
    # functions of position): 55 joint "seeds" by farthest-point sampling on the template, every joint regresses
    # from its 32 nearest vertices, every vertex is skinned to its 4 nearest joints.
    vt = v_template.double()
    seeds = [int(((vt - vt.mean(0)) ** 2).sum(1).argmin())]
    dmin = ((vt - vt[seeds[0]]) ** 2).sum(1)
    for _ in range(J - 1):
        seeds.append(int(dmin.argmax()))
        dmin = torch.minimum(dmin, ((vt - vt[seeds[-1]]) ** 2).sum(1))
    d2 = ((vt[:, None, :] - vt[seeds][None, :, :]) ** 2).sum(-1)          # (V, J)
    J_regressor = torch.zeros(J, V)
    for j in range(J):
        idx = torch.topk(-d2[:, j], 32).indices
        J_regressor[j, idx] = torch.softmax(-d2[idx, j] / (2 * 0.03 ** 2), 0).float()
    near = torch.topk(-d2, 4, dim=1).indices                               # (V, 4)
    w = torch.softmax(-torch.gather(d2, 1, near) / (2 * 0.08 ** 2), 1).float().clamp_min(1e-4)
    w = w / w.sum(1, keepdim=True)
    lbs_weights = torch.zeros(V, J)
    lbs_weights.scatter_(1, near, w)

    def sparse_rows(nrows, nnz):
        R = torch.zeros(nrows, V)
        for r in range(nrows):
            idx = torch.randperm(V, generator=g)[:nnz]
            w_ = torch.softmax(torch.randn(nnz, generator=g), 0)
            R[r, idx] = w_
        return R

    def dirichlet(*shape):
        x = -torch.log(torch.rand(*shape, 3, generator=g).clamp_min(1e-6))
        return x / x.sum(-1, keepdim=True)

    F = faces.shape[0]
    lmk_faces_idx = torch.randint(0, F, (51,), generator=g)
    lmk_bary = dirichlet(51)
    dyn_faces = torch.randint(0, F, (79, 17), generator=g)
    dyn_bary = dirichlet(79, 17)
    extra = sparse_rows(14, 32)
    # J14 overwrite: positions of the 14 LSP-style names inside the 123-name SMPL-X keypoint
    # list (body_models.py:185-196).  The name table is data the reference does not need for
    # arithmetic; a fixed seeded choice of 14 distinct body-joint slots keeps the semantics
    # (scatter of regressed joints over existing ones).
    source_idxs = torch.tensor([8, 5, 2, 1, 4, 7, 21, 19, 17, 16, 18, 20, 12, 15])
    target_idxs = torch.arange(14)
    kin = []
    c = 15
    while c != -1:
        kin.append(c)
        c = SMPLX_PARENTS[c]
    return dict(
        v_template=v_template.to(dtype), shapedirs=shapedirs_all[:, :, :10].contiguous().to(dtype),
        expr_dirs=shapedirs_all[:, :, 10:].contiguous().to(dtype), posedirs=posedirs.to(dtype),
        J_regressor=J_regressor.to(dtype), lbs_weights=lbs_weights.to(dtype),
        parents=torch.tensor(SMPLX_PARENTS, dtype=torch.long), faces_tensor=faces,
        lmk_faces_idx=lmk_faces_idx, lmk_bary_coords=lmk_bary.to(dtype),
        dynamic_lmk_faces_idx=dyn_faces, dynamic_lmk_bary_coords=dyn_bary.to(dtype),
        neck_kin_chain=torch.tensor(kin, dtype=torch.long),
        extra_joint_regressor=extra.to(dtype), source_idxs=source_idxs, target_idxs=target_idxs,
        head_vertices_ids=torch.zeros(0, dtype=torch.long))


_RESIDUAL_TAIL_BN = re.compile(r'(branches\.\d+\.\d+\.bn2|layer1\.\d+\.bn3|conv_layers\.\d+\.bn3)\.weight$')


def _key_gen(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f'{seed}:{key}'.encode()).digest()
    return torch.Generator().manual_seed(int.from_bytes(h[:7], 'little'))


def make_state_dict(template: dict, seed: int = 1) -> dict:
    """Fills every float tensor of ``template`` (a state_dict: name -> tensor) with seeded values.

    conv / linear weights: N(0, sqrt(2 / fan_in)) (He; the reference's own std=0.001
    init, hrnet.py:505, collapses activations to zero and is useless for parity);
    biases N(0, 0.05); BatchNorm: weight U(0.5, 1.5), bias N(0, 0.1),
    running_mean N(0, 0.1), running_var U(0.5, 1.5).
    Each tensor depends only on (seed, key, shape) so two structurally identical
    modules receive identical weights regardless of registration order."""
    out = {}
    for k, v in template.items():
        g = _key_gen(seed, k)
        if not torch.is_floating_point(v):
            out[k] = v.clone()
            continue
        name = k.rsplit('.', 1)[-1]
        is_bn = (k.replace('.' + name, '') + '.running_var') in template
        if name == 'running_var':
            t = torch.rand(v.shape, generator=g) + 0.5
        elif name == 'running_mean':
            t = torch.randn(v.shape, generator=g) * 0.1
        elif name == 'weight' and is_bn:
            t = torch.rand(v.shape, generator=g) + 0.5
            if _RESIDUAL_TAIL_BN.search(k):
                t = t * 0.2        # last BN of a residual branch: keeps activations O(1)
            elif 'fuse_layers' in k:
                t = t * 0.4
        elif name == 'bias' and is_bn:
            t = torch.randn(v.shape, generator=g) * 0.1
        elif name == 'weight' and v.dim() >= 2:
            fan_in = v[0].numel()
            gain = 0.5 if k.endswith('downsample.weight') else 2.0   # conv_layers.N.downsample has no BN
            t = torch.randn(v.shape, generator=g) * math.sqrt(gain / fan_in)
        elif name == 'bias':
            t = torch.randn(v.shape, generator=g) * 0.05
        else:
            t = v.clone().float()
        out[k] = t.to(v.dtype)
    return out


def make_head_state_dict(feat_dim=2048, param_dim=145, hidden=(1024, 1024), seed=1) -> dict:
    """regressor.* entries: MLP of networks.py:308-400 (default nn.Linear init on hidden
    layers, xavier_uniform(gain=0.01) on the output layer, 378-382)."""
    sd = {}
    dims = [feat_dim + param_dim] + list(hidden)
    for i in range(len(hidden)):
        g = _key_gen(seed, f'regressor.module.layer_{i:03d}.0')
        bound = 1.0 / math.sqrt(dims[i])
        sd[f'regressor.module.layer_{i:03d}.0.weight'] = (torch.rand(dims[i + 1], dims[i], generator=g) * 2 - 1) * bound
        sd[f'regressor.module.layer_{i:03d}.0.bias'] = (torch.rand(dims[i + 1], generator=g) * 2 - 1) * bound
    g = _key_gen(seed, 'regressor.module.output_layer')
    a = 0.01 * math.sqrt(6.0 / (dims[-1] + param_dim))
    sd['regressor.module.output_layer.weight'] = (torch.rand(param_dim, dims[-1], generator=g) * 2 - 1) * a
    sd['regressor.module.output_layer.bias'] = (torch.rand(param_dim, generator=g) * 2 - 1) / math.sqrt(dims[-1])
    sd['regressor.mean_param'] = mean_params().view(1, -1)
    return sd


def mean_params() -> torch.Tensor:
    """[global_rot 0:6 ; body_pose 6:132 ; betas 132:142 ; camera 142:145].

    global_rot mean = identity-6D with element 3 flipped to -1 (body_heads.py:101-108);
    body_pose mean = identity-6D x 21 (no all_means.pkl); betas 0; camera
    [softplus^-1(0.9), 0, 0] (camera_projection.py:73-78)."""
    ident6 = torch.tensor([1., 0., 0., 1., 0., 0.])
    g = ident6.clone()
    g[3] = -1
    cam = torch.tensor([math.log(math.exp(0.9) - 1), 0., 0.])
    return torch.cat([g, ident6.repeat(21), torch.zeros(10), cam])


def load_landmarks(path=None) -> dict:
    import json
    return json.load(open(path or os.path.join(GOLDEN_DIR, 'measurement_landmarks.json')))


def make_exp_cfg(smplx: dict = None, landmarks: dict = None) -> dict:
    """Plain-dict equivalent of regressor/configs/b2a_expose_hrnet_demo.yaml:175-231 (SHAPY_A), with the
    licensed assets replaced by the synthetic SMPL-X tensors and the measurement landmarks given inline."""
    return {
        'use_adv_training': False,
        'network': {'type': 'SMPLXRegressor', 'smplx': {
            'type': 'iterative-mlp', 'num_stages': 3, 'pose_last_stage': True, 'feature_key': 'concat',
            'predict_hands': False, 'predict_face': False, 'compute_measurements': True,
            'meas_landmarks': landmarks or load_landmarks(), 'use_b2a': False, 'use_a2b': False,
            'backbone': {'type': 'hrnet', 'pretrained': False, 'hrnet': {'pretrained_path': ''}},
            'mlp': {'layers': [1024, 1024], 'dropout': 0.5, 'gain': 0.01, 'normalization': {'type': 'none'},
                    'activation': {'type': 'none'}},
            'camera': {'pos_func': 'softplus', 'weak_persp': {'regress_translation': True, 'regress_scale': True}}}},
        'body_model': {'type': 'smplx', 'model_folder': '', 'smplx': {
            'data_struct': smplx or make_smplx(), 'betas': {'num': 10}, 'expression': {'num': 10},
            'use_face_contour': True, 'global_rot': {'type': 'cont_rot_repr'}, 'body_pose': {'type': 'cont_rot_repr'}}},
        'losses': {'body': {}},
    }


def build_synthetic_regressor(seed: int = 1, device=None):
    """SMPLXRegressor (shapy_b200 mirror) with the seeded synthetic checkpoint loaded."""
    from .human_shape.models import build_model
    model = build_model(make_exp_cfg())['network']
    sd = model.state_dict()
    bb = {k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}
    new = {'backbone.' + k: v for k, v in make_state_dict(bb, seed=seed).items()}
    new.update(make_head_state_dict(seed=seed))
    missing, unexpected = model.load_state_dict(new, strict=False)
    assert not unexpected, unexpected
    model.eval()
    if device is not None:
        model = model.to(device)
    return model
