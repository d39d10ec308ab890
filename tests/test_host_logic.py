"""CPU tests of the host-side logic: state-dict compatibility with the reference, the compiled HRNet op
program (interpreted with torch ops against the oracle), C-ABI exports.  No GPU, no compute through the
library."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import net_oracle
from shapy_b200 import _lib, synth


@pytest.fixture(scope='module')
def regressor():
    return synth.build_synthetic_regressor()


def test_backbone_state_dict_matches_reference(regressor, golden_dir):
    ref = json.load(open(os.path.join(golden_dir, 'hrnet_keys.json')))
    mine = [[k[len('backbone.'):], list(v.shape)] for k, v in regressor.state_dict().items()
            if k.startswith('backbone.')]
    assert len(mine) == ref['n'] == 1967
    assert mine == ref['keys']


def test_checkpoint_key_families(regressor):
    keys = set(regressor.state_dict())
    for k in ['regressor.module.layer_000.0.weight', 'regressor.module.layer_001.0.bias',
              'regressor.module.output_layer.weight', 'regressor.mean_param', 'param_mean', 'global_rot_idxs',
              'body_pose_mean', 'betas_idxs', 'camera_mean', 'global_rot_decoder.mean', 'body_pose_decoder.mean',
              'model.v_template', 'model.shapedirs', 'model.posedirs', 'model.J_regressor', 'model.lbs_weights',
              'model.parents', 'model.faces_tensor', 'model.expr_dirs', 'model.lmk_faces_idx',
              'model.dynamic_lmk_bary_coords', 'model.neck_kin_chain', 'model.extra_joint_regressor',
              'model.source_idxs', 'model.target_idxs', 'model.head_vertices_ids',
              'body_measurements.left_heel_bc', 'body_measurements.head_top_bc', 'body_measurements.chest_bcs',
              'body_measurements.belly_bcs', 'body_measurements.hips_bcs']:
        assert k in keys, k
    assert regressor.param_mean.shape == (1, 145)
    assert torch.equal(regressor.param_mean.view(-1), synth.mean_params())
    assert regressor.betas_idxs.tolist() == list(range(132, 142))


def _run_program(convs, ops, slots, feat_slot, x):
    """Interprets the op program with fp32 torch ops (BN applied after the conv, as the reference does)."""
    B, _, H, W = x.shape
    mem = [None] * len(slots)

    def conv(c, inp):
        y = F.conv2d(inp, c['weight'], c['bias'], stride=c['stride'], padding=c['ksize'] // 2)
        if c['bn'] is not None:
            bn = c['bn']
            y = F.batch_norm(y, bn['mean'], bn['var'], bn['weight'], bn['bias'], False, 0.0, bn['eps'])
        return y
    feats = None
    for o in ops:
        if o['kind'] in (_lib.OP_STEM, _lib.OP_CONV):
            c = convs[o['conv']]
            inp = x if o['kind'] == _lib.OP_STEM else mem[o['in_slot']]
            y = conv(c, inp)
            if o['res_slot'] >= 0:
                y = y + mem[o['res_slot']]
            if o['relu']:
                y = F.relu(y)
            s = slots[o['out_slot']]
            if s['channels'] == c['cout']:
                mem[o['out_slot']] = y
            else:
                if mem[o['out_slot']] is None or mem[o['out_slot']].shape[1] != s['channels']:
                    mem[o['out_slot']] = torch.zeros(B, s['channels'], H // s['div'], W // s['div'])
                mem[o['out_slot']][:, o['out_coff']:o['out_coff'] + c['cout']] = y
        elif o['kind'] == _lib.OP_FUSE:
            acc = None
            for t, sh in zip(o['fuse_in'], o['fuse_shift']):
                v = mem[t]
                if sh:
                    v = F.interpolate(v, scale_factor=2 ** sh, mode='nearest')
                acc = v if acc is None else acc + v
            if o['relu']:
                acc = F.relu(acc)
            s = slots[o['out_slot']]
            if s['channels'] == acc.shape[1]:
                mem[o['out_slot']] = acc
            else:
                mem[o['out_slot']][:, o.get('out_coff', 0):o.get('out_coff', 0) + acc.shape[1]] = acc
        elif o['kind'] == _lib.OP_POOL:
            feats = mem[o['in_slot']].mean(dim=(2, 3))
    return feats, mem


def test_hrnet_program_matches_oracle(regressor):
    bb = regressor.backbone
    convs, ops, slots, feat_slot, layer_slots = bb.build_program()
    assert sum(o['kind'] in (_lib.OP_STEM, _lib.OP_CONV) for o in ops) == 331
    assert sum(o['kind'] == _lib.OP_FUSE for o in ops) == 1 * 2 + 4 * 3 + 3 * 4 + 1
    assert len(slots) < 50, len(slots)           # liveness-based reuse (within a lane) keeps the workspace small
    assert {o['lane'] for o in ops} == {0, 1, 2, 3}     # the four branches are emitted on four lanes
    sd = {k[len('backbone.'):]: v for k, v in regressor.state_dict().items() if k.startswith('backbone.')}
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        ref = net_oracle.hrnet_forward(sd, x)
        feats, mem = _run_program(convs, ops, slots, feat_slot, x)
    assert torch.allclose(feats, ref['concat'], rtol=1e-4, atol=1e-5)
    for name, s in layer_slots.items():
        assert torch.allclose(mem[s], ref[name], rtol=1e-4, atol=1e-5), name


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads and exports exactly the functions include/shapy_b200.h declares."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, 'include', 'shapy_b200.h')).read()
    declared = set(re.findall(r'\b(shapy_[a-z0-9_]+)\s*\(', header))
    declared -= {'shapy_b200'}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.lib()           # raises if the .so is missing or a symbol is not exported
    assert L.shapy_version() >= 100
    for name in declared:
        assert hasattr(L, name)


def test_ops_refuse_cpu_tensors(regressor):
    with pytest.raises(RuntimeError):
        regressor(torch.zeros(1, 3, 64, 64))
    from shapy_b200 import ops
    with pytest.raises(RuntimeError):
        ops.decode_rot6d(torch.zeros(2, 6))
    with pytest.raises(RuntimeError):
        ops.mesh_to_mesh_forward(torch.zeros(1, 2, 3, 3), torch.zeros(1, 4, 3, 3))


def test_j14_names_resolve_without_the_reference_package(tmp_path):
    """A real SHAPY config passes `j14_regressor_path`; the 14 LSP-style names must resolve to the public SMPL-X body
    joint indices even when human_shape.data.utils is not importable (the overwrite was silently skipped)."""
    import numpy as np
    from shapy_b200 import synth
    from shapy_b200.human_shape.models.body_models.body_models import SMPLX
    ds = {k: v for k, v in synth.make_smplx().items() if k not in ('extra_joint_regressor', 'source_idxs', 'target_idxs')}
    path = str(tmp_path / 'SMPLX_to_J14.npy')
    np.save(path, np.random.default_rng(0).random((14, synth.NUM_VERTS)).astype(np.float32))
    m = SMPLX(data_struct=ds, j14_regressor_path=path)
    assert m.use_joint_regressor
    assert m.source_idxs.tolist() == [1, 2, 4, 5, 7, 8, 12, 15, 16, 17, 18, 19, 20, 21]
    names = ['right_ankle', 'right_knee', 'right_hip', 'left_hip', 'left_knee', 'left_ankle', 'right_wrist', 'right_elbow',
             'right_shoulder', 'left_shoulder', 'left_elbow', 'left_wrist', 'neck', 'head']
    assert [names[t] for t in m.target_idxs.tolist()] == [m.keypoint_names[s] for s in m.source_idxs.tolist()]
