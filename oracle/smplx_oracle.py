"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference SMPL-X body model.

Pinned by tests/test_oracle_pins.py against (a) the reference's own lbs.py run in the
build container (tests/golden/ref_smplx.npz, written by tools/make_golden.py) and
(b) img_00.npz for the 6D decoder and the camera.  Real-weights parity of lbs()
itself is UNPINNED (the licensed SMPL-X file is not in the reference repo).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
Plain float32 torch-on-CPU, written op by op after:

  decode_6d          regressor/human_shape/models/common/pose_utils.py:138-153
  lbs chain          regressor/human_shape/models/body_models/lbs.py:99-196, 199-295
  transform_mat      regressor/human_shape/models/body_models/utils.py:14-24
  dynamic landmarks  lbs.py:20-49, rot_mat_to_euler utils/rotation_utils.py:86-92
  landmarks          lbs.py:52-94
  SMPLX.forward glue body_models.py:628-767 (identity hands/jaw/eyes, zero expression,
                     J14 overwrite 738-744, v_shaped recomputed without expression 763-765)
  camera             models/camera/camera_projection.py:181-213, iterative_regressor.py:715-728
"""
import math

import torch
import torch.nn.functional as F


def decode_6d(x: torch.Tensor) -> torch.Tensor:
    """(B, 6n) -> (B, n, 3, 3).  The 6 numbers are a row-major 3x2 matrix."""
    B = x.shape[0]
    m = x.reshape(-1, 3, 2)
    a1, a2 = m[:, :, 0], m[:, :, 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-12)
    d = (b1 * a2).sum(1, keepdim=True)
    u = a2 - d * b1
    b2 = u / u.norm(dim=1, keepdim=True).clamp_min(1e-12)
    b3 = torch.linalg.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1).reshape(B, -1, 3, 3)


def full_pose_of(global_rot, body_pose):
    B = global_rot.shape[0]
    eye = torch.eye(3, dtype=global_rot.dtype).view(1, 1, 3, 3)
    # [global, body(21), jaw, leye, reye, lhand(15), rhand(15)]  body_models.py:687-690
    return torch.cat([global_rot, body_pose, eye.expand(B, 33, 3, 3)], dim=1)


def rigid_chain(rot_mats, J, parents):
    """batch_rigid_transform (lbs.py:242-295) without TorchScript."""
    B, N = rot_mats.shape[:2]
    rel = J.clone()
    rel[:, 1:] = J[:, 1:] - J[:, parents[1:]]
    T = torch.zeros(B, N, 4, 4, dtype=J.dtype)
    T[:, :, :3, :3] = rot_mats
    T[:, :, :3, 3] = rel
    T[:, :, 3, 3] = 1
    chain = [T[:, 0]]
    for i in range(1, N):
        chain.append(torch.bmm(chain[int(parents[i])], T[:, i]))
    G = torch.stack(chain, dim=1)
    posed = G[:, :, :3, 3].clone()
    Jh = F.pad(J, [0, 1], value=0.0).unsqueeze(-1)
    A = G.clone()
    A[:, :, :, 3] = G[:, :, :, 3] - torch.matmul(G, Jh)[..., 0]
    return posed, A


def dynamic_landmarks(full_pose, dyn_faces, dyn_bary, neck_kin_chain):
    rel = torch.eye(3, dtype=full_pose.dtype).expand(full_pose.shape[0], 3, 3)
    for idx in neck_kin_chain.tolist():
        rel = torch.matmul(full_pose[:, idx], rel)
    sy = torch.sqrt(rel[:, 0, 0] * rel[:, 0, 0] + rel[:, 1, 0] * rel[:, 1, 0])
    ang = torch.atan2(-rel[:, 2, 0], sy)
    y = torch.round(torch.clamp(-ang * 180.0 / math.pi, max=39)).long()
    neg = (y < 0).long()
    big = (y < -39).long()
    negv = big * 78 + (1 - big) * (39 - y)
    y = neg * negv + (1 - neg) * y
    return dyn_faces[y], dyn_bary[y]


def smplx_forward(model: dict, betas, global_rot, body_pose) -> dict:
    B = betas.shape[0]
    full_pose = full_pose_of(global_rot, body_pose)
    shapedirs = torch.cat([model['shapedirs'], model['expr_dirs']], dim=-1)
    comps = torch.cat([betas, torch.zeros(B, model['expr_dirs'].shape[-1])], dim=-1)
    v_shaped_e = model['v_template'] + torch.einsum('bl,mkl->bmk', comps, shapedirs)
    J = torch.einsum('bik,ji->bjk', v_shaped_e, model['J_regressor'])
    ident = torch.eye(3)
    pose_feature = (full_pose[:, 1:] - ident).reshape(B, -1)
    v_posed = v_shaped_e + torch.matmul(pose_feature, model['posedirs']).view(B, -1, 3)
    posed_J, A = rigid_chain(full_pose, J, model['parents'])
    T = torch.einsum('vj,bjmn->bvmn', model['lbs_weights'], A)
    vh = F.pad(v_posed, [0, 1], value=1.0).unsqueeze(-1)
    vertices = torch.matmul(T, vh)[:, :, :3, 0]
    # landmarks
    lf = model['lmk_faces_idx'].unsqueeze(0).expand(B, -1)
    lb = model['lmk_bary_coords'].unsqueeze(0).expand(B, -1, -1)
    df, db = dynamic_landmarks(full_pose, model['dynamic_lmk_faces_idx'], model['dynamic_lmk_bary_coords'],
                               model['neck_kin_chain'])
    lf = torch.cat([lf, df], 1)
    lb = torch.cat([lb, db], 1)
    tri = model['faces_tensor'][lf]                     # (B, L, 3)
    lv = torch.stack([vertices[b][tri[b]] for b in range(B)])   # (B, L, 3, 3)
    landmarks = (lv * lb.unsqueeze(-1)).sum(2)
    joints = torch.cat([posed_J, landmarks], dim=1)
    reg = torch.einsum('ji,bik->bjk', model['extra_joint_regressor'], vertices)
    joints[:, model['source_idxs']] = reg[:, model['target_idxs']]
    v_shaped = model['v_template'] + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])
    return dict(vertices=vertices, joints=joints, v_shaped=v_shaped, A=A, posed_joints=posed_J)


def forward_shape(model: dict, betas) -> torch.Tensor:
    """SMPL.forward_shape (body_models.py:292-302)."""
    return model['v_template'] + torch.einsum('bl,mkl->bmk', betas, model['shapedirs'])


def weak_persp(joints, camera):
    """scale = softplus(camera[:, 0]); proj = scale * (joints_xy + camera[:, 1:3])."""
    s = F.softplus(camera[:, 0]).view(-1, 1, 1)
    return s * (joints[:, :, :2] + camera[:, 1:3].view(-1, 1, 2))
