#!/usr/bin/env python
"""Writes tests/golden/* by running the REFERENCE's own code (build container only).

    python tools/make_golden.py            # needs /root/reference

Everything written here is a small fixture that travels to the GPU box, where
/root/reference does not exist.  Fixtures:

  img00_body.npz        real SHAPY_A output shipped by the reference
                        (samples/shapy_fit_for_virtual_measurements/img_00.npz): v_shaped,
                        faces (int32), raw/decoded poses, camera, joints, proj_joints,
                        golden measurements; plus 3 more real v_shaped bodies from
                        regressor/hbw_evaluation/example_shapy_prediction.npz.
  measurement_landmarks.json   the 5 (face_idx, bc) landmarks BodyMeasurements reads from
                        mesh-mesh-intersection/data/{measurement_defitions,smplx_measurements}.yaml
  ref_smplx.npz         reference lbs()/SMPLX.forward outputs on the seeded synthetic model
  ref_head.npz          reference IterativeRegression + ContinuousRotReprDecoder outputs
  ref_hrnet.npz         reference HighResolutionNet outputs on the seeded synthetic checkpoint
  hrnet_keys.json       the 1 967 state-dict keys/shapes of the reference backbone (sha256 + list)
  ref_measure.json      oracle/measure (quirk-faithful op.cu emulation) results on the 4 real bodies
  b2a.npz               reference Polynomial (B2A head) outputs with seeded male / female weights, routed by gender
  p2p.npz               reference v2vhdError (P2P metric) outputs on seeded sparse point regressors and meshes
  v2v.npz               reference PointError(TranslationAlignment / NoAlignment) outputs (`v2v_t` / `v2v`) on two seeded meshes
  preprocess.npz        reference input stage (transf_utils.crop with cv2 + ToTensor + Normalize) on seeded uint8 images:
                        crop windows, crops with OpenCV's portable path (IPP off) and with this container's IPP build
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

G = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(G, exist_ok=True)
REF = ref_shim.REF


def body_fixture():
    d = ref_shim.load_img00()
    ex = np.load(os.path.join(REF, 'regressor/hbw_evaluation/example_shapy_prediction.npz'))
    np.savez_compressed(
        os.path.join(G, 'img00_body.npz'),
        v_shaped=d['v_shaped'].astype(np.float32), faces=d['faces'].astype(np.int32),
        vertices=d['vertices'].astype(np.float32),
        raw_global_rot=d['raw_global_rot'], global_rot=d['global_rot'],
        raw_body_pose=d['raw_body_pose'], body_pose=d['body_pose'], betas=d['betas'],
        camera=d['camera'], joints=d['joints'], proj_joints=d['proj_joints'],
        meas_names=np.array(['mass', 'height', 'chest', 'waist', 'hips']),
        meas_values=np.array([d['measurements'][k] for k in ['mass', 'height', 'chest', 'waist', 'hips']],
                             dtype=np.float64),
        extra_v_shaped=ex['v_shaped'].astype(np.float32))
    import yaml
    mv = yaml.safe_load(open(os.path.join(REF, 'mesh-mesh-intersection/data/smplx_measurements.yaml')))
    md = yaml.safe_load(open(os.path.join(REF, 'mesh-mesh-intersection/data/measurement_defitions.yaml')))
    lm = {}
    for out_name, src in [('head_top', 'HeadTop'), ('left_heel', 'HeelLeft'), ('chest', md['CW_p'][0]),
                          ('waist', md['BW_p'][0]), ('hips', md['IW_p'][0])]:
        lm[out_name] = dict(name=src, face_idx=int(mv[src]['face_idx']), bc=[float(x) for x in mv[src]['bc']])
    json.dump(lm, open(os.path.join(G, 'measurement_landmarks.json'), 'w'), indent=1)
    print('img00_body.npz, measurement_landmarks.json', lm)


def smplx_fixture():
    from shapy_b200 import synth
    ref = ref_shim.load()
    model = synth.make_smplx()
    g = torch.Generator().manual_seed(11)
    B = 3
    betas = torch.randn(B, 10, generator=g)
    raw = torch.randn(B, 22 * 6, generator=g) * 0.4 + synth.mean_params()[:132]
    dec1, dec21 = ref.pu.ContinuousRotReprDecoder(1), ref.pu.ContinuousRotReprDecoder(21)
    with torch.no_grad():
        grot = dec1(raw[:, :6].contiguous())
        bpose = dec21(raw[:, 6:].contiguous())
        out = ref_shim.smplx_forward_ref(model, betas, grot, bpose)
        # config 1: T-pose, B=1
        eye = torch.eye(3).view(1, 1, 3, 3)
        out_t = ref_shim.smplx_forward_ref(model, betas[:1], eye.clone(), eye.expand(1, 21, -1, -1).contiguous())
        # large head rotation to exercise the dynamic-contour LUT on both sides
        raw2 = raw.clone()
        aa = torch.tensor([[0.0, 0.9, 0.0], [0.0, -1.2, 0.1], [0.2, 0.3, 0.0]])
        R = ref.rot.batch_rodrigues(aa)
        # neck joint 12 sits at body_pose index 11 -> flat offset 6 + 11 * 6
        raw2[:, 6 + 11 * 6: 6 + 12 * 6] = R[:, :, :2].reshape(B, 6)
        grot2, bpose2 = dec1(raw2[:, :6].contiguous()), dec21(raw2[:, 6:].contiguous())
        out2 = ref_shim.smplx_forward_ref(model, betas, grot2, bpose2)
    np.savez_compressed(
        os.path.join(G, 'ref_smplx.npz'), betas=betas.numpy(), raw=raw.numpy(), global_rot=grot.numpy(),
        body_pose=bpose.numpy(), vertices=out['vertices'].numpy(), joints=out['joints'].numpy(),
        v_shaped=out['v_shaped'].numpy(), t_vertices=out_t['vertices'].numpy(), t_joints=out_t['joints'].numpy(),
        raw2=raw2.numpy(), joints2=out2['joints'].numpy(), vertices2_sub=out2['vertices'][:, ::97].numpy())
    print('ref_smplx.npz', out['vertices'].shape, out['joints'].shape)


def head_fixture():
    from shapy_b200 import synth
    ref = ref_shim.load()
    mlp = ref.net.MLP(2048 + 145, 145, layers=[1024, 1024], activation={'type': 'none'},
                      normalization={'type': 'none'}, dropout=0.5, gain=0.01)
    it = ref.net.IterativeRegression(mlp, synth.mean_params().view(1, -1), num_stages=3).eval()
    sd = synth.make_head_state_dict()
    it.load_state_dict({k[len('regressor.'):]: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(12)
    feats = torch.randn(5, 2048, generator=g).abs() * 0.5
    with torch.no_grad():
        params, _ = it(feats)
    np.savez_compressed(os.path.join(G, 'ref_head.npz'), feats=feats.numpy(),
                        params=np.stack([p.numpy() for p in params]))
    print('ref_head.npz', params[-1].shape)


def hrnet_fixture():
    from shapy_b200 import synth
    m = ref_shim.build_hrnet()
    sd0 = m.state_dict()
    keys = [[k, list(v.shape)] for k, v in sd0.items()]
    digest = hashlib.sha256('\n'.join(f'{k} {s}' for k, s in keys).encode()).hexdigest()
    json.dump(dict(sha256=digest, n=len(keys), keys=keys), open(os.path.join(G, 'hrnet_keys.json'), 'w'))
    m.load_state_dict(synth.make_state_dict(sd0, seed=1))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 224, 224, generator=g)
    with torch.no_grad():
        out = m(x)
        x64 = torch.randn(1, 3, 64, 96, generator=g)
        out64 = m(x64)
    np.savez_compressed(
        os.path.join(G, 'ref_hrnet.npz'), concat=out['concat'].numpy(),
        layer1_sub=out['layer1'][:, ::7, ::5, ::5].numpy(), layer4=out['layer4'][:, ::16].numpy(),
        concat64=out64['concat'].numpy())
    print('ref_hrnet.npz', out['concat'].shape, float(out['concat'].abs().mean()), len(keys), digest[:12])


def preprocess_fixture():
    """Runs the reference's own crop() (regressor/human_shape/utils/transf_utils.py, loaded by path; cv2 is present in the
    build container) followed by the ToTensor / Normalize arithmetic of data/transforms/transforms.py:603-624,710-733."""
    import importlib.util
    import cv2
    import torchvision.transforms.functional as F
    spec = importlib.util.spec_from_file_location('ref_transf_utils', os.path.join(ref_shim.HS, 'utils', 'transf_utils.py'))
    tu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tu)
    mean, std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]      # config/datasets_defaults.py:37-38
    rng = np.random.default_rng(20240923)
    out = {'mean': np.float32(mean), 'std': np.float32(std)}
    cases = [  # (H, W, center, scale, size): inside, overhanging every border, strong down- and up-scaling
        (96, 128, (64.0, 48.0), 0.40, 64), (96, 128, (5.5, 90.25), 0.55, 64), (120, 90, (80.0, 10.0), 0.9, 32),
        (70, 70, (35.0, 35.0), 0.12, 64), (150, 200, (100.7, 75.2), 1.35, 32), (64, 48, (24.0, 32.0), 0.32, 64),
    ]
    for i, (H, W, center, scale, size) in enumerate(cases):
        # image-like content: smooth gradients plus noise, uint8
        yy, xx = np.mgrid[0:H, 0:W]
        img = np.stack([(xx * 255.0 / W), (yy * 255.0 / H), ((xx + yy) * 255.0 / (H + W))], -1)
        img = np.clip(img + rng.normal(0, 25, img.shape), 0, 255).astype(np.uint8)
        imgf = np.clip(img.astype(np.float32) / 255.0, 0, 1)                      # read_img, img_utils.py:57-61
        c = np.array(center, dtype=np.float32)
        res = {}
        for tag, ipp in (('portable', False), ('ipp', True)):
            cv2.ipp.setUseIPP(ipp)
            crop = tu.crop(imgf, c, scale, [size, size])
            t = F.to_tensor(crop)
            t = torch.clamp(t, 0, 1)
            res[tag] = F.normalize(t, mean=mean, std=std).numpy()
        cv2.ipp.setUseIPP(True)
        ul = np.array(tu.transform([1, 1], c, scale, [size, size], invert=1)) - 1
        br = np.array(tu.transform([size + 1, size + 1], c, scale, [size, size], invert=1)) - 1
        out[f'img{i}'] = img
        out[f'center{i}'] = c
        out[f'scale{i}'] = np.float64(scale)
        out[f'size{i}'] = np.int64(size)
        out[f'ul{i}'] = ul.astype(np.int64)
        out[f'br{i}'] = br.astype(np.int64)
        out[f'out{i}'] = res['portable']
        out[f'ipp_dev{i}'] = np.float64(np.abs(res['portable'] - res['ipp']).max())   # this container's IPP build vs portable
    out['n'] = np.int64(len(cases))
    np.savez_compressed(os.path.join(G, 'preprocess.npz'), **out)
    print('preprocess.npz', os.path.getsize(os.path.join(G, 'preprocess.npz')), 'bytes; max |portable - ipp| =',
          max(float(out[f'ipp_dev{i}']) for i in range(len(cases))))


def b2a_fixture():
    """Runs the reference's own Polynomial module (attributes/attributes/attributes_betas/polynomial.py, loaded by path behind
    a stub of attributes.utils.typing) with seeded male / female weights, and routes by gender exactly as
    regressor/human_shape/models/common/iterative_regressor.py:761-776 does."""
    import importlib.util
    import types
    from loguru import logger
    logger.remove()
    for name in ('attributes', 'attributes.utils', 'attributes.utils.typing'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules['attributes.utils.typing'].Tensor = torch.Tensor
    sys.modules['attributes.utils.typing'].Array = np.ndarray
    spec = importlib.util.spec_from_file_location(
        'ref_polynomial', os.path.join(REF, 'attributes', 'attributes', 'attributes_betas', 'polynomial.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(31)
    males, females = mod.Polynomial(10, 15, degree=2), mod.Polynomial(10, 15, degree=2)
    for p in (males, females):                      # ratings live in [1, 5]: give the synthetic heads that scale
        with torch.no_grad():
            p.linear.weight.mul_(0.5)
            p.linear.bias.copy_(torch.rand(15) * 4 + 1)
    betas = torch.randn(9, 10) * 1.5
    gender_strs = ['male', 'Female', 'f', '', None, 'M', 'neutral', 'female', 'm']
    genders = np.array([x.lower()[0] if (x is not None and x != '') else 'n' for x in gender_strs])
    gm, gf = np.where(genders == 'm')[0], np.where(genders == 'f')[0]
    with torch.no_grad():
        am, af = males(betas[gm, :]), females(betas[gf, :])
        attributes = torch.zeros(betas.shape[0], am.shape[1])
        attributes[gm, :] = am
        attributes[gf, :] = af
    np.savez_compressed(
        os.path.join(G, 'b2a.npz'), betas=betas.numpy(), genders=np.array([g if g is not None else '<none>' for g in gender_strs]),
        w_male=males.linear.weight.detach().numpy(), b_male=males.linear.bias.detach().numpy(),
        w_female=females.linear.weight.detach().numpy(), b_female=females.linear.bias.detach().numpy(),
        indices_000=males.indices_000.numpy(), indices_001=males.indices_001.numpy(), attributes=attributes.numpy())
    print('b2a.npz', os.path.getsize(os.path.join(G, 'b2a.npz')), 'bytes')


def _load_ref_metrics():
    """The reference's regressor/human_shape/utils/metrics.py loaded by path behind stubs of open3d / .np_utils / .typing."""
    import importlib.util
    import types
    from loguru import logger
    logger.remove()
    if 'ref_hs.utils.metrics' in sys.modules:
        return sys.modules['ref_hs.utils.metrics']
    for name in ('open3d', 'ref_hs', 'ref_hs.utils', 'ref_hs.utils.np_utils', 'ref_hs.utils.typing'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules['ref_hs.utils.np_utils'].np2o3d_pcl = None
    for n in ('Tensor', 'Array', 'IntList'):
        setattr(sys.modules['ref_hs.utils.typing'], n, object)
    spec = importlib.util.spec_from_file_location('ref_hs.utils.metrics', os.path.join(ref_shim.HS, 'utils', 'metrics.py'))
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = 'ref_hs.utils'
    sys.modules['ref_hs.utils.metrics'] = mod
    spec.loader.exec_module(mod)
    return mod


def v2v_fixture():
    """Runs the reference's own PointError with TranslationAlignment / NoAlignment (metrics.py:232-277, 335-366), i.e. the
    `v2v_t` / `v2v` evaluation metrics (evaluation.py:192-224, 595-602), on two seeded meshes of the same topology."""
    mod = _load_ref_metrics()
    rng = np.random.default_rng(78)
    B, V = 3, 1500
    est = rng.normal(0, 0.4, (B, V, 3)).astype(np.float32)
    gt = (est + rng.normal(0, 0.01, (B, V, 3)) + rng.normal(0, 0.05, (B, 1, 3))).astype(np.float32)
    out = {}
    for tag, name in (('aligned', 'translation'), ('raw', 'none')):
        pe = mod.PointError(mod.build_alignment(name))
        out[f'error_{tag}'] = np.asarray(pe(est.copy(), gt.copy()), np.float64)
    np.savez_compressed(os.path.join(G, 'v2v.npz'), est=est, gt=gt, **out)
    print('v2v.npz', os.path.getsize(os.path.join(G, 'v2v.npz')), 'bytes; mean aligned error', out['error_aligned'].mean())


def p2p_fixture():
    """Runs the reference's own v2vhdError (regressor/human_shape/utils/metrics.py:368-456, loaded by path behind stubs of
    open3d / .np_utils / .typing) on seeded sparse point regressors (3 barycentric weights per point) and meshes, in
    float64 on the CPU as the evaluator does (evaluation.py:258-260)."""
    import pickle
    import tempfile
    import scipy.sparse as sp
    mod = _load_ref_metrics()
    rng = np.random.default_rng(77)
    P, V1, V2, B = 700, 400, 250, 3

    def regressor(V):
        cols = rng.integers(0, V, (P, 3))
        w = rng.dirichlet([1, 1, 1], P)
        # float32-representable weights stored as float64, the dtype of the reference's regressor pickles (the evaluator
        # multiplies them with .double() vertices)
        return sp.csr_matrix((w.reshape(-1).astype(np.float32).astype(np.float64), (np.repeat(np.arange(P), 3), cols.reshape(-1))),
                             shape=(P, V))
    r_in, r_tg = regressor(V1), regressor(V2)
    r_in.sum_duplicates()
    r_tg.sum_duplicates()
    tmp = tempfile.mkdtemp()
    paths = []
    for name, r in (('in', r_in), ('tg', r_tg)):
        paths.append(os.path.join(tmp, name + '.pkl'))
        with open(paths[-1], 'wb') as f:
            pickle.dump(r, f)
    v_in = (rng.normal(0, 0.3, (B, V1, 3)) + [0.1, -0.2, 0.05]).astype(np.float32)
    v_tg = (rng.normal(0, 0.3, (B, V2, 3)) + [-0.3, 0.4, 0.0]).astype(np.float32)
    out = {}
    for align in (True, False):
        metric = mod.v2vhdError(paths[0], paths[1], align=align)
        if align:
            mean, err = metric(torch.from_numpy(v_in).double(), torch.from_numpy(v_tg).double())
        else:       # metrics.py:449-452 only defines t under `if self.align`: the unaligned variant is the same formula with t = 0
            a = metric.sparse_batch_mm(metric.input_point_regressor.double(), torch.from_numpy(v_in).double())
            c = metric.sparse_batch_mm(metric.target_point_regressor.double(), torch.from_numpy(v_tg).double())
            err = torch.sqrt(torch.pow(a - c, 2).sum(axis=-1))
            mean = err.mean(1)
        out['mean_aligned' if align else 'mean_raw'] = mean.numpy()
        out['error_aligned' if align else 'error_raw'] = err.numpy()
    for name, r in (('in', r_in), ('tg', r_tg)):
        out[f'{name}_row_ptr'], out[f'{name}_col'], out[f'{name}_val'] = (r.indptr.astype(np.int32), r.indices.astype(np.int32),
                                                                             r.data.astype(np.float32))
    np.savez_compressed(os.path.join(G, 'p2p.npz'), v_in=v_in, v_tg=v_tg, **out)
    print('p2p.npz', os.path.getsize(os.path.join(G, 'p2p.npz')), 'bytes; mean aligned error', out['mean_aligned'])


if __name__ == '__main__':
    which = sys.argv[1:] or ['body', 'smplx', 'head', 'hrnet', 'preprocess', 'b2a', 'p2p', 'v2v']
    torch.set_num_threads(8)
    if 'body' in which:
        body_fixture()
    if 'smplx' in which:
        smplx_fixture()
    if 'head' in which:
        head_fixture()
    if 'hrnet' in which:
        hrnet_fixture()
    if 'preprocess' in which:
        preprocess_fixture()
    if 'b2a' in which:
        b2a_fixture()
    if 'p2p' in which:
        p2p_fixture()
    if 'v2v' in which:
        v2v_fixture()
