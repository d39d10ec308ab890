"""End-to-end: image -> HRNet -> head -> SMPL-X -> measurements through the reference-facing module
(`SMPLXRegressor.forward`), against the CPU oracle pipeline on the same seeded synthetic inputs.
Bar (BASELINE.json north_star): vertices, betas, measurements within 1e-4 relative fp32."""
import numpy as np
import pytest
import torch

from oracle import measure_oracle, net_oracle, smplx_oracle
from shapy_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def test_full_regressor_vs_oracle():
    model = synth.build_synthetic_regressor()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    smplx = synth.make_smplx()
    lm = synth.load_landmarks()
    B = 3
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    # ---- oracle
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        feats = net_oracle.hrnet_forward({k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')}, x)['concat']
        params = net_oracle.head_forward(sd, feats)
        p = params[-1]
        grot, bpose = smplx_oracle.decode_6d(p[:, :6]), smplx_oracle.decode_6d(p[:, 6:132])
        betas, cam = p[:, 132:142], p[:, 142:145]
        body = smplx_oracle.smplx_forward(smplx, betas, grot, bpose)
        proj = smplx_oracle.weak_persp(body['joints'], cam)
    faces = smplx['faces_tensor'].numpy()
    meas = [measure_oracle.measure(body['v_shaped'][b].numpy(), faces, lm) for b in range(B)]
    # ---- product path
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(x.cuda(), targets=None, full_imgs=None, device=torch.device('cuda'))
    st = out['stage_02']
    assert out['num_stages'] == 3 and out['stage_keys'] == ['stage_00', 'stage_01', 'stage_02']
    assert rel(out['features'], feats) < TOL
    for k in range(3):
        sk = out[f'stage_{k:02d}']
        assert rel(sk['betas'], params[k][:, 132:142]) < TOL
        assert rel(sk['raw_body_pose'], params[k][:, 6:132]) < TOL
        assert rel(sk['body_pose'], smplx_oracle.decode_6d(params[k][:, 6:132])) < TOL
        assert isinstance(sk['faces'], np.ndarray) and sk['faces'].dtype == np.int64
    assert rel(st['vertices'], body['vertices']) < TOL
    assert rel(st['v_shaped'], body['v_shaped']) < TOL
    assert rel(st['joints']._t, body['joints']) < TOL
    assert rel(out['proj_joints']._t, proj) < TOL
    assert rel(st['global_rot'], grot) < TOL
    assert rel(st['camera'], cam) < TOL
    for name in ('mass', 'height', 'chest', 'waist', 'hips'):
        ref = torch.tensor([m[name] for m in meas])
        assert rel(out['measurements'][name], ref) < TOL, name
        assert out['stage_02']['measurements'][name] is out['measurements'][name]
    # deepcopy / float() / state_dict round trip as evaluate.py does (evaluation.py:647-651)
    import copy
    m2 = copy.deepcopy(model).float().eval()
    with torch.no_grad():
        out2 = m2(x.cuda())
    assert torch.equal(out2['stage_02']['vertices'], st['vertices'])


def test_host_pipeline_matches_direct_calls():
    """shapy_b200.pipeline.HostPipeline (overlapped H2D / forward / D2H) returns exactly what direct calls return."""
    import torch
    from shapy_b200 import synth
    from shapy_b200.pipeline import HostPipeline, pack_result
    model = synth.build_synthetic_regressor().cuda().eval()
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(3, 3, 224, 224, generator=g).pin_memory() for _ in range(4)]
    outs = [dict(vertices=torch.empty(3, 10475, 3).pin_memory(), betas=torch.empty(3, 10).pin_memory(),
                 measurements=torch.empty(3, 5).pin_memory()) for _ in range(4)]
    pipe = HostPipeline(model, 'cuda')
    for x, o in zip(xs, outs):
        pipe.submit(x, o)
    pipe.drain()
    torch.cuda.synchronize()
    for x, o in zip(xs, outs):
        with torch.no_grad():
            ref = pack_result(model(x.cuda()))
        for k in o:
            assert torch.equal(o[k], ref[k].cpu()), k


def rms_rel(a, b):
    a, b = a.detach().double().cpu(), torch.as_tensor(b).double().cpu()
    return float(((a - b) ** 2).mean().sqrt() / (b ** 2).mean().sqrt())


def test_full_regressor_B64_rows_vs_oracle():
    """BASELINE configs[2] as benchmarked: B = 64, 224 x 224, full regressor.  The CPU oracle runs on 6 of the 64
    images (the path has no cross-sample term); every output of those rows within 1e-4 (max-relative) and 5e-5 RMS."""
    model = synth.build_synthetic_regressor()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    smplx, lm = synth.make_smplx(), synth.load_landmarks()
    B = 64
    rows = [0, 9, 31, 32, 46, 63]
    x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        feats = net_oracle.hrnet_forward({k[len('backbone.'):]: v for k, v in sd.items() if k.startswith('backbone.')},
                                         x[rows])['concat']
        p = net_oracle.head_forward(sd, feats)[-1]
        body = smplx_oracle.smplx_forward(smplx, p[:, 132:142], smplx_oracle.decode_6d(p[:, :6]),
                                          smplx_oracle.decode_6d(p[:, 6:132]))
    faces = smplx['faces_tensor'].numpy()
    meas = [measure_oracle.measure(body['v_shaped'][i].numpy(), faces, lm) for i in range(len(rows))]
    model = model.cuda().eval()
    with torch.no_grad():
        out = model(x.cuda())
    st = out['stage_02']
    checks = dict(features=(out['features'][rows], feats), betas=(st['betas'][rows], p[:, 132:142]),
                  vertices=(st['vertices'][rows], body['vertices']), v_shaped=(st['v_shaped'][rows], body['v_shaped']),
                  joints=(st['joints']._t[rows], body['joints']))
    for k, (a, b) in checks.items():
        assert rel(a, b) < TOL, (k, rel(a, b))
        assert rms_rel(a, b) < 5e-5, (k, rms_rel(a, b))
    for name in ('mass', 'height', 'chest', 'waist', 'hips'):
        ref = torch.tensor([m[name] for m in meas])
        got = out['measurements'][name][rows].cpu()
        assert float(((got - ref).abs() / ref.abs()).max()) < TOL, name
