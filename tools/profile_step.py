"""One profiled step of the hot path for ncu (cudaProfilerStart/Stop around the step; run with
`ncu --profile-from-start off ...`).  args: [batch] [what: full|hrnet|lbs]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapy_b200 import synth, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
what = sys.argv[2] if len(sys.argv) > 2 else 'full'
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = synth.build_synthetic_regressor()
model.backbone.precision_mode = mode
model = model.cuda().eval()
x = torch.randn(B, 3, 224, 224, device='cuda')


def step():
    with torch.no_grad():
        if what == 'hrnet':
            return model.backbone(x)['concat']
        if what == 'lbs':
            packed = model.model.packed('cuda')
            betas = torch.randn(B, 10, device='cuda')
            rot = ops.decode_rot6d(torch.randn(B, 132, device='cuda') * 0.3 + synth.mean_params()[:132].cuda())
            ops.smplx_forward(packed, betas, rot)
            b4096 = torch.randn(4096, 10, device='cuda').clamp(-3, 3)
            vs = ops.smplx_forward_shape(packed, b4096)
            return ops.measure(model.body_measurements.landmarks(), v_shaped=vs, faces_i32=model.model.faces_i32)
        return model(x)


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
