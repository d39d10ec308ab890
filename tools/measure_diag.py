"""Which capacity do the 4 096 random-beta bodies of BASELINE configs[3] hit in the measurement kernel?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shapy_b200 import ops, synth
smplx, lm = synth.make_smplx(), synth.load_landmarks()
packed = ops.SmplxModel(dict(smplx), 'cuda')
betas = torch.randn(4096, 10, generator=torch.Generator().manual_seed(3)).clamp(-3, 3)
vs = ops.smplx_forward_shape(packed, betas.cuda())
faces = smplx['faces_tensor'].to(torch.int32).cuda()
out, pts, cnt, status = ops.measure(ops.make_landmarks(lm), v_shaped=vs, faces_i32=faces, return_points=True)
torch.cuda.synchronize()
print('status', int(status.item()), 'nan rows', int(torch.isnan(out).any(1).sum()), 'max points per plane', cnt.max(0).values.tolist(),
      'mean', cnt.float().mean(0).tolist())
bad = torch.isnan(out).any(1).nonzero().flatten()[:8].tolist()
print('first bad bodies', bad, [cnt[b].tolist() for b in bad], [float(betas[b].abs().max()) for b in bad])
