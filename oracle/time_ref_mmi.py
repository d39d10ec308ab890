"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- times the REFERENCE's own mesh_to_mesh_forward kernel (oracle/_ref, the
one-token-patched build of mesh-mesh-intersection/src/mesh_mesh_intersect_cuda_op.cu) on this GPU, in its own process
(the kernel exit(0)s on CUDA errors).  It is "the kernel to beat" of SURVEY.md 8d for BASELINE configs[3].

    python oracle/time_ref_mmi.py in.npz out.json [reps]
in.npz: query (B,Q,3,3) f32, target (B,F,3,3) f32, max_collisions.  The reference loops over the batch itself
(op.cu:1011) and synchronises the device after every launch, so wall clock == device time; one call processes all B.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(inp, outp, reps):
    import torch
    from oracle import build_ref
    mod = build_ref.load()
    d = np.load(inp)
    q = torch.from_numpy(d['query']).cuda().contiguous()
    t = torch.from_numpy(d['target']).cuda().contiguous()
    m = int(d['max_collisions'])
    for _ in range(2):
        mod.mesh_to_mesh_forward(q, t, m)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        mod.mesh_to_mesh_forward(q, t, m)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    json.dump({'ms_per_call': ms, 'bodies': int(q.shape[0]), 'ms_per_body': ms / q.shape[0], 'reps': reps}, open(outp, 'w'))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 5)
