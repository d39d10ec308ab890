"""TEST INFRASTRUCTURE ONLY -- runs the REFERENCE's mesh_to_mesh_forward (oracle/_ref, built by oracle/build_ref.py) in
its own process and stores the outputs.  A separate process because the reference kernel calls exit(0) on any CUDA
error (op.cu:81-85), synchronises the device after every launch and uses raw cudaMalloc / the legacy default stream.

    python oracle/run_ref_mmi.py in.npz out.npz     # in: query (B,Q,3,3) f32, target (B,F,3,3) f32, max_collisions
One body per call, so the reference's scratch buffer of barycentrics starts zeroed for every body (op.cu:1002-1011)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(inp, outp):
    import torch
    from oracle import build_ref
    mod = build_ref.load()
    d = np.load(inp)
    q, t, m = d['query'], d['target'], int(d['max_collisions'])
    faces, bcs = [], []
    for b in range(q.shape[0]):
        f, c = mod.mesh_to_mesh_forward(torch.from_numpy(q[b:b + 1]).cuda().contiguous(),
                                        torch.from_numpy(t[b:b + 1]).cuda().contiguous(), m)
        torch.cuda.synchronize()
        faces.append(f.cpu().numpy())
        bcs.append(c.cpu().numpy())
    np.savez(outp, faces=np.concatenate(faces, 0), bcs=np.concatenate(bcs, 0), ok=np.int64(1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
