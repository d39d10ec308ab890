"""Output writers of the demo, batched and off the critical path.

Host mirror of regressor/demo.py:
    weak_persp_to_blender   demo.py:70-106    weak-perspective camera -> translation / focal length / shifts per person
    per-image outputs       demo.py:305-353   `<name>.npz` (np.savez_compressed: fname, every entry of the last stage,
                                              the camera entries) and `<name>.ply` (vertices + camera translation, faces)
The reference writes the files of a batch one after the other on the main thread, after a blocking `.cpu()` of every
tensor.  Here the arrays arrive in pinned host buffers (shapy_b200.pipeline.HostPipeline) and a thread pool writes
them while the next batches run.  The .ply files are binary little-endian PLY with the element / property layout
trimesh's exporter uses (float x, y, z; list uchar int vertex_indices), readable by trimesh / MeshLab / Blender.
Pure host I/O: no GPU work and no arithmetic beyond demo.py:85-102.
"""
import os
import threading
import os.path as osp
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Optional

import numpy as np

try:  # tensors are optional here: the writer accepts numpy arrays as well
    import torch
except ImportError:  # pragma: no cover
    torch = None


def _np(x):
    if torch is not None and torch.is_tensor(x):
        return x.detach().cpu().numpy()
    if hasattr(x, '_t'):                      # KeypointTensor (demo.py:343-344)
        return x._t.detach().cpu().numpy()
    return x


def weak_persp_to_blender(targets, camera_scale, camera_transl, H, W, sensor_width=36, focal_length=5000):
    """demo.py:70-106, operation by operation."""
    camera_scale, camera_transl = _np(camera_scale), _np(camera_transl)
    output = defaultdict(lambda: [])
    for ii, target in enumerate(targets):
        orig_bbox_size = target.get_field('orig_bbox_size')
        bbox_center = target.get_field('orig_center')
        z = 2 * focal_length / (camera_scale[ii] * orig_bbox_size)
        transl = [camera_transl[ii, 0].item(), camera_transl[ii, 1].item(), z.item()]
        shift_x = - (bbox_center[0] / W - 0.5)
        shift_y = (bbox_center[1] - 0.5 * H) / W
        focal_length_in_mm = focal_length / W * sensor_width
        output['shift_x'].append(shift_x)
        output['shift_y'].append(shift_y)
        output['transl'].append(transl)
        output['focal_length_in_mm'].append(focal_length_in_mm)
        output['focal_length_in_px'].append(focal_length)
        output['center'].append(bbox_center)
        output['sensor_width'].append(sensor_width)
    for key in output:
        output[key] = np.array(output[key])
    return output


def write_ply(path: str, vertices: np.ndarray, faces: np.ndarray) -> None:
    """Binary little-endian PLY: float32 vertices (V, 3), int32 triangles (F, 3)."""
    v = np.ascontiguousarray(vertices, dtype='<f4').reshape(-1, 3)
    f = np.ascontiguousarray(faces, dtype='<i4').reshape(-1, 3)
    header = ('ply\nformat binary_little_endian 1.0\ncomment shapy_b200\n'
              f'element vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n'
              f'element face {len(f)}\nproperty list uchar int vertex_indices\nend_header\n').encode('ascii')
    rec = np.empty(len(f), dtype=[('n', 'u1'), ('idx', '<i4', (3,))])
    rec['n'] = 3
    rec['idx'] = f
    tmp = f'{path}.{os.getpid()}.{threading.get_ident()}.tmp'     # unique: two jobs may target the same file
    with open(tmp, 'wb') as fh:
        fh.write(header)
        fh.write(v.tobytes())
        fh.write(rec.tobytes())
    os.replace(tmp, path)


def read_ply(path: str):
    """Reader for the files write_ply produces (tests, round trips)."""
    with open(path, 'rb') as fh:
        data = fh.read()
    end = data.index(b'end_header\n') + len(b'end_header\n')
    head = data[:end].decode('ascii').splitlines()
    nv = int([l for l in head if l.startswith('element vertex')][0].split()[-1])
    nf = int([l for l in head if l.startswith('element face')][0].split()[-1])
    v = np.frombuffer(data, dtype='<f4', count=nv * 3, offset=end).reshape(nv, 3)
    rec = np.frombuffer(data, dtype=[('n', 'u1'), ('idx', '<i4', (3,))], count=nf, offset=end + nv * 12)
    return v.copy(), rec['idx'].copy()


def output_dir(output_folder: str, target) -> str:
    """demo.py:311-317: `<folder>/<f1>/<f2>` when the target carries a `filename`, else `<folder>`."""
    filename = target.get_field('filename', '') if hasattr(target, 'get_field') else ''
    if filename != '':
        f1, f2 = filename.split('/')[-3:-1]
        return osp.join(output_folder, f1, f2)
    return output_folder


class ResultWriter:
    """Writes the per-image .npz / .ply files of demo.py:305-353 from host arrays on a pool of threads."""

    def __init__(self, output_folder: str, save_params: bool = True, save_mesh: bool = True, workers: int = 4):
        self.output_folder, self.save_params, self.save_mesh = output_folder, save_params, save_mesh
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.pending = []
        self._last = {}

    def _write_one(self, idx, target, stage_out: Dict, hd_params: Dict, faces, after=None):
        if after is not None:
            after.result()                 # an earlier job writes the same files: keep the reference's order (last one wins)
        fname = target.get_field('fname')
        path = output_dir(self.output_folder, target)
        imgfname = fname.split('.')[0]
        os.makedirs(path, exist_ok=True)
        if self.save_mesh and 'vertices' in stage_out:
            write_ply(osp.join(path, f'{imgfname}.ply'), stage_out['vertices'][0][idx] + hd_params['transl'][idx], faces)
        if self.save_params:
            out_params = dict(fname=fname)
            for key, (val, per_image) in stage_out.items():
                out_params[key] = val[idx] if per_image else val          # demo.py:340-345
            for key, val in hd_params.items():
                out_params[key] = val[idx].item() if np.isscalar(val[idx]) else val[idx]
            tmp = osp.join(path, f'{imgfname}.{os.getpid()}.{threading.get_ident()}.tmp.npz')
            np.savez_compressed(tmp, **out_params)
            os.replace(tmp, osp.join(path, f'{imgfname}.npz'))

    def submit(self, targets, stage_out: Dict, hd_params: Dict, faces: Optional[np.ndarray] = None):
        """Queues the files of one batch.  stage_out: the last stage's entries as HOST arrays (numpy, CPU tensors or
        KeypointTensor); they must stay unchanged until flush()/close() -- hand over copies of reused pinned buffers."""
        # tensors / KeypointTensors are stored per image, everything else (faces, dicts) as it is (demo.py:340-345)
        targets = list(targets)
        host = {}
        for k, v in stage_out.items():
            # per image: tensors / KeypointTensors (demo.py:340-345) and numpy arrays with one row per target; the mesh
            # topology (`faces`) and anything else is stored whole
            per_image = (torch is not None and torch.is_tensor(v)) or hasattr(v, '_t') or (
                isinstance(v, np.ndarray) and k != 'faces' and v.ndim >= 1 and len(v) == len(targets))
            host[k] = (_np(v), per_image)
        if 'faces' in host:
            faces = host['faces'][0]
        if self.save_mesh and 'vertices' in host and faces is None:
            raise ValueError('ResultWriter.submit: save_mesh needs the mesh topology (stage_out[\'faces\'] or faces=)')
        hd = {k: np.asarray(v) for k, v in hd_params.items()}
        for idx, target in enumerate(targets):
            key = (output_dir(self.output_folder, target), target.get_field('fname').split('.')[0])
            job = self.pool.submit(self._write_one, idx, target, host, hd, faces, self._last.get(key))
            self._last[key] = job          # jobs with the same output files run in submission order
            self.pending.append(job)

    def flush(self):
        for f in self.pending:
            f.result()                     # re-raises a writer's exception
        self.pending = []
        self._last = {}

    def close(self):
        self.flush()
        self.pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
