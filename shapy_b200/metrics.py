"""Point-to-point evaluation metric on the GPU.

Host mirror of regressor/human_shape/utils/metrics.py:368-456 (`v2vhdError`): same constructor (two pickled scipy sparse
point regressors), same registered buffers (`input_point_regressor`, `target_point_regressor`, sparse COO) and the same
call `(input_points, target_points) -> (error.mean(1), error)`.  The reference evaluates it in float64 on the CPU after
copying every predicted mesh to the host (evaluation.py:227-265); here the meshes stay on the device and the result is
fp32 (`shapy_p2p_error`, csrc/metrics.cu).  There is no CPU fallback.
"""
import pickle

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class v2vhdError(nn.Module):
    def __init__(self, input_point_regressor_path: str = '', target_point_regressor_path: str = '', align: bool = True) -> None:
        super().__init__()
        self.align = align
        self._csr = {}
        for name, path in (('input', input_point_regressor_path), ('target', target_point_regressor_path)):
            with open(path, 'rb') as f:
                m = pickle.load(f)
            self.register_buffer(f'{name}_point_regressor', self.to_pytorch(m))
            c = m.tocsr()
            c.sum_duplicates()
            self._csr[name] = (torch.from_numpy(c.indptr.astype(np.int32)), torch.from_numpy(c.indices.astype(np.int32)),
                               torch.from_numpy(c.data.astype(np.float32)), c.shape)
        if self._csr['input'][3][0] != self._csr['target'][3][0]:
            raise ValueError('v2vhdError: the two point regressors must produce the same number of points')
        self._dev = {}

    def to_pytorch(self, point_regressor):
        """metrics.py:398-414: scipy sparse -> torch sparse COO."""
        point_regressor = point_regressor.tocoo()
        indices = np.vstack((point_regressor.row, point_regressor.col))
        return torch.sparse_coo_tensor(indices, point_regressor.data, point_regressor.shape)

    def _arrays(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {n: tuple(t.to(device) for t in self._csr[n][:3]) for n in ('input', 'target')}
        return self._dev[key]

    def __call__(self, input_points, target_points):
        for name, t in (('input_points', input_points), ('target_points', target_points)):
            if not (torch.is_tensor(t) and t.is_cuda):
                raise RuntimeError(f'shapy_b200 v2vhdError: {name} must be a CUDA tensor (there is no CPU fallback)')
        P, V1 = self._csr['input'][3]
        _, V2 = self._csr['target'][3]
        if input_points.shape[1:] != (V1, 3) or target_points.shape[1:] != (V2, 3) or input_points.shape[0] != target_points.shape[0]:
            raise ValueError(f'v2vhdError: expected (B, {V1}, 3) and (B, {V2}, 3) vertices')
        dev = input_points.device
        a = self._arrays(dev)
        vi = input_points.contiguous().float()
        vt = target_points.to(dev).contiguous().float()
        B = vi.shape[0]
        error = torch.empty(B, P, dtype=torch.float32, device=dev)
        mean = torch.empty(B, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            L = _lib.lib()
            nbytes = L.shapy_p2p_workspace_bytes(B, P)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.shapy_p2p_error(*[_lib.ptr(t) for t in a['input']], *[_lib.ptr(t) for t in a['target']], _lib.ptr(vi),
                                         _lib.ptr(vt), B, P, V1, V2, int(bool(self.align)), _lib.ptr(error), _lib.ptr(mean), _lib.ptr(ws),
                                         nbytes, _lib.stream_ptr()), 'p2p_error')
        return mean, error
