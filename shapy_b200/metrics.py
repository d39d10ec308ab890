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
        if B == 0:
            return mean, error
        with torch.cuda.device(dev):
            L = _lib.lib()
            nbytes = L.shapy_p2p_workspace_bytes(B, P)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.shapy_p2p_error(*[_lib.ptr(t) for t in a['input']], *[_lib.ptr(t) for t in a['target']], _lib.ptr(vi),
                                         _lib.ptr(vt), B, P, V1, V2, int(bool(self.align)), _lib.ptr(error), _lib.ptr(mean), _lib.ptr(ws),
                                         nbytes, _lib.stream_ptr()), 'p2p_error')
        return mean, error


class NoAlignment:
    """metrics.py:85-98."""
    name = 'none'

    def __repr__(self):
        return 'NoAlignment'


class TranslationAlignment:
    """metrics.py:232-277: est + (mean(gt) - mean(est))."""
    name = 'translation'

    def __repr__(self):
        return 'ScaleAlignment'       # sic: the reference's __repr__ (metrics.py:236-237)


def build_alignment(name: str, **kwargs):
    """metrics.py:14-29 for the alignments evaluated on the device; Procrustes / root / scale are not built."""
    if name == 'translation':
        return TranslationAlignment()
    if name in ('no', 'none'):
        return NoAlignment()
    raise NotImplementedError(f'shapy_b200.metrics: alignment `{name}` is not built (translation / none are)')


class PointError:
    """metrics.py:335-366 for same-topology meshes on the device: `PointError(build_alignment('translation'))(est, gt)`
    returns the (B, V) per-vertex error the reference's `_compute_v2v` stores (evaluation.py:192-224), as a CUDA fp32
    tensor; `.last_mean` keeps the per-body mean of the same launch.  The reference copies both meshes to the host."""

    def __init__(self, alignment_object, name: str = ''):
        if not isinstance(alignment_object, (TranslationAlignment, NoAlignment)):
            raise NotImplementedError('shapy_b200 PointError: only translation / no alignment run on the device')
        self._alignment, self._name, self.last_mean = alignment_object, name, None

    name = property(lambda self: self._name)

    def __repr__(self):
        return f'PointError: Alignment = {self._alignment}'

    def set_alignment(self, alignment_object):
        self.__init__(alignment_object, self._name)

    def __call__(self, est_points, gt_points):
        for nm, t in (('est_points', est_points), ('gt_points', gt_points)):
            if not (torch.is_tensor(t) and t.is_cuda):
                raise RuntimeError(f'shapy_b200 PointError: {nm} must be a CUDA tensor (there is no CPU fallback)')
        if est_points.shape != gt_points.shape or est_points.dim() != 3 or est_points.shape[2] != 3:
            raise ValueError('PointError: expected two (B, P, 3) tensors of the same shape')
        dev = est_points.device
        a, b = est_points.contiguous().float(), gt_points.to(dev).contiguous().float()
        B, V = a.shape[:2]
        error = torch.empty(B, V, dtype=torch.float32, device=dev)
        mean = torch.empty(B, dtype=torch.float32, device=dev)
        if B == 0 or V == 0:
            self.last_mean = mean.fill_(float('nan')) if B else mean
            return error
        with torch.cuda.device(dev):
            L = _lib.lib()
            nbytes = L.shapy_p2p_workspace_bytes(B, V)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(L.shapy_v2v_error(_lib.ptr(a), _lib.ptr(b), B, V, int(isinstance(self._alignment, TranslationAlignment)),
                                         _lib.ptr(error), _lib.ptr(mean), _lib.ptr(ws), nbytes, _lib.stream_ptr()), 'v2v_error')
        self.last_mean = mean
        return error
