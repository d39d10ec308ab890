"""TEST INFRASTRUCTURE ONLY -- CPU restatement of BodyMeasurements + the mesh-mesh op.

  BodyMeasurements.forward / compute_mass / compute_height / compute_peripheries
      mesh-mesh-intersection/body_measurements/body_measurements.py:99-246
  mesh_to_mesh_forward semantics: oracle/mmi_oracle.c (restates op.cu)
  hull: scipy.spatial.ConvexHull (qhull), exactly as body_measurements.py:165

Pinned by the reference's golden measurements in img_00.npz (mass 56.868896,
height 1.6437092, chest 0.8745367, waist 0.76514757, hips 0.95468146) to <= 1e-6
relative -- tests/test_oracle_pins.py.  Only tests/, smoke() and bench.py's CPU
baseline legs may import this.
"""
import ctypes
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
DENSITY = 985.0


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, '_build', 'libmmi_oracle.so')
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build()
        _LIB = ctypes.CDLL(path)
        _LIB.mmi_oracle_forward.restype = ctypes.c_int
        _LIB.mmi_oracle_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def mesh_to_mesh_forward(query: np.ndarray, target: np.ndarray, max_collisions: int = 16):
    """query (B,Q,3,3), target (B,F,3,3) float32 -> faces (B,Q*M) int64 (-1 empty), bcs (B,Q*M,2,3)."""
    query = np.ascontiguousarray(query, np.float32)
    target = np.ascontiguousarray(target, np.float32)
    B, Q = query.shape[:2]
    F = target.shape[1]
    M = max_collisions
    faces = np.full((B, Q * M), -1, np.int64)
    bcs = np.zeros((B, Q * M, 2, 3), np.float32)
    L = lib()
    for b in range(B):
        L.mmi_oracle_forward(query[b].ctypes.data, target[b].ctypes.data, Q, F, M, faces[b].ctypes.data,
                             bcs[b].ctypes.data)
    return faces, bcs


def load_landmarks(path=None):
    path = path or os.path.join(os.path.dirname(_HERE), 'tests', 'golden', 'measurement_landmarks.json')
    return json.load(open(path))


def plane_quad(h):
    v = np.array([[-1, h, -1], [1, h, -1], [1, h, 1], [-1, h, 1]], np.float32)
    return v[np.array([[0, 1, 2], [0, 2, 3]])]


def periphery(tris: np.ndarray, face_idx: int, bc, max_collisions=256):
    """One body, one measurement.  tris (F,3,3) float32.  Returns (value, points (n,3))."""
    from scipy.spatial import ConvexHull
    bc = np.asarray(bc, np.float32)
    vertex = (tris[face_idx] * bc.reshape(3, 1)).sum(axis=0, dtype=np.float32)
    q = plane_quad(np.float32(vertex[1]))
    faces, bcs = mesh_to_mesh_forward(q[None], tris[None], max_collisions)
    faces, bcs = faces[0], bcs[0]
    valid = np.where(faces > 0)[0]                      # NB excludes face 0, body_measurements.py:161
    pts = (tris[faces[valid]][:, None] * bcs[valid][..., None]).sum(axis=-2, dtype=np.float32)  # (n,2,3)
    flat = pts.reshape(-1, 3)
    hull = ConvexHull(flat[:, [0, 2]])
    seg = flat[hull.simplices.reshape(-1)].reshape(-1, 2, 3)
    d = (seg[:, 1] - seg[:, 0]).astype(np.float32)
    value = np.sqrt((d * d).sum(-1, dtype=np.float32)).sum(dtype=np.float32)
    return float(value), flat, float(vertex[1])


def mass(tris: np.ndarray) -> float:
    x, y, z = tris[..., 0], tris[..., 1], tris[..., 2]
    vol = (-x[:, 2] * y[:, 1] * z[:, 0] + x[:, 1] * y[:, 2] * z[:, 0] + x[:, 2] * y[:, 0] * z[:, 1]
           - x[:, 0] * y[:, 2] * z[:, 1] - x[:, 1] * y[:, 0] * z[:, 2] + x[:, 0] * y[:, 1] * z[:, 2])
    return float(np.abs(vol.sum(dtype=np.float32)) / np.float32(6.0) * np.float32(DENSITY))


def height(tris: np.ndarray, lm) -> float:
    ht = (tris[lm['head_top']['face_idx']] * np.asarray(lm['head_top']['bc'], np.float32).reshape(3, 1)).sum(0)
    hl = (tris[lm['left_heel']['face_idx']] * np.asarray(lm['left_heel']['bc'], np.float32).reshape(3, 1)).sum(0)
    return float(np.abs(np.float32(ht[1]) - np.float32(hl[1])))


def measure(v_shaped: np.ndarray, faces: np.ndarray, lm=None) -> dict:
    """v_shaped (V,3) float32, faces (F,3) -> {'mass','height','chest','waist','hips'}."""
    lm = lm or load_landmarks()
    tris = np.ascontiguousarray(v_shaped[faces].astype(np.float32))
    out = dict(mass=mass(tris), height=height(tris, lm))
    for name in ('chest', 'waist', 'hips'):
        out[name] = periphery(tris, lm[name]['face_idx'], lm[name]['bc'])[0]
    return out
