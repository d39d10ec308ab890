"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's point-to-point evaluation metric.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Reference: regressor/human_shape/utils/metrics.py:368-456 (v2vhdError)
    sparse_batch_mm   416-428   HD points = point_regressor (P x V, sparse) @ vertices, per body
    get_translation   430-439   t = gt_points.mean(1) - pred_points.mean(1)   (align)
    __call__          441-456   diff = hd_input + t - hd_target; error = sqrt((diff ** 2).sum(-1)); return error.mean(1), error
(the evaluator feeds it float64 CPU tensors, evaluation.py:258-260).  Same arithmetic in numpy for
regressor/hbw_evaluation/evaluate_hbw.py:44-58.

Pinned: tests/golden/p2p.npz holds outputs of the reference's own v2vhdError (loaded by path in the build container by
tools/make_golden.py) on seeded sparse regressors and meshes; tests/test_metrics_cpu.py checks this file against them.
"""
import numpy as np


def csr_points(row_ptr, col, val, verts):
    """(B, V, 3) float -> (B, P, 3) float64: sparse regressor times vertices."""
    verts = np.asarray(verts, np.float64)
    P = len(row_ptr) - 1
    out = np.zeros((verts.shape[0], P, 3), np.float64)
    rows = np.repeat(np.arange(P), np.diff(row_ptr))
    np.add.at(out, (slice(None), rows), np.asarray(val, np.float64)[None, :, None] * verts[:, np.asarray(col)])
    return out


def p2p_error(in_csr, tg_csr, input_vertices, target_vertices, align=True):
    """Returns (mean error (B,), error (B, P)) in float64, metrics.py:441-456."""
    a = csr_points(*in_csr, input_vertices)
    c = csr_points(*tg_csr, target_vertices)
    t = c.mean(1) - a.mean(1) if align else np.zeros((a.shape[0], 3))
    diff = a + t[:, None, :] - c
    err = np.sqrt((diff ** 2).sum(-1))
    return err.mean(1), err


def v2v_error(est, gt, align=True):
    """PointError(TranslationAlignment() | NoAlignment()), regressor/human_shape/utils/metrics.py:232-277 + 31-52:
    est' = est + (gt.mean(1) - est.mean(1)); error = sqrt(((est' - gt) ** 2).sum(-1)) -> (B, V), float64."""
    est, gt = np.asarray(est, np.float64), np.asarray(gt, np.float64)
    t = gt.mean(1, keepdims=True) - est.mean(1, keepdims=True) if align else 0.0
    return np.sqrt(((est + t - gt) ** 2).sum(-1))
