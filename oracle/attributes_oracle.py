"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's B2A attribute head.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Reference:
    Polynomial            attributes/attributes/attributes_betas/polynomial.py:21-140
        _combinations     55-59   itertools.combinations_with_replacement(range(n), i) for i in 1..degree
        build_polynomial_coeffs 61-69   A = cat([prod(X[:, indices_k], -1) for k in range(degree)], -1)
        forward           137-140 linear(A)
    per-gender routing    regressor/human_shape/models/common/iterative_regressor.py:761-776
        genders 'm' -> b2a_males, 'f' -> b2a_females, anything else -> a row of zeros

Pinned: tests/golden/b2a.npz holds outputs of the reference's own Polynomial module (loaded by path in the build
container by tools/make_golden.py) on seeded weights and betas; tests/test_attributes_cpu.py checks this file
against them.
"""
from itertools import combinations_with_replacement

import numpy as np


def feature_indices(n: int, degree: int = 2):
    """The index tuples in the order polynomial.py:38-59 registers them (degree 1 first, then degree 2)."""
    out = []
    for d in range(1, degree + 1):
        out += list(combinations_with_replacement(range(n), d))
    return out


def polynomial_features(x: np.ndarray, degree: int = 2) -> np.ndarray:
    """(B, n) -> (B, n_feat) float32, polynomial.py:61-69."""
    x = np.asarray(x, np.float32)
    cols = [np.prod(x[:, list(idx)], axis=-1, dtype=np.float32) for idx in feature_indices(x.shape[1], degree)]
    return np.stack(cols, -1).astype(np.float32)


def polynomial_forward(x, weight, bias, degree: int = 2) -> np.ndarray:
    """polynomial.py:137-140: Linear on the features (float64 accumulation; the bar against fp32 GEMMs is 1e-6 relative)."""
    a = polynomial_features(x, degree).astype(np.float64)
    return (a @ np.asarray(weight, np.float64).T + np.asarray(bias, np.float64)).astype(np.float32)


def b2a_by_gender(betas, genders, male, female) -> np.ndarray:
    """iterative_regressor.py:761-776.  genders: sequence of str or None; male / female: (weight, bias)."""
    betas = np.asarray(betas, np.float32)
    g = np.array([x.lower()[0] if (x is not None and x != '') else 'n' for x in genders])
    out = np.zeros((betas.shape[0], np.asarray(male[0]).shape[0]), np.float32)
    m, f = np.where(g == 'm')[0], np.where(g == 'f')[0]
    if len(m):
        out[m] = polynomial_forward(betas[m], *male)
    if len(f):
        out[f] = polynomial_forward(betas[f], *female)
    return out


def a2b_features(rating, selected_idx, mmts: dict, selected_mmts, bodytalk_meas_preprocess: bool = False) -> np.ndarray:
    """attributes/attributes/attributes_betas/a2b.py:569-592 (no noise): rating[:, selected] followed by one column per
    selected measurement name (height * 100 and cube root of mass / weight under bodytalk_meas_preprocess)."""
    fv = np.asarray(rating, np.float32)[:, list(selected_idx)]
    for name in selected_mmts:
        m = np.asarray(mmts[name], np.float32).reshape(-1, 1)
        if bodytalk_meas_preprocess:
            if 'height' in name:
                m = m * 100
            if 'mass' in name or 'weight' in name:
                m = np.power(m, np.float32(1.0 / 3.0))
        fv = np.hstack((fv, m))
    return fv.astype(np.float32)


def a2b_by_gender(feat_m, feat_f, genders, male, female, linear: bool = False) -> np.ndarray:
    """regressor/human_shape/models/common/iterative_regressor.py:837-850: betas_ref rows of the males from
    a2b_males(male features), of the females from a2b_females(female features), zeros elsewhere."""
    g = np.array([x.lower()[0] if (x is not None and x != '') else 'n' for x in genders])
    out = np.zeros((len(g), np.asarray(male[0]).shape[0]), np.float32)

    def net(x, w, b):
        if linear:
            return (np.asarray(x, np.float64) @ np.asarray(w, np.float64).T + np.asarray(b, np.float64)).astype(np.float32)
        return polynomial_forward(x, w, b)
    m, f = np.where(g == 'm')[0], np.where(g == 'f')[0]
    if len(m):
        out[m] = net(np.asarray(feat_m)[m], *male)
    if len(f):
        out[f] = net(np.asarray(feat_f)[f], *female)
    return out
