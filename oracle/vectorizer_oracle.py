"""CPU oracle for MatrixVectorizer and the evaluation reductions -- TEST
INFRASTRUCTURE, not product code.  numpy restatement with closed-form index
maps; pinned against the unmodified reference by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np


def vectorize_index_map(n: int, include_diagonal: bool = False):
    """utils/matrix_vectorizer.py:47-56.  Walk columns left to right; inside a
    column take rows above the diagonal top to bottom, then (``include_diagonal``)
    the SUB-diagonal element (col+1, col) when it exists.  Returns (rows, cols)."""
    rows, cols = [], []
    for c in range(n):
        r = np.arange(0, c)
        rows.append(r)
        cols.append(np.full(c, c))
        if include_diagonal and c + 1 < n:
            rows.append(np.array([c + 1]))
            cols.append(np.array([c]))
    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    return rows.astype(np.int64), cols.astype(np.int64)


def vectorize(matrix: np.ndarray, include_diagonal: bool = False) -> np.ndarray:
    """utils/matrix_vectorizer.py:24-58 -- numpy in, numpy out, dtype kept."""
    n = matrix.shape[0]
    r, c = vectorize_index_map(n, include_diagonal)
    if r.size == 0:
        return np.array([])        # the reference builds np.array(list): empty -> float64
    return np.asarray(matrix)[r, c]


def anti_vectorize(vector: np.ndarray, n: int, include_diagonal: bool = False) -> np.ndarray:
    """utils/matrix_vectorizer.py:61-95 -- ROW-major strict upper fill
    (torch.triu_indices order), mirrored; ``include_diagonal`` puts
    vector[n_el : n_el + n] on the MAIN diagonal.  Output is float32; raises
    AssertionError when the vector is shorter than n(n-1)/2."""
    r, c = np.triu_indices(n, k=1)          # row-major, same order as torch.triu_indices
    n_el = r.shape[0]
    assert vector.shape[0] >= n_el, "Vector size doesn't match required elements"
    out = np.zeros((n, n), dtype=np.float32)
    out[r, c] = vector[:n_el]
    out[c, r] = vector[:n_el]
    if include_diagonal:
        d = np.arange(n)
        out[d, d] = vector[n_el:n_el + n]
    return out


def eval_reductions(pred: np.ndarray, gt: np.ndarray):
    """utils/evaluation.py:96-112 -- MAE, Pearson r and Jensen-Shannon DISTANCE
    (natural log, inputs normalised to unit sum) of the concatenated vectorised
    upper triangles; fp64 arithmetic as the accuracy yardstick.
    ``pred`` / ``gt``: [N, n, n]."""
    p = np.concatenate([vectorize(m) for m in pred]).astype(np.float64)
    g = np.concatenate([vectorize(m) for m in gt]).astype(np.float64)
    return eval_reductions_flat(p, g)


def eval_reductions_flat(p: np.ndarray, g: np.ndarray):
    p = p.astype(np.float64)
    g = g.astype(np.float64)
    mae = np.abs(p - g).mean()
    pm, gm = p - p.mean(), g - g.mean()
    pcc = float((pm * gm).sum() / np.sqrt((pm * pm).sum() * (gm * gm).sum()))
    # scipy.spatial.distance.jensenshannon(gt, pred): normalise, m = (a+b)/2,
    # sqrt((KL(a||m) + KL(b||m)) / 2) with rel_entr(0, y) = 0.
    a, b = g / g.sum(), p / p.sum()
    m = (a + b) / 2

    def rel_entr(x, y):
        out = np.zeros_like(x)
        pos = (x > 0) & (y > 0)
        out[pos] = x[pos] * np.log(x[pos] / y[pos])
        out[(x < 0) | ((x > 0) & (y <= 0))] = np.inf
        return out

    js = np.sqrt((rel_entr(a, m).sum() + rel_entr(b, m).sum()) / 2.0)
    return float(mae), pcc, float(js)
