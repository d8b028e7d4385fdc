"""The on-disk edge of the GSR-Net path: vectorised CSV rows -> adjacency matrices
and predictions -> the submission vector, on the GPU.

    extract_matrices        utils/preprocessing.py:60-75 (``extract_data`` after ``pd.read_csv``):
                            NaN -> 0, every row anti-vectorised into a symmetric [n, n] matrix
    upper_triangular_flatten training_gsr.ipynb:112-117: ``matrix[np.triu_indices(268, k=1)]`` (ROW-major)
    submission_vector       training_gsr.ipynb:119-124 (row-major, the GSR-Net notebook) or
                            utils/data.py:214-231 (``MatrixVectorizer.vectorize``: COLUMN-major)
    write_submission        the two-column ``ID,Predicted`` CSV both call sites write

The reference's ``extract_data`` passes numpy rows to ``MatrixVectorizer.anti_vectorize`` and fails
under the pinned torch ("CSV loaders"); numpy and torch inputs are both accepted here.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels as K


def _device_rows(rows) -> torch.Tensor:
    if isinstance(rows, torch.Tensor):
        t = rows.to(device="cuda", dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).cuda()
    if t.dim() == 1:
        t = t[None]
    return torch.nan_to_num(t, nan=0.0, posinf=float("inf"), neginf=float("-inf")).contiguous()     # np.nan_to_num(nan=0.0): NaN only


def extract_matrices(rows, is_hr: bool = False, matrix_size: int | None = None, as_numpy: bool = True):
    """[N, n(n-1)/2] vectorised rows -> [N, n, n] symmetric matrices with zero diagonal
    (n = 268 if ``is_hr`` else 160, as the reference hard-codes, unless ``matrix_size`` is given).
    Returns float64 numpy like the reference's ``np.zeros`` target, or the CUDA fp32 tensor."""
    n = int(matrix_size) if matrix_size is not None else (268 if is_hr else 160)
    t = _device_rows(rows)
    if t.shape[-1] < n * (n - 1) // 2:
        raise AssertionError("Vector size doesn't match required elements")
    mats = K.anti_vectorize(t, n, False)
    return mats.cpu().numpy().astype(np.float64) if as_numpy else mats


def upper_triangular_flatten(matrix):
    """``matrix[np.triu_indices(n, k=1)]`` for one matrix or a batch; numpy in -> numpy out."""
    is_np = not isinstance(matrix, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(matrix, dtype=np.float32)).cuda() if is_np else matrix
    out = K.triu_flatten(t)
    return out.cpu().numpy() if is_np else out


def submission_vector(preds, order: str = "row"):
    """Concatenated per-subject upper triangles of the predictions [N, n, n]:
    ``order='row'`` as training_gsr.ipynb writes them, ``order='column'`` as utils/data.py does."""
    is_np = not isinstance(preds, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32)).cuda() if is_np else preds.contiguous()
    if t.dim() == 2:
        t = t[None]
    if order == "row":
        flat = K.triu_flatten(t)
    elif order == "column":
        flat = K.vectorize(t, False)
    else:
        raise ValueError("submission_vector: order must be 'row' or 'column'")
    flat = flat.reshape(-1)
    return flat.cpu().numpy() if is_np else flat


def write_submission(path: str, predicted) -> int:
    """``ID,Predicted`` with IDs 1..len, as ``pd.DataFrame(...).to_csv(path, index=False)`` writes it."""
    v = predicted.detach().cpu().numpy() if isinstance(predicted, torch.Tensor) else np.asarray(predicted)
    v = v.reshape(-1)
    with open(path, "w") as fh:
        fh.write("ID,Predicted\n")
        for i, x in enumerate(v.tolist(), 1):
            fh.write(f"{i},{x!r}\n")
    return int(v.size)
