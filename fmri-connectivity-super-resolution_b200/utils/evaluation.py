"""The MAE / PCC / Jensen-Shannon reductions of utils/evaluation.py:96-112 on the GPU.

The reference vectorises every prediction / ground-truth pair with MatrixVectorizer.vectorize,
concatenates and calls sklearn / scipy on the host.  All three numbers are invariant under the
order of that concatenation, so the kernels read the strict upper triangles in place (half of
each matrix is touched, nothing is materialised) with fp64 accumulators.  The topological
metrics of the same function (networkx centralities) are out of scope.

Precision contract: the kernels read **float32** matrices (the model's output dtype and what the
challenge CSVs hold to their printed digits) and accumulate in float64.  float64 inputs are
rounded to float32 first -- a ~6e-8 relative perturbation of every element -- so against
sklearn / scipy fed the float64 originals the three numbers agree to ~1e-7 relative; the
1e-10 / 1e-8 tolerances of the tests are against float64 sums of the same float32 inputs.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import kernels as K


def _as_device_batch(x):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    elif isinstance(x, (list, tuple)):
        x = torch.stack([torch.as_tensor(np.asarray(m), dtype=torch.float32) if not isinstance(m, torch.Tensor) else m.float()
                         for m in x])
    x = x.float()
    if x.dim() == 2:
        x = x[None]
    return x.cuda()


def evaluation_reductions(pred_matrices, gt_matrices, return_tensor: bool = False):
    """(mae, pcc, js_distance) over the concatenated vectorised upper triangles of
    pred_matrices / gt_matrices ([N, n, n] numpy, list of matrices, or torch)."""
    pred, gt = _as_device_batch(pred_matrices), _as_device_batch(gt_matrices)
    res = K.eval_reduce(pred, gt, upper=True)
    if return_tensor:
        return res
    mae, pcc, js = res.tolist()
    return mae, pcc, js


def evaluation_reductions_flat(pred_1d, gt_1d):
    """Same three numbers from already vectorised / concatenated 1-D arrays
    (``pred_1d`` / ``gt_1d`` of utils/evaluation.py:102-103)."""
    p = torch.as_tensor(np.asarray(pred_1d, dtype=np.float32) if not isinstance(pred_1d, torch.Tensor) else pred_1d).float().cuda()
    g = torch.as_tensor(np.asarray(gt_1d, dtype=np.float32) if not isinstance(gt_1d, torch.Tensor) else gt_1d).float().cuda()
    return tuple(K.eval_reduce(p.reshape(-1), g.reshape(-1), upper=False).tolist())


def print_metrics(gt_matrices, pred_matrices, fold_i=None):
    """The three-number subset of the reference's print_metrics (utils/evaluation.py:13,106-116)."""
    mae, pcc, js = evaluation_reductions(pred_matrices, gt_matrices)
    print("MAE: ", mae)
    print("PCC: ", pcc)
    print("Jensen-Shannon Distance: ", js)
    return {"mae": mae, "pcc": pcc, "js_dis": js}
