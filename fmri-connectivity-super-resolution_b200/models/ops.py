"""Graph U-Net of the GSR-Net path -- same classes, constructor signatures, attribute names and
state_dict keys as the reference's models/ops.py; the arithmetic runs in libgsr_sm100 kernels.

Kept quirks: ``GCN`` ignores its adjacency argument (models/ops.py:56-61);
the pooled adjacency ``A[idx][:, idx]`` (:46-47) is returned by ``GraphPool`` but never consumed
inside ``GraphUnet``, which therefore skips that gather; pool sizes use Python's round().
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_
from .. import kernels as K

device = torch.device("cuda")


def _to_device(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_cuda else t.to(device)


def _check_index(idx: torch.Tensor, n: int, what: str) -> None:
    """Caller-supplied indices (the public GraphUnpool / gather entry points): torch raises IndexError for an index
    outside [-n, n); the kernels take raw non-negative row numbers, so anything outside [0, n) is rejected here (one
    device read).  Indices produced by gsr_pool_topk inside GraphPool / GraphUnet are valid by construction."""
    if idx.dtype != torch.int64:
        raise IndexError(f"{what}: tensors used as indices must be int64, got {idx.dtype}")
    if idx.numel():
        lo, hi = int(idx.min()), int(idx.max())
        if lo < 0 or hi >= n:
            raise IndexError(f"{what}: index {lo if lo < 0 else hi} is out of bounds for dimension 0 with size {n}")


class GraphUnpool(nn.Module):
    """models/ops.py:7-24."""

    def __init__(self):
        super().__init__()

    def forward(self, A, X, idx):
        A = A.to_dense() if A.is_sparse else A
        A, X, idx = _to_device(A), _to_device(X), _to_device(idx)
        _check_index(idx, A.shape[-2], "GraphUnpool")
        new_X = F_.GraphUnpoolFn.apply(X, idx, A.shape[-2])
        return A, new_X


class GraphPool(nn.Module):
    """models/ops.py:26-48."""

    def __init__(self, k, in_dim):
        super().__init__()
        self.k = k
        self.proj = nn.Linear(in_dim, 1)
        self.sigmoid = nn.Sigmoid()

    def pool_size(self, num_nodes: int) -> int:
        return max(1, int(round(self.k * num_nodes)))          # :42 (banker's rounding, as Python does)

    def select(self, X):
        """Scores, top-k and the scaled row gather -- everything but the adjacency gather."""
        new_X, idx, scores = F_.GraphPoolFn.apply(X, self.proj.weight, self.proj.bias, self.pool_size(X.shape[0]))
        return new_X, idx, scores

    def forward(self, A, X):
        A = A.to_dense() if A.is_sparse else A
        A, X = _to_device(A), _to_device(X)
        if A.shape[-2] != X.shape[0]:
            raise RuntimeError(f"GraphPool: A has {A.shape[-2]} nodes but X has {X.shape[0]} rows")
        new_X, idx, _ = self.select(X)
        A = K.gather_rows_cols(A, idx)                          # :46-47
        return A, new_X, idx


class GCN(nn.Module):
    """models/ops.py:50-61 -- Dropout(p=0) + Linear; the adjacency is ignored."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.proj = nn.Linear(in_dim, out_dim)
        self.drop = nn.Dropout(p=0)

    def forward(self, A, X):
        X = _to_device(X)
        return F_.LinearFn.apply(X, self.proj.weight, self.proj.bias)


class GraphUnet(nn.Module):
    """models/ops.py:63-115."""

    def __init__(self, ks, in_dim, out_dim, dim=268):
        super().__init__()
        self.ks = ks
        self.start_gcn = GCN(in_dim, dim)
        self.bottom_gcn = GCN(dim, dim)
        self.end_gcn = GCN(2 * dim, out_dim)
        self.down_gcns = nn.ModuleList()
        self.up_gcns = nn.ModuleList()
        self.pools = nn.ModuleList()
        self.unpools = nn.ModuleList()
        self.l_n = len(ks)
        for i in range(self.l_n):
            self.down_gcns.append(GCN(dim, dim))
            self.up_gcns.append(GCN(dim, dim))
            self.pools.append(GraphPool(ks[i], dim))
            self.unpools.append(GraphUnpool())
        self.pool_indices = []          # idx of the last forward (int64, device), one per level
        self.pool_scores = []

    def forward(self, A, X):
        X = _to_device(X)
        sizes, indices_list, down_outs = [], [], []
        self.pool_indices, self.pool_scores = [], []
        X = self.start_gcn(A, X)
        start_gcn_outs = X
        org_X = X
        for i in range(self.l_n):
            X = self.down_gcns[i](A, X)
            sizes.append(X.shape[0])
            down_outs.append(X)
            X, idx, scores = self.pools[i].select(X)            # A[idx][:, idx] is dead here (:46-47)
            indices_list.append(idx)
            self.pool_indices.append(idx)
            self.pool_scores.append(scores)
        X = self.bottom_gcn(A, X)
        for i in range(self.l_n):
            up_idx = self.l_n - i - 1
            X = F_.GraphUnpoolFn.apply(X, indices_list[up_idx], sizes[up_idx])
            X = self.up_gcns[i](A, X)
            X = X.add(down_outs[up_idx])
        X = torch.cat([X, org_X], 1)
        X = self.end_gcn(A, X)
        return X, start_gcn_outs
