"""GSRLayer and GraphConvolution -- same classes / signatures / parameter names as the
reference's models/layers.py; arithmetic in libgsr_sm100 kernels (no zero-padded mat-muls)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as F_
from .. import kernels as K

device = torch.device("cuda")


class GSRLayer(nn.Module):
    """models/layers.py:11-63.  forward(A, X) -> (adj, |X_out|); keeps ``self.f_d``."""

    def __init__(self, hr_dim, lr_dim):
        super().__init__()
        self.hr_dim = hr_dim
        self.lr_dim = lr_dim
        self.weights = torch.nn.Parameter(torch.randn(hr_dim, lr_dim))     # :18
        self.check_eigh = True      # torch.linalg.eigh's error contract costs one device sync
        self.eigh_info = None       # per-graph sweep counts / error codes of the last forward
        self.eig_val_lr = None
        self.U_lr = None
        self._F = None

    @property
    def f_d(self):
        """models/layers.py:55-56: |W U^T f| with unit diagonal (materialised on demand)."""
        return None if self._F is None else K.abs_diag1(self._F)

    def forward(self, A, X, U=None, check_eigh=None):
        """``U`` (optional, extension) injects a precomputed eigenbasis of A, e.g. one cached
        across epochs: the LR eigendecomposition depends on the data only.
        ``check_eigh`` overrides ``self.check_eigh`` for this call."""
        A = A if A.is_cuda else A.to(device)
        f = X if X.is_cuda else X.to(device)
        if A.shape[-1] != self.lr_dim or f.shape[0] != self.lr_dim:
            raise RuntimeError(f"GSRLayer: expected A [..., {self.lr_dim}, {self.lr_dim}] and X [{self.lr_dim}, d], "
                               f"got {tuple(A.shape)} and {tuple(f.shape)}")
        if f.shape[1] != self.hr_dim:
            raise RuntimeError(f"GSRLayer: X must have {self.hr_dim} feature columns (normalize_adj needs a square f_d)")
        if U is None:
            self.eig_val_lr, U, self.eigh_info = F_.eigh(A, "U", check_errors=self.check_eigh if check_eigh is None else bool(check_eigh))      # :27
        self.U_lr = U
        adj, Z, self._F = F_.GSRLayerFn.apply(self.weights, A.detach(), f, U)
        return adj, Z


class GraphConvolution(nn.Module):
    """models/layers.py:66-88: adj @ (input @ W); dropout and activation are constructor
    arguments the reference never applies (:84, :87 are commented out)."""

    def __init__(self, in_features, out_features, dropout=0., act=F.relu):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.dropout = dropout
        self.act = act
        self.weight = torch.nn.Parameter(torch.FloatTensor(in_features, out_features))
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.xavier_uniform_(self.weight)

    def forward(self, input, adj):
        input = input if input.is_cuda else input.to(device)
        adj = adj if adj.is_cuda else adj.to(device)
        return F_.GraphConvFn.apply(input, adj, self.weight)
