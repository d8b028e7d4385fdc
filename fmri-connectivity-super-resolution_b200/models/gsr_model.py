"""GSRNet -- drop-in for the reference's models/gsr_model.py:10-37.

Same constructor (``GSRNet(ks, args)`` with args.lr_dim / hr_dim / hidden_dim), same
sub-module and parameter names (31 state_dict keys), same 4-tuple return and cached
attributes.  Extension: ``forward`` also accepts a subject batch [B, lr, lr] and then returns
[B, hr, hr] predictions; the Graph U-Net runs ONCE per call because its output does not depend
on the subject (it is fed the identity and its GCN ignores the adjacency).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as F_
from .layers import GSRLayer, GraphConvolution
from .ops import GraphUnet

device = torch.device("cuda")


class GSRNet(nn.Module):
    def __init__(self, ks, args):
        super().__init__()
        self.lr_dim = args.lr_dim
        self.hr_dim = args.hr_dim
        self.hidden_dim = args.hidden_dim
        self.layer = GSRLayer(self.hr_dim, self.lr_dim)
        self.net = GraphUnet(ks, self.lr_dim, self.hr_dim)
        self.gc1 = GraphConvolution(self.hr_dim, self.hidden_dim, 0, act=F.relu)
        self.gc2 = GraphConvolution(self.hidden_dim, self.hr_dim, 0, act=F.relu)
        self.overlap_unet = True        # batched forward: U-Net on a side stream under the eigendecomposition
        self._side_stream = None

    def forward(self, lr, U=None, check_eigh=None):
        """``U`` (extension): a precomputed eigenbasis of the normalised LR adjacency.  ``check_eigh`` (extension):
        override ``self.layer.check_eigh`` for this call (False = no host synchronisation; the device-side Jacobi
        fallback for QL failures runs either way)."""
        hidden2, net_outs, start_gcn_outs, outputs = self.forward_hidden(lr, U=U, check_eigh=check_eigh)
        z = F_.sym_diag_abs(hidden2)                                                  # :32-37
        return z, net_outs, start_gcn_outs, outputs

    def forward_hidden(self, lr, U=None, check_eigh=None):
        """``forward`` up to and including :30 -- ``hidden2`` instead of the symmetrised prediction.  The training step
        feeds it to ``functional.sym_diag_abs_l1_per_subject``, which forms :32-37 and the L1 term against the HR target
        in one pass (the prediction is never written)."""
        lr = lr if lr.is_cuda else lr.to(device)
        check = self.layer.check_eigh if check_eigh is None else bool(check_eigh)
        if self.layer.weights.device.type != "cuda":
            raise RuntimeError("GSRNet: move the model to the GPU first (model.to('cuda')); there is no CPU path")
        I = torch.eye(self.lr_dim, device=lr.device)                                  # :24
        A = F_.normalize_adj(lr)                                                      # :25
        if U is None and self.overlap_unet and lr.dim() == 3:
            # The Graph U-Net depends on the weights only and the LR eigendecomposition on the data only:
            # the ~25 small U-Net launches run on a side stream underneath the eigh kernels (autograd
            # replays the same stream assignment in the backward pass).
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.net_outs, self.start_gcn_outs = self.net(A, I)                   # :27
            self.layer.eig_val_lr, U, self.layer.eigh_info = F_.eigh(A, "U", check_errors=check)   # layers.py:27
            main.wait_stream(side)
            self.net_outs.record_stream(main)
            self.start_gcn_outs.record_stream(main)
        else:
            self.net_outs, self.start_gcn_outs = self.net(A, I)                       # :27
        self.outputs, self.Z = self.layer(A, self.net_outs, U=U, check_eigh=check)    # :28
        # :29-30  hidden1 = gc1(Z, outputs); hidden2 = gc2(hidden1, outputs) -- one autograd node for the two layers
        self.hidden1, self.hidden2 = F_.GraphConvPairFn.apply(self.Z, self.outputs, self.gc1.weight, self.gc2.weight)
        return self.hidden2, self.net_outs, self.start_gcn_outs, self.outputs
