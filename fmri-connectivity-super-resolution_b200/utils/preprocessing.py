"""normalize_adj_torch / pad_HR_adj / unpad -- drop-ins for utils/preprocessing.py:10-33."""
from __future__ import annotations

import torch

from .. import functional as F_

device = torch.device("cuda")


def pad_HR_adj(label, split):
    """utils/preprocessing.py:10-14: zero-pad by ``split`` on every side, unit diagonal.  The
    reference bounces through CPU numpy every step; this stays on the device."""
    label = label if label.is_cuda else label.to(device)
    out = torch.nn.functional.pad(label.detach().float(), (split, split, split, split))
    out.diagonal(dim1=-2, dim2=-1).fill_(1.0)
    return out


def normalize_adj_torch(mx):
    """utils/preprocessing.py:16-25 (differentiable; [n, n] or [B, n, n])."""
    mx = mx if mx.is_cuda else mx.to(device)
    return F_.normalize_adj(mx)


def unpad(data, split):
    """utils/preprocessing.py:27-33: centre slice of the last two dimensions."""
    idx_0 = data.shape[-2] - split
    idx_1 = data.shape[-1] - split
    return data[..., split:idx_0, split:idx_1]
