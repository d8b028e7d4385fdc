"""CPU oracle for the GSR-Net hot path -- TEST INFRASTRUCTURE, not product code.

A functional torch-CPU (fp32, autograd) restatement of what the reference
executes for one subject, operation by operation, so that its results AND its
CPU cost are those of the reference (the padded mat-muls and the diag-matmul
normalisation are kept on purpose).  Every function cites the reference lines
it follows (paths relative to the reference repository root).

Two documented wrappers make the result well defined where the reference leans
on unspecified library behaviour; the golden
generator applies the same wrappers to the real reference:
  * ``eigh_canonical``  -- ``torch.linalg.eigh(A, 'U')`` followed by the sign
    rule "largest-|.| component of every eigenvector (lowest row index on
    ties) is positive";
  * ``stable_topk``     -- rank by (score descending, index ascending).

Pinned against the unmodified reference by tests/test_oracle_golden.py through
the fixtures written by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

DEFAULT_KS = (0.9, 0.7, 0.6, 0.5)   # training_gsr.ipynb:70
UNET_DIM = 268                      # models/ops.py:64 (dim=268 default)


# --------------------------------------------------------------------------
# documented wrappers
# --------------------------------------------------------------------------
def canonicalize_signs(U: Tensor) -> Tensor:
    """Flip eigenvector columns so that the entry of largest magnitude (lowest
    row index on exact ties) is positive.  Works on [..., n, n]."""
    mag = U.abs()
    # argmax returns the first maximal index along the dim -> lowest row on ties
    row = mag.argmax(dim=-2, keepdim=True)
    pivot = torch.gather(U, -2, row)
    sign = torch.where(pivot < 0, -torch.ones_like(pivot), torch.ones_like(pivot))
    return U * sign


def eigh_canonical(A: Tensor) -> Tuple[Tensor, Tensor]:
    """models/layers.py:27 and utils/gsr_train.py:34 -- ``eigh(A, UPLO='U')``
    (ascending eigenvalues, upper triangle read) plus the sign rule above."""
    w, U = torch.linalg.eigh(A, UPLO="U")
    return w, canonicalize_signs(U)


def stable_topk(scores: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """models/ops.py:42 ``torch.topk(scores, k)`` with the stated tie rule:
    score descending, index ascending inside a tie."""
    vals, idx = torch.sort(scores, descending=True, stable=True)
    return vals[:k], idx[:k]


def pool_size(k: float, num_nodes: int) -> int:
    """models/ops.py:42 -- ``max(1, int(round(k * n)))`` (Python banker's
    rounding: round(30.5) == 30)."""
    return max(1, int(round(k * num_nodes)))


# --------------------------------------------------------------------------
# utils/preprocessing.py
# --------------------------------------------------------------------------
def normalize_adj(mx: Tensor) -> Tensor:
    """utils/preprocessing.py:16-25.  r = rowsum^-1/2 (inf -> 0);
    out = (mx @ diag(r))^T @ diag(r), i.e. out[i,j] = (mx[j,i]*r_i)*r_j.
    No self loops are added and the result is the TRANSPOSE-scaled matrix."""
    rowsum = mx.sum(1)
    r = torch.pow(rowsum, -0.5).flatten()
    r[torch.isinf(r)] = 0.0
    d = torch.diag(r)
    t = torch.matmul(mx, d)
    t = torch.transpose(t, 0, 1)
    return torch.matmul(t, d)


def pad_hr_adj(label: Tensor, split: int) -> Tensor:
    """utils/preprocessing.py:10-14 -- zero pad by ``split`` on every side and
    set the diagonal to one (the reference does it through numpy on the host)."""
    out = torch.nn.functional.pad(label.detach().clone(), (split, split, split, split))
    out.fill_diagonal_(1.0)
    return out


def unpad(data: Tensor, split: int) -> Tensor:
    """utils/preprocessing.py:27-33 -- centre slice."""
    return data[split:data.shape[0] - split, split:data.shape[1] - split]


# --------------------------------------------------------------------------
# models/ops.py
# --------------------------------------------------------------------------
def _linear(sd: StateDict, prefix: str, X: Tensor) -> Tensor:
    """models/ops.py:56-61 -- ``GCN`` is Dropout(p=0) + nn.Linear; the
    adjacency argument is ignored."""
    return torch.nn.functional.linear(X, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])


def graph_pool(sd: StateDict, prefix: str, k: float, A: Tensor, X: Tensor):
    """models/ops.py:34-48.  scores = sigmoid(proj(X)/100); top-k rows of X
    scaled by their score; A gathered on rows then columns."""
    scores = torch.nn.functional.linear(X, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])
    scores = torch.squeeze(scores)
    scores = torch.sigmoid(scores / 100)
    kk = pool_size(k, A.shape[0])
    values, idx = stable_topk(scores, kk)
    new_X = X[idx, :] * values.unsqueeze(-1)
    A = A[idx, :][:, idx]
    return A, new_X, idx, scores


def graph_unpool(A: Tensor, X: Tensor, idx: Tensor) -> Tensor:
    """models/ops.py:13-24 -- zeros(n, d) with rows ``idx`` set to X."""
    new_X = torch.zeros([A.shape[0], X.shape[1]], dtype=X.dtype)
    new_X[idx] = X
    return new_X


def graph_unet(sd: StateDict, A: Tensor, X: Tensor, ks: Sequence[float] = DEFAULT_KS,
               prefix: str = "net", trace: dict | None = None):
    """models/ops.py:84-115 -- 4-level U-Net: start, (down, pool)x4, bottom,
    (unpool, up, +skip)x4, concat with the start output, end."""
    l_n = len(ks)
    adj_ms, indices, down_outs = [], [], []
    X = _linear(sd, f"{prefix}.start_gcn", X)
    start_gcn_outs = X
    org_X = X
    for i in range(l_n):
        X = _linear(sd, f"{prefix}.down_gcns.{i}", X)
        adj_ms.append(A)
        down_outs.append(X)
        A, X, idx, scores = graph_pool(sd, f"{prefix}.pools.{i}", ks[i], A, X)
        indices.append(idx)
        if trace is not None:
            trace.setdefault("pool_scores", []).append(scores.detach().clone())
            trace.setdefault("pool_idx", []).append(idx.detach().clone())
    X = _linear(sd, f"{prefix}.bottom_gcn", X)
    for i in range(l_n):
        up = l_n - i - 1
        A, idx = adj_ms[up], indices[up]
        X = graph_unpool(A, X, idx)
        X = _linear(sd, f"{prefix}.up_gcns.{i}", X)
        X = X.add(down_outs[up])
    X = torch.cat([X, org_X], 1)
    X = _linear(sd, f"{prefix}.end_gcn", X)
    return X, start_gcn_outs


# --------------------------------------------------------------------------
# models/layers.py
# --------------------------------------------------------------------------
def gsr_layer(W: Tensor, A: Tensor, f: Tensor, U: Tensor | None = None, trace: dict | None = None):
    """models/layers.py:20-63.  Eigen-decompose the LR adjacency, project with
    the (268,160) weight through the padded identity / padded eigenvector
    matrices, abs, unit diagonal, normalise, adj @ adj^T, symmetrise, unit
    diagonal, abs.  Returns (adj, |X|).  ``U`` may be injected so that non-eigh
    kernels can be tested with a shared eigenbasis."""
    hr_dim, lr_dim = W.shape
    if U is None:
        _, U = eigh_canonical(A)
    eye = torch.eye(lr_dim)
    s_d = torch.cat((eye, eye[: hr_dim - lr_dim]), 0)                     # :30-34
    a = torch.matmul(W, s_d.T)                                            # :37
    pad = hr_dim - lr_dim
    U_pad = torch.nn.functional.pad(U, (0, pad, 0, pad))                  # :40-42
    b = torch.matmul(a, torch.t(U_pad))                                   # :45
    f_pad = torch.nn.functional.pad(f, (0, 0, 0, hr_dim - f.shape[0])) if f.shape[0] < hr_dim else f
    f_d = torch.matmul(b, f_pad)                                          # :54
    f_d = torch.abs(f_d)
    f_d = f_d.fill_diagonal_(1)                                           # :56
    adj = normalize_adj(f_d)                                              # :58
    X = torch.mm(adj, adj.t())                                            # :59
    X = (X + X.t()) / 2
    idx = torch.eye(hr_dim, dtype=torch.bool)
    X[idx] = 1                                                            # :62
    if trace is not None:
        trace["U"] = U.detach().clone()
        trace["f_d"] = f_d.detach().clone()
    return adj, torch.abs(X)


def graph_convolution(weight: Tensor, x: Tensor, adj: Tensor) -> Tensor:
    """models/layers.py:83-88 -- adj @ (x @ W); no bias / activation / dropout."""
    return torch.mm(adj, torch.mm(x, weight))


# --------------------------------------------------------------------------
# models/gsr_model.py
# --------------------------------------------------------------------------
def gsrnet_forward(sd: StateDict, lr: Tensor, ks: Sequence[float] = DEFAULT_KS,
                   U: Tensor | None = None, trace: dict | None = None):
    """models/gsr_model.py:23-37 -- returns (|z|, net_outs, start_gcn_outs, adj)."""
    lr_dim = sd["layer.weights"].shape[1]
    hr_dim = sd["layer.weights"].shape[0]
    I = torch.eye(lr_dim)
    A = normalize_adj(lr)
    net_outs, start_gcn_outs = graph_unet(sd, A, I, ks, trace=trace)
    outputs, Z = gsr_layer(sd["layer.weights"], A, net_outs, U=U, trace=trace)
    hidden1 = graph_convolution(sd["gc1.weight"], Z, outputs)
    hidden2 = graph_convolution(sd["gc2.weight"], hidden1, outputs)
    z = (hidden2 + hidden2.t()) / 2
    z[torch.eye(hr_dim, dtype=torch.bool)] = 1
    if trace is not None:
        trace["A"] = A.detach().clone()
        trace["Z"] = Z.detach().clone()
        trace["hidden1"] = hidden1.detach().clone()
        trace["hidden2"] = hidden2.detach().clone()
    return torch.abs(z), net_outs, start_gcn_outs, outputs


# --------------------------------------------------------------------------
# utils/gsr_train.py
# --------------------------------------------------------------------------
def l1(a: Tensor, b: Tensor) -> Tensor:
    """utils/gsr_train.py:9 -- nn.L1Loss() == mean |a - b|."""
    return (a - b).abs().mean()


def hr_target_basis(hr: Tensor, lr_dim: int, padding: int = 0) -> Tensor:
    """utils/gsr_train.py:33-36 -- eigenvectors of the 160 SMALLEST eigenvalues
    of the (padded, unit-diagonal) HR target."""
    _, U_hr = eigh_canonical(pad_hr_adj(hr, padding))
    return U_hr[:, :lr_dim]


def train_loss(sd: StateDict, lr: Tensor, hr: Tensor, lmbda: float = 16.0, padding: int = 0,
               ks: Sequence[float] = DEFAULT_KS, U: Tensor | None = None,
               U_hr_reduced: Tensor | None = None, trace: dict | None = None):
    """utils/gsr_train.py:30-43 for one subject: returns (loss, error, outputs)."""
    out, net_outs, start_gcn_outs, adj = gsrnet_forward(sd, lr, ks, U=U, trace=trace)
    model_outputs = unpad(out, padding)
    if U_hr_reduced is None:
        U_hr_reduced = hr_target_basis(hr, sd["layer.weights"].shape[1], padding)
    loss = lmbda * l1(net_outs, start_gcn_outs) + l1(sd["layer.weights"], U_hr_reduced) + l1(model_outputs, hr)
    error = l1(model_outputs, hr)
    return loss, error, (out, net_outs, start_gcn_outs, adj)


def step_grads(sd: StateDict, lrs: Sequence[Tensor], hrs: Sequence[Tensor], lmbda: float = 16.0,
               padding: int = 0, ks: Sequence[float] = DEFAULT_KS,
               Us: Sequence[Tensor] | None = None):
    """Frozen-weight fwd+bwd over several subjects: returns the
    per-subject losses / errors and the MEAN over subjects of the parameter
    gradients (what one batched step of the CUDA path produces)."""
    params = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in sd.items())
    losses, errors = [], []
    for i, (lr, hr) in enumerate(zip(lrs, hrs)):
        loss, err, _ = train_loss(params, lr, hr, lmbda, padding, ks, U=None if Us is None else Us[i])
        (loss / len(lrs)).backward()
        losses.append(float(loss))
        errors.append(float(err))
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v)).clone())
                        for k, v in params.items())
    return losses, errors, grads


def test_errors(sd: StateDict, lrs, hrs, padding: int = 0, ks: Sequence[float] = DEFAULT_KS) -> List[float]:
    """utils/gsr_train.py:64-81 -- per-subject MAE of the prediction against the
    HR target with unit diagonal; all-zero samples are skipped."""
    errs = []
    for lr, hr in zip(lrs, hrs):
        if not bool(lr.any()) or not bool(hr.any()):
            continue
        hr1 = hr.clone()
        hr1.fill_diagonal_(1.0)
        out, _, _, _ = gsrnet_forward(sd, lr, ks)
        errs.append(float(l1(unpad(out, padding), hr1)))
    return errs


# --------------------------------------------------------------------------
# deterministic inputs / weights shared by tests, fixtures and bench
# --------------------------------------------------------------------------
def synth_connectome(n: int, seed: int) -> Tensor:
    """x ~ U[0,1)^{n x n}; M = (x + x^T)/2; zero diagonal."""
    g = torch.Generator().manual_seed(int(seed))
    x = torch.rand(n, n, generator=g, dtype=torch.float32)
    m = (x + x.t()) / 2
    m.fill_diagonal_(0.0)
    return m


def state_dict_keys(n_levels: int = 4) -> List[str]:
    """The 31 state_dict keys of the reference GSRNet, in its registration
    order (models/gsr_model.py:18-21, models/ops.py:68-82)."""
    keys = ["layer.weights"]
    for name in ("start_gcn", "bottom_gcn", "end_gcn"):
        keys += [f"net.{name}.proj.weight", f"net.{name}.proj.bias"]
    for grp in ("down_gcns", "up_gcns", "pools"):
        for i in range(n_levels):
            keys += [f"net.{grp}.{i}.proj.weight", f"net.{grp}.{i}.proj.bias"]
    keys += ["gc1.weight", "gc2.weight"]
    return keys


def make_state_dict(seed: int, lr_dim: int = 160, hr_dim: int = 268, hidden_dim: int = 268,
                    dim: int = UNET_DIM, n_levels: int = 4) -> StateDict:
    """Seeded weights with the reference's shapes and init FAMILIES (randn for
    GSRLayer, models/layers.py:18; Xavier-uniform for GraphConvolution, :81;
    nn.Linear default elsewhere) drawn from one explicit generator, so that the
    same numbers can be rebuilt anywhere without the reference."""
    g = torch.Generator().manual_seed(int(seed))

    def uni(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    shapes = {
        "start_gcn": (dim, lr_dim), "bottom_gcn": (dim, dim), "end_gcn": (hr_dim, 2 * dim),
    }
    sd: StateDict = OrderedDict()
    sd["layer.weights"] = torch.randn(hr_dim, lr_dim, generator=g, dtype=torch.float32)
    for name, (o, i) in shapes.items():
        b = 1.0 / math.sqrt(i)
        sd[f"net.{name}.proj.weight"] = uni((o, i), b)
        sd[f"net.{name}.proj.bias"] = uni((o,), b)
    for grp in ("down_gcns", "up_gcns"):
        for l in range(n_levels):
            b = 1.0 / math.sqrt(dim)
            sd[f"net.{grp}.{l}.proj.weight"] = uni((dim, dim), b)
            sd[f"net.{grp}.{l}.proj.bias"] = uni((dim,), b)
    for l in range(n_levels):
        b = 1.0 / math.sqrt(dim)
        sd[f"net.pools.{l}.proj.weight"] = uni((1, dim), b)
        sd[f"net.pools.{l}.proj.bias"] = uni((1,), b)
    a1 = math.sqrt(6.0 / (hr_dim + hidden_dim))
    sd["gc1.weight"] = uni((hr_dim, hidden_dim), a1)
    sd["gc2.weight"] = uni((hidden_dim, hr_dim), a1)
    assert list(sd.keys()) == state_dict_keys(n_levels)
    return sd
