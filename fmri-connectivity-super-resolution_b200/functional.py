"""torch.autograd.Functions over the sm_100a kernels -- the differentiable ops of the GSR-Net path.

Every op accepts the reference's 2-D tensors and, as an extension, a leading subject
dimension ([B, n, n]).  Shared operands (weights, the U-Net output) stay 2-D; their
gradients are reduced over the subjects on the device (gsr_bgemm_reduce_f32).
Backward formulas: SURVEY.md Appendix A (restated from the reference's autograd graph).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import kernels as K


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


def _wgrad(A: torch.Tensor, B: torch.Tensor, transA: bool = False, transB: bool = False) -> torch.Tensor:
    """sum over subjects of op(A[b]) @ op(B[b]) (or the plain product for 2-D operands)."""
    if A.dim() == 3:
        return K.bgemm_reduce(A, B, transA=transA, transB=transB)
    return K.bgemm(A, B, transA=transA, transB=transB)


# ----------------------------------------------------------------------------- a1
class NormalizeAdjFn(Function):
    """utils/preprocessing.py:16-25."""

    @staticmethod
    def forward(ctx, mx):
        out, r, _ = K.normalize_adj_fwd(mx, pre=False)
        ctx.save_for_backward(mx, r)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        mx, r = ctx.saved_tensors
        return K.normalize_adj_bwd(mx, r, _c(g), pre=False)


def normalize_adj(mx: torch.Tensor) -> torch.Tensor:
    return NormalizeAdjFn.apply(mx)


# ----------------------------------------------------------------------------- a2
class EighError(torch.linalg.LinAlgError):
    pass


def raise_on_eigh_info(info: torch.Tensor) -> None:
    """torch.linalg.eigh's error contract (synchronises: reads `info` back)."""
    worst = int(info.min().item())
    if worst == -1:
        raise EighError("linalg.eigh: the input contains non-finite values")
    if worst == -2:
        raise EighError("linalg.eigh: the Jacobi iteration failed to converge")


def eigh(A: torch.Tensor, UPLO: str = "U", check_errors: bool = True, canonical_sign: bool = True):
    """models/layers.py:27 / utils/gsr_train.py:34 -- eigh(A, UPLO='U'): ascending eigenvalues,
    eigenvectors in columns, sign-canonicalised.  Not differentiable (the reference never
    differentiates through it: its input carries no gradient)."""
    if UPLO != "U":
        raise NotImplementedError("eigh: only UPLO='U' is on the GSR-Net path")
    if A.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("eigh: gradients through the eigendecomposition are not on the GSR-Net path")
    A = A.detach()
    # QL failures (more than 30 iterations per eigenvalue / 2 n^2 rotations: info == -2) are redone by the Jacobi
    # kernel inside gsr_eigh_tridiag_f32 itself, on the device and in the same stream -- also when nothing here
    # synchronises (training steps, CUDA graphs).  `check_errors` only adds torch's raise-on-failure contract.
    w, U, info = K.eigh(A, canonical_sign=canonical_sign)
    if check_errors:
        raise_on_eigh_info(info)                        # torch.linalg.eigh's error contract costs one sync
    return w, U, info


# ----------------------------------------------------------------------------- Linear (models/ops.py:60)
class LinearFn(Function):
    @staticmethod
    def forward(ctx, X, W, b):
        ctx.save_for_backward(X, W)
        return K.bgemm(X, W, transB=True, bias=b)

    @staticmethod
    @once_differentiable
    def backward(ctx, gY):
        X, W = ctx.saved_tensors
        gY = _c(gY)
        gYp = gY
        if gY.dim() == 2 and gY.shape[1] % 4 != 0 and gY.shape[0] >= K.TC_MIN_ROWS and gY.numel() >= (1 << 20):
            # rows of a [B, 35 778] gradient (the notebook MLP's output layer) are not 16-byte aligned, which keeps both of its
            # products off the tensor-core kernels: a copy with the rows padded to a multiple of 4 floats (one pass over gY)
            buf = torch.empty((gY.shape[0], (gY.shape[1] + 3) // 4 * 4), device=gY.device, dtype=gY.dtype)
            gYp = buf[:, :gY.shape[1]]
            gYp.copy_(gY)
        gX = K.bgemm(gYp, W) if ctx.needs_input_grad[0] else None
        gW = K.bgemm(gYp, X, transA=True) if ctx.needs_input_grad[1] else None
        gb = K.colsum(gY) if ctx.needs_input_grad[2] else None
        return gX, gW, gb


class BatchNormActFn(Function):
    """models/mlp.py:14-16 (and :22-24): BatchNorm1d -> Dropout(p) -> ReLU in one pass.  ``mean`` / ``var`` are the
    batch statistics (training) or the running estimates (eval); the dropout mask is recomputed from ``seed``."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, var, eps, p_drop, seed, relu, batch_stats):
        ctx.save_for_backward(x, gamma, beta, mean, var)
        ctx.cfg = (eps, p_drop, seed, relu, batch_stats)
        return K.bn_act_fwd(x, mean, var, gamma, beta, eps, p_drop, seed, relu)

    @staticmethod
    @once_differentiable
    def backward(ctx, gy):
        x, gamma, beta, mean, var = ctx.saved_tensors
        eps, p_drop, seed, relu, batch_stats = ctx.cfg
        dx, dgamma, dbeta = K.bn_act_bwd(x, mean, var, gamma, beta, _c(gy), eps, p_drop, seed, relu, batch_stats)
        return dx, dgamma, dbeta, None, None, None, None, None, None, None


def batch_norm_act(x, gamma, beta, running_mean, running_var, training, momentum=0.1, eps=1e-5, p_drop=0.0,
                   relu=False, seed=None):
    """BatchNorm1d (+ Dropout + ReLU) on x [B, H].  Training: batch statistics, running estimates updated in place,
    dropout active with a seed drawn from torch's default generator unless given.  Eval: running estimates, no
    dropout."""
    if training:
        if x.shape[0] == 1:           # nn.BatchNorm1d's contract: batch statistics of one sample are undefined
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(x.shape)}")
        if momentum is None:
            raise NotImplementedError("batch_norm_act: momentum=None (cumulative moving average) is not on the MLP baseline's "
                                      "path (models/mlp.py uses nn.BatchNorm1d's default momentum 0.1)")
        mean, var = K.bn_stats(x.detach(), running_mean, running_var, momentum)
        if p_drop > 0 and seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    else:
        mean, var, p_drop = running_mean, running_var, 0.0
    relu = relu if isinstance(relu, float) else bool(relu)        # a float is the negative slope of a LeakyReLU
    return BatchNormActFn.apply(x, gamma, beta, mean, var, float(eps), float(p_drop), int(seed or 0), relu, bool(training))


class AddActFn(Function):
    """training_template.ipynb:176-180: ``F.leaky_relu(block(x) + residual, negative_slope)`` in one pass."""

    @staticmethod
    def forward(ctx, a, b, negative_slope):
        out = K.add_act_fwd(a, b, negative_slope)
        ctx.save_for_backward(out)
        ctx.slope = negative_slope
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        gi = K.act_bwd(out, _c(g), ctx.slope)
        return gi, gi, None


class SpectralScaleFn(Function):
    """The weight a ``spectral_norm(nn.Linear)`` uses (training_template.ipynb:125,134,143; torch.nn.utils.spectral_norm
    ``compute_weight``): W / sigma with sigma = u . (W v); u, v are the power-iteration buffers (constants for autograd).
    sigma itself is torch's own ``dot(u, mv(W, v))`` (a gemv on the weight: parameter plumbing, as in the hook); the two
    passes over the weight that follow and the three of the backward are the library's."""

    @staticmethod
    def forward(ctx, W, u, v):
        sigma = torch.dot(u, torch.mv(W, v))
        W_eff = K.scale_div(W, sigma)
        ctx.save_for_backward(W_eff, u, v, sigma)
        return W_eff

    @staticmethod
    @once_differentiable
    def backward(ctx, G):
        W_eff, u, v, sigma = ctx.saved_tensors
        return K.spectral_bwd(_c(G), W_eff, u, v, sigma), None, None


class SigmoidFn(Function):
    """training_template.ipynb:132: nn.Sigmoid of the output layer."""

    @staticmethod
    def forward(ctx, x):
        y = K.sigmoid_fwd(x)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return K.sigmoid_bwd(y, _c(g))


class TriuFlattenFn(Function):
    """``x[:, mask]`` with the strict-upper-triangle mask (training_template.ipynb:121,166): row-major triangle of every
    matrix; backward scatters into the upper triangle only (no mirror: the mask selects those entries alone)."""

    @staticmethod
    def forward(ctx, dense):
        ctx.n = dense.shape[-1]
        return K.triu_flatten(dense)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        full = K.anti_vectorize(_c(g), ctx.n)            # symmetric fill ...
        return torch.triu(full, diagonal=1)              # ... of which the mask's gradient is the upper half


class SymFromTriuFn(Function):
    """``matrix[:, mask] = x; matrix + matrix^T`` (training_template.ipynb:186-190): symmetric zero-diagonal matrices from
    row-major upper triangles -- gsr_anti_vectorize; backward = upper + lower triangle of the gradient, flattened."""

    @staticmethod
    def forward(ctx, vec, n):
        ctx.n = n
        return K.anti_vectorize(vec, n)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        g = _c(g)
        return K.triu_flatten(g + g.transpose(-1, -2)), None


# ----------------------------------------------------------------------------- a8 / a9
class GraphPoolFn(Function):
    """models/ops.py:35-45: scores, top-k (stated tie rule), gather * score."""

    @staticmethod
    def forward(ctx, X, pw, pb, k):
        scores = K.pool_scores(X, pw, pb)
        vals, idx = K.pool_topk(scores, k)
        new_X = K.pool_gather_scale(X, idx, vals)
        ctx.save_for_backward(X, pw, scores, idx)
        ctx.mark_non_differentiable(idx, scores)
        return new_X, idx, scores

    @staticmethod
    @once_differentiable
    def backward(ctx, g_new, _gi, _gs):
        X, pw, scores, idx = ctx.saved_tensors
        gX, gp, gc = K.pool_bwd(X, pw, scores, idx, _c(g_new))
        return gX, gp.view_as(pw), gc, None


class GraphUnpoolFn(Function):
    """models/ops.py:22-23."""

    @staticmethod
    def forward(ctx, X, idx, n):
        ctx.save_for_backward(idx)
        return K.unpool_fwd(X, idx, n)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        return K.unpool_bwd(_c(g), idx), None, None


# ----------------------------------------------------------------------------- a3 / a4
class GSRLayerFn(Function):
    """models/layers.py:20-63 without the dead zero padding:
        (lambda, U) = eigh(A);  P = W U^T;  F = P f;  G = |F|, G_ii = 1
        adj = normalize(G);  X = adj adj^T;  X = (X + X^T)/2, X_ii = 1;  Z = |X|
    Returns (adj, Z, F).  F is the raw product (f_d = |F| with unit diagonal)."""

    @staticmethod
    def forward(ctx, W, A, f, U):
        Pt = K.bgemm(U, W, transB=True)                 # P^T = U W^T         [B, lr, hr]  (stacked over subjects)
        F = K.bgemm(Pt, f, transA=True)                 # F   = P f           [B, hr, hr]
        adj, r, _ = K.normalize_adj_fwd(F, pre=True)    # abs + unit diagonal fused in front
        X0 = K.bgemm(adj, adj, transB=True)             # adj adj^T
        Z = K.sym_diag_abs_fwd(X0)
        ctx.save_for_backward(W, f, U, Pt, F, r, adj, X0)
        ctx.mark_non_differentiable(F)
        ctx.set_materialize_grads(False)                # no 294 MB of zeros for the gradient slot of F
        return adj, Z, F

    @staticmethod
    @once_differentiable
    def backward(ctx, g_adj, g_Z, _gF):
        W, f, U, Pt, F, r, adj, X0 = ctx.saved_tensors
        if g_adj is None and g_Z is None:
            return None, None, None, None
        g_Z = _c(g_Z) if g_Z is not None else torch.zeros_like(X0)
        # dX0 = (dX + dX^T)/2 with dX = dZ * sgn(X), zero diagonal;  d_adj += (dX0 + dX0^T) adj = 2 dX0 adj.  The gradient
        # that arrives for adj itself (the graph convolutions') is NOT added in the product's epilogue (an addend read per
        # accumulator row made that product three times as slow): the normalise backward sums the two as it reads them.
        gX0 = K.sym_diag_abs_bwd(X0, g_Z)
        d_adj = K.bgemm(gX0, adj, alpha=2.0)
        gF = K.normalize_adj_bwd(F, r, d_adj, pre=True,          # through normalise, abs and the diagonal fill
                                 grad_out2=_c(g_adj) if g_adj is not None else None)
        gP = K.bgemm(gF, f, transB=True)                         # dP = dF f^T     [B, hr, lr] (stacked)
        gf = _wgrad(Pt, gF) if ctx.needs_input_grad[2] else None           # sum_b P^T dF
        gW = _wgrad(gP, U) if ctx.needs_input_grad[0] else None            # sum_b dP U
        return gW, None, gf, None


# ----------------------------------------------------------------------------- a5
class GraphConvFn(Function):
    """models/layers.py:83-88: adj @ (x @ W); no bias, activation or dropout."""

    @staticmethod
    def forward(ctx, x, adj, W):
        S = K.bgemm(x, W)                                # support  (stacked over subjects)
        H = K.bgemm(adj, S)
        ctx.save_for_backward(x, adj, W, S)
        return H

    @staticmethod
    @once_differentiable
    def backward(ctx, gH):
        x, adj, W, S = ctx.saved_tensors
        gH = _c(gH)
        gS = K.bgemm(adj, gH, transA=True)               # adj^T dH
        g_adj = K.bgemm(gH, S, transB=True) if ctx.needs_input_grad[1] else None   # dH S^T
        gW = _wgrad(x, gS, transA=True) if ctx.needs_input_grad[2] else None       # sum_b x^T dS
        gx = K.bgemm(gS, W, transB=True) if ctx.needs_input_grad[0] else None      # dS W^T
        return gx, g_adj, gW


class GraphConvPairFn(Function):
    """models/gsr_model.py:29-30: hidden1 = gc1(Z, adj); hidden2 = gc2(hidden1, adj) as one node.  Same products as
    two GraphConvFn nodes; the two contributions to the gradient of ``adj`` are accumulated in the tensor-memory
    accumulator of one launch (``bgemm_dual``) instead of a separate 3 x 294 MB addition by autograd."""

    @staticmethod
    def forward(ctx, Z, adj, W1, W2):
        S1 = K.bgemm(Z, W1)
        H1 = K.bgemm(adj, S1)
        S2 = K.bgemm(H1, W2)
        H2 = K.bgemm(adj, S2)
        ctx.save_for_backward(Z, adj, W1, W2, S1, H1, S2)
        ctx.set_materialize_grads(False)
        return H1, H2

    @staticmethod
    @once_differentiable
    def backward(ctx, gH1, gH2):
        Z, adj, W1, W2, S1, H1, S2 = ctx.saved_tensors
        g_adj = gW2 = None
        if gH2 is not None:
            gH2 = _c(gH2)
            gS2 = K.bgemm(adj, gH2, transA=True)
            gW2 = _wgrad(H1, gS2, transA=True) if ctx.needs_input_grad[3] else None
            gH1 = K.bgemm(gS2, W2, transB=True, add=_c(gH1) if gH1 is not None else None)
        if gH1 is None:                                  # (cannot happen after the branch above: kept for symmetry)
            return None, None, None, gW2
        gH1 = _c(gH1)
        gS1 = K.bgemm(adj, gH1, transA=True)
        if ctx.needs_input_grad[1]:
            # d adj = dH2 S2^T + dH1 S1^T: ONE launch, both products accumulated in tensor memory
            g_adj = K.bgemm_dual(gH2, S2, gH1, S1, transB=True) if gH2 is not None else K.bgemm(gH1, S1, transB=True)
        gW1 = _wgrad(Z, gS1, transA=True) if ctx.needs_input_grad[2] else None
        gZ = K.bgemm(gS1, W1, transB=True) if ctx.needs_input_grad[0] else None
        return gZ, g_adj, gW1, gW2


# ----------------------------------------------------------------------------- a6
class SymDiagAbsFn(Function):
    """models/gsr_model.py:32-37: z = (z + z^T)/2; z[diag] = 1; |z|."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return K.sym_diag_abs_fwd(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return K.sym_diag_abs_bwd(x, _c(g))


def sym_diag_abs(x: torch.Tensor) -> torch.Tensor:
    return SymDiagAbsFn.apply(x)


class SymDiagAbsL1Fn(Function):
    """models/gsr_model.py:32-37 followed by the L1 term against the HR target (utils/gsr_train.py:39-43) as one
    autograd node: [B] per-subject means of | |(x + x^T)/2| (unit diagonal) - y |.  Neither the prediction nor its
    gradient is materialised; the gradient of x is bit identical to ``l1_loss_per_subject(sym_diag_abs(x), y)``'s."""

    @staticmethod
    def forward(ctx, x, y):
        x, y = x.contiguous(), y.contiguous()
        ctx.save_for_backward(x, y)
        return K.sym_diag_abs_l1_fwd(x, y)[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        return K.sym_diag_abs_l1_bwd(x, y, _c(g)), None


def sym_diag_abs_l1_per_subject(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Per-subject L1 of the symmetrised prediction against ``y`` (no gradient for ``y``: it is data)."""
    if y.requires_grad:
        raise NotImplementedError("sym_diag_abs_l1_per_subject: the target is data (no gradient); use sym_diag_abs + l1_loss_per_subject")
    if not K.sym_diag_abs_l1_supported(x.contiguous(), y.contiguous()):
        return l1_loss_per_subject(sym_diag_abs(x), y)
    return SymDiagAbsL1Fn.apply(x, y)


# ----------------------------------------------------------------------------- a13
class L1LossFn(Function):
    """nn.L1Loss() (utils/gsr_train.py:9): mean |x - y|, per subject when `batch` > 1.
    y may be shared by the subjects (y.dim() == x.dim() - 1)."""

    @staticmethod
    def forward(ctx, x, y, batch):
        shared = y.numel() != x.numel()
        if shared and y.numel() * batch != x.numel():
            raise ValueError(f"l1_loss: shapes {tuple(x.shape)} and {tuple(y.shape)} do not match")
        x, y = x.contiguous(), y.contiguous()
        ctx.save_for_backward(x, y)
        ctx.batch, ctx.shared = batch, shared
        return K.l1_loss_fwd(x, y, batch, shared)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        if ctx.shared and ctx.batch > 1 and not ctx.needs_input_grad[0]:
            # |W - U_hr| of the batched step: only the shared weights need a gradient -- one pass over x
            return None, (K.l1_loss_bwd_shared_y(x, y, _c(g), ctx.batch) if ctx.needs_input_grad[1] else None), None
        gx = K.l1_loss_bwd(x, y, _c(g), ctx.batch, ctx.shared)
        gy = None
        if ctx.needs_input_grad[1]:
            gy = (-gx).reshape(y.shape) if not ctx.shared else -gx.reshape((ctx.batch,) + tuple(y.shape)).sum(0)
        return (gx if ctx.needs_input_grad[0] else None), gy, None


def l1_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Scalar mean |x - y| over all elements (nn.L1Loss default reduction)."""
    return L1LossFn.apply(x, y, 1)[0]


def l1_loss_per_subject(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """[B] per-subject means for x = [B, ...]; y is [B, ...] or shared [...]."""
    return L1LossFn.apply(x, y, x.shape[0])
