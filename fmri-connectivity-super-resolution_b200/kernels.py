"""Tensor-level wrappers around the C ABI (include/gsr_sm100.h).

Each function checks device / dtype / layout, allocates outputs with torch (the
library never allocates) and launches on torch's current CUDA stream.  Inputs
may carry a leading subject dimension.  No function here has a non-CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from ._lib import check, lib

EPI_NONE, EPI_BIAS, EPI_ADD = 0, 1, 2


class KernelTimer:
    """CUDA-event pairs around individual launches on the launching stream (bench.py sets
    ``GEMM_TIMER`` to one of these to measure the contraction kernel inside the timed region)."""

    def __init__(self):
        self.records = []

    def start(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, e0, work: float, tag: str = "") -> None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((e0, e1, work, tag))

    def totals(self, tag: Optional[str] = None):
        """(milliseconds, work, launches) of all records, or of the records carrying `tag`."""
        torch.cuda.synchronize()
        recs = [r for r in self.records if tag is None or r[3] == tag]
        ms = sum(a.elapsed_time(b) for a, b, _, _ in recs)
        return ms, sum(w for _, _, w, _ in recs), len(recs)


GEMM_TIMER = None
EIGH_TIMER = None      # same, around gsr_eigh_jacobi_f32 (work = number of graphs)
EW_TIMER = None        # same, around the memory-bound n x n kernels (work = algorithmic bytes)

# "tc" (default): tcgen05 3xTF32 (gsr_tcgemm_f32) wherever the operands meet its alignment
# contract and the product is large enough to fill a 128-row tile, the SIMT FP32 kernel
# (gsr_bgemm_f32) otherwise; "simt": always the SIMT kernel.  Same results to fp32 tolerance.
GEMM_BACKEND = os.environ.get("GSR_GEMM_BACKEND", "tc")
TC_MIN_ROWS = 256          # smaller products (the Graph U-Net linears) cannot fill a 128-row tensor-core tile
# One product of two 2-D operands (no subject dimension: the Graph U-Net's Linear layers and their gradients, models/ops.py:60)
# with at most LIN_MAX_OUT outputs and K <= LIN_MAX_K is a latency problem: gsr_bgemm_f32 runs it on its 32 x 32-tile
# whole-panel kernel (csrc/bgemm.cu, lin::lin_kernel) -- the weight gradients (268 x 268 x <= 160) included, which the row rule
# alone would send to a tensor-core tile or two.  Operands with a subject dimension keep the row rule whatever the batch, so
# that a shard of the batch runs on the same kernels as the whole batch.
LIN_MAX_OUT, LIN_MAX_K = 1 << 18, 768


def set_gemm_backend(name: str) -> None:
    global GEMM_BACKEND
    if name not in ("simt", "tc"):
        raise ValueError("gemm backend must be 'simt' or 'tc'")
    GEMM_BACKEND = name


def _tc_ok(A: torch.Tensor, B: torch.Tensor, lda: int, ldb: int, sA: int, sB: int, rows: int = 1 << 30) -> bool:
    return (GEMM_BACKEND == "tc" and rows >= TC_MIN_ROWS and lda % 4 == 0 and ldb % 4 == 0 and sA % 4 == 0 and sB % 4 == 0
            and A.data_ptr() % 16 == 0 and B.data_ptr() % 16 == 0)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (the GSR-Net kernels have no CPU path), got {t.device}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.int64:
        raise TypeError(f"{name}: expected a CUDA int64 tensor, got {t.dtype} on {t.device}")
    return t if t.is_contiguous() else t.contiguous()


def _square_batch(t: torch.Tensor, name: str) -> Tuple[int, int]:
    if t.dim() not in (2, 3) or t.shape[-1] != t.shape[-2]:
        raise ValueError(f"{name}: expected [n, n] or [batch, n, n], got {tuple(t.shape)}")
    return (1 if t.dim() == 2 else t.shape[0]), t.shape[-1]


# --------------------------------------------------------------------- a1
def normalize_adj_fwd(mx: torch.Tensor, pre: bool = False, want_m: bool = False):
    mx = _f32(mx, "normalize_adj: mx")
    batch, n = _square_batch(mx, "normalize_adj: mx")
    out = torch.empty_like(mx)
    r = torch.empty(mx.shape[:-1], device=mx.device, dtype=torch.float32)
    m = torch.empty_like(mx) if want_m else None
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_normalize_adj_fwd(mx.data_ptr(), out.data_ptr(), r.data_ptr(),
                                      m.data_ptr() if want_m else None, batch, n, int(pre), _stream()),
          "gsr_normalize_adj_fwd")
    if timer is not None:
        timer.stop(e0, 8.0 * batch * n * n)
    return out, r, m


def normalize_adj_bwd(mx: torch.Tensor, r: torch.Tensor, grad_out: torch.Tensor, pre: bool = False,
                      grad_out2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Backward of normalize_adj_fwd.  `grad_out2`: a second contribution to the incoming gradient, summed with
    `grad_out` element by element as both passes read it (the sum is never materialised)."""
    mx, r, grad_out = _f32(mx, "mx"), _f32(r, "r"), _f32(grad_out, "grad_out")
    batch, n = _square_batch(mx, "normalize_adj_bwd: mx")
    if grad_out2 is not None:
        grad_out2 = _f32(grad_out2, "grad_out2")
        if tuple(grad_out2.shape) != tuple(grad_out.shape):
            raise ValueError("normalize_adj_bwd: grad_out2 shape != grad_out shape")
        if n % 4 != 0 or any(t.data_ptr() % 16 for t in (mx, grad_out, grad_out2)):     # the strip kernel takes one gradient
            grad_out, grad_out2 = grad_out + grad_out2, None
    grad = torch.empty_like(mx)
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    ws_floats = lib().gsr_normalize_adj_bwd_workspace(batch, n)
    ws = torch.empty((max(ws_floats, 1),), device=mx.device, dtype=torch.float32)
    if grad_out2 is None:
        check(lib().gsr_normalize_adj_bwd_tiles(mx.data_ptr(), r.data_ptr(), grad_out.data_ptr(), grad.data_ptr(),
                                                ws.data_ptr(), batch, n, int(pre), _stream()), "gsr_normalize_adj_bwd_tiles")
    else:
        check(lib().gsr_normalize_adj_bwd_tiles2(mx.data_ptr(), r.data_ptr(), grad_out.data_ptr(), grad_out2.data_ptr(),
                                                 grad.data_ptr(), ws.data_ptr(), batch, n, int(pre), _stream()),
              "gsr_normalize_adj_bwd_tiles2")
    if timer is not None:
        timer.stop(e0, (12.0 if grad_out2 is None else 20.0) * batch * n * n)
    return grad


# --------------------------------------------------------------------- a2
def eigh_jacobi(A: torch.Tensor, canonical_sign: bool = True, tol: float = 0.0, max_sweeps: int = 0):
    """Returns (w, U, info) -- see gsr_eigh_jacobi_f32.  No synchronisation."""
    A = _f32(A, "eigh: A")
    batch, n = _square_batch(A, "eigh: A")
    w = torch.empty(A.shape[:-1], device=A.device, dtype=torch.float32)
    U = torch.empty_like(A)
    info = torch.empty((batch,), device=A.device, dtype=torch.int32)
    ws_floats = lib().gsr_eigh_workspace(batch, n)
    ws = torch.empty((ws_floats,), device=A.device, dtype=torch.float32) if ws_floats else None
    timer = EIGH_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_eigh_jacobi_f32(A.data_ptr(), w.data_ptr(), U.data_ptr(), info.data_ptr(),
                                    ws.data_ptr() if ws is not None else None, batch, n,
                                    int(canonical_sign), float(tol), int(max_sweeps), _stream()),
          "gsr_eigh_jacobi_f32")
    if timer is not None:
        timer.stop(e0, float(batch))
        timer.last_info = info
        timer.last_method = "jacobi"
    return w, U, info


EIGH_TRIDIAG_MAX_N = 288
EIGH_METHOD = os.environ.get("GSR_EIGH_METHOD", "auto")     # "auto" | "jacobi" | "tridiag"


def eigh_tridiag(A: torch.Tensor, canonical_sign: bool = True, _return_workspace: bool = False):
    """Returns (w, U, info) -- see gsr_eigh_tridiag_f32 (n <= 288).  No synchronisation."""
    A = _f32(A, "eigh: A")
    batch, n = _square_batch(A, "eigh: A")
    w = torch.empty(A.shape[:-1], device=A.device, dtype=torch.float32)
    U = torch.empty_like(A)
    info = torch.empty((batch,), device=A.device, dtype=torch.int32)
    ws_bytes = lib().gsr_eigh_tridiag_workspace(batch, n)
    if ws_bytes == 0 and batch > 0:
        raise ValueError(f"eigh_tridiag: n = {n} is outside 1..{EIGH_TRIDIAG_MAX_N}")
    ws = torch.empty((max(ws_bytes, 16),), device=A.device, dtype=torch.uint8)
    timer = EIGH_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_eigh_tridiag_f32(A.data_ptr(), w.data_ptr(), U.data_ptr(), info.data_ptr(), ws.data_ptr(),
                                     ws_bytes, batch, n, int(canonical_sign), _stream()),
          "gsr_eigh_tridiag_f32")
    if timer is not None:
        timer.stop(e0, float(batch))
        timer.last_info = info
        timer.last_method = "tridiag"
    if _return_workspace:
        return w, U, info, ws
    return w, U, info


def eigh(A: torch.Tensor, canonical_sign: bool = True, method: Optional[str] = None):
    """torch.linalg.eigh(A, 'U') on the device: the tridiagonal-QL kernels up to 288 nodes (register-resident
    for the 160-node LR graphs, shared-memory variants for the 268-node HR target); the Jacobi kernel is the
    fallback (functional.eigh) and the explicit ``method='jacobi'`` route."""
    method = method or EIGH_METHOD
    n = A.shape[-1]
    if method == "tridiag" or (method == "auto" and n <= EIGH_TRIDIAG_MAX_N):
        return eigh_tridiag(A, canonical_sign=canonical_sign)
    return eigh_jacobi(A, canonical_sign=canonical_sign)


def set_eigh_method(name: str) -> None:
    global EIGH_METHOD
    if name not in ("auto", "jacobi", "tridiag"):
        raise ValueError("eigh method must be 'auto', 'jacobi' or 'tridiag'")
    EIGH_METHOD = name


# --------------------------------------------------------------------- GEMMs
def _f32_rows(t: torch.Tensor, name: str):
    """(tensor, leading dimension): a 2-D operand may be a view with padded rows (unit column stride, row stride >= columns:
    e.g. ``buf[:, :35778]`` of a [B, 35780] buffer, whose rows start 16-byte aligned); anything else is made contiguous."""
    if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 \
            and t.stride(0) >= t.shape[1] and t.shape[0] > 1:
        return t, t.stride(0)
    t = _f32(t, name)
    return t, t.shape[-1]


# K-split of ONE skinny product with a long inner dimension (the notebook MLP's 512 x 1024 x 12 720 / 35 778 layers): the
# 256-row x 256-column tiles of such a product are few (8) and each walks all of K, so the 74 CTA pairs idle.  The K slices
# run as the items of one batched call into a [splits, M, N] workspace and gsr_splitk_reduce_f32 sums them in a fixed order.
SPLITK_MIN_K, SPLITK_MIN_SLICE, SPLITK_MAX_TILES = 1024, 256, 24


def _splitk_plan(M: int, N: int, K: int):
    """(splits, slice length, remainder) or None.  The slice is a multiple of 4 (16-byte aligned slice starts); the last
    K - splits * slice < 4 * splits columns are a product of their own, which also carries the epilogue."""
    if K < SPLITK_MIN_K or M < TC_MIN_ROWS:
        return None
    tiles = ((M + 255) // 256) * (1 if N <= 272 else (N + 255) // 256)
    if tiles > SPLITK_MAX_TILES:
        return None
    splits = min(max(2, round(74 / tiles)), K // SPLITK_MIN_SLICE)
    if splits < 2:
        return None
    kc = (K // splits) & ~3
    return splits, kc, K - splits * kc


def bgemm(A: torch.Tensor, B: torch.Tensor, transA: bool = False, transB: bool = False, alpha: float = 1.0,
          bias: Optional[torch.Tensor] = None, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[b] = alpha * op(A[b]) @ op(B[b]) (+ bias[n] | + add[b]).  2-D operands are shared by
    all subjects.  A 3-D non-transposed A against a shared B is run as ONE stacked GEMM
    (M = batch * rows): the weight tile is reused across subjects and no tile is ragged."""
    (A, lda), (B, ldb) = _f32_rows(A, "bgemm: A"), _f32_rows(B, "bgemm: B")
    if A.dim() not in (2, 3) or B.dim() not in (2, 3):
        raise ValueError("bgemm: operands must be 2-D or 3-D")
    ba = A.shape[0] if A.dim() == 3 else 1
    bb = B.shape[0] if B.dim() == 3 else 1
    if A.dim() == 3 and B.dim() == 3 and ba != bb:
        raise ValueError(f"bgemm: batch mismatch {ba} vs {bb}")
    batch = max(ba, bb)
    ar, ac = A.shape[-2], A.shape[-1]
    br, bc = B.shape[-2], B.shape[-1]
    M, K = (ac, ar) if transA else (ar, ac)
    K2, N = (bc, br) if transB else (br, bc)
    if K != K2:
        raise RuntimeError(f"bgemm: inner dimensions differ: op(A) is {M}x{K}, op(B) is {K2}x{N}")
    batched_out = A.dim() == 3 or B.dim() == 3
    out = torch.empty((batch, M, N) if batched_out else (M, N), device=A.device, dtype=torch.float32)
    sA = ar * ac if A.dim() == 3 else 0
    sB = br * bc if B.dim() == 3 else 0
    epi, extra, ld_extra, s_extra = EPI_NONE, None, 0, 0
    if bias is not None and add is not None:
        raise ValueError("bgemm: bias and add are mutually exclusive")
    if bias is not None:
        bias = _f32(bias, "bgemm: bias")
        if bias.numel() != N:
            raise ValueError("bgemm: bias length != N")
        epi, extra = EPI_BIAS, bias
    if add is not None:
        add = _f32(add, "bgemm: add")
        if tuple(add.shape) != tuple(out.shape):
            raise ValueError(f"bgemm: add shape {tuple(add.shape)} != output shape {tuple(out.shape)}")
        epi, extra, ld_extra, s_extra = EPI_ADD, add, N, M * N
    Mk, bk, sAk, sCk, sEk = M, batch, sA, M * N, s_extra
    if A.dim() == 3 and B.dim() == 2 and not transA:          # stack the subjects along M
        Mk, bk, sAk, sCk, sEk = batch * M, 1, 0, 0, 0
    timer = GEMM_TIMER
    e0 = timer.start() if timer is not None else None
    tiny = A.dim() == 2 and B.dim() == 2 and M * N <= LIN_MAX_OUT and K <= LIN_MAX_K
    use_tc = not tiny and _tc_ok(A, B, lda, ldb, sAk, sB, Mk * bk)
    plan = _splitk_plan(Mk, N, K) if (use_tc and bk == 1 and N % 4 == 0) else None
    if plan is not None:
        splits, kc, krem = plan
        k0 = splits * kc
        if krem:          # the last columns of K first, as a product of their own with the epilogue; the reduction adds to it
            Ar = A.reshape(Mk, K)[:, k0:] if not transA else A[k0:, :]
            Br = B[:, k0:] if transB else B[k0:, :]
            check(lib().gsr_bgemm_f32(Ar.data_ptr(), Br.data_ptr(), out.data_ptr(), Mk, N, krem, lda, ldb, N, 0, 0, 0,
                                      int(transA), int(transB), 1, float(alpha), epi,
                                      extra.data_ptr() if extra is not None else None, ld_extra, 0, _stream()), "gsr_bgemm_f32")
        ws = torch.empty((splits, Mk, N), device=A.device, dtype=torch.float32)
        check(lib().gsr_tcgemm_f32(A.data_ptr(), B.data_ptr(), ws.data_ptr(), Mk, N, kc, lda, ldb, N,
                                   kc * lda if transA else kc, kc if transB else kc * ldb, Mk * N,
                                   int(transA), int(transB), splits, 1.0, EPI_NONE, None, 0, 0, _stream()), "gsr_tcgemm_f32 (K slices)")
        if krem:
            check(lib().gsr_splitk_reduce_f32(ws.data_ptr(), out.data_ptr(), Mk, N, N, splits, float(alpha), EPI_ADD,
                                              out.data_ptr(), N, _stream()), "gsr_splitk_reduce_f32")
        else:
            check(lib().gsr_splitk_reduce_f32(ws.data_ptr(), out.data_ptr(), Mk, N, N, splits, float(alpha), epi,
                                              extra.data_ptr() if extra is not None else None, ld_extra, _stream()),
                  "gsr_splitk_reduce_f32")
    else:
        fn = lib().gsr_tcgemm_f32 if use_tc else lib().gsr_bgemm_f32
        check(fn(A.data_ptr(), B.data_ptr(), out.data_ptr(), Mk, N, K, lda, ldb, N,
                 sAk, sB, sCk, int(transA), int(transB), bk, float(alpha),
                 epi, extra.data_ptr() if extra is not None else None, ld_extra, sEk, _stream()),
              "gsr_bgemm_f32 / gsr_tcgemm_f32")
    if timer is not None:
        timer.stop(e0, 2.0 * M * N * K * batch, "tc" if use_tc else "simt")
    return out


def bgemm_dual(A1: torch.Tensor, B1: torch.Tensor, A2: torch.Tensor, B2: torch.Tensor, transA: bool = False,
               transB: bool = False, alpha: float = 1.0) -> torch.Tensor:
    """C[b] = alpha * (op(A1[b]) @ op(B1[b]) + op(A2[b]) @ op(B2[b])): two per-subject products of one shape summed in the
    accumulator of ONE launch (gsr_tcgemm_dual_f32) -- the first product is never written out and read back as an addend.
    Falls back to two launches (the second with add=) off the tensor-core path."""
    A1, B1, A2, B2 = _f32(A1, "bgemm_dual: A1"), _f32(B1, "bgemm_dual: B1"), _f32(A2, "bgemm_dual: A2"), _f32(B2, "bgemm_dual: B2")
    if all(t.dim() == 2 for t in (A1, B1, A2, B2)):      # one subject (the reference's batch-1 calls)
        return bgemm_dual(A1[None], B1[None], A2[None], B2[None], transA=transA, transB=transB, alpha=alpha)[0]
    if any(t.dim() != 3 for t in (A1, B1, A2, B2)) or len({t.shape[0] for t in (A1, B1, A2, B2)}) != 1:
        raise ValueError("bgemm_dual: all four operands must be [batch, ., .] with equal batch (or all 2-D)")
    batch = A1.shape[0]

    def dims(A, B):
        ar, ac, br, bc = A.shape[-2], A.shape[-1], B.shape[-2], B.shape[-1]
        M, K = (ac, ar) if transA else (ar, ac)
        Kb, N = (bc, br) if transB else (br, bc)
        if K != Kb:
            raise RuntimeError(f"bgemm_dual: inner dimensions differ: op(A) is {M}x{K}, op(B) is {Kb}x{N}")
        return M, N, K, ac, bc, ar * ac, br * bc

    M, N, K1, lda1, ldb1, sA1, sB1 = dims(A1, B1)
    M2, N2, K2, lda2, ldb2, sA2, sB2 = dims(A2, B2)
    if (M, N) != (M2, N2):
        raise RuntimeError(f"bgemm_dual: the two products differ in shape: {M}x{N} vs {M2}x{N2}")
    if not (_tc_ok(A1, B1, lda1, ldb1, sA1, sB1, M * batch) and _tc_ok(A2, B2, lda2, ldb2, sA2, sB2, M * batch)):
        return bgemm(A2, B2, transA=transA, transB=transB, alpha=alpha,
                     add=bgemm(A1, B1, transA=transA, transB=transB, alpha=alpha))
    out = torch.empty((batch, M, N), device=A1.device, dtype=torch.float32)
    timer = GEMM_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_tcgemm_dual_f32(A1.data_ptr(), B1.data_ptr(), A2.data_ptr(), B2.data_ptr(), out.data_ptr(),
                                    M, N, K1, K2, lda1, ldb1, lda2, ldb2, N, sA1, sB1, sA2, sB2, M * N,
                                    int(transA), int(transB), batch, float(alpha), _stream()), "gsr_tcgemm_dual_f32")
    if timer is not None:
        timer.stop(e0, 2.0 * M * N * (K1 + K2) * batch, "tc")
    return out


def bgemm_reduce(A: torch.Tensor, B: torch.Tensor, transA: bool = False, transB: bool = False,
                 alpha: float = 1.0, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = alpha * sum_b op(A[b]) @ op(B[b]) (+ add): weight gradients over the subject batch."""
    A, B = _f32(A, "bgemm_reduce: A"), _f32(B, "bgemm_reduce: B")
    if A.dim() != 3 or B.dim() != 3 or A.shape[0] != B.shape[0]:
        raise ValueError("bgemm_reduce: both operands must be [batch, ., .] with equal batch")
    batch = A.shape[0]
    ar, ac = A.shape[-2], A.shape[-1]
    br, bc = B.shape[-2], B.shape[-1]
    M, K = (ac, ar) if transA else (ar, ac)
    K2, N = (bc, br) if transB else (br, bc)
    if K != K2:
        raise RuntimeError(f"bgemm_reduce: inner dimensions differ ({K} vs {K2})")
    out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    splits = C.c_int(0)
    use_tc = _tc_ok(A, B, ac, bc, ar * ac, br * bc)
    ws_fn = lib().gsr_tcgemm_reduce_workspace if use_tc else lib().gsr_bgemm_reduce_workspace
    ws_floats = ws_fn(M, N, batch, C.byref(splits))
    ws = torch.empty((ws_floats,), device=A.device, dtype=torch.float32)
    epi, extra = EPI_NONE, None
    if add is not None:
        add = _f32(add, "bgemm_reduce: add")
        if tuple(add.shape) != (M, N):
            raise ValueError("bgemm_reduce: add shape != (M, N)")
        epi, extra = EPI_ADD, add
    timer = GEMM_TIMER
    e0 = timer.start() if timer is not None else None
    fn = lib().gsr_tcgemm_reduce_f32 if use_tc else lib().gsr_bgemm_reduce_f32
    check(fn(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, ac, bc, N,
             ar * ac, br * bc, int(transA), int(transB), batch, float(alpha),
             epi, extra.data_ptr() if extra is not None else None, N,
             ws.data_ptr(), splits.value, _stream()), "gsr_bgemm_reduce_f32 / gsr_tcgemm_reduce_f32")
    if timer is not None:
        timer.stop(e0, 2.0 * M * N * K * batch, "tc" if use_tc else "simt")
    return out


def colsum(X: torch.Tensor) -> torch.Tensor:
    X = _f32(X, "colsum: X")
    X2 = X.reshape(-1, X.shape[-1])
    out = torch.empty((X2.shape[1],), device=X.device, dtype=torch.float32)
    check(lib().gsr_colsum_f32(X2.data_ptr(), out.data_ptr(), X2.shape[0], X2.shape[1], X2.shape[1], _stream()),
          "gsr_colsum_f32")
    return out


# --------------------------------------------------------------------- sym / diag / abs
def sym_diag_abs_fwd(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "sym_diag_abs: x")
    batch, n = _square_batch(x, "sym_diag_abs: x")
    out = torch.empty_like(x)
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_sym_diag_abs_fwd(x.data_ptr(), out.data_ptr(), batch, n, _stream()), "gsr_sym_diag_abs_fwd")
    if timer is not None:
        timer.stop(e0, 8.0 * batch * n * n)
    return out


def sym_diag_abs_bwd(x: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    x, grad_out = _f32(x, "x"), _f32(grad_out, "grad_out")
    batch, n = _square_batch(x, "sym_diag_abs_bwd: x")
    grad = torch.empty_like(x)
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_sym_diag_abs_bwd(x.data_ptr(), grad_out.data_ptr(), grad.data_ptr(), batch, n, _stream()),
          "gsr_sym_diag_abs_bwd")
    if timer is not None:
        timer.stop(e0, 12.0 * batch * n * n)
    return grad


def sym_diag_abs_l1_supported(x: torch.Tensor, y: torch.Tensor) -> bool:
    """The fused pass takes contiguous fp32 [B, n, n] with n % 4 == 0 (the tile kernels' float4 rows)."""
    return (x.dim() == 3 and x.shape == y.shape and x.shape[-1] == x.shape[-2] and x.shape[-1] % 4 == 0
            and x.shape[0] <= 65535 and x.is_contiguous() and y.is_contiguous()
            and x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0)


def sym_diag_abs_l1_fwd(x: torch.Tensor, y: torch.Tensor, want_pred: bool = False):
    """(per-subject mean | |(x + x^T)/2| (unit diagonal) - y |  [B], the prediction or None): one pass over x and y."""
    x, y = _f32(x, "sym_diag_abs_l1: x"), _f32(y, "sym_diag_abs_l1: y")
    batch, n = _square_batch(x, "sym_diag_abs_l1: x")
    loss = torch.empty((batch,), device=x.device, dtype=torch.float32)
    part = torch.empty((max(batch, 1) * lib().gsr_sym_diag_abs_l1_partials(n),), device=x.device, dtype=torch.float64)
    pred = torch.empty_like(x) if want_pred else None
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_sym_diag_abs_l1_fwd(x.data_ptr(), y.data_ptr(), pred.data_ptr() if want_pred else None, loss.data_ptr(),
                                        part.data_ptr(), batch, n, _stream()), "gsr_sym_diag_abs_l1_fwd")
    if timer is not None:
        timer.stop(e0, (12.0 if want_pred else 8.0) * batch * n * n)
    return loss, pred


def sym_diag_abs_l1_bwd(x: torch.Tensor, y: torch.Tensor, grad_loss: torch.Tensor) -> torch.Tensor:
    x, y, grad_loss = _f32(x, "x"), _f32(y, "y"), _f32(grad_loss, "grad_loss")
    batch, n = _square_batch(x, "sym_diag_abs_l1_bwd: x")
    if grad_loss.numel() != batch:
        raise ValueError("sym_diag_abs_l1_bwd: grad_loss must hold one value per subject")
    grad = torch.empty_like(x)
    timer = EW_TIMER
    e0 = timer.start() if timer is not None else None
    check(lib().gsr_sym_diag_abs_l1_bwd(x.data_ptr(), y.data_ptr(), grad_loss.data_ptr(), grad.data_ptr(), batch, n, _stream()),
          "gsr_sym_diag_abs_l1_bwd")
    if timer is not None:
        timer.stop(e0, 12.0 * batch * n * n)
    return grad


def abs_diag1(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "abs_diag1: x")
    batch, n = _square_batch(x, "abs_diag1: x")
    out = torch.empty_like(x)
    check(lib().gsr_abs_diag1_fwd(x.data_ptr(), out.data_ptr(), batch, n, _stream()), "gsr_abs_diag1_fwd")
    return out


# --------------------------------------------------------------------- graph U-Net ops
def pool_scores(X: torch.Tensor, p: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    X, p, c = _f32(X, "pool: X"), _f32(p, "pool: proj.weight"), _f32(c, "pool: proj.bias")
    n, d = X.shape
    if p.numel() != d or c.numel() != 1:
        raise ValueError("pool_scores: projection shape mismatch")
    s = torch.empty((n,), device=X.device, dtype=torch.float32)
    check(lib().gsr_pool_scores(X.data_ptr(), p.data_ptr(), c.data_ptr(), s.data_ptr(), n, d, _stream()),
          "gsr_pool_scores")
    return s


def pool_topk(scores: torch.Tensor, k: int):
    scores = _f32(scores, "topk: scores")
    n = scores.numel()
    idx = torch.empty((k,), device=scores.device, dtype=torch.int64)
    vals = torch.empty((k,), device=scores.device, dtype=torch.float32)
    check(lib().gsr_pool_topk(scores.data_ptr(), n, int(k), idx.data_ptr(), vals.data_ptr(), _stream()),
          "gsr_pool_topk")
    return vals, idx


def pool_gather_scale(X: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor) -> torch.Tensor:
    X, idx, vals = _f32(X, "X"), _i64(idx, "idx"), _f32(vals, "values")
    k, d = idx.numel(), X.shape[1]
    out = torch.empty((k, d), device=X.device, dtype=torch.float32)
    check(lib().gsr_pool_gather_scale_fwd(X.data_ptr(), idx.data_ptr(), vals.data_ptr(), out.data_ptr(), k, d, _stream()),
          "gsr_pool_gather_scale_fwd")
    return out


def pool_bwd(X, p, scores, idx, grad_new_X):
    X, p, scores, idx, g = _f32(X, "X"), _f32(p, "p"), _f32(scores, "scores"), _i64(idx, "idx"), _f32(grad_new_X, "grad")
    n, d = X.shape
    k = idx.numel()
    gX = torch.empty_like(X)
    gp = torch.empty((d,), device=X.device, dtype=torch.float32)
    gc = torch.empty((1,), device=X.device, dtype=torch.float32)
    ws = torch.empty((n,), device=X.device, dtype=torch.float32)
    check(lib().gsr_pool_bwd(X.data_ptr(), p.data_ptr(), scores.data_ptr(), idx.data_ptr(), g.data_ptr(),
                             gX.data_ptr(), gp.data_ptr(), gc.data_ptr(), ws.data_ptr(), n, k, d, _stream()),
          "gsr_pool_bwd")
    return gX, gp, gc


def gather_rows_cols(A: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    A, idx = _f32(A, "A"), _i64(idx, "idx")
    batch, n = _square_batch(A, "gather_rows_cols: A")
    k = idx.numel()
    out = torch.empty(A.shape[:-2] + (k, k), device=A.device, dtype=torch.float32)
    check(lib().gsr_gather_rows_cols(A.data_ptr(), idx.data_ptr(), out.data_ptr(), batch, n, k, _stream()),
          "gsr_gather_rows_cols")
    return out


def unpool_fwd(X: torch.Tensor, idx: torch.Tensor, n: int) -> torch.Tensor:
    X, idx = _f32(X, "X"), _i64(idx, "idx")
    k, d = X.shape
    if idx.numel() != k:
        raise ValueError("unpool: idx length != rows of X")
    out = torch.empty((n, d), device=X.device, dtype=torch.float32)
    check(lib().gsr_unpool_fwd(X.data_ptr(), idx.data_ptr(), out.data_ptr(), n, k, d, _stream()), "gsr_unpool_fwd")
    return out


def unpool_bwd(grad_new_X: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    g, idx = _f32(grad_new_X, "grad"), _i64(idx, "idx")
    k, d = idx.numel(), g.shape[1]
    out = torch.empty((k, d), device=g.device, dtype=torch.float32)
    check(lib().gsr_unpool_bwd(g.data_ptr(), idx.data_ptr(), out.data_ptr(), k, d, _stream()), "gsr_unpool_bwd")
    return out


# --------------------------------------------------------------------- L1 loss
def l1_loss_fwd(x: torch.Tensor, y: torch.Tensor, batch: int, shared_y: bool) -> torch.Tensor:
    """x viewed as [batch, count]; y is [batch, count] or, with shared_y, [count]."""
    x, y = _f32(x, "l1: x"), _f32(y, "l1: y")
    count = x.numel() // batch
    loss = torch.empty((batch,), device=x.device, dtype=torch.float32)
    chunks = lib().gsr_l1_partials(count)
    part = torch.empty((batch * chunks,), device=x.device, dtype=torch.float64)
    check(lib().gsr_l1_loss_fwd(x.data_ptr(), y.data_ptr(), 0 if shared_y else count, loss.data_ptr(),
                                part.data_ptr(), batch, count, _stream()), "gsr_l1_loss_fwd")
    return loss


def l1_loss_bwd(x: torch.Tensor, y: torch.Tensor, grad_loss: torch.Tensor, batch: int, shared_y: bool) -> torch.Tensor:
    x, y, grad_loss = _f32(x, "l1: x"), _f32(y, "l1: y"), _f32(grad_loss, "l1: grad")
    count = x.numel() // batch
    gx = torch.empty_like(x)
    check(lib().gsr_l1_loss_bwd(x.data_ptr(), y.data_ptr(), 0 if shared_y else count, grad_loss.data_ptr(),
                                gx.data_ptr(), batch, count, _stream()), "gsr_l1_loss_bwd")
    return gx


def l1_loss_bwd_shared_y(x: torch.Tensor, y: torch.Tensor, grad_loss: torch.Tensor, batch: int) -> torch.Tensor:
    """Gradient of the shared y of a per-subject L1 loss (x [batch, ...], y [...]), without the gradient of x."""
    x, y, grad_loss = _f32(x, "l1: x"), _f32(y, "l1: y"), _f32(grad_loss, "l1: grad")
    count = x.numel() // batch
    gy = torch.empty_like(y)
    check(lib().gsr_l1_loss_bwd_shared_y(x.data_ptr(), y.data_ptr(), grad_loss.data_ptr(), gy.data_ptr(), batch, count,
                                         _stream()), "gsr_l1_loss_bwd_shared_y")
    return gy


# --------------------------------------------------------------------- vectoriser
def vectorize_len(n: int, include_diagonal: bool) -> int:
    return int(lib().gsr_vectorize_len(int(n), int(include_diagonal)))


def vectorize(matrix: torch.Tensor, include_diagonal: bool = False) -> torch.Tensor:
    if not matrix.is_cuda:
        raise RuntimeError("vectorize: expected a CUDA tensor")
    matrix = matrix.contiguous()
    batch, n = _square_batch(matrix, "vectorize: matrix")
    L = vectorize_len(n, include_diagonal)
    out = torch.empty(matrix.shape[:-2] + (L,), device=matrix.device, dtype=matrix.dtype)
    if L:
        check(lib().gsr_vectorize(matrix.data_ptr(), out.data_ptr(), batch, n, matrix.element_size(),
                                  int(include_diagonal), _stream()), "gsr_vectorize")
    return out


def anti_vectorize(vec: torch.Tensor, n: int, include_diagonal: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out`` (optional): a contiguous float32 [..., n, n] buffer to fill, e.g. the input buffer of a captured step."""
    vec = _f32(vec, "anti_vectorize: vector")
    if vec.dim() not in (1, 2):
        raise ValueError("anti_vectorize: expected [len] or [batch, len]")
    batch = 1 if vec.dim() == 1 else vec.shape[0]
    if out is None:
        out = torch.empty(vec.shape[:-1] + (n, n), device=vec.device, dtype=torch.float32)
    elif (not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous()
          or tuple(out.shape) != tuple(vec.shape[:-1]) + (n, n)):
        raise ValueError("anti_vectorize: `out` must be a contiguous CUDA float32 tensor of shape [..., n, n]")
    check(lib().gsr_anti_vectorize(vec.data_ptr(), out.data_ptr(), batch, int(n), vec.shape[-1],
                                   int(include_diagonal), _stream()), "gsr_anti_vectorize")
    return out


def triu_flatten(matrix: torch.Tensor) -> torch.Tensor:
    """matrix[np.triu_indices(n, k=1)] for every matrix of the batch (row-major strict upper triangle)."""
    matrix = _f32(matrix, "triu_flatten: matrix")
    batch, n = _square_batch(matrix, "triu_flatten: matrix")
    out = torch.empty(matrix.shape[:-2] + (n * (n - 1) // 2,), device=matrix.device, dtype=torch.float32)
    if out.numel():
        check(lib().gsr_triu_flatten(matrix.data_ptr(), out.data_ptr(), batch, n, _stream()), "gsr_triu_flatten")
    return out


# --------------------------------------------------------------------- evaluation
def eval_reduce(pred: torch.Tensor, gt: torch.Tensor, upper: bool) -> torch.Tensor:
    """Returns a device fp64 tensor [mae, pcc, js_distance]."""
    pred, gt = _f32(pred, "eval: pred"), _f32(gt, "eval: gt")
    if pred.shape != gt.shape:
        raise ValueError("eval_reduce: shape mismatch")
    if upper:
        if pred.dim() == 2:
            pred, gt = pred[None], gt[None]
        batch, n = pred.shape[0], pred.shape[-1]
    else:
        batch, n = 1, pred.numel()
        if n >= 2 ** 31:
            n = pred.shape[-1]
            batch = pred.numel() // n
    ws = torch.empty((lib().gsr_eval_workspace(),), device=pred.device, dtype=torch.float64)
    res = torch.empty((3,), device=pred.device, dtype=torch.float64)
    check(lib().gsr_eval_reduce(pred.data_ptr(), gt.data_ptr(), batch, n, int(upper), ws.data_ptr(),
                                res.data_ptr(), _stream()), "gsr_eval_reduce")
    return res


# --------------------------------------------------------------------- N2: BatchNorm1d + Dropout + ReLU
def bn_stats(x: torch.Tensor, running_mean=None, running_var=None, momentum: float = 0.1):
    """Batch mean and biased variance of x [B, H]; updates the running estimates in place when given."""
    x = _f32(x, "bn_stats: x")
    B, H = x.shape
    mean, var = torch.empty((H,), device=x.device), torch.empty((H,), device=x.device)
    check(lib().gsr_bn_stats_f32(x.data_ptr(), B, H, mean.data_ptr(), var.data_ptr(),
                                 running_mean.data_ptr() if running_mean is not None else None,
                                 running_var.data_ptr() if running_var is not None else None,
                                 float(momentum), _stream()), "gsr_bn_stats_f32")
    return mean, var


def _slope(relu) -> float:
    """`relu` of the BatchNorm pass: True / False (nn.ReLU or nothing, models/mlp.py) or the negative slope of a
    LeakyReLU as a float (0.01 in the notebook MLP)."""
    if relu is True:
        return 0.0
    if relu is False or relu is None:
        return 1.0
    return float(relu)


def bn_act_fwd(x, mean, var, gamma, beta, eps: float, p_drop: float, seed: int, relu) -> torch.Tensor:
    x = _f32(x, "bn_act_fwd: x")
    B, H = x.shape
    y = torch.empty_like(x)
    check(lib().gsr_bn_act_fwd_f32(x.data_ptr(), mean.data_ptr(), var.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                   float(eps), float(p_drop), int(seed), _slope(relu), y.data_ptr(), B, H, _stream()),
          "gsr_bn_act_fwd_f32")
    return y


def bn_act_bwd(x, mean, var, gamma, beta, dy, eps: float, p_drop: float, seed: int, relu, batch_stats: bool):
    x, dy = _f32(x, "bn_act_bwd: x"), _f32(dy, "bn_act_bwd: dy")
    B, H = x.shape
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.empty((H,), device=x.device), torch.empty((H,), device=x.device)
    check(lib().gsr_bn_act_bwd_f32(x.data_ptr(), mean.data_ptr(), var.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                   dy.data_ptr(), float(eps), float(p_drop), int(seed), _slope(relu), int(batch_stats),
                                   dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), B, H, _stream()),
          "gsr_bn_act_bwd_f32")
    return dx, dgamma, dbeta


def add_act_fwd(a: torch.Tensor, b: torch.Tensor, negative_slope: float) -> torch.Tensor:
    a, b = _f32(a, "add_act: a"), _f32(b, "add_act: b")
    if a.shape != b.shape:
        raise ValueError("add_act: shape mismatch")
    out = torch.empty_like(a)
    check(lib().gsr_add_act_fwd_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), float(negative_slope), _stream()), "gsr_add_act_fwd_f32")
    return out


def act_bwd(out: torch.Tensor, grad_out: torch.Tensor, negative_slope: float) -> torch.Tensor:
    out, grad_out = _f32(out, "act_bwd: out"), _f32(grad_out, "act_bwd: grad")
    g = torch.empty_like(out)
    check(lib().gsr_act_bwd_f32(out.data_ptr(), grad_out.data_ptr(), g.data_ptr(), out.numel(), float(negative_slope), _stream()), "gsr_act_bwd_f32")
    return g


def scale_div(x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """x / sigma with sigma a DEVICE scalar (spectral_norm: W / sigma); one streaming pass, nothing synchronises."""
    x, sigma = _f32(x, "scale_div: x"), _f32(sigma, "scale_div: sigma")
    if sigma.numel() != 1:
        raise ValueError("scale_div: sigma must hold one element")
    y = torch.empty_like(x)
    check(lib().gsr_scale_div_f32(x.data_ptr(), sigma.data_ptr(), y.data_ptr(), x.numel(), _stream()), "gsr_scale_div_f32")
    return y


def spectral_bwd(grad_w_eff: torch.Tensor, w_eff: torch.Tensor, u: torch.Tensor, v: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    """Gradient of W for W_eff = W / sigma, sigma = u . (W v):  G / sigma - (sum(G o W_eff) / sigma) u v^T."""
    g, w = _f32(grad_w_eff, "spectral_bwd: grad"), _f32(w_eff, "spectral_bwd: w_eff")
    u, v, sigma = _f32(u, "spectral_bwd: u"), _f32(v, "spectral_bwd: v"), _f32(sigma, "spectral_bwd: sigma")
    if w.dim() != 2 or tuple(g.shape) != tuple(w.shape) or u.numel() != w.shape[0] or v.numel() != w.shape[1] or sigma.numel() != 1:
        raise ValueError("spectral_bwd: expected grad / w_eff [rows, cols], u [rows], v [cols], sigma [1]")
    out = torch.empty_like(w)
    ws = torch.empty((lib().gsr_spectral_bwd_workspace(),), device=w.device, dtype=torch.float64)
    check(lib().gsr_spectral_bwd_f32(g.data_ptr(), w.data_ptr(), u.data_ptr(), v.data_ptr(), sigma.data_ptr(), out.data_ptr(),
                                     w.shape[0], w.shape[1], ws.data_ptr(), _stream()), "gsr_spectral_bwd_f32")
    return out


def sigmoid_fwd(x: torch.Tensor) -> torch.Tensor:
    x = _f32(x, "sigmoid: x")
    y = torch.empty_like(x)
    check(lib().gsr_sigmoid_fwd_f32(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "gsr_sigmoid_fwd_f32")
    return y


def sigmoid_bwd(y: torch.Tensor, grad_y: torch.Tensor) -> torch.Tensor:
    y, grad_y = _f32(y, "sigmoid_bwd: y"), _f32(grad_y, "sigmoid_bwd: grad")
    g = torch.empty_like(y)
    check(lib().gsr_sigmoid_bwd_f32(y.data_ptr(), grad_y.data_ptr(), g.data_ptr(), y.numel(), _stream()), "gsr_sigmoid_bwd_f32")
    return g


# --------------------------------------------------------------------- flat gradient bucket
class FlatBucket:
    """Table (pointers + offsets, passed to the kernels by value) of a list of fp32 tensors, for the one-launch
    pack / unpack around the NCCL all-reduce of the subject-data-parallel step."""

    def __init__(self, tensors):
        self.tensors = [t for t in tensors]
        for t in self.tensors:
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise TypeError("FlatBucket: tensors must be contiguous CUDA float32")
        dev = self.tensors[0].device
        offs = [0]
        for t in self.tensors:
            offs.append(offs[-1] + t.numel())
        self.total = offs[-1]
        # host tables: gsr_flat_pack / _unpack pass them to the kernel by value (capturable in a CUDA graph)
        self.ptrs = (C.c_void_p * len(self.tensors))(*[t.data_ptr() for t in self.tensors])
        self.offsets = (C.c_longlong * len(offs))(*offs)
        self.flat = torch.empty((self.total,), device=dev, dtype=torch.float32)

    def pack(self, scale: float = 1.0) -> torch.Tensor:
        check(lib().gsr_flat_pack(self.ptrs, self.offsets, len(self.tensors),
                                  self.flat.data_ptr(), self.total, float(scale), _stream()), "gsr_flat_pack")
        return self.flat

    def unpack(self, scale: float = 1.0) -> None:
        check(lib().gsr_flat_unpack(self.ptrs, self.offsets, len(self.tensors),
                                    self.flat.data_ptr(), self.total, float(scale), _stream()), "gsr_flat_unpack")
