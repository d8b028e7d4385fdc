#!/usr/bin/env python
"""tcgen05 3xTF32 GEMM against float64 and against the SIMT kernel, plus timings at B = 1024."""
import importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200")
K = pkg.kernels
K.TC_MIN_ROWS = 0
g = torch.Generator().manual_seed(0)


def check(M, N, Kd, tA, tB, batch, sharedB, tag=""):
    A = torch.randn((batch, Kd, M) if tA else (batch, M, Kd), generator=g)
    B = (torch.randn((N, Kd) if tB else (Kd, N), generator=g) if sharedB
         else torch.randn((batch, N, Kd) if tB else (batch, Kd, N), generator=g))
    opA = A.transpose(-1, -2) if tA else A
    opB = B.transpose(-1, -2) if tB else B
    want = opA.double() @ opB.double()
    scale = float((opA.abs().double() @ opB.abs().double()).max())
    K.set_gemm_backend("tc")
    got = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB).cpu().double()
    K.set_gemm_backend("simt")
    ref = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB).cpu().double()
    e_tc, e_simt = float((got - want).abs().max()) / scale, float((ref - want).abs().max()) / scale
    print(json.dumps({"shape": [M, N, Kd], "tA": tA, "tB": tB, "batch": batch, "sharedB": sharedB,
                      "err_tc": e_tc, "err_simt": e_simt, "ok": e_tc < 2e-6}), flush=True)
    return e_tc < 2e-6


ok = True
for shape in [(128, 16, 32), (128, 160, 32), (128, 272, 32), (128, 288, 64), (40, 300, 70), (268, 268, 268), (268, 160, 160), (160, 268, 268), (300, 100, 52)]:
    for tA in (False, True):
        for tB in (False, True):
            ok &= check(*shape, tA, tB, 2, False)
ok &= check(268, 268, 268, False, False, 3, True)
ok &= check(268, 160, 268, False, True, 3, True)
# reduce
A = torch.randn(37, 268, 160, generator=g); B = torch.randn(37, 268, 268, generator=g)
want = (A.transpose(1, 2).double() @ B.double()).sum(0)
K.set_gemm_backend("tc"); got = K.bgemm_reduce(A.cuda(), B.cuda(), transA=True).cpu().double(); K.set_gemm_backend("simt")
scale = float((A.transpose(1, 2).abs().double() @ B.abs().double()).sum(0).max())
print("reduce err", float((got - want).abs().max()) / scale)
ok &= float((got - want).abs().max()) / scale < 2e-6
print("ALL OK" if ok else "FAILURES")


def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


Bn = 1024
gd = torch.Generator(device="cuda").manual_seed(1)
H = torch.rand(Bn, 268, 268, device="cuda", generator=gd); H2 = torch.rand(Bn, 268, 268, device="cuda", generator=gd)
W = torch.randn(268, 268, device="cuda", generator=gd)
f = 2 * 268 ** 3 * Bn
for backend in ("simt", "tc"):
    K.set_gemm_backend(backend)
    for name, fn in (("NN per-subject", lambda: K.bgemm(H, H2)), ("NT per-subject", lambda: K.bgemm(H, H2, transB=True)),
                     ("TN per-subject", lambda: K.bgemm(H, H2, transA=True)), ("NN stacked shared W", lambda: K.bgemm(H, W)),
                     ("reduce TN", lambda: K.bgemm_reduce(H, H2, transA=True))):
        ms = timeit(fn)
        print(json.dumps({"backend": backend, "gemm": name, "ms": round(ms, 4), "tflops_fp32_equiv": round(f / ms / 1e9, 1)}), flush=True)
K.set_gemm_backend("simt")
