"""TMA-fed persistent CTA-pair tcgemm (csrc/tcgemm_tma.cu, gsr_set_tuning key tcgemm_tma): accuracy against float64 for every
operand layout / epilogue / reduction the step uses, bit-comparison with the thread-staged pair kernel where the arithmetic
is the same, and time per launch at 1024 subjects."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
lib = pkg.lib()
g = torch.Generator(device="cuda").manual_seed(0)
def timeit(fn, reps=7):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
Bs = 9
X = torch.randn(Bs, 268, 268, device="cuda", generator=g); Y = torch.randn(Bs, 268, 268, device="cuda", generator=g)
W = torch.randn(268, 268, device="cuda", generator=g); Wl = torch.randn(268, 160, device="cuda", generator=g)
U = torch.randn(Bs, 160, 160, device="cuda", generator=g); add = torch.randn(Bs, 268, 268, device="cuda", generator=g)
Pt = torch.randn(Bs, 160, 268, device="cuda", generator=g); f = torch.randn(160, 268, device="cuda", generator=g)
bias = torch.randn(268, device="cuda", generator=g)
X512 = torch.randn(3, 512, 268, device="cuda", generator=g)
cases = {
    "NN per-subject": (lambda: K.bgemm(X, Y), lambda: X.double() @ Y.double()),
    "NT per-subject": (lambda: K.bgemm(X, Y, transB=True), lambda: X.double() @ Y.double().transpose(1, 2)),
    "TN per-subject": (lambda: K.bgemm(X, Y, transA=True), lambda: X.double().transpose(1, 2) @ Y.double()),
    "NN + add, alpha": (lambda: K.bgemm(X, Y, alpha=2.0, add=add), lambda: 2 * (X.double() @ Y.double()) + add.double()),
    "NT + add": (lambda: K.bgemm(X, Y, transB=True, add=add), lambda: X.double() @ Y.double().transpose(1, 2) + add.double()),
    "NN stacked W": (lambda: K.bgemm(X, W), lambda: X.double() @ W.double()),
    "NT stacked W": (lambda: K.bgemm(X, W, transB=True), lambda: X.double() @ W.double().T),
    "NT stacked W + bias": (lambda: K.bgemm(X, W, transB=True, bias=bias), lambda: X.double() @ W.double().T + bias.double()),
    "NT stacked add": (lambda: K.bgemm(X, W, transB=True, add=add), lambda: X.double() @ W.double().T + add.double()),
    "NT stacked U.W^T (K=160)": (lambda: K.bgemm(U, Wl, transB=True), lambda: U.double() @ Wl.double().T),
    "NT stacked N=160 (K=268)": (lambda: K.bgemm(X, Wl.T.contiguous(), transB=True), lambda: X.double() @ Wl.double()),
    "NN stacked N=160": (lambda: K.bgemm(X, Wl), lambda: X.double() @ Wl.double()),
    "TN Pt^T f (shared B, K=160)": (lambda: K.bgemm(Pt, f, transA=True), lambda: Pt.double().transpose(1, 2) @ f.double()),
    "NN per-subject M=512": (lambda: K.bgemm(X512, Y[:3]), lambda: X512.double() @ Y[:3].double()),
}
for tma in (0, 1):
    pkg.set_tuning("tcgemm_tma", tma)
    for name, (fn, ref) in cases.items():
        out = fn(); torch.cuda.synchronize()
        r = ref()
        err = float((out.double() - r).abs().max()); scale = float(r.abs().max())
        print(json.dumps({"tma": tma, "case": name, "max_err": err, "rel": err / scale, "finite": bool(torch.isfinite(out).all())}), flush=True)
Br = 187
Xr = torch.randn(Br, 268, 268, device="cuda", generator=g); Yr = torch.randn(Br, 268, 268, device="cuda", generator=g)
Wadd = torch.randn(268, 268, device="cuda", generator=g)
Ur = torch.randn(Br, 160, 160, device="cuda", generator=g); gP = torch.randn(Br, 268, 160, device="cuda", generator=g)
rcases = {
    "reduce TN": (lambda: K.bgemm_reduce(Xr, Yr, transA=True), lambda: (Xr.double().transpose(1, 2) @ Yr.double()).sum(0)),
    "reduce NN + add": (lambda: K.bgemm_reduce(Xr, Yr, alpha=0.5, add=Wadd), lambda: 0.5 * (Xr.double() @ Yr.double()).sum(0) + Wadd.double()),
    "reduce NT": (lambda: K.bgemm_reduce(Xr, Yr, transB=True), lambda: (Xr.double() @ Yr.double().transpose(1, 2)).sum(0)),
    "reduce NN gP.U (N=160, K=160)": (lambda: K.bgemm_reduce(gP, Ur), lambda: (gP.double() @ Ur.double()).sum(0)),
    "reduce TN M=160": (lambda: K.bgemm_reduce(Xr[:, :, :160].contiguous(), Yr, transA=True), lambda: (Xr[:, :, :160].double().transpose(1, 2) @ Yr.double()).sum(0)),
}
for tma in (0, 1):
    pkg.set_tuning("tcgemm_tma", tma)
    for name, (fn, ref) in rcases.items():
        out = fn(); torch.cuda.synchronize()
        r = ref()
        err = float((out.double() - r).abs().max()); scale = float(r.abs().max())
        print(json.dumps({"tma": tma, "case": name, "max_err": err, "rel": err / scale}), flush=True)
if "--time" in sys.argv:
    B = 1024
    X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g)
    U = torch.rand(B, 160, 160, device="cuda", generator=g)
    fl = 2 * 268 ** 3 * B
    for tma in (0, 1):
        pkg.set_tuning("tcgemm_tma", tma)
        r = {"tma": tma, "unit": "[ms, TFLOP/s fp32-equivalent]"}
        for name, fn, flops in (("NN per-subject", lambda: K.bgemm(X, Y), fl), ("NT per-subject", lambda: K.bgemm(X, Y, transB=True), fl),
                                ("TN per-subject", lambda: K.bgemm(X, Y, transA=True), fl), ("NN stacked W", lambda: K.bgemm(X, W), fl),
                                ("NT stacked W", lambda: K.bgemm(X, W, transB=True), fl),
                                ("NT stacked U.W^T", lambda: K.bgemm(U, Wl, transB=True), 2 * 160 * 160 * 268 * B),
                                ("reduce TN", lambda: K.bgemm_reduce(X, Y, transA=True), fl), ("reduce NN", lambda: K.bgemm_reduce(X, Y), fl)):
            ms = timeit(fn); r[name] = [round(ms, 4), round(flops / ms / 1e9, 1)]
        print(json.dumps(r), flush=True)
pkg.set_tuning("tcgemm_tma", 1)
if "--sweep" in sys.argv:
    for pf in (0, 3, 6, 9, 12):
        pkg.set_tuning("tcgemm_prefetch", pf)
        r = {"prefetch_ahead": pf}
        for name, fn, flops in (("NN per-subject", lambda: K.bgemm(X, Y), fl), ("NT per-subject", lambda: K.bgemm(X, Y, transB=True), fl),
                                ("NN stacked W", lambda: K.bgemm(X, W), fl), ("NT stacked W", lambda: K.bgemm(X, W, transB=True), fl),
                                ("reduce TN", lambda: K.bgemm_reduce(X, Y, transA=True), fl)):
            ms = timeit(fn); r[name] = [round(ms, 4), round(flops / ms / 1e9, 1)]
        print(json.dumps(r), flush=True)
    pkg.set_tuning("tcgemm_prefetch", 0)
# MLP-sized products: N > 272 -> 256-column tiles in the TMA kernel
for tma in (0, 1):
    pkg.set_tuning("tcgemm_tma", tma)
    Xm = torch.randn(512, 1024, device="cuda", generator=g); Wm = torch.randn(35778, 1024, device="cuda", generator=g) * 0.03
    bm = torch.randn(35778, device="cuda", generator=g)
    X1 = torch.randn(512, 12720, device="cuda", generator=g); W1 = torch.randn(1024, 12720, device="cuda", generator=g) * 0.01
    for name, fn, ref, flops in (
            ("MLP out 512x1024 . (35778x1024)^T + bias", lambda: K.bgemm(Xm, Wm, transB=True, bias=bm), lambda: Xm.double() @ Wm.double().T + bm.double(), 2 * 512 * 1024 * 35778),
            ("MLP in 512x12720 . (1024x12720)^T", lambda: K.bgemm(X1, W1, transB=True), lambda: X1.double() @ W1.double().T, 2 * 512 * 12720 * 1024),
            ("MLP dX 512x35778 . 35778x1024", lambda: K.bgemm(K.bgemm(Xm, Wm, transB=True), Wm), None, 2 * 512 * 1024 * 35778),
            ("MLP dW (512x35778)^T . 512x1024", lambda: K.bgemm(K.bgemm(Xm, Wm, transB=True), Xm, transA=True), None, 2 * 512 * 1024 * 35778)):
        out = fn(); torch.cuda.synchronize()
        rec = {"tma": tma, "case": name}
        if ref is not None:
            r = ref(); rec["rel"] = float((out.double() - r).abs().max()) / float(r.abs().max())
        ms = timeit(fn, reps=5); rec["ms"] = round(ms, 4); rec["tflops"] = round(flops / ms / 1e9, 1)
        print(json.dumps(rec), flush=True)
pkg.set_tuning("tcgemm_tma", 1)
