"""CTA-pair tcgemm and the FP32 ragged-row kernels (gsr_set_tuning keys tcgemm_pairs / tcgemm_ragged_rows): accuracy
against float64 and time against the one-CTA kernel."""
import ctypes, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
lib = pkg.lib()
def dbg(key, value): pkg.set_tuning({1: "tcgemm_pairs", 2: "tcgemm_ragged_rows"}[key], value)
g = torch.Generator(device="cuda").manual_seed(0)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
Bs = 6
X = torch.randn(Bs, 268, 268, device="cuda", generator=g); Y = torch.randn(Bs, 268, 268, device="cuda", generator=g)
W = torch.randn(268, 268, device="cuda", generator=g); Wl = torch.randn(268, 160, device="cuda", generator=g)
U = torch.randn(Bs, 160, 160, device="cuda", generator=g); add = torch.randn(Bs, 268, 268, device="cuda", generator=g)
cases = {
    "NN per-subject": (lambda: K.bgemm(X, Y), lambda: X.double() @ Y.double()),
    "NT per-subject": (lambda: K.bgemm(X, Y, transB=True), lambda: X.double() @ Y.double().transpose(1, 2)),
    "TN per-subject": (lambda: K.bgemm(X, Y, transA=True), lambda: X.double().transpose(1, 2) @ Y.double()),
    "NN + add, alpha": (lambda: K.bgemm(X, Y, alpha=2.0, add=add), lambda: 2 * (X.double() @ Y.double()) + add.double()),
    "NN stacked W": (lambda: K.bgemm(X, W), lambda: X.double() @ W.double()),
    "NT stacked W": (lambda: K.bgemm(X, W, transB=True), lambda: X.double() @ W.double().T),
    "NT stacked U.W^T (K=160)": (lambda: K.bgemm(U, Wl, transB=True), lambda: U.double() @ Wl.double().T),
    "NT stacked N=160 (K=268)": (lambda: K.bgemm(X, Wl.T.contiguous(), transB=True), lambda: X.double() @ Wl.double()),
}
MODES = ((0, 0), (1, 0), (1, 1))          # (CTA pairs, FP32 kernel for the ragged rows)
for pairs, tail in MODES:
    dbg(1, pairs); dbg(2, tail)
    for name, (fn, ref) in cases.items():
        out = fn(); torch.cuda.synchronize()
        r = ref()
        err = float((out.double() - r).abs().max()); scale = float(r.abs().max())
        print(json.dumps({"pairs": pairs, "tail": tail, "case": name, "max_err": err, "rel": err / scale}), flush=True)
Br = 37
Xr = torch.randn(Br, 268, 268, device="cuda", generator=g); Yr = torch.randn(Br, 268, 268, device="cuda", generator=g)
Wadd = torch.randn(268, 268, device="cuda", generator=g)
rcases = {
    "reduce TN": (lambda: K.bgemm_reduce(Xr, Yr, transA=True), lambda: (Xr.double().transpose(1, 2) @ Yr.double()).sum(0)),
    "reduce NN + add": (lambda: K.bgemm_reduce(Xr, Yr, alpha=0.5, add=Wadd), lambda: 0.5 * (Xr.double() @ Yr.double()).sum(0) + Wadd.double()),
    "reduce NT": (lambda: K.bgemm_reduce(Xr, Yr, transB=True), lambda: (Xr.double() @ Yr.double().transpose(1, 2)).sum(0)),
    "reduce TN M=160": (lambda: K.bgemm_reduce(Xr[:, :, :160].contiguous(), Yr, transA=True), lambda: (Xr[:, :, :160].double().transpose(1, 2) @ Yr.double()).sum(0)),
}
for pairs, tail in MODES:
    dbg(1, pairs); dbg(2, tail)
    for name, (fn, ref) in rcases.items():
        out = fn(); torch.cuda.synchronize()
        r = ref()
        err = float((out.double() - r).abs().max()); scale = float(r.abs().max())
        print(json.dumps({"pairs": pairs, "tail": tail, "case": name, "max_err": err, "rel": err / scale}), flush=True)
B = 1024
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g)
fl = 2 * 268 ** 3 * B
for pairs, tail in MODES:
    dbg(1, pairs); dbg(2, tail)
    r = {"pairs": pairs, "tail": tail}
    for name, fn in (("NN per-subject", lambda: K.bgemm(X, Y)), ("NT per-subject", lambda: K.bgemm(X, Y, transB=True)),
                     ("TN per-subject", lambda: K.bgemm(X, Y, transA=True)), ("NN stacked W", lambda: K.bgemm(X, W)),
                     ("NT stacked W", lambda: K.bgemm(X, W, transB=True)),
                     ("reduce TN", lambda: K.bgemm_reduce(X, Y, transA=True)), ("reduce NN", lambda: K.bgemm_reduce(X, Y))):
        ms = timeit(fn); r[name] = [round(ms, 4), round(fl / ms / 1e9, 1)]
    print(json.dumps(r), flush=True)
dbg(1, 1); dbg(2, 1)
