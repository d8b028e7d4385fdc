"""The Linear products of the notebook MLP (BASELINE configs[3]: 12720 -> 1024 x 2 -> 35778, batch 512), one by one:
device time and TFLOP/s of forward (X W^T + b), dX (gY W) and dW (gY^T X).  One JSON line."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
B = int(os.environ.get("GSR_B", "512"))
g = torch.Generator(device="cuda").manual_seed(0)


def t(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
for fin, fout in ((12720, 1024), (1024, 1024), (1024, 35778)):
    X = torch.randn(B, fin, device="cuda", generator=g); W = torch.randn(fout, fin, device="cuda", generator=g) * 0.01
    b = torch.randn(fout, device="cuda", generator=g); gY = torch.randn(B, fout, device="cuda", generator=g)
    fl = 2.0 * B * fin * fout
    for name, fn in (("fwd", lambda: K.bgemm(X, W, transB=True, bias=b)), ("dX", lambda: K.bgemm(gY, W)),
                     ("dW", lambda: K.bgemm(gY, X, transA=True))):
        ms = t(fn)
        res[f"{fin}->{fout} {name}"] = {"ms": round(ms, 3), "tflops": round(fl / ms / 1e9, 1)}
print(json.dumps(res))
