"""Per-subject 268-row products at the BASELINE batch (CTA-pair tensor kernel for rows 0..255 + FP32 ragged rows 256..267):
device time per call, with the ragged rows on the FP32 kernels and on the one-CTA tensor kernel.  One JSON line."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
g = torch.Generator(device="cuda").manual_seed(0)
B = int(os.environ.get("GSR_B", "1024"))
X = torch.rand(B, 268, 268, device="cuda", generator=g) - 0.5; Y = torch.rand(B, 268, 268, device="cuda", generator=g) - 0.5
P = torch.rand(B, 268, 160, device="cuda", generator=g) - 0.5; Q = torch.rand(B, 160, 268, device="cuda", generator=g) - 0.5


def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)


calls = {"NN 268x268x268": lambda: K.bgemm(X, Y), "TN 268x268x268": lambda: K.bgemm(X, Y, transA=True),
         "NT 268x268x268 (tail_bt)": lambda: K.bgemm(X, Y, transB=True), "NN 268x268x160": lambda: K.bgemm(P, Q),
         "TN 268x268x160": lambda: K.bgemm(Q, Q, transA=True), "NN 268x160x268": lambda: K.bgemm(X, P)}
res = {"FP32 ragged rows": {k: t(f) for k, f in calls.items()}}
ref = {k: f()[:, 256:].double() for k, f in calls.items()}
pkg.set_tuning("tcgemm_ragged_rows", 0)
res["one-CTA tensor kernel for the ragged rows"] = {k: t(f) for k, f in calls.items()}
res["max |FP32 rows - tensor rows| / max |C|"] = {k: float((f()[:, 256:].double() - ref[k]).abs().max() / ref[k].abs().max()) for k, f in calls.items()}
pkg.set_tuning("tcgemm_ragged_rows", 1)
print(json.dumps(res))
