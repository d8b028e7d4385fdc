"""The memory-bound n x n kernels of a step once each at the BASELINE batch (for ncu): normalise forward / backward (with
the two-gradient backward the step uses), sym-diag-abs forward / backward, the fused sym-diag-abs + L1 pass."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
for kv in os.environ.get("GSR_TUNE", "").split(","):           # e.g. GSR_TUNE=normalize_bwd_cluster=1
    if kv: pkg.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
B = int(os.environ.get("GSR_B", "1024"))
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.rand(B, 268, 268, device="cuda", generator=g) + 0.01; g1 = torch.randn(B, 268, 268, device="cuda", generator=g); g2 = torch.randn(B, 268, 268, device="cuda", generator=g)
gl = torch.full((B,), 1.0 / B, device="cuda")
for _ in range(2):
    out, r, _ = K.normalize_adj_fwd(x, pre=True)
    K.normalize_adj_bwd(x, r, g1, pre=True, grad_out2=g2)
    z = K.sym_diag_abs_fwd(x)
    K.sym_diag_abs_bwd(x, g1)
    K.sym_diag_abs_l1_fwd(g1, x)                                # the fused final pass: sym-diag-abs + L1 against the target
    K.sym_diag_abs_l1_bwd(g1, x, gl)
torch.cuda.synchronize()
print("done")

# outside a profiler: device time of each call (CUDA events, 10 calls)
import json
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)
if not os.environ.get("GSR_NO_TIMES"):
    print(json.dumps({"batch": B, "us": {
        "normalize_fwd(pre)": t(lambda: K.normalize_adj_fwd(x, pre=True)),
        "normalize_bwd(pre)": t(lambda: K.normalize_adj_bwd(x, r, g1, pre=True)),
        "normalize_bwd(pre, two gradients)": t(lambda: K.normalize_adj_bwd(x, r, g1, pre=True, grad_out2=g2)),
        "sym_diag_abs_fwd": t(lambda: K.sym_diag_abs_fwd(x)),
        "sym_diag_abs_bwd": t(lambda: K.sym_diag_abs_bwd(x, g1)),
        "sym_diag_abs_l1_fwd (no prediction written)": t(lambda: K.sym_diag_abs_l1_fwd(g1, x)),
        "sym_diag_abs_l1_bwd": t(lambda: K.sym_diag_abs_l1_bwd(g1, x, gl)),
        "l1_loss_fwd + l1_loss_bwd (the two launches the fused pass replaces, next to sym_diag_abs fwd / bwd)":
            t(lambda: (K.l1_loss_fwd(x, g1, B, False), K.l1_loss_bwd(x, g1, gl, B, False)))}}))
