"""Per-subject 268^3 products with and without the addend epilogue (B = 1024): device time per launch (CUDA events,
20 launches after 3 warm-ups).  The two addend products of the backward pass (functional.py: d adj) are the case."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
K.set_gemm_backend("tc")
g = torch.Generator(device="cuda").manual_seed(1)
B = int(os.environ.get("GSR_B", "1024"))
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g)
add = torch.randn(B, 268, 268, device="cuda", generator=g)
cases = {
    "NN": lambda: K.bgemm(X, Y),
    "NN + add, alpha": lambda: K.bgemm(X, Y, alpha=2.0, add=add),
    "NT": lambda: K.bgemm(X, Y, transB=True),
    "NT + add": lambda: K.bgemm(X, Y, transB=True, add=add),
}
for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(json.dumps({"case": name, "ms_per_launch_incl_ragged_rows": round(ms, 4), "tflops": round(2 * B * 268 ** 3 / ms / 1e9, 1)}))
