"""The memory-bound n x n kernels of a step once each at the BASELINE batch (for ncu): normalise forward / backward (with
the two-gradient backward the step uses), sym-diag-abs forward / backward, L1 partial / backward."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
B = int(os.environ.get("GSR_B", "1024"))
g = torch.Generator(device="cuda").manual_seed(2)
x = torch.rand(B, 268, 268, device="cuda", generator=g) + 0.01; g1 = torch.randn(B, 268, 268, device="cuda", generator=g); g2 = torch.randn(B, 268, 268, device="cuda", generator=g)
for _ in range(2):
    out, r, _ = K.normalize_adj_fwd(x, pre=True)
    K.normalize_adj_bwd(x, r, g1, pre=True, grad_out2=g2)
    z = K.sym_diag_abs_fwd(x)
    K.sym_diag_abs_bwd(x, g1)
torch.cuda.synchronize()
print("done")
