import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
K.set_gemm_backend(os.environ.get("GSR_BACKEND", "tc"))
g = torch.Generator(device="cuda").manual_seed(1)
B = int(os.environ.get("GSR_B", "1024"))
H = torch.rand(B, 268, 268, device="cuda", generator=g); H2 = torch.rand(B, 268, 268, device="cuda", generator=g)
W = torch.randn(268, 268, device="cuda", generator=g)
for _ in range(3):
    K.bgemm(H, W); K.bgemm(H, H2); K.bgemm(H, H2, transA=True)
torch.cuda.synchronize()
print("done")
