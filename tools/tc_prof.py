import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
for kv in os.environ.get("GSR_TUNE", "").split(","):           # e.g. GSR_TUNE=eigh_chunks=1 under ncu (serialised kernels)
    if kv: pkg.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
K.set_gemm_backend(os.environ.get("GSR_BACKEND", "tc"))
g = torch.Generator(device="cuda").manual_seed(1)
B = int(os.environ.get("GSR_B", "1024"))
H = torch.rand(B, 268, 268, device="cuda", generator=g); H2 = torch.rand(B, 268, 268, device="cuda", generator=g)
W = torch.randn(268, 268, device="cuda", generator=g)
for _ in range(3):
    K.bgemm(H, W); K.bgemm(H, H2); K.bgemm(H, H2, transA=True)
torch.cuda.synchronize()
print("done")
