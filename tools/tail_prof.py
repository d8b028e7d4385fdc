"""The FP32 kernel of the ragged rows (csrc/tcgemm.cu, tail_kernel) at the BASELINE batch, one launch per operand
layout inside a cudaProfilerStart/Stop window:  ncu --profile-from-start off -k regex:tail_kernel ..."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
g = torch.Generator(device="cuda").manual_seed(0)
B = int(os.environ.get("GSR_B", "1024"))
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g)
def all_layouts():
    K.bgemm(X, Y); K.bgemm(X, Y, transB=True); K.bgemm(X, Y, transA=True); K.bgemm_reduce(X, Y, transA=True); K.bgemm_reduce(X, Y)
all_layouts(); torch.cuda.synchronize()
torch.cuda.profiler.start(); all_layouts(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("profiled")
