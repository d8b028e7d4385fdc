import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
for kv in os.environ.get("GSR_TUNE", "").split(","):           # e.g. GSR_TUNE=eigh_chunks=1 under ncu (serialised kernels)
    if kv: pkg.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 160, 160, device="cuda", generator=gen); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
for _ in range(2): K.eigh_tridiag(A)
torch.cuda.synchronize()

if os.environ.get("GSR_PROF_HR"):
    x = torch.rand(1024, 268, 268, device="cuda", generator=gen); hr = (x + x.transpose(1, 2)) / 2
    hr.diagonal(dim1=1, dim2=2).fill_(1.0)
    for _ in range(2): K.eigh_tridiag(hr)
    torch.cuda.synchronize()
