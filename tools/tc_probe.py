import importlib, os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
dbg = pkg.lib().gsr_tcgemm_debug; dbg.restype, dbg.argtypes = ctypes.c_int, [ctypes.c_int, ctypes.c_int]
K.set_gemm_backend("tc")
M, N, Kd = 128, 288, 32
A = torch.zeros(M, Kd, device="cuda"); B = torch.zeros(Kd, N, device="cuda")     # A K-major, B MN-major
dbg(0, 1); lo = K.bgemm(A, B).cpu()
dbg(0, 2); hi = K.bgemm(A, B).cpu()
dbg(0, 0)
idx = (hi * 1024 + lo).long()          # idx[m][n] = word index in the B tile read as B(n, k = m % 8)
torch.set_printoptions(linewidth=250)
for k in range(8):
    print("k", k, "n=0..39:", idx[k, :40].tolist())
print("n=0, k=0..7:", idx[:8, 0].tolist())
print("n=32..35 k=0..3:", idx[:4, 32:36].tolist())
print("n=160..163 k=0..1:", idx[:2, 160:164].tolist())
print("rows 8..15 equal rows 0..7:", bool((idx[8:16] == idx[:8]).all()))
