"""Accuracy + timing of the 268-node HR-target decomposition: tridiagonal route vs the Jacobi kernel."""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from oracle import gsr_oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eigh_check import report, timed
for n in (161, 200, 225, 268, 288):
    A = torch.stack([O.normalize_adj(O.synth_connectome(n, 70 + s)) for s in range(3)]).cuda()
    w, U, info = K.eigh_tridiag(A); report("tridiag", A, w, U, info)
hrs = torch.stack([O.pad_hr_adj(O.synth_connectome(268, 1500 + s), 0) for s in range(4)]).cuda()
w, U, info = K.eigh_tridiag(hrs); report("tridiag HR targets", hrs, w, U, info)
B = int(os.environ.get("GSR_B", "1024"))
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(B, 268, 268, device="cuda", generator=gen); hr = (x + x.transpose(1, 2)) / 2
hr.diagonal(dim1=1, dim2=2).fill_(1.0)
(w, U, info), ms = timed(K.eigh_tridiag, hr, reps=2); report(f"tridiag {B}x268", hr, w, U, info, ms)
(w, U, info), ms = timed(K.eigh_jacobi, hr, reps=1); report(f"jacobi {B}x268", hr, w, U, info, ms)
