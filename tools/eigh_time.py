import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from oracle import gsr_oracle as O
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 160, 160, device="cuda", generator=g); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
w64, U64 = torch.linalg.eigh(A[:8].double().cpu(), UPLO="U"); U64 = O.canonicalize_signs(U64)
w64a, U64a = torch.linalg.eigh(A[:64].double().cpu(), UPLO="U"); U64a = O.canonicalize_signs(U64a)
for b2 in (1, 2, 6):
    pkg.set_tuning("eigh_max_polish", b2)
    for _ in range(2): K.eigh_jacobi(A)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record(); w, U, info = K.eigh_jacobi(A); b.record(); torch.cuda.synchronize()
    Ad, Ud, wd = A.double(), U.double(), w.double()
    As = torch.triu(Ad) + torch.triu(Ad, 1).transpose(1, 2)
    res = float((As @ Ud - Ud * wd[:, None, :]).abs().amax()); orth = float((Ud.transpose(1, 2) @ Ud - torch.eye(160, device="cuda", dtype=torch.float64)).abs().amax())
    errs = sorted(float((U[s].cpu().double() - U64a[s]).abs().max()) for s in range(64)); uerr = errs[-1]; umed = errs[32]
    print(json.dumps({"max_polish": b2, "ms": round(a.elapsed_time(b), 3), "us_per_graph": round(a.elapsed_time(b) * 1e3 / 1024, 2),
                      "sweeps_mean": float((info & 0xFF).float().mean()), "polish_mean": float((info >> 8).float().mean()), "residual": res, "orth": orth, "U_err_vs_f64_max": uerr, "U_err_median": umed}), flush=True)
