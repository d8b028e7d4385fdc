"""K2 with lane = graph (tql_scalar_lanes_kernel) against the one-warp-per-graph kernel and a float64 decomposition:
ragged batches (32 k + r graphs), special matrices in one batch, timing of the whole decomposition at 1024 x 160
and 1024 x 268 with either K2."""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from oracle import gsr_oracle as O
from eigh_check import report, timed

def both(A):
    pkg.set_tuning("eigh_k2_lanes", 0); r0 = K.eigh_tridiag(A)
    pkg.set_tuning("eigh_k2_lanes", 2); r1 = K.eigh_tridiag(A)         # 2: lane = graph also for one launch per stage
    return r0, r1

for n in (1, 2, 3, 5, 17, 31, 64, 100, 159, 160, 200, 268):
    B = 41
    A = torch.stack([O.normalize_adj(O.synth_connectome(n, 50 + s)) if n > 1 else torch.tensor([[0.7 + s]]) for s in range(B)]).cuda()
    (w0, U0, i0), (w1, U1, i1) = both(A)
    report(f"lanes n={n}", A, w1, U1, i1)
    print(json.dumps({"n": n, "w_lanes_vs_warp": float((w0 - w1).abs().max()), "U_lanes_vs_warp": float((U0 - U1).abs().max()),
                      "info_equal": bool(torch.equal(i0 < 0, i1 < 0))}), flush=True)

# special matrices in ONE batch (lanes diverge as far as they can): zeros, identity multiples, ones, block diagonal, NaN, a tiny and a huge scale
n = 64
g = torch.Generator().manual_seed(3)
mats = []
for s in range(40):
    M = torch.randn(n, n, generator=g); M = (M + M.T) / 2
    if s == 1: M = torch.zeros(n, n)
    if s == 2: M = torch.eye(n) * 3.0
    if s == 3: M = torch.ones(n, n)
    if s == 4: M[:20, 20:] = 0; M[20:, :20] = 0                       # exact block diagonal
    if s == 5: M[3, 7] = float("nan")
    if s == 6: M = M * 1e-12
    if s == 7: M = M * 1e9
    if s == 8: M = torch.diag(torch.arange(n, dtype=torch.float32))     # already diagonal
    mats.append(M)
A = torch.stack(mats).cuda()
(w0, U0, i0), (w1, U1, i1) = both(A)
ok = [s for s in range(40) if s != 5]
Ad = A[ok].double(); As = torch.triu(Ad) + torch.triu(Ad, 1).transpose(1, 2)
w64 = torch.linalg.eigvalsh(As.cpu())
scale = As.abs().amax(dim=(1, 2)).clamp_min(1e-30).cpu()
print(json.dumps({"special": True, "info": i1.tolist()[:10], "nan_row_is_nan": bool(torch.isnan(w1[5]).all()),
                  "w_rel_err_max": float(((w1[ok].cpu().double() - w64).abs().amax(1) / scale).max()),
                  "orth": float((U1[ok].double().transpose(1, 2) @ U1[ok].double() - torch.eye(n, device="cuda", dtype=torch.float64)).abs().amax()),
                  "w_lanes_vs_warp_rel": float(((w0[ok] - w1[ok]).abs().amax(1).cpu() / scale).max())}), flush=True)

gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 160, 160, device="cuda", generator=gen); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
ref = None
for lanes, chunks, warps in ((0, 1, 2), (2, 1, 2), (1, 8, 4), (1, 8, 2)):
    pkg.set_tuning("eigh_k2_lanes", lanes); pkg.set_tuning("eigh_chunks", chunks); pkg.set_tuning("eigh_k2_warps", warps)
    (w, U, info), ms = timed(K.eigh_tridiag, A, reps=5); report(f"tridiag 1024x160 k2_lanes={lanes} chunks={chunks} k2_warps={warps}", A, w, U, info, ms)
    if lanes >= 1:
        if ref is None: ref = (w, U)
        else: print(json.dumps({"chunks": chunks, "bit_identical_to_unchunked": bool(torch.equal(w, ref[0]) and torch.equal(U, ref[1]))}), flush=True)
pkg.set_tuning("eigh_k2_lanes", 1); pkg.set_tuning("eigh_chunks", 8)
for Bq in (256, 300, 1000):           # ragged chunking
    (w, U, info), ms = timed(K.eigh_tridiag, A[:Bq]); report(f"tridiag {Bq}x160 chunks=8", A[:Bq], w, U, info, ms)
    print(json.dumps({"batch": Bq, "equal_to_full_batch_rows": bool(torch.equal(w, ref[0][:Bq]) and torch.equal(U, ref[1][:Bq]))}), flush=True)
x = torch.rand(1024, 268, 268, device="cuda", generator=gen); hr = (x + x.transpose(1, 2)) / 2
for lanes in (0, 2):
    pkg.set_tuning("eigh_k2_lanes", lanes)
    (w, U, info), ms = timed(K.eigh_tridiag, hr); report(f"tridiag 1024x268 k2_lanes={lanes}", hr, w, U, info, ms)
