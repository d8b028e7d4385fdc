import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from oracle import gsr_oracle as O
n, B = 160, 8
A = torch.stack([O.normalize_adj(O.synth_connectome(n, 50 + s)) for s in range(B)]).cuda()
w, U, info, ws = K.eigh_tridiag(A, _return_workspace=True)
torch.cuda.synchronize()
cap = ((2 * n * n + 63) // 64) * 64 + 64
al = lambda x: (x + 255) // 256 * 256
o_qt = al(B * cap * 8)
o_d = o_qt + al(B * n * n * 4); o_e = o_d + al(B * n * 4)
Qt = ws[o_qt:o_qt + B * n * n * 4].view(torch.float32).view(B, n, n).double()
I = torch.eye(n, device="cuda", dtype=torch.float64)
print("Q orth", float((Qt @ Qt.transpose(1, 2) - I).abs().amax()))
e = ws[o_e:o_e + B * n * 4].view(torch.float32).view(B, n).double()
# T from A: Q^T A Q should be tridiagonal; check vs float64 reduction quality: eigenvalues of T (we lost d: overwritten) -> check Q^T A Q off-band
As = torch.triu(A.double()) + torch.triu(A.double(), 1).transpose(1, 2)
T = Qt @ As @ Qt.transpose(1, 2)
band = torch.tril(torch.triu(torch.ones(n, n, device="cuda", dtype=torch.float64), -1), 1)
print("off-band max", float((T * (1 - band)).abs().amax()), "e vs T subdiag", float((T.diagonal(1, 1, 2).abs() - e[:, :n - 1].abs()).abs().amax()))
Ud = U.double()
print("U orth", float((Ud.transpose(1, 2) @ Ud - I).abs().amax()))
# Z_T = Q^T U should be orthogonal too; measure the rotation part alone
ZT = Qt @ Ud
print("ZT orth (rotations only, approx)", float((ZT.transpose(1, 2) @ ZT - I).abs().amax()))
o_sc = o_e + al(B * n * 4); o_cnt = o_sc + al(B * 4)
cnt = ws[o_cnt:o_cnt + B * 16].view(torch.int32).view(B, 4)
print(cnt.tolist()[:2])
st = ws[0:B * cap * 8].view(torch.float32).view(B, cap, 2).double()
nrot = int(cnt[0, 1])
cs = st[0, :nrot]
nrm = (cs ** 2).sum(1) - 1
print("c2+s2-1: mean", float(nrm.mean()), "std", float(nrm.std()), "max", float(nrm.abs().max()))
G = (Ud.transpose(1, 2) @ Ud - I)[0]
print("diag max", float(G.diagonal().abs().max()), "offdiag max", float((G - torch.diag(G.diagonal())).abs().max()))
dg = G.diagonal()
print("diag by column index (sorted eigen order) first/last 5:", dg[:5].tolist(), dg[-5:].tolist(), "mean", float(dg.mean()))
# replay the stream in float64 and float32 on the host
import numpy as np
hdr_o = o_cnt + al(B * 16)
hcap = 30 * n + 2
hdr = ws[hdr_o:hdr_o + B * hcap * 4].view(torch.int32).view(B, hcap)[0].cpu().numpy()
csn = st[0, :nrot].cpu().numpy()
for dt in (np.float64, np.float32):
    Z = Qt[0].cpu().numpy().T.astype(dt).copy()      # Z = Q, rows = rows of Q
    pos = 0
    for it in range(int(cnt[0, 0])):
        l = hdr[it] & 0xffff; m = hdr[it] >> 16
        for i in range(m - 1, l - 1, -1):
            c = dt(csn[pos, 0]); s = dt(csn[pos, 1]); pos += 1
            h = Z[:, i + 1].copy()
            Z[:, i + 1] = s * Z[:, i] + c * h
            Z[:, i] = c * Z[:, i] - s * h
    Zd = Z.astype(np.float64)
    print(dt.__name__, "host replay orth", np.abs(Zd.T @ Zd - np.eye(n)).max())
