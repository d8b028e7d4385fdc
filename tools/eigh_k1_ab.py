"""K1 (Householder tridiagonalisation, thread = matrix row) with 6 .. 10 of the ten 16-column blocks of a row in
registers (6, 7: 3 CTAs per SM at 128 registers) (gsr_set_tuning("eigh_k1_regblocks", k); the rest in shared memory): results of the variants against each
other (the arithmetic is the same, so bit identity is expected), against float64, and the time of the whole
decomposition of 1024 graphs of 160 nodes with one launch per stage and as the chunked pipeline.
Since the fourth session of round 2 the Q launch works in COLUMN form (csrc/eigh_tridiag.cu, form_q_col): q_launch = 0 (K3 forms
Q in row form) and q_launch = 1 agree up to the rounding of Q (2e-5 on the bases, asserted by
test_eigh_pipeline_matches_one_launch_per_stage), so the `q_launch_bit_identical` lines below print false; the K1 register-block
variants stay bit identical to each other.
"""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from oracle import gsr_oracle as O
from eigh_check import report, timed

for n in (1, 2, 17, 100, 127, 128, 129, 143, 144, 145, 159, 160):
    A = torch.stack([O.normalize_adj(O.synth_connectome(n, 50 + s)) if n > 1 else torch.tensor([[0.7 + s]]) for s in range(9)]).cuda()
    if n >= 17:
        A[3] = A[3] * 1e-12; A[4] = A[4] * 1e9; A[5, 1, n - 1] = float("nan")      # rescaled rows, one NaN in the shared-memory columns
    res = {}
    for rb in (10, 9, 8, 7, 6):
        pkg.set_tuning("eigh_k1_regblocks", rb)
        res[rb] = K.eigh_tridiag(A)
    ok = [s for s in range(9) if not (n >= 17 and s == 5)]
    same = all(torch.equal(res[10][0][ok], res[rb][0][ok]) and torch.equal(res[10][1][ok], res[rb][1][ok]) and torch.equal(res[10][2], res[rb][2]) for rb in (9, 8, 7, 6))
    print(json.dumps({"n": n, "regblocks_9_8_7_6_bit_identical_to_10": bool(same), "info": res[8][2].tolist()}), flush=True)
    keep = [s for s in ok if not (n >= 17 and s in (3, 4))]
    report(f"k1 regblocks=8 n={n}", A[keep], res[8][0][keep], res[8][1][keep], res[8][2][keep])

gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 160, 160, device="cuda", generator=gen); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
for chunks in (1, 8):
    ref = None
    for q_launch in ((0,) if chunks == 1 else (0, 1)):
        for rb in (10, 9, 8, 7, 6):
            pkg.set_tuning("eigh_k1_regblocks", rb); pkg.set_tuning("eigh_chunks", chunks); pkg.set_tuning("eigh_q_launch", q_launch)
            (w, U, info), ms = timed(K.eigh_tridiag, A, reps=5)
            report(f"tridiag 1024x160 k1_regblocks={rb} q_launch={q_launch} chunks={chunks}", A, w, U, info, ms)
            if ref is None: ref = (w, U)
            else: print(json.dumps({"regblocks": rb, "q_launch": q_launch, "chunks": chunks, "bit_identical_to_first_of_this_chunking": bool(torch.equal(w, ref[0]) and torch.equal(U, ref[1]))}), flush=True)
# ragged batches through the pipeline with Q formed by its own launch: 32 k + r graphs, a NaN graph in the last chunk
pkg.set_tuning("eigh_k1_regblocks", 8); pkg.set_tuning("eigh_chunks", 8)
for B in (384, 400, 777, 1025):
    Ab = A[:B].clone() if B <= 1024 else torch.cat([A, A[:B - 1024]])
    Ab[B - 3, 0, 5] = float("nan")
    out = {}
    for q_launch in (0, 1):
        pkg.set_tuning("eigh_q_launch", q_launch)
        out[q_launch] = K.eigh_tridiag(Ab)
    ok = [s for s in range(B) if s != B - 3]
    print(json.dumps({"batch": B, "q_launch_bit_identical": bool(torch.equal(out[0][0][ok], out[1][0][ok]) and torch.equal(out[0][1][ok], out[1][1][ok]) and torch.equal(out[0][2], out[1][2])),
                      "nan_graph_info": int(out[1][2][B - 3])}), flush=True)
pkg.set_tuning("eigh_q_launch", 1)
