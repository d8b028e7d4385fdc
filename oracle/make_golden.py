"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Run in the build container (where /root/reference is
mounted):   python -m oracle.make_golden
The fixtures pin oracle/gsr_oracle.py and oracle/vectorizer_oracle.py to the
reference's real behaviour; they are small (strided samples and SHA-256 digests
stand in for the large tensors) and travel to the GPU box, the reference does
not.  Inputs / weights are regenerated from seeds by the same helper functions
(oracle.gsr_oracle.synth_connectome / make_state_dict) that the tests use.
"""
from __future__ import annotations

import hashlib
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

from . import gsr_oracle as O
from . import ref_shim

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GRAD_STRIDE = 13          # large gradient tensors are stored as flat[::GRAD_STRIDE]
GRAD_FULL_MAX = 4096


class Args:               # training_gsr.ipynb:51-59
    def __init__(self, lr_dim, hr_dim, hidden_dim, padding=0, lmbda=16, epochs=1):
        self.lr_dim, self.hr_dim, self.hidden_dim = lr_dim, hr_dim, hidden_dim
        self.padding, self.lmbda, self.epochs = padding, lmbda, epochs


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_grad(g: torch.Tensor) -> np.ndarray:
    a = g.detach().numpy().reshape(-1)
    return a.copy() if a.size <= GRAD_FULL_MAX else a[::GRAD_STRIDE].copy()


def reference_step(ref, model, lr, hr, args):
    """utils/gsr_train.py:24-43 driven around the reference model (the module
    itself cannot be imported: it loads CSVs at import time)."""
    crit = torch.nn.L1Loss()
    model.train()
    model.zero_grad()
    outs, net_outs, start_gcn_outs, layer_outs = model(lr)
    outs_u = ref.unpad(outs, args.padding)
    padded_hr = ref.pad_HR_adj(hr, args.padding)
    _, U_hr = torch.linalg.eigh(padded_hr, UPLO="U")
    U_hr_reduced = U_hr[:, :model.layer.lr_dim]
    loss = args.lmbda * crit(net_outs, start_gcn_outs) + crit(model.layer.weights, U_hr_reduced) + crit(outs_u, hr)
    error = crit(outs_u, hr)
    loss.backward()
    return loss, error, (outs, net_outs, start_gcn_outs, layer_outs), U_hr_reduced


def gsrnet_fixture(ref, name, lr_dim, hr_dim, n_subj, weight_seed, data_seed, row_stride):
    args = Args(lr_dim, hr_dim, hr_dim)
    sd = O.make_state_dict(weight_seed, lr_dim, hr_dim, hr_dim)
    model = ref.GSRNet(list(O.DEFAULT_KS), args)
    assert list(model.state_dict().keys()) == list(sd.keys()), "state_dict keys differ from the reference"
    model.load_state_dict(sd)
    out = OrderedDict(lr_dim=lr_dim, hr_dim=hr_dim, n_subj=n_subj, weight_seed=weight_seed,
                      data_seed=data_seed, row_stride=row_stride, grad_stride=GRAD_STRIDE,
                      keys=np.array(list(sd.keys())))
    acc = None
    # raw torch.topk tie behaviour is recorded once for the report in DESIGN.md
    with ref_shim.canonical_library_calls():
        for s in range(n_subj):
            lr = O.synth_connectome(lr_dim, data_seed + s)
            hr = O.synth_connectome(hr_dim, data_seed + 500 + s)
            loss, err, (z, net_outs, start, adj), U_hr_red = reference_step(ref, model, lr, hr, args)
            A = ref.normalize_adj_torch(lr)
            w, U = torch.linalg.eigh(A, UPLO="U")
            p = f"s{s}_"
            out[p + "loss"] = float(loss)
            out[p + "error"] = float(err)
            out[p + "out"] = z.detach().numpy()[::row_stride].copy()
            out[p + "adj"] = adj.detach().numpy()[::row_stride].copy()
            out[p + "out_sha"] = sha(z.detach().numpy())
            out[p + "Z"] = model.Z.detach().numpy()[::row_stride].copy()
            out[p + "f_d"] = model.layer.f_d.detach().numpy()[::row_stride].copy()
            out[p + "eigvals"] = w.numpy().copy()
            out[p + "U"] = U.numpy()[::row_stride].copy()
            out[p + "U_hr_reduced"] = U_hr_red.numpy()[::row_stride].copy()
            # the reference's FULL bases (LAPACK, sign rule applied): the own-eigenbasis GPU tests align the kernel's
            # eigenvectors to U_full column by column and feed U_hr_full to the |W - U_hr| term, so that every
            # subject is asserted whatever the sign-pivot margin of its eigenvectors
            out[p + "U_full"] = U.numpy().copy()
            out[p + "U_hr_full"] = U_hr_red.numpy().copy()
            if s == 0:
                out["net_outs"] = net_outs.detach().numpy().copy()
                out["start_gcn_outs"] = start.detach().numpy().copy()
            grads = OrderedDict((k, v.grad.detach().clone()) for k, v in model.named_parameters())
            if s == 0:
                for k, g in grads.items():
                    out["s0_grad/" + k] = sample_grad(g)
            acc = grads if acc is None else OrderedDict((k, acc[k] + grads[k]) for k in acc)
        for k, g in acc.items():
            out["mean_grad/" + k] = sample_grad(g / n_subj)
        # pooling trace of the (subject independent) U-Net
        A0 = ref.normalize_adj_torch(O.synth_connectome(lr_dim, data_seed))
        X = model.net.start_gcn(A0, torch.eye(lr_dim))
        Acur = A0
        for i in range(4):
            X = model.net.down_gcns[i](Acur, X)
            scores = torch.sigmoid(torch.squeeze(model.net.pools[i].proj(X)) / 100)
            Acur, X, idx = model.net.pools[i](Acur, X)
            out[f"pool{i}_scores"] = scores.detach().numpy().copy()
            out[f"pool{i}_idx"] = idx.numpy().copy()
            if i == 0:
                out["pool0_A_gathered_sha"] = sha(Acur.detach().numpy())
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, "written:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:8]})


def ops_fixture(ref):
    out = OrderedDict()
    g = torch.Generator().manual_seed(7)
    # normalize_adj_torch: symmetric, non-symmetric, a zero row (inf -> 0), a negative row sum (NaN)
    for tag, n in (("sym12", 12), ("sym160", 160)):
        m = O.synth_connectome(n, 11 + n)
        out[f"norm_{tag}_in"] = m.numpy()
        out[f"norm_{tag}_out"] = ref.normalize_adj_torch(m.clone()).numpy()
    m = torch.rand(9, 9, generator=g)
    out["norm_nonsym9_in"] = m.numpy().copy()
    out["norm_nonsym9_out"] = ref.normalize_adj_torch(m.clone()).numpy()
    m = torch.rand(6, 6, generator=g); m[2] = 0
    out["norm_zerorow6_in"] = m.numpy().copy()
    out["norm_zerorow6_out"] = ref.normalize_adj_torch(m.clone()).numpy()
    m = torch.rand(5, 5, generator=g); m[1] = -m[1]
    out["norm_negrow5_in"] = m.numpy().copy()
    out["norm_negrow5_out"] = ref.normalize_adj_torch(m.clone()).numpy()
    # normalize backward (autograd of the reference function) on a non-symmetric input
    m = torch.rand(10, 10, generator=g).requires_grad_(True)
    T = torch.randn(10, 10, generator=g)
    (ref.normalize_adj_torch(m) * T).sum().backward()
    out["normbwd10_in"], out["normbwd10_T"], out["normbwd10_grad"] = m.detach().numpy().copy(), T.numpy().copy(), m.grad.numpy().copy()
    # pad_HR_adj / unpad
    h = O.synth_connectome(8, 3)
    out["pad_in"] = h.numpy().copy()
    out["pad2_out"] = ref.pad_HR_adj(h.clone(), 2).numpy()
    out["pad0_out"] = ref.pad_HR_adj(h.clone(), 0).numpy()
    out["unpad2_out"] = ref.unpad(torch.from_numpy(out["pad2_out"]), 2).numpy().copy()
    # GraphConvolution / GraphUnpool / GraphPool on small shapes
    gc = ref.GraphConvolution(7, 5, 0.0)
    x, adj = torch.randn(6, 7, generator=g), torch.rand(6, 6, generator=g)
    out["gc_w"], out["gc_x"], out["gc_adj"] = gc.weight.detach().numpy().copy(), x.numpy().copy(), adj.numpy().copy()
    out["gc_out"] = gc(x, adj).detach().numpy().copy()
    unp = ref.GraphUnpool()
    Xs, idx = torch.randn(3, 4, generator=g), torch.tensor([4, 0, 2])
    out["unpool_X"], out["unpool_idx"] = Xs.numpy().copy(), idx.numpy().copy()
    out["unpool_out"] = unp(torch.zeros(6, 6), Xs, idx)[1].numpy().copy()
    pool = ref.GraphPool(0.5, 4)
    Xp, Ap = torch.randn(9, 4, generator=g), torch.rand(9, 9, generator=g)
    with ref_shim.canonical_library_calls():
        A2, X2, idx2 = pool(Ap, Xp)
    out["pool_w"], out["pool_b"] = pool.proj.weight.detach().numpy().copy(), pool.proj.bias.detach().numpy().copy()
    out["pool_X"], out["pool_A"] = Xp.numpy().copy(), Ap.numpy().copy()
    out["pool_A_out"], out["pool_X_out"], out["pool_idx"] = A2.detach().numpy().copy(), X2.detach().numpy().copy(), idx2.numpy().copy()
    # ties: identical scores must come out in index order under the stated rule
    with ref_shim.canonical_library_calls():
        _, tidx = torch.topk(torch.full((160,), 0.5), 144)
    out["tie_idx_stable"] = tidx.numpy().copy()
    out["tie_idx_raw_torch"] = torch.topk(torch.full((160,), 0.5), 144).indices.numpy().copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ops.npz"), **out)
    print("ops written")


def vectorizer_fixture(ref):
    MV = ref.MatrixVectorizer
    out = OrderedDict()
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5, 8):
        for dt in (np.float64, np.float32, np.int64):
            m = (rng.random((n, n)) * 100).astype(dt)
            for diag in (False, True):
                out[f"vec_n{n}_{np.dtype(dt).name}_d{int(diag)}_in"] = m
                out[f"vec_n{n}_{np.dtype(dt).name}_d{int(diag)}_out"] = MV.vectorize(m, include_diagonal=diag)
    for n in (160, 268):
        m = rng.random((n, n))
        for diag in (False, True):
            v = MV.vectorize(m, include_diagonal=diag)
            out[f"vec_n{n}_d{int(diag)}_seed"] = 100 + n
            out[f"vec_n{n}_d{int(diag)}_len"] = v.shape[0]
            out[f"vec_n{n}_d{int(diag)}_sha"] = sha(v)
            out[f"vec_n{n}_d{int(diag)}_in_sha"] = sha(m)
        out[f"vec_n{n}_in"] = m.astype(np.float32)          # fp32 copy, digest below is of the fp32 result
        out[f"vec_n{n}_f32_d0_sha"] = sha(MV.vectorize(m.astype(np.float32)))
        out[f"vec_n{n}_f32_d1_sha"] = sha(MV.vectorize(m.astype(np.float32), include_diagonal=True))
    for n in (1, 2, 3, 5, 8, 160, 268):
        n_el = n * (n - 1) // 2
        v = torch.from_numpy(rng.random(n_el + n + 3)).to(torch.float32)
        out[f"anti_n{n}_in"] = v.numpy()
        for diag in (False, True):
            r = MV.anti_vectorize(v, n, include_diagonal=diag)
            assert r.dtype == torch.float32
            key = f"anti_n{n}_d{int(diag)}"
            if n <= 8:
                out[key + "_out"] = r.numpy()
            else:
                out[key + "_sha"] = sha(r.numpy())
    # dtype behaviour: a float64 vector is rejected by index_put (dtype mismatch)
    try:
        MV.anti_vectorize(torch.zeros(10, dtype=torch.float64), 5)
        out["anti_f64_raises"] = "none"
    except Exception as e:                                   # noqa
        out["anti_f64_raises"] = type(e).__name__
    # error behaviour: short vector asserts; numpy input raises TypeError
    try:
        MV.anti_vectorize(torch.zeros(5), 5)
        out["anti_short_raises"] = "none"
    except AssertionError:
        out["anti_short_raises"] = "AssertionError"
    try:
        MV.anti_vectorize(np.zeros(10), 5)
        out["anti_numpy_raises"] = "none"
    except Exception as e:                                   # noqa
        out["anti_numpy_raises"] = type(e).__name__
    np.savez_compressed(os.path.join(GOLDEN_DIR, "vectorizer.npz"), **out)
    print("vectorizer written")


def eval_fixture(ref):
    """utils/evaluation.py:96-112 -- the three reductions, through the same
    sklearn / scipy calls on the reference vectoriser's output."""
    from sklearn.metrics import mean_absolute_error
    from scipy.stats import pearsonr
    from scipy.spatial.distance import jensenshannon
    MV = ref.MatrixVectorizer
    out = OrderedDict()
    for tag, N, n in (("small", 6, 12), ("hr", 3, 268)):
        gts, preds = [], []
        for i in range(N):
            gt = O.synth_connectome(n, 900 + i)
            g = torch.Generator().manual_seed(1900 + i)
            noise = torch.randn(n, n, generator=g) * 0.1
            pred = torch.clamp(gt + (noise + noise.t()) / 2, 0, 1)
            pred.fill_diagonal_(0)
            gts.append(gt.numpy()); preds.append(pred.numpy())
        p1 = np.concatenate([MV.vectorize(m) for m in preds])
        g1 = np.concatenate([MV.vectorize(m) for m in gts])
        out[f"{tag}_N"], out[f"{tag}_n"] = N, n
        out[f"{tag}_mae"] = float(mean_absolute_error(g1, p1))
        out[f"{tag}_pcc"] = float(pearsonr(g1, p1)[0])
        out[f"{tag}_js"] = float(jensenshannon(g1, p1))
        out[f"{tag}_mae64"] = float(mean_absolute_error(g1.astype(np.float64), p1.astype(np.float64)))
        out[f"{tag}_pcc64"] = float(pearsonr(g1.astype(np.float64), p1.astype(np.float64))[0])
        out[f"{tag}_js64"] = float(jensenshannon(g1.astype(np.float64), p1.astype(np.float64)))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "eval.npz"), **out)
    print("eval written")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = ref_shim.load()
    vectorizer_fixture(ref)
    ops_fixture(ref)
    eval_fixture(ref)
    gsrnet_fixture(ref, "gsrnet_small", lr_dim=136, hr_dim=268, n_subj=4, weight_seed=3, data_seed=40, row_stride=8)
    gsrnet_fixture(ref, "gsrnet_full", lr_dim=160, hr_dim=268, n_subj=8, weight_seed=0, data_seed=1000, row_stride=8)


if __name__ == "__main__":
    sys.exit(main())
