"""Pin the CPU oracle (oracle/*.py, a restatement) to the UNMODIFIED reference
through the fixtures that oracle/make_golden.py wrote from it.  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import gsr_oracle as O
from oracle import vectorizer_oracle as VO


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sample_grad(g, stride):
    a = g.detach().numpy().reshape(-1)
    return a if a.size <= 4096 else a[::stride]


# ---------------------------------------------------------------- vectoriser
def test_vectorize_small_bit_exact(golden):
    G = golden("vectorizer")
    n_cases = 0
    for key in G.files:
        if key.startswith("vec_") and key.endswith("_in") and "_d" in key:
            out = G[key[:-3] + "_out"]
            diag = key.split("_d")[1][0] == "1"
            got = VO.vectorize(G[key], include_diagonal=diag)
            assert got.dtype == out.dtype and got.shape == out.shape, key
            assert np.array_equal(got, out), key
            n_cases += 1
    assert n_cases == 30


@pytest.mark.parametrize("n", [160, 268])
def test_vectorize_large_digest(golden, n):
    G = golden("vectorizer")
    m = G[f"vec_n{n}_in"]
    assert sha(VO.vectorize(m)) == str(G[f"vec_n{n}_f32_d0_sha"])
    assert sha(VO.vectorize(m, include_diagonal=True)) == str(G[f"vec_n{n}_f32_d1_sha"])
    assert int(G[f"vec_n{n}_d0_len"]) == n * (n - 1) // 2
    assert int(G[f"vec_n{n}_d1_len"]) == n * (n - 1) // 2 + n - 1


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 160, 268])
def test_anti_vectorize(golden, n):
    G = golden("vectorizer")
    v = G[f"anti_n{n}_in"]
    for diag in (False, True):
        got = VO.anti_vectorize(v, n, include_diagonal=diag)
        assert got.dtype == np.float32
        if n <= 8:
            assert np.array_equal(got, G[f"anti_n{n}_d{int(diag)}_out"])
        else:
            assert sha(got) == str(G[f"anti_n{n}_d{int(diag)}_sha"])


def test_vectorizer_error_behaviour(golden):
    G = golden("vectorizer")
    assert str(G["anti_short_raises"]) == "AssertionError"
    with pytest.raises(AssertionError):
        VO.anti_vectorize(np.zeros(5, np.float32), 5)
    # recorded for INTEGRATION.md: numpy / float64 inputs are rejected by the reference
    assert str(G["anti_numpy_raises"]) == "TypeError"
    assert str(G["anti_f64_raises"]) == "RuntimeError"


def test_vectorize_antivectorize_are_not_inverses():
    m = VO.anti_vectorize(np.arange(10, dtype=np.float32), 5)
    assert not np.array_equal(VO.vectorize(m), np.arange(10, dtype=np.float32))


# ---------------------------------------------------------------- small ops
@pytest.mark.parametrize("tag", ["sym12", "sym160", "nonsym9", "zerorow6", "negrow5"])
def test_normalize_adj(golden, tag):
    G = golden("ops")
    got = O.normalize_adj(torch.from_numpy(G[f"norm_{tag}_in"]).clone()).numpy()
    assert np.array_equal(got, G[f"norm_{tag}_out"], equal_nan=True)
    if tag == "negrow5":
        assert np.isnan(got).any()
    if tag == "zerorow6":
        assert np.all(got[2] == 0) and np.all(got[:, 2] == 0)


def test_normalize_adj_backward(golden):
    G = golden("ops")
    m = torch.from_numpy(G["normbwd10_in"]).clone().requires_grad_(True)
    (O.normalize_adj(m) * torch.from_numpy(G["normbwd10_T"])).sum().backward()
    assert np.allclose(m.grad.numpy(), G["normbwd10_grad"], rtol=0, atol=1e-6)


def test_pad_unpad(golden):
    G = golden("ops")
    h = torch.from_numpy(G["pad_in"])
    assert np.array_equal(O.pad_hr_adj(h, 2).numpy(), G["pad2_out"])
    assert np.array_equal(O.pad_hr_adj(h, 0).numpy(), G["pad0_out"])
    assert np.array_equal(O.unpad(torch.from_numpy(G["pad2_out"]), 2).numpy(), G["unpad2_out"])
    assert np.array_equal(h.numpy(), G["pad_in"])          # input not mutated


def test_gc_pool_unpool(golden):
    G = golden("ops")
    got = O.graph_convolution(torch.from_numpy(G["gc_w"]), torch.from_numpy(G["gc_x"]), torch.from_numpy(G["gc_adj"]))
    assert np.array_equal(got.numpy(), G["gc_out"])
    up = O.graph_unpool(torch.zeros(6, 6), torch.from_numpy(G["unpool_X"]), torch.from_numpy(G["unpool_idx"]))
    assert np.array_equal(up.numpy(), G["unpool_out"])
    sd = {"p.proj.weight": torch.from_numpy(G["pool_w"]), "p.proj.bias": torch.from_numpy(G["pool_b"])}
    A2, X2, idx, _ = O.graph_pool(sd, "p", 0.5, torch.from_numpy(G["pool_A"]), torch.from_numpy(G["pool_X"]))
    assert np.array_equal(idx.numpy(), G["pool_idx"])
    assert np.array_equal(A2.numpy(), G["pool_A_out"])
    assert np.array_equal(X2.numpy(), G["pool_X_out"])


def test_topk_tie_rule(golden):
    G = golden("ops")
    _, idx = O.stable_topk(torch.full((160,), 0.5), 144)
    assert np.array_equal(idx.numpy(), np.arange(144))
    assert np.array_equal(idx.numpy(), G["tie_idx_stable"])
    # the raw library call does NOT follow index order: that is why the rule is stated
    assert not np.array_equal(G["tie_idx_raw_torch"], np.arange(144))
    assert [O.pool_size(k, n) for k, n in zip(O.DEFAULT_KS, (160, 144, 101, 61))] == [144, 101, 61, 30]


# ---------------------------------------------------------------- evaluation
@pytest.mark.parametrize("tag", ["small", "hr"])
def test_eval_reductions(golden, tag):
    G = golden("eval")
    N, n = int(G[f"{tag}_N"]), int(G[f"{tag}_n"])
    gts, preds = [], []
    for i in range(N):
        gt = O.synth_connectome(n, 900 + i)
        g = torch.Generator().manual_seed(1900 + i)
        noise = torch.randn(n, n, generator=g) * 0.1
        pred = torch.clamp(gt + (noise + noise.t()) / 2, 0, 1)
        pred.fill_diagonal_(0)
        gts.append(gt.numpy()); preds.append(pred.numpy())
    mae, pcc, js = VO.eval_reductions(np.stack(preds), np.stack(gts))
    assert mae == pytest.approx(float(G[f"{tag}_mae64"]), rel=1e-12)
    assert pcc == pytest.approx(float(G[f"{tag}_pcc64"]), rel=1e-10)
    assert js == pytest.approx(float(G[f"{tag}_js64"]), rel=1e-10)
    # the reference feeds fp32 to sklearn/scipy: its own numbers sit within fp32 noise of the fp64 yardstick
    assert mae == pytest.approx(float(G[f"{tag}_mae"]), rel=2e-6)
    assert pcc == pytest.approx(float(G[f"{tag}_pcc"]), rel=2e-5)
    assert js == pytest.approx(float(G[f"{tag}_js"]), rel=2e-5)


# ---------------------------------------------------------------- GSR-Net
@pytest.mark.parametrize("name", ["gsrnet_small", "gsrnet_full"])
def test_gsrnet_forward_backward(golden, name):
    G = golden(name)
    lr_dim, hr_dim, n_subj = int(G["lr_dim"]), int(G["hr_dim"]), int(G["n_subj"])
    rs, gs = int(G["row_stride"]), int(G["grad_stride"])
    sd = O.make_state_dict(int(G["weight_seed"]), lr_dim, hr_dim, hr_dim)
    assert list(sd.keys()) == [str(k) for k in G["keys"]]
    torch.set_num_threads(4)
    lrs = [O.synth_connectome(lr_dim, int(G["data_seed"]) + s) for s in range(n_subj)]
    hrs = [O.synth_connectome(hr_dim, int(G["data_seed"]) + 500 + s) for s in range(n_subj)]
    for s in range(n_subj):
        params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        trace = {}
        loss, err, (out, net_outs, start, adj) = O.train_loss(params, lrs[s], hrs[s], trace=trace)
        p = f"s{s}_"
        # same library and op sequence: bit-identical on the machine that wrote the
        # fixtures; tolerances below only absorb MKL code-path differences between CPUs
        # (eigenvectors are the ill-conditioned quantity).
        close = lambda a, b, tol: np.allclose(a, b, rtol=0, atol=tol)
        assert close(out.detach().numpy()[::rs], G[p + "out"], 2e-6)
        assert close(adj.detach().numpy()[::rs], G[p + "adj"], 2e-6)
        assert close(trace["Z"].numpy()[::rs], G[p + "Z"], 2e-6)
        assert close(trace["f_d"].numpy()[::rs], G[p + "f_d"], 5e-5)
        assert close(trace["U"].numpy()[::rs], G[p + "U"], 5e-3)
        assert float(loss.detach()) == pytest.approx(float(G[p + "loss"]), abs=2e-6)
        assert float(err.detach()) == pytest.approx(float(G[p + "error"]), abs=2e-6)
        if s == 0:
            assert close(net_outs.detach().numpy(), G["net_outs"], 2e-6)
            assert close(start.detach().numpy(), G["start_gcn_outs"], 1e-6)
            for i in range(4):
                assert np.array_equal(trace["pool_idx"][i].numpy(), G[f"pool{i}_idx"])
                assert close(trace["pool_scores"][i].numpy(), G[f"pool{i}_scores"], 1e-7)
            loss.backward()
            for k, v in params.items():
                ref = G["s0_grad/" + k]
                got = sample_grad(v.grad, gs)
                assert np.allclose(got, ref, rtol=0, atol=1e-7 + 2e-3 * np.abs(ref).max()), k
    _, _, grads = O.step_grads(sd, lrs, hrs)
    for k, g in grads.items():
        ref = G["mean_grad/" + k]
        assert np.allclose(sample_grad(g, gs), ref, rtol=0, atol=1e-7 + 2e-3 * np.abs(ref).max()), k


def test_hr_target_basis_matches_reference(golden):
    G = golden("gsrnet_full")
    hr = O.synth_connectome(268, int(G["data_seed"]) + 500)
    got = O.hr_target_basis(hr, 160).numpy()[::int(G["row_stride"])]
    assert np.allclose(got, G["s0_U_hr_reduced"], rtol=0, atol=5e-3)
