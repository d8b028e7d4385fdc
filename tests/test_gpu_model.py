"""GPU parity tests, module level: the drop-in modules (GSRLayer, GraphConvolution, GraphUnet,
GSRNet, the train step) against the CPU oracle on identical seeded inputs and weights, and
against the golden fixtures written from the unmodified reference.  Run with ``-m gpu``.

Where the eigenbasis matters the comparison is made twice: with the oracle's LAPACK basis
injected into the CUDA path (isolates every non-eigh kernel at 1e-5-level tolerances) and
with the CUDA path's own Jacobi basis (tolerances of SURVEY H1).
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import gsr_oracle as O

pytestmark = pytest.mark.gpu


class Args:
    def __init__(self, lr_dim=160, hr_dim=268, hidden_dim=268, padding=0, lmbda=16, epochs=1):
        self.lr_dim, self.hr_dim, self.hidden_dim = lr_dim, hr_dim, hidden_dim
        self.padding, self.lmbda, self.epochs = padding, lmbda, epochs


def make_model(pkg, sd, lr_dim=160, hr_dim=268):
    model = pkg.GSRNet(list(O.DEFAULT_KS), Args(lr_dim, hr_dim, hr_dim))
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd)
    return model.cuda()


def rel_fro(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def own_basis_aligned(pkg, lr, U_ref):
    """The kernels' own eigendecomposition of the normalised LR adjacency, its eigenvectors sign-aligned column by
    column to the reference's LAPACK basis ``U_ref`` (sign of <u_gpu_k, u_ref_k>).  Eigenvector signs are not
    canonical and the stated sign rule is decided by rounding noise whenever the two largest
    |components| of a vector are within ~1e-4 of each other, which is the normal case for 160-node connectomes
    (72 % of random subjects have such a vector) -- so the comparison with the reference fixture is
    made in the reference's sign convention, for EVERY subject, and the number of flipped columns is reported.
    The sign rule itself is asserted on the raw kernel output.  Returns (w, U_raw, U_aligned, flips)."""
    A = pkg.normalize_adj_torch(lr.cuda())
    w, U, info = pkg.functional.eigh(A, "U")
    assert int(info.min()) > 0
    piv = torch.gather(U, -2, U.abs().argmax(dim=-2, keepdim=True))
    assert bool((piv > 0).all()), "sign rule: the largest-|.| component of every eigenvector is positive"
    dots = (U.double().cpu() * U_ref.double()).sum(0)
    assert float(dots.abs().min()) > 0.9, "an eigenvector of the kernel does not match its LAPACK counterpart"
    sign = torch.where(dots < 0, -1.0, 1.0).float()
    return w, U, U * sign.cuda(), int((dots < 0).sum())


@pytest.fixture(scope="module")
def setup(cuda_pkg):
    torch.set_num_threads(4)
    sd = O.make_state_dict(0)
    model = make_model(cuda_pkg, sd)
    return cuda_pkg, sd, model


# ------------------------------------------------------------------ U-Net (subject independent)
def test_graph_unet_forward_and_pool_indices(setup, golden):
    pkg, sd, model = setup
    G = golden("gsrnet_full")
    A = pkg.normalize_adj_torch(O.synth_connectome(160, 1000).cuda())
    net_outs, start = model.net(A, torch.eye(160, device="cuda"))
    assert np.allclose(start.detach().cpu().numpy(), G["start_gcn_outs"], rtol=0, atol=2e-6)
    assert np.allclose(net_outs.detach().cpu().numpy(), G["net_outs"], rtol=0, atol=1e-5)
    # Pool levels: compare per ORIGINAL node.  The order inside the selected set never changes
    # the U-Net output (every op is row-wise and the unpool scatters back by index), but it
    # permutes the rows of the next level, so scores / indices are mapped back to node ids.
    ids_ref, ids_got = torch.arange(160), torch.arange(160)
    for i in range(4):
        s_ref = torch.from_numpy(G[f"pool{i}_scores"])
        s_got = model.net.pool_scores[i].cpu()
        idx_ref = torch.from_numpy(G[f"pool{i}_idx"])
        idx_got = model.net.pool_indices[i].cpu()
        assert idx_got.dtype == torch.int64 and idx_got.numel() == idx_ref.numel()
        by_node_ref = dict(zip(ids_ref.tolist(), s_ref.tolist()))
        by_node_got = dict(zip(ids_got.tolist(), s_got.tolist()))
        common = set(by_node_ref) & set(by_node_got)
        assert len(common) >= len(by_node_ref) - 2, f"level {i}: different node sets enter the pool"
        assert max(abs(by_node_ref[k] - by_node_got[k]) for k in common) < 3e-7, f"level {i}"
        k = idx_ref.numel()
        srt = torch.sort(s_ref, descending=True).values
        boundary = float(srt[k - 1])
        sel_ref, sel_got = ids_ref[idx_ref], ids_got[idx_got]
        # The selected SET is exact up to nodes whose reference score sits within the score noise
        # (3e-7) of the k-th score: sigmoid(x/100) squeezes 160 scores into ~3500 fp32 bins, so
        # near-ties at the boundary are normal (this fixture has one 1 ulp apart) and no
        # implementation of the preceding mat-mul can promise their order.
        for node in set(sel_ref.tolist()) ^ set(sel_got.tolist()):
            if node in by_node_ref:
                assert abs(by_node_ref[node] - boundary) < 1e-6, f"level {i}: node {node} is not a near-tie"
        # and the ORDER (bit-exact idx) wherever neighbouring reference scores are separated
        if torch.equal(ids_ref, ids_got) and float(torch.diff(srt[:k + 1]).abs().min()) > 1e-6:
            assert torch.equal(idx_got, idx_ref), f"level {i}"
        ids_ref, ids_got = sel_ref, sel_got


def test_graph_unet_backward(setup):
    pkg, sd, model = setup
    g = torch.Generator().manual_seed(21)
    T1, T2 = torch.randn(160, 268, generator=g), torch.randn(160, 268, generator=g)
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    A = O.normalize_adj(O.synth_connectome(160, 1000))
    X, s0 = O.graph_unet(params, A, torch.eye(160))
    ((X * T1).sum() + (s0 * T2).sum()).backward()
    model.zero_grad()
    Xc, s0c = model.net(A.cuda(), torch.eye(160, device="cuda"))
    ((Xc * T1.cuda()).sum() + (s0c * T2.cuda()).sum()).backward()
    for k, p in model.named_parameters():
        if k.startswith("net."):
            assert rel_fro(p.grad.cpu(), params[k].grad) < 2e-5, k


# ------------------------------------------------------------------ GSR layer / graph convolution
def test_gsr_layer_injected_basis(setup):
    pkg, sd, model = setup
    g = torch.Generator().manual_seed(5)
    A = O.normalize_adj(O.synth_connectome(160, 1001))
    f = torch.randn(160, 268, generator=g) * 0.1
    _, U = O.eigh_canonical(A)
    W = sd["layer.weights"].clone().requires_grad_(True)
    fr = f.clone().requires_grad_(True)
    trace = {}
    adj, Z = O.gsr_layer(W, A, fr, U=U, trace=trace)
    T1, T2 = torch.randn(268, 268, generator=g), torch.randn(268, 268, generator=g)
    ((adj * T1).sum() + (Z * T2).sum()).backward()
    layer = model.layer
    layer.zero_grad()
    fc = f.cuda().requires_grad_(True)
    adj_c, Z_c = layer(A.cuda(), fc, U=U.cuda())
    ((adj_c * T1.cuda()).sum() + (Z_c * T2.cuda()).sum()).backward()
    # atol: entries of W.U^T.f that cancel to ~0 carry the contraction's absolute error (3xTF32 tensor-core
    # path: <= 1e-6 * sum|a||b|), scaled by r_i r_j ~ 1/3000
    assert torch.allclose(adj_c.detach().cpu(), adj.detach(), rtol=1e-5, atol=5e-8)
    assert torch.allclose(Z_c.detach().cpu(), Z.detach(), rtol=1e-5, atol=1e-7)
    assert torch.allclose(layer.f_d.cpu(), trace["f_d"], rtol=1e-5, atol=2e-5)
    assert rel_fro(layer.weights.grad.cpu(), W.grad) < 1e-4
    assert rel_fro(fc.grad.cpu(), fr.grad) < 1e-4


def test_gsr_layer_own_eigh(setup):
    pkg, sd, model = setup
    A = O.normalize_adj(O.synth_connectome(160, 1000))
    f = O.graph_unet(sd, A, torch.eye(160))[0]
    adj, Z = O.gsr_layer(sd["layer.weights"], A, f)
    with torch.no_grad():
        adj_c, Z_c = model.layer(A.cuda(), f.cuda())
    assert int(model.layer.eigh_info.min()) > 0
    # identical weights, different (but equally accurate) eigenbases
    assert torch.allclose(adj_c.cpu(), adj, rtol=0, atol=2e-6)
    assert torch.allclose(Z_c.cpu(), Z, rtol=0, atol=2e-6)


def test_graph_convolution(setup, golden):
    pkg, sd, model = setup
    G = golden("ops")
    gc = pkg.GraphConvolution(7, 5, 0.0).cuda()
    with torch.no_grad():
        gc.weight.copy_(torch.from_numpy(G["gc_w"]))
    out = gc(torch.from_numpy(G["gc_x"]).cuda(), torch.from_numpy(G["gc_adj"]).cuda())
    assert np.allclose(out.detach().cpu().numpy(), G["gc_out"], rtol=1e-5, atol=1e-6)
    # batched fwd + bwd against autograd of the oracle
    g = torch.Generator().manual_seed(2)
    x, adj, W = torch.randn(3, 268, 268, generator=g), torch.rand(3, 268, 268, generator=g), torch.randn(268, 268, generator=g) * 0.05
    T = torch.randn(3, 268, 268, generator=g)
    xr, ar, Wr = x.clone().requires_grad_(True), adj.clone().requires_grad_(True), W.clone().requires_grad_(True)
    ref = torch.stack([O.graph_convolution(Wr, xr[b], ar[b]) for b in range(3)])
    (ref * T).sum().backward()
    gc2 = pkg.GraphConvolution(268, 268).cuda()
    with torch.no_grad():
        gc2.weight.copy_(W)
    xc, ac = x.cuda().requires_grad_(True), adj.cuda().requires_grad_(True)
    out = gc2(xc, ac)
    (out * T.cuda()).sum().backward()
    # two chained contractions; the tensor-core accumulator rounds toward zero at every K = 8 step, which
    # shows as a ~3e-6 relative bias on all-positive sums (the SIMT FP32 backend sits at 3e-7 here)
    assert rel_fro(out.detach().cpu(), ref.detach()) < 6e-6
    assert rel_fro(xc.grad.cpu(), xr.grad) < 6e-6
    assert rel_fro(ac.grad.cpu(), ar.grad) < 6e-6
    assert rel_fro(gc2.weight.grad.cpu(), Wr.grad) < 6e-6


# ------------------------------------------------------------------ GSRNet
def _oracle_bases(lrs, hrs):
    Us = [O.eigh_canonical(O.normalize_adj(lr))[1] for lr in lrs]
    Uh = [O.hr_target_basis(hr, 160) for hr in hrs]
    return Us, Uh


def test_gsrnet_forward_backward_single_subject_injected(setup):
    """B = 1, reference call signature, LAPACK bases injected on both sides."""
    pkg, sd, model = setup
    lr, hr = O.synth_connectome(160, 1000), O.synth_connectome(268, 1500)
    (U,), (Uh,) = _oracle_bases([lr], [hr])
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    trace = {}
    loss, err, (out, net_outs, start, adj) = O.train_loss(params, lr, hr, U=U, U_hr_reduced=Uh, trace=trace)
    loss.backward()
    model.zero_grad()
    out_c, net_c, start_c, adj_c = model(lr.cuda(), U=U.cuda())
    assert out_c.shape == (268, 268) and net_c.shape == (160, 268) and adj_c.shape == (268, 268)
    crit = pkg.functional.l1_loss
    loss_c = 16 * crit(net_c, start_c) + crit(model.layer.weights, Uh.cuda()) + crit(out_c, hr.cuda())
    loss_c.backward()
    assert torch.allclose(out_c.detach().cpu(), out.detach(), rtol=0, atol=1e-5)
    assert torch.allclose(adj_c.detach().cpu(), adj.detach(), rtol=0, atol=1e-6)
    assert torch.allclose(net_c.detach().cpu(), net_outs.detach(), rtol=0, atol=1e-5)
    assert torch.allclose(model.Z.detach().cpu(), trace["Z"], rtol=0, atol=1e-5)
    assert torch.allclose(model.hidden2.detach().cpu(), trace["hidden2"], rtol=0, atol=1e-5)
    assert float(out_c.detach().diagonal().min()) == 1.0 and float(out_c.min()) >= 0.0
    assert float(loss_c) == pytest.approx(float(loss), abs=2e-6)
    for k, p in model.named_parameters():
        assert rel_fro(p.grad.cpu(), params[k].grad) < 2e-3, k          # L1 / abs sign kinks


def test_gsrnet_own_eigh_matches_golden(setup, golden):
    """Whole path on the GPU with the kernels' own eigendecomposition against the fixture written by the unmodified
    reference, every subject asserted (no margin guard, no skip): outputs in the reference's sign convention
    (own_basis_aligned), and -- for the subjects whose raw kernel basis already has the reference's signs -- the
    plain ``model(lr)`` call bit-identical to that."""
    pkg, sd, model = setup
    G = golden("gsrnet_full")
    rs, n_subj = int(G["row_stride"]), int(G["n_subj"])
    assert n_subj >= 8
    checked, flips_total, raw_identical = 0, 0, 0
    for s in range(n_subj):
        p = f"s{s}_"
        lr = O.synth_connectome(160, int(G["data_seed"]) + s)
        w, U_raw, U_al, flips = own_basis_aligned(pkg, lr, torch.from_numpy(G[p + "U_full"]))
        flips_total += flips
        with torch.no_grad():
            out, net_outs, start, adj = model(lr.cuda(), U=U_al)
        assert np.allclose(out.cpu().numpy()[::rs], G[p + "out"], rtol=0, atol=2e-5), s
        assert np.allclose(adj.cpu().numpy()[::rs], G[p + "adj"], rtol=0, atol=5e-6), s
        assert np.allclose(model.Z.cpu().numpy()[::rs], G[p + "Z"], rtol=0, atol=2e-5), s
        assert np.allclose(model.layer.f_d.cpu().numpy()[::rs], G[p + "f_d"], rtol=1e-3, atol=4e-4), s    # |F| ~ 0.36
        assert np.allclose(w.cpu().numpy(), G[p + "eigvals"], rtol=0, atol=3e-6), s
        assert np.allclose(U_al.cpu().numpy()[::rs], G[p + "U"], rtol=0, atol=2e-3), s
        if flips == 0:
            with torch.no_grad():
                out2, _, _, adj2 = model(lr.cuda())             # own eigh AND own sign rule, nothing injected
            assert torch.equal(out2, out) and torch.equal(adj2, adj)
            raw_identical += 1
        checked += 1
    assert checked == n_subj
    print(f"own-basis fixture parity: {checked} subjects asserted, {flips_total} of {160 * n_subj} eigenvector columns "
          f"needed a sign flip, {raw_identical} subjects matched with the raw kernel signs")


def test_gsrnet_batched_equals_per_subject_oracle(setup):
    """B = 4 in one call == the oracle run subject by subject with frozen weights and its
    gradients averaged."""
    pkg, sd, model = setup
    B = 4
    lrs = [O.synth_connectome(160, 2000 + s) for s in range(B)]
    hrs = [O.synth_connectome(268, 2500 + s) for s in range(B)]
    Us, Uh = _oracle_bases(lrs, hrs)
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    losses, errs, outs = [], [], []
    for s in range(B):
        loss, err, (out, _, _, _) = O.train_loss(params, lrs[s], hrs[s], U=Us[s], U_hr_reduced=Uh[s])
        (loss / B).backward()
        losses.append(float(loss)); errs.append(float(err)); outs.append(out.detach())
    model.zero_grad()
    step = pkg.GSRTrainStep(model, lmbda=16, padding=0)
    lr_b, hr_b = torch.stack(lrs).cuda(), torch.stack(hrs).cuda()
    loss_c, err_c = step(lr_b, hr_b, U_hr_reduced=torch.stack(Uh).cuda(), U_lr=torch.stack(Us).cuda())
    assert float(loss_c) == pytest.approx(np.mean(losses), abs=2e-6)
    assert float(err_c) == pytest.approx(np.mean(errs), abs=2e-6)
    with torch.no_grad():
        out_b = model(lr_b, U=torch.stack(Us).cuda())[0]
    assert out_b.shape == (B, 268, 268)
    assert torch.allclose(out_b.cpu(), torch.stack(outs), rtol=0, atol=1e-5)
    for k, p in model.named_parameters():
        assert rel_fro(p.grad.cpu(), params[k].grad) < 2e-3, k
    # batch of B == B calls of batch 1 on the GPU itself (same kernels, same order): near bit-level
    model.zero_grad()
    for s in range(B):
        l1, _ = step.forward_loss(lr_b[s:s + 1], hr_b[s:s + 1], U_hr_reduced=torch.stack(Uh).cuda()[s:s + 1],
                                  U_lr=torch.stack(Us).cuda()[s:s + 1])
        (l1 / B).backward()
    g_loop = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad()
    step(lr_b, hr_b, U_hr_reduced=torch.stack(Uh).cuda(), U_lr=torch.stack(Us).cuda())
    for k, p in model.named_parameters():
        assert rel_fro(p.grad, g_loop[k]) < 1e-5, k


def test_gsrnet_mean_gradients_match_golden(setup, golden):
    """Loss and batch-mean parameter gradients with the kernels' own LR eigendecomposition (sign-aligned to the
    reference's basis) against the reference-written fixture: all subjects in one batched step, nothing skipped."""
    pkg, sd, model = setup
    G = golden("gsrnet_full")
    n_subj, gs = int(G["n_subj"]), int(G["grad_stride"])
    lrs = [O.synth_connectome(160, int(G["data_seed"]) + s) for s in range(n_subj)]
    hrs = [O.synth_connectome(268, int(G["data_seed"]) + 500 + s) for s in range(n_subj)]
    Uh = torch.stack([torch.from_numpy(G[f"s{s}_U_hr_full"]) for s in range(n_subj)]).cuda()
    Ul = torch.stack([own_basis_aligned(pkg, lrs[s], torch.from_numpy(G[f"s{s}_U_full"]))[2] for s in range(n_subj)])
    model.zero_grad()
    step = pkg.GSRTrainStep(model, lmbda=16, padding=0, check_errors=True)
    loss, err = step(torch.stack(lrs).cuda(), torch.stack(hrs).cuda(), U_hr_reduced=Uh, U_lr=Ul)
    want_loss = np.mean([float(G[f"s{s}_loss"]) for s in range(n_subj)])
    want_err = np.mean([float(G[f"s{s}_error"]) for s in range(n_subj)])
    assert float(err) == pytest.approx(want_err, abs=5e-6)
    assert float(loss) == pytest.approx(want_loss, abs=5e-5)
    checked = 0
    for k, p in model.named_parameters():
        ref = G["mean_grad/" + k]
        a = p.grad.detach().cpu().numpy().reshape(-1)
        got = a if a.size <= 4096 else a[::gs]
        rel = np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30)
        assert rel < 2e-2, (k, rel)           # gradients move O(1e-3..1e-2) with the basis
        checked += 1
    assert checked == 33
    # the GPU's own HR-target basis (268-node tridiagonal route) gives the same |W - U_hr| term up to its sign noise
    Uh_own = pkg.hr_target_basis(torch.stack(hrs).cuda(), 160)
    dots = (Uh_own.double() * Uh.double()).sum(1)
    assert float(dots.abs().min()) > 0.9
    l_own = pkg.functional.l1_loss_per_subject(Uh_own * torch.where(dots < 0, -1.0, 1.0).float()[:, None, :], model.layer.weights)
    l_ref = pkg.functional.l1_loss_per_subject(Uh, model.layer.weights)
    assert torch.allclose(l_own, l_ref, rtol=0, atol=2e-6)
    # single-subject gradients (subject 0) against the fixture as well
    model.zero_grad()
    loss0, _ = step(lrs[0][None].cuda(), hrs[0][None].cuda(), U_hr_reduced=Uh[:1], U_lr=Ul[:1])
    assert float(loss0) == pytest.approx(float(G["s0_loss"]), abs=5e-5)
    for k, p in model.named_parameters():
        ref = G["s0_grad/" + k]
        a = p.grad.detach().cpu().numpy().reshape(-1)
        got = a if a.size <= 4096 else a[::gs]
        assert np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30) < 2e-2, k


def test_gsrnet_small_lr_dim_fixture(cuda_pkg, golden):
    """lr_dim = 136 (ragged against the 32-row register tiling of the eigh kernel): every subject asserted."""
    G = golden("gsrnet_small")
    sd = O.make_state_dict(int(G["weight_seed"]), 136, 268, 268)
    model = make_model(cuda_pkg, sd, 136, 268)
    rs, n_subj = int(G["row_stride"]), int(G["n_subj"])
    checked = 0
    for s in range(n_subj):
        lr = O.synth_connectome(136, int(G["data_seed"]) + s)
        w, U_raw, U_al, flips = own_basis_aligned(cuda_pkg, lr, torch.from_numpy(G[f"s{s}_U_full"]))
        with torch.no_grad():
            out, _, _, adj = model(lr.cuda(), U=U_al)
        assert np.allclose(out.cpu().numpy()[::rs], G[f"s{s}_out"], rtol=0, atol=2e-5), s
        assert np.allclose(adj.cpu().numpy()[::rs], G[f"s{s}_adj"], rtol=0, atol=5e-6), s
        assert np.allclose(w.cpu().numpy(), G[f"s{s}_eigvals"], rtol=0, atol=3e-6), s
        checked += 1
    assert checked == n_subj >= 4


def test_train_and_test_loops_match_oracle(cuda_pkg, capsys):
    """utils/gsr_train.py train() / test() driven exactly as the notebook does (numpy float64 in,
    per-subject Adam steps, print-only) against the same loop over the oracle."""
    sd = O.make_state_dict(0)
    model = make_model(cuda_pkg, sd)
    n_subj = 3
    lrs = np.stack([O.synth_connectome(160, 3000 + s).numpy().astype(np.float64) for s in range(n_subj)])
    hrs = np.stack([O.synth_connectome(268, 3500 + s).numpy().astype(np.float64) for s in range(n_subj)])
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    assert cuda_pkg.train(model, opt, lrs, hrs, Args(epochs=1)) is None
    printed = capsys.readouterr().out
    assert printed.startswith("Epoch: 1, Loss: ")
    loss_got = float(printed.split("Loss: ")[1].split(",")[0])
    err_got = float(printed.split("Error (MAE): ")[1])
    # oracle loop: same init, same optimiser, its own LAPACK bases
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    opt_ref = torch.optim.Adam(list(params.values()), lr=1e-4)
    losses, errs = [], []
    for s in range(n_subj):
        opt_ref.zero_grad()
        loss, err, _ = O.train_loss(params, torch.from_numpy(lrs[s]).float(), torch.from_numpy(hrs[s]).float())
        loss.backward()
        opt_ref.step()
        losses.append(float(loss)); errs.append(float(err))
    assert loss_got == pytest.approx(np.mean(losses), abs=1e-4)
    assert err_got == pytest.approx(np.mean(errs), abs=1e-5)
    # parameters moved the same way (Adam's first steps are +-lr per element: compare in bulk)
    moved = 0.0
    for k, p in model.named_parameters():
        moved = max(moved, float((p.detach().cpu() - sd[k]).abs().max()))
        agree = float(((p.detach().cpu() - sd[k]).sign() == (params[k].detach() - sd[k]).sign()).float().mean())
        assert agree > 0.97, (k, agree)
    assert 1e-4 < moved < 1e-3
    # test(): per-subject MAE against unit-diagonal targets, all-zero samples skipped, caller's
    # array mutated in place (utils/gsr_train.py:70)
    frozen = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    lr_t = np.concatenate([lrs, np.zeros((1, 160, 160))]); hr_t = np.concatenate([hrs, np.zeros((1, 268, 268))])
    assert cuda_pkg.test(model, lr_t, hr_t, Args()) is None
    out = capsys.readouterr().out.strip().splitlines()
    maes = [float(l.split("MAE: ")[1]) for l in out if l.startswith("MAE: ")]
    assert len(maes) == n_subj and out[-1].startswith("Test error MAE: ")
    assert np.all(np.diagonal(hr_t[:n_subj], axis1=1, axis2=2) == 1.0) and np.all(hr_t[n_subj] == 0)
    ref = O.test_errors(frozen, [torch.from_numpy(a).float() for a in lrs], [torch.from_numpy(a).float() for a in hrs])
    assert np.allclose(maes, ref, rtol=0, atol=1e-5)


def test_train_batched_runs_and_caches(cuda_pkg, capsys):
    sd = O.make_state_dict(0)
    model = make_model(cuda_pkg, sd)
    lrs = np.stack([O.synth_connectome(160, 4000 + s).numpy() for s in range(6)])
    hrs = np.stack([O.synth_connectome(268, 4500 + s).numpy() for s in range(6)])
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    hist = cuda_pkg.train_batched(model, opt, lrs, hrs, Args(epochs=3), batch_size=4, verbose=False)
    assert len(hist) == 3 and all(np.isfinite(hist)) and hist[2] < hist[0]
    # first-epoch loss == mean of the oracle's per-subject losses at the initial weights (first batch)
    losses = [float(O.train_loss(sd, torch.from_numpy(lrs[s]), torch.from_numpy(hrs[s]))[0]) for s in range(4)]
    model2 = make_model(cuda_pkg, sd)
    step = cuda_pkg.GSRTrainStep(model2)
    loss, _ = step(torch.from_numpy(lrs[:4]).cuda(), torch.from_numpy(hrs[:4]).cuda())     # own LR and HR eigh
    assert float(loss) == pytest.approx(np.mean(losses), abs=1e-4)


def test_model_requires_gpu_and_library(cuda_pkg):
    model = cuda_pkg.GSRNet(list(O.DEFAULT_KS), Args())
    with pytest.raises(RuntimeError):
        model(O.synth_connectome(160, 1))           # parameters still on the CPU: loud, no fallback
    with pytest.raises(RuntimeError):
        cuda_pkg.kernels.normalize_adj_fwd(torch.rand(4, 4))


def test_full_batch_step_properties(setup):
    """BASELINE config 2 size (1024 subjects): one fwd+bwd; size-independent properties."""
    pkg, sd, model = setup
    B = 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(B, 160, 160, device="cuda", generator=g)
    lr = (x + x.transpose(1, 2)) / 2
    lr.diagonal(dim1=1, dim2=2).zero_()
    y = torch.rand(B, 268, 268, device="cuda", generator=g)
    hr = (y + y.transpose(1, 2)) / 2
    hr.diagonal(dim1=1, dim2=2).zero_()
    Uh = torch.zeros(B, 268, 160, device="cuda")
    model.zero_grad()
    step = pkg.GSRTrainStep(model, check_errors=True)
    loss, err = step(lr, hr, U_hr_reduced=Uh)
    out = model(lr[:8])[0].detach()
    assert bool(torch.isfinite(loss)) and bool(torch.isfinite(err))
    assert torch.equal(out, out.transpose(1, 2)) and float(out.min()) >= 0
    assert float((out.diagonal(dim1=1, dim2=2) - 1).abs().max()) == 0
    # permutation equivariance over the subjects: same mean loss / gradients for a shuffled batch
    g_ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    perm = torch.randperm(B, device="cuda")
    model.zero_grad()
    loss2, _ = step(lr[perm], hr[perm], U_hr_reduced=Uh)
    assert float(loss2) == pytest.approx(float(loss), rel=1e-5)
    for k, p in model.named_parameters():
        assert rel_fro(p.grad, g_ref[k]) < 1e-4, k


def _synth_batch(n, B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(B, n, n, device="cuda", generator=g)
    m = (x + x.transpose(1, 2)) / 2
    m.diagonal(dim1=1, dim2=2).zero_()
    return m.contiguous()


def test_full_batch_matches_oracle_on_sampled_subjects(setup):
    """BASELINE configs[1] (1024 subjects in one call, own eigendecompositions): 8 subjects of the batch against the
    oracle (outputs, adjacency, eigenvalues, per-subject loss terms), the oracle fed the GPU's eigenbases of those
    subjects so that the comparison is tight; plus the batch-mean gradient == mean of the 8 chunk gradients."""
    pkg, sd, model = setup
    F_ = pkg.functional
    B = 1024
    lr, hr = _synth_batch(160, B, 31), _synth_batch(268, B, 32)
    Uh = pkg.hr_target_basis(hr, 160)
    with torch.no_grad():
        out, net_outs, start, adj = model(lr)
        err_b = F_.l1_loss_per_subject(out, hr)
        w_b = F_.l1_loss_per_subject(Uh, model.layer.weights)
        reg = F_.l1_loss(net_outs, start)
    assert int(model.layer.eigh_info.min()) > 0
    U_lr, w_lr = model.layer.U_lr, model.layer.eig_val_lr
    checked = 0
    for s in (0, 1, 255, 256, 511, 700, 1022, 1023):
        lr_s, hr_s = lr[s].cpu(), hr[s].cpu()
        loss_ref, err_ref, (out_ref, _, _, adj_ref) = O.train_loss(sd, lr_s, hr_s, U=U_lr[s].cpu(), U_hr_reduced=Uh[s].cpu())
        assert torch.allclose(out[s].cpu(), out_ref, rtol=0, atol=1e-5), s
        assert torch.allclose(adj[s].cpu(), adj_ref, rtol=0, atol=1e-6), s
        assert float(err_b[s]) == pytest.approx(float(err_ref), abs=2e-6), s
        assert float(16 * reg + w_b[s] + err_b[s]) == pytest.approx(float(loss_ref), abs=5e-6), s
        w64 = torch.linalg.eigvalsh(O.normalize_adj(lr_s).double(), UPLO="U")
        assert float((w_lr[s].cpu().double() - w64).abs().max()) < 3e-6, s
        A64 = O.normalize_adj(lr_s).double()
        A64 = torch.triu(A64) + torch.triu(A64, 1).T
        Us = U_lr[s].cpu().double()
        assert float((A64 @ Us - Us * w_lr[s].cpu().double()).abs().max()) < 5e-6, s
        checked += 1
    assert checked == 8
    # size-independent property at the full size: gradients are linear in the subjects
    step = pkg.GSRTrainStep(model)
    model.zero_grad()
    step(lr, hr, U_hr_reduced=Uh, U_lr=U_lr)
    g_full = {k: p.grad.clone() for k, p in model.named_parameters()}
    acc = {k: torch.zeros_like(v) for k, v in g_full.items()}
    for c in range(8):
        sl = slice(128 * c, 128 * (c + 1))
        model.zero_grad()
        step(lr[sl], hr[sl], U_hr_reduced=Uh[sl], U_lr=U_lr[sl])
        for k, p in model.named_parameters():
            acc[k] += p.grad / 8
    # (the subject-reduced products accumulate 14 subjects x 34 k-steps in tensor memory at B = 1024 against 2 x 34
    # per chunk; the tensor core's accumulator rounds toward zero at every step, which shows as a ~3e-5 relative
    # drift between the two groupings -- measured 3.6e-5 on gc1.weight)
    for k in g_full:
        assert rel_fro(g_full[k], acc[k]) < 1e-4, k


def test_batched_gradients_match_oracle_through_the_split_reduction(setup):
    """B = 96 subjects (> 74: the subject-reduced weight-gradient products run their full 74-way split with the
    deterministic partial reduce in tensor memory) against the oracle's accumulated per-subject gradients, all 33
    tensors, own LR eigendecomposition fed to both sides."""
    pkg, sd, model = setup
    B = 96
    lr, hr = _synth_batch(160, B, 41), _synth_batch(268, B, 42)
    Uh = pkg.hr_target_basis(hr, 160)
    model.zero_grad()
    step = pkg.GSRTrainStep(model, check_errors=True)
    loss, err = step(lr, hr, U_hr_reduced=Uh)
    U_lr = model.layer.U_lr
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    losses, errs = [], []
    for s in range(B):
        l, e, _ = O.train_loss(params, lr[s].cpu(), hr[s].cpu(), U=U_lr[s].cpu(), U_hr_reduced=Uh[s].cpu())
        (l / B).backward()
        losses.append(float(l)); errs.append(float(e))
    assert float(loss) == pytest.approx(np.mean(losses), abs=5e-6)
    assert float(err) == pytest.approx(np.mean(errs), abs=2e-6)
    for k, p in model.named_parameters():
        assert rel_fro(p.grad.cpu(), params[k].grad) < 2e-3, k


def test_fused_final_pass_equals_the_materialised_prediction(setup):
    """GSRTrainStep's default (models/gsr_model.py:32-37 and the L1 term against the HR target in one pass, the prediction
    never written) against the step that materialises model(lr) and calls l1_loss on it: every parameter gradient BIT
    identical, loss and error to rounding (the per-subject sums are folded in a different order)."""
    pkg, sd, model = setup
    B = 5
    lr, hr = _synth_batch(160, B, 61), _synth_batch(268, B, 62)
    Uh = pkg.hr_target_basis(hr, 160)
    model.zero_grad()
    loss_a, err_a = pkg.GSRTrainStep(model, fuse_final=False)(lr, hr, U_hr_reduced=Uh)
    U_lr = model.layer.U_lr
    g_a = [p.grad.clone() for p in model.parameters()]
    model.zero_grad()
    loss_b, err_b = pkg.GSRTrainStep(model)(lr, hr, U_hr_reduced=Uh, U_lr=U_lr)
    assert float(err_a) == pytest.approx(float(err_b), rel=1e-6)
    assert float(loss_a) == pytest.approx(float(loss_b), rel=1e-6)
    for a, p in zip(g_a, model.parameters()):
        assert torch.equal(a, p.grad)


@pytest.mark.parametrize("G", [2, 8])
def test_subject_shards_reduce_to_the_one_gpu_gradient(setup, G):
    """SURVEY section 4 shard emulation on one GPU: the batch split into G rank shards (shard_bounds), every shard's
    gradients packed by FlatBucket.pack(1 / G) exactly as SubjectParallel.sync_gradients does, the flat buckets
    summed (what the all-reduce computes) and unpacked -- must equal the gradient of the whole batch on one GPU."""
    pkg, sd, model = setup
    B = 16
    lr, hr = _synth_batch(160, B, 51), _synth_batch(268, B, 52)
    Uh = pkg.hr_target_basis(hr, 160)
    step = pkg.GSRTrainStep(model)
    model.zero_grad()
    step(lr, hr, U_hr_reduced=Uh)
    U_lr = model.layer.U_lr
    g_one = [p.grad.clone() for p in model.parameters()]
    total = None
    for r in range(G):
        lo, hi = pkg.shard_bounds(B, r, G)
        assert hi - lo == B // G
        model.zero_grad()
        step(lr[lo:hi], hr[lo:hi], U_hr_reduced=Uh[lo:hi], U_lr=U_lr[lo:hi])
        grads = [p.grad for p in model.parameters()]
        bucket = pkg.kernels.FlatBucket(grads)
        assert bucket.total == sum(p.numel() for p in model.parameters())
        flat = bucket.pack(1.0 / G).clone()
        total = flat if total is None else total + flat
    bucket.flat.copy_(total)
    bucket.unpack(1.0)
    for a, p in zip(g_one, model.parameters()):
        assert rel_fro(p.grad, a) < 2e-6


@pytest.mark.gpu
def test_graphed_train_step_equals_eager_steps(cuda_pkg):
    """SURVEY 8 f row N3: the captured CUDA graph (forward, backward, Adam) replays to the same losses and
    parameters as the eager steps, bit for bit, and capture leaves parameters / optimiser state untouched."""
    pkg = cuda_pkg
    import bench
    B = 6
    lr = bench.synth_batch(torch, 160, B, 11, torch.device("cuda"))
    hr = bench.synth_batch(torch, 268, B, 12, torch.device("cuda"))
    U = pkg.hr_target_basis(hr, 160)

    def build():
        torch.manual_seed(3)
        m = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], bench.Args()).cuda()
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)

    m1, o1 = build()
    eager = pkg.GSRTrainStep(m1)
    m2, o2 = build()
    before = [p.detach().clone() for p in m2.parameters()]
    graphed = pkg.GraphedTrainStep(m2, o2, B)
    assert all(torch.equal(a, b) for a, b in zip(before, m2.parameters()))
    assert graphed.launches_per_step > 50
    for i in range(3):
        o1.zero_grad(set_to_none=True)
        l1, e1 = eager(lr, hr, U_hr_reduced=U)
        o1.step()
        l2, e2 = graphed(lr, hr, U)
        assert float(l1) == float(l2) and float(e1) == float(e2), i
        assert all(torch.equal(a, b) for a, b in zip(m1.parameters(), m2.parameters())), i
    with pytest.raises(ValueError):
        pkg.GraphedTrainStep(m1, torch.optim.Adam(m1.parameters(), lr=1e-3), B)
    # two input-buffer sets (a loader fills one while the other replays): alternating them continues the same run
    m5, o5 = build()
    g5 = pkg.GraphedTrainStep(m5, o5, B, slots=2)
    m6, o6 = build()
    eager6 = pkg.GSRTrainStep(m6)
    for i in range(4):
        o6.zero_grad(set_to_none=True)
        l6, _ = eager6(lr, hr, U_hr_reduced=U)
        o6.step()
        l5, _ = g5(lr, hr, U, slot=i % 2)
        assert float(l5) == float(l6), i
    assert all(torch.equal(a, b) for a, b in zip(m5.parameters(), m6.parameters()))
    # the HR-target eigendecomposition inside the graph, as the reference computes it every step
    m3, o3 = build()
    g3 = pkg.GraphedTrainStep(m3, o3, B, hr_basis_in_graph=True)
    l3, _ = g3(lr, hr)
    m4, o4 = build()
    o4.zero_grad(set_to_none=True)
    l4, _ = pkg.GSRTrainStep(m4)(lr, hr)
    assert float(l3) == float(l4)


@pytest.mark.gpu
def test_flat_adam_equals_torch_adam(cuda_pkg):
    """SURVEY 8 f row N3: gsr_adam_step (one launch for all tensors, device-side step counter) against
    torch.optim.Adam on the same gradients, with and without weight decay, 45 tensors (two table chunks)."""
    pkg = cuda_pkg
    g = torch.Generator().manual_seed(8)
    shapes = [(int(a), int(b)) for a, b in torch.randint(1, 40, (45, 2), generator=g)]
    for wd in (0.0, 0.01):
        init = [torch.randn(s, generator=g) for s in shapes]
        p_ref = [torch.nn.Parameter(t.clone().cuda()) for t in init]
        p_new = [torch.nn.Parameter(t.clone().cuda()) for t in init]
        o_ref = torch.optim.Adam(p_ref, lr=1e-2, weight_decay=wd)
        o_new = pkg.FlatAdam(p_new, lr=1e-2, weight_decay=wd)
        for step in range(6):
            grads = [torch.randn(s, generator=g).cuda() for s in shapes]
            for a, b, gr in zip(p_ref, p_new, grads):
                a.grad, b.grad = gr.clone(), gr.clone()
            o_ref.step(); o_new.step()
            worst = max(float((a - b).abs().max()) for a, b in zip(p_ref, p_new))
            assert worst < 2e-6, (wd, step, worst)
        assert float(o_new.param_groups[0]["step_tensor"]) == 6.0
    with pytest.raises(RuntimeError):
        q = torch.nn.Parameter(torch.zeros(3)); q.grad = torch.zeros(3)
        pkg.FlatAdam([q]).step()
    # load_state_dict on an optimiser that has already stepped: the loaded moments (host copies here, as after
    # torch.load(map_location="cpu")) must be the ones the next step uses, and fresh .grad tensors every step
    # (zero_grad(set_to_none=True)) must never leave a stale pointer table behind
    init = [torch.randn(s, generator=g) for s in shapes[:7]]
    p_a = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    p_b = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    o_a, o_b = pkg.FlatAdam(p_a, lr=1e-2), pkg.FlatAdam(p_b, lr=1e-2)
    def step_both(opts_params):
        grads = [torch.randn(s, generator=g).cuda() for s in shapes[:7]]
        for opt, params in opts_params:
            opt.zero_grad(set_to_none=True)
            for q, gr in zip(params, grads):
                q.grad = gr.clone()
            opt.step()
    for _ in range(3):
        step_both([(o_a, p_a), (o_b, p_b)])
    import copy
    sd_a = copy.deepcopy(o_a.state_dict())
    for st in sd_a["state"].values():
        for kk in list(st):
            if torch.is_tensor(st[kk]):
                st[kk] = st[kk].cpu()
    for gsd in sd_a["param_groups"]:
        gsd["step_tensor"] = gsd["step_tensor"].cpu()
    step_both([(o_b, p_b)])                                   # o_b moves on, then is rolled back to o_a's state
    with torch.no_grad():
        for a, b in zip(p_a, p_b):
            b.copy_(a)
    o_b.load_state_dict(sd_a)
    for _ in range(2):
        step_both([(o_a, p_a), (o_b, p_b)])
    assert all(torch.equal(a, b) for a, b in zip(p_a, p_b))
    assert float(o_b.param_groups[0]["step_tensor"]) == 5.0 and o_b.param_groups[0]["step_tensor"].is_cuda


@pytest.mark.gpu
def test_graphed_step_with_flat_adam(cuda_pkg):
    """The whole step on the library's kernels: GraphedTrainStep + FlatAdam follows eager steps with torch's Adam."""
    pkg = cuda_pkg
    import bench
    B = 4
    lr = bench.synth_batch(torch, 160, B, 21, torch.device("cuda"))
    hr = bench.synth_batch(torch, 268, B, 22, torch.device("cuda"))
    U = pkg.hr_target_basis(hr, 160)
    torch.manual_seed(5); m1 = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], bench.Args()).cuda()
    torch.manual_seed(5); m2 = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], bench.Args()).cuda()
    o1 = torch.optim.Adam(m1.parameters(), lr=1e-4)
    graphed = pkg.GraphedTrainStep(m2, pkg.FlatAdam(m2.parameters(), lr=1e-4), B)
    eager = pkg.GSRTrainStep(m1)
    for i in range(3):
        o1.zero_grad(set_to_none=True)
        l1, _ = eager(lr, hr, U_hr_reduced=U)
        o1.step()
        l2, _ = graphed(lr, hr, U)
        assert abs(float(l1) - float(l2)) < 1e-5 * abs(float(l1)), i
    worst = max(float((a - b).abs().max()) for a, b in zip(m1.parameters(), m2.parameters()))
    assert worst < 5e-6
