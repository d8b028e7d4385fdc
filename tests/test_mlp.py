"""SURVEY 8 f, row N2 -- the MLP baseline (models/mlp.py).  CPU part: oracle/mlp_oracle.py against the fixture
written from the unmodified reference (oracle/make_golden_mlp.py) and the host-side mirror; GPU part: the CUDA path
against the oracle and the same fixture, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import mlp_oracle as MO


def golden_state(G):
    return {k[len("state."):]: G[k] for k in G.files if k.startswith("state.")}


def grad_scale(G):
    """Largest parameter gradient of the fixture: the bias of a Linear that feeds a training-mode BatchNorm has a
    mathematically zero gradient (fp32 noise in the reference), so tolerances are relative to the whole set."""
    return max(float(np.abs(G[k]).max()) for k in G.files if k.startswith("grad."))


def case(G):
    return {k[len("case."):]: int(G[k]) for k in G.files if k.startswith("case.")}


# ------------------------------------------------------------------ CPU: oracle pinned to the reference
def test_oracle_eval_forward_equals_reference(golden):
    G = golden("mlp")
    y, stats = MO.forward(golden_state(G), G["x"], training=False)
    assert stats == {}
    assert np.abs(y - G["y_eval"]).max() < 2e-5 * np.abs(G["y_eval"]).max()


def test_oracle_training_step_equals_reference(golden):
    """Training mode with the reference's Dropout probability set to 0 (torch's mask cannot be reproduced)."""
    G = golden("mlp")
    st = golden_state(G)
    y, stats, cache = MO.forward(st, G["x"], training=True, p=0.0, return_cache=True)
    assert np.abs(y - G["y_train_p0"]).max() < 2e-5 * np.abs(G["y_train_p0"]).max()
    for k, v in stats.items():
        assert np.allclose(v, G["after." + k], rtol=1e-5, atol=1e-6), k
    grads, gx = MO.backward(st, cache, G["grad_out"])
    assert np.abs(gx - G["grad_x"]).max() < 1e-4 * np.abs(G["grad_x"]).max()
    gscale = grad_scale(G)
    for k, v in grads.items():
        ref = G["grad." + k]
        assert np.abs(v - ref).max() < 1e-4 * max(np.abs(ref).max(), 1e-2 * gscale), k


def test_oracle_dropout_mask_properties():
    u = MO.dropout_uniform(12345, 1 << 20)
    assert u.dtype == np.float32 and u.min() >= 0 and u.max() < 1
    assert abs((u >= np.float32(0.1)).mean() - 0.9) < 2e-3
    assert np.array_equal(MO.dropout_uniform(12345, 100), u[:100])           # counter based: a prefix is a prefix
    assert not np.array_equal(MO.dropout_uniform(12346, 100), u[:100])
    # known answers (splitmix64 of seed + e * golden, top 24 bits)
    z = MO.dropout_uniform(0, 2)
    assert z[0] == 0.0 and 0 < z[1] < 1


def test_mirror_has_the_reference_interface(pkg, golden):
    G = golden("mlp")
    c = case(G)
    torch.manual_seed(c["seed"])
    m = pkg.SuperResMLP(c["input_size"], c["output_size"], c["hidden_dim"], c["n_layers"])
    st = golden_state(G)
    assert list(m.state_dict().keys()) == list(st.keys())
    # same seeded initialisation as the reference (Linear weights / biases; the fixture perturbs only the BatchNorm tensors)
    for k, v in m.state_dict().items():
        if v.ndim == 2 or (k.endswith(".bias") and f"{k[:-5]}.running_mean" not in st):
            assert np.array_equal(v.numpy(), st[k]), k
    assert m.device.type == "cpu"
    with pytest.raises(RuntimeError):                                         # no CPU path
        m(torch.zeros(2, c["input_size"]))
    # models/mlp.py:9,27: a non-square output_size constructs (h = int(sqrt)) -- 35778 -> 189 -- as in the reference
    assert pkg.SuperResMLP(10, 35778, 4, 1).layers[-1].unflattened_size == (189, 189)


# ------------------------------------------------------------------ GPU: kernels and module
def load_model(pkg, G):
    c = case(G)
    m = pkg.SuperResMLP(c["input_size"], c["output_size"], c["hidden_dim"], c["n_layers"])
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in golden_state(G).items()})
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("B,H", [(9, 52), (64, 214), (1, 7), (130, 33)])
def test_bn_kernels_against_the_oracle(cuda_pkg, B, H):
    K = cuda_pkg.kernels
    g = torch.Generator().manual_seed(B * 1000 + H)
    x = torch.randn(B, H, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.3 * torch.randn(H, generator=g), 0.2 * torch.randn(H, generator=g)
    rm, rv = 0.1 * torch.randn(H, generator=g), 0.5 + torch.rand(H, generator=g)
    dy = torch.randn(B, H, generator=g)
    rm_d, rv_d = rm.clone().cuda(), rv.clone().cuda()
    mean, var = K.bn_stats(x.cuda(), rm_d, rv_d, 0.1)
    xd = x.double().numpy()
    assert np.allclose(mean.cpu().numpy(), xd.mean(0), rtol=1e-5, atol=1e-6)
    assert np.allclose(var.cpu().numpy(), xd.var(0), rtol=1e-4, atol=1e-6)
    assert np.allclose(rm_d.cpu().numpy(), 0.9 * rm.numpy() + 0.1 * xd.mean(0), rtol=1e-5, atol=1e-6)
    assert np.allclose(rv_d.cpu().numpy(), 0.9 * rv.numpy() + 0.1 * xd.var(0) * B / max(B - 1, 1), rtol=1e-4, atol=1e-6)
    for p, seed, relu, training in ((0.1, 77, True, True), (0.0, 0, True, True), (0.3, 5, False, True), (0.0, 0, True, False)):
        m_, v_ = (mean, var) if training else (rm.cuda(), rv.cuda())
        y = K.bn_act_fwd(x.cuda(), m_, v_, gamma.cuda(), beta.cuda(), 1e-5, p, seed, relu).cpu().numpy()
        keep = MO.keep_mask(seed, (B, H), p)
        md, vd = m_.cpu().double().numpy(), v_.cpu().double().numpy()
        rstd = 1 / np.sqrt(vd + 1e-5)
        xhat = (xd - md) * rstd
        z = np.where(keep, (xhat * gamma.double().numpy() + beta.double().numpy()) / (1 - p), 0.0)
        want = np.maximum(z, 0) if relu else z
        assert np.abs(y - want).max() < 1e-5 * max(np.abs(want).max(), 1)
        if p > 0:
            assert np.array_equal(y == 0, ~keep | (want == 0))                  # the mask itself, bit exact
        dx, dga, dbe = K.bn_act_bwd(x.cuda(), m_, v_, gamma.cuda(), beta.cuda(), dy.cuda(), 1e-5, p, seed, relu, training)
        gq = np.where(keep & ((z > 0) | (not relu)), dy.double().numpy() / (1 - p), 0.0)
        wb, wg = gq.sum(0), (gq * xhat).sum(0)
        gg = gq - wb / B - xhat * wg / B if training else gq
        wx = gamma.double().numpy() * rstd * gg
        assert np.abs(dbe.cpu().numpy() - wb).max() < 1e-4 * max(np.abs(wb).max(), 1)
        assert np.abs(dga.cpu().numpy() - wg).max() < 1e-4 * max(np.abs(wg).max(), 1)
        assert np.abs(dx.cpu().numpy() - wx).max() < 1e-4 * max(np.abs(wx).max(), 1)


@pytest.mark.gpu
def test_mlp_eval_and_training_equal_the_reference(cuda_pkg, golden):
    G = golden("mlp")
    m = load_model(cuda_pkg, G)
    x = torch.from_numpy(G["x"])
    m.eval()
    with torch.no_grad():
        y = m(x).cpu().numpy()
    assert y.shape == G["y_eval"].shape and np.abs(y - G["y_eval"]).max() < 2e-5 * np.abs(G["y_eval"]).max()
    m.train()
    for layer in m.layers:
        if isinstance(layer, torch.nn.Dropout):
            layer.p = 0.0                                                     # as the fixture
    xr = x.cuda().requires_grad_(True)
    y = m(xr)
    (y * torch.from_numpy(G["grad_out"]).cuda()).sum().backward()
    assert np.abs(y.detach().cpu().numpy() - G["y_train_p0"]).max() < 2e-5 * np.abs(G["y_train_p0"]).max()
    assert np.abs(xr.grad.cpu().numpy() - G["grad_x"]).max() < 1e-4 * np.abs(G["grad_x"]).max()
    for k, p in m.named_parameters():
        ref = G["grad." + k]
        assert np.abs(p.grad.cpu().numpy() - ref).max() < 1e-4 * max(np.abs(ref).max(), 1e-2 * grad_scale(G)), k
    for k, v in m.state_dict().items():
        if "running" in k:
            assert np.allclose(v.cpu().numpy(), G["after." + k], rtol=1e-5, atol=1e-6), k
        if "num_batches" in k:
            assert int(v) == int(G["after." + k])


@pytest.mark.gpu
def test_mlp_training_with_dropout_equals_the_oracle(cuda_pkg, golden):
    G = golden("mlp")
    m = load_model(cuda_pkg, G)
    st = golden_state(G)
    m.train()
    m.dropout_seed = 424242
    xr = torch.from_numpy(G["x"]).cuda().requires_grad_(True)
    y = m(xr)
    w = torch.from_numpy(G["grad_out"]).cuda()
    (y * w).sum().backward()
    yo, _, cache = MO.forward(st, G["x"], training=True, p=0.1, seed=424242, return_cache=True)
    assert np.abs(y.detach().cpu().numpy() - yo).max() < 2e-5 * np.abs(yo).max()
    grads, gx = MO.backward(st, cache, G["grad_out"])
    assert np.abs(xr.grad.cpu().numpy() - gx).max() < 1e-4 * np.abs(gx).max()
    gs = max(float(np.abs(v).max()) for v in grads.values())
    for k, p in m.named_parameters():
        assert np.abs(p.grad.cpu().numpy() - grads[k]).max() < 1e-4 * max(np.abs(grads[k]).max(), 1e-2 * gs), k
    # torch's generator drives the seed when none is fixed: two calls draw different masks, a re-seeded call repeats
    m.dropout_seed = None
    torch.manual_seed(1); a = m(xr.detach()); b = m(xr.detach()); torch.manual_seed(1); c = m(xr.detach())
    assert not torch.equal(a, b) and torch.equal(a, c)


@pytest.mark.gpu
def test_mlp_challenge_size(cuda_pkg):
    """BASELINE configs[3] shape: 12 720 LR features in (160 nodes), 268 x 268 = 71 824 out (the 35 778-feature
    output of the config line cannot be unflattened, models/mlp.py:9,27 -- same failure as the reference)."""
    torch.manual_seed(0)
    m = cuda_pkg.SuperResMLP(12720, 268 * 268, 256, 2).cuda().eval()
    x = torch.rand(64, 12720)
    with torch.no_grad():
        y = m(x)
        ref = x.cuda().double()
        for layer in m.layers:
            if isinstance(layer, torch.nn.Linear):
                ref = ref @ layer.weight.double().t() + layer.bias.double()
            elif isinstance(layer, torch.nn.BatchNorm1d):
                ref = (ref - layer.running_mean.double()) / torch.sqrt(layer.running_var.double() + layer.eps) * layer.weight.double() + layer.bias.double()
            elif isinstance(layer, torch.nn.ReLU):
                ref = ref.clamp_min(0)
    assert y.shape == (64, 268, 268)
    assert float((y.double().reshape(64, -1) - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    bad = cuda_pkg.SuperResMLP(12720, 35778, 64, 1).cuda().eval()
    with pytest.raises(RuntimeError):
        bad(x)


@pytest.mark.gpu
def test_batch_norm_contract_matches_torch(cuda_pkg):
    """a single training sample raises as nn.BatchNorm1d does (its batch variance is undefined), and
    momentum=None is refused loudly instead of being silently read as 0."""
    F_ = cuda_pkg.functional
    x = torch.randn(1, 8, device="cuda")
    g, b, rm, rv = torch.ones(8, device="cuda"), torch.zeros(8, device="cuda"), torch.zeros(8, device="cuda"), torch.ones(8, device="cuda")
    with pytest.raises(ValueError):
        F_.batch_norm_act(x, g, b, rm, rv, training=True)
    with pytest.raises(NotImplementedError):
        F_.batch_norm_act(torch.randn(4, 8, device="cuda"), g, b, rm, rv, training=True, momentum=None)
    y = F_.batch_norm_act(x, g, b, rm, rv, training=False)          # eval mode: one sample is fine
    assert torch.allclose(y, x / torch.sqrt(torch.tensor(1.0 + 1e-5)), atol=1e-6)


# ------------------------------------------------------------------ the notebook MLP (BASELINE configs[3]: 12 720 -> 35 778)
def _notebook_model(pkg, G):
    c = {k.split(".", 1)[1]: int(G[k]) for k in G.files if k.startswith("case.")}
    torch.manual_seed(c["seed"])
    model = pkg.SuperResMLPNotebook(c["nodes_in"], c["nodes_out"], c["hidden"], c["n_layers"])
    return model, c


def test_notebook_mlp_keys_and_seeded_init_equal_the_reference(pkg, golden):
    """CPU: same state_dict keys (spectral-norm ``weight_orig`` / ``weight_u`` / ``weight_v`` included) and, built from the
    same seed, the same initial parameters as the unmodified notebook class -- including its quirk that the Xavier call
    lands on the derived ``weight`` and ``weight_orig`` keeps nn.Linear's default initialisation."""
    G = golden("mlp_notebook")
    model, c = _notebook_model(pkg, G)
    ref_keys = [k[len("state."):] for k in G.files if k.startswith("state.")]
    assert list(model.state_dict().keys()) == ref_keys
    for k, v in model.state_dict().items():
        if k.endswith("weight_orig") or k.endswith("weight_u") or k.endswith("weight_v") or k.endswith("_mask"):
            assert np.array_equal(v.numpy(), G["state." + k]), k
    assert (c["nodes_in"] * (c["nodes_in"] - 1)) // 2 == model.input_layer[1].in_features
    with pytest.raises(RuntimeError):
        model(torch.zeros(2, c["nodes_in"], c["nodes_in"]))    # parameters on the CPU: loud, no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(1024, 12720), (37, 50), (130, 1024)])
def test_spectral_scale_against_torch_autograd(cuda_pkg, rows, cols):
    """W / sigma with sigma = u . (W v) and its gradient (gsr_scale_div_f32, gsr_spectral_bwd_f32) against torch's own
    float64 autograd of the same expression (torch.nn.utils.spectral_norm.compute_weight)."""
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(rows + cols)
    W = torch.randn(rows, cols, generator=g) * 0.05
    u = torch.nn.functional.normalize(torch.randn(rows, generator=g), dim=0)
    v = torch.nn.functional.normalize(torch.randn(cols, generator=g), dim=0)
    G = torch.randn(rows, cols, generator=g)
    Wc = W.cuda().requires_grad_(True)
    out = F_.SpectralScaleFn.apply(Wc, u.cuda(), v.cuda())
    out.backward(G.cuda())
    Wd = W.double().requires_grad_(True)
    ref = Wd / torch.dot(u.double(), torch.mv(Wd, v.double()))
    ref.backward(G.double())
    assert float((out.detach().double().cpu() - ref.detach()).abs().max()) <= 2e-6 * float(ref.detach().abs().max())
    assert float((Wc.grad.double().cpu() - Wd.grad).abs().max()) <= 2e-5 * float(Wd.grad.abs().max())
    out2 = F_.SpectralScaleFn.apply(Wc, u.cuda(), v.cuda())
    Wc.grad = None
    out2.backward(G.cuda())
    assert torch.equal(out2, out)                                # deterministic, two-stage fp64 sum included
    again = Wc.grad.clone(); Wc.grad = None
    F_.SpectralScaleFn.apply(Wc, u.cuda(), v.cuda()).backward(G.cuda())
    assert torch.equal(Wc.grad, again)


@pytest.mark.gpu
def test_notebook_mlp_matches_the_reference_fixture(cuda_pkg, golden):
    """Eval output, training-mode (p = 0) output, every parameter gradient, the updated BatchNorm statistics and the
    spectral-norm power-iteration buffers against the unmodified notebook class (oracle/make_golden_mlp_notebook.py)."""
    G = golden("mlp_notebook")
    model, c = _notebook_model(cuda_pkg, G)
    model.load_state_dict({k[len("state."):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("state.")})
    model.cuda()
    x = torch.from_numpy(G["x"]).cuda()
    model.eval()
    n0 = cuda_pkg.launch_count()
    with torch.no_grad():
        y = model(x)
    assert cuda_pkg.launch_count() - n0 >= 8                    # triu, 4 linears, BatchNorm passes, residual adds, sigmoid, anti-vectorise
    assert y.shape == (c["batch"], c["nodes_out"], c["nodes_out"])
    assert torch.equal(y, y.transpose(1, 2)) and float(y.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    # the fixture's logits reach |z| ~ 80 (sigma is estimated by ONE power iteration at construction, so W / sigma is
    # large): an fp32-rounding-level error of 2e-7 |z| in the logit is up to 0.25 * 1.6e-5 = 4e-6 after the sigmoid
    assert np.allclose(y.cpu().numpy(), G["y_eval"], rtol=2e-4, atol=1e-5)
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    y = model(x)
    (y * torch.from_numpy(G["grad_out"]).cuda()).sum().backward()
    assert np.allclose(y.detach().cpu().numpy(), G["y_train_p0"], rtol=2e-4, atol=1e-5)
    checked = 0
    # (the Linear biases in front of a training-mode BatchNorm have a mathematically ZERO gradient -- 1e-6-level noise
    # on both sides -- hence the absolute term, scaled by the largest gradient norm of the model)
    scale = max(float(np.linalg.norm(G[k])) for k in G.files if k.startswith("grad.") and k != "grad_out")
    for k, p in model.named_parameters():
        ref = G["grad." + k]
        got = p.grad.cpu().numpy()
        assert np.linalg.norm(got - ref) <= 1e-3 * np.linalg.norm(ref) + 1e-5 * scale, k
        checked += 1
    assert checked == len([k for k in G.files if k.startswith("grad.") and k != "grad_out"])
    for k, v in model.state_dict().items():
        if "after." + k in G.files:
            assert np.allclose(v.cpu().numpy(), G["after." + k], rtol=1e-5, atol=1e-6), k
    # dropout active: the counter-based mask is reproducible from the seed and keeps ~90 % of the activations
    model.dropout_seed = 5
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.1
    snap = {k: v.clone() for k, v in model.state_dict().items()}      # a training forward moves u / v and the statistics
    with torch.no_grad():
        a = model(x)
        model.load_state_dict(snap)
        b = model(x)
        model.load_state_dict(snap)
        model.dropout_seed = 6
        c2 = model(x)
    assert torch.equal(a, b) and not torch.equal(a, c2)


@pytest.mark.gpu
def test_notebook_mlp_full_size_against_torch_eager(cuda_pkg):
    """BASELINE configs[3] shape: 160-node LR -> 268-node HR, 12 720 -> 256 -> 35 778 features.  The same module tree run
    through torch's own eager kernels (the reference's execution: nn.Sequential calls) is the fp32 reference."""
    torch.manual_seed(4)
    model = cuda_pkg.SuperResMLPNotebook(160, 268, 256, 2).cuda()
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(64, 160, 160, device="cuda", generator=g)
    x = (x + x.transpose(1, 2)) / 2
    # a freshly constructed model has done ONE power iteration: sigma is far too small, W / sigma blows the logits up to
    # ~1e6 and every sigmoid saturates.  Twenty training-mode forwards (torch's own modules) let the spectral norms
    # converge, as the first steps of the reference's training do; the comparison is then made on a sane model.
    model.train()
    with torch.no_grad():
        for _ in range(20):
            v = model.input_layer(x[:, model.input_mask])
            for block in model.residual_blocks:
                v = torch.nn.functional.leaky_relu(block(v) + v, negative_slope=0.01)
            model.output_layer(v)
    model.eval()
    with torch.no_grad():
        y = model(x)
        v = x[:, model.input_mask]                                                    # training_template.ipynb:166-187 in torch
        v = model.input_layer(v)
        for block in model.residual_blocks:
            v = torch.nn.functional.leaky_relu(block(v) + v, negative_slope=0.01)
        v = model.output_layer(v)
        ref = torch.zeros(64, 268, 268, device="cuda")
        ref[:, model.output_mask] = v
        ref = ref + ref.transpose(1, 2)
    assert y.shape == (64, 268, 268) and model.output_layer[0].out_features == 35778 and model.input_layer[1].in_features == 12720
    assert float((y - ref).abs().max()) < 5e-5
    # training step: gradients of the three GEMM weights against torch autograd through the same modules
    model.train()
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    w = torch.randn(64, 268, 268, device="cuda", generator=g)
    (model(x) * w).sum().backward()
    got = {k: p.grad.clone() for k, p in model.named_parameters()}
    gscale = max(float(v.norm()) for v in got.values())
    model.load_state_dict(sd)                                                          # same power-iteration start
    model.zero_grad()
    v = x[:, model.input_mask]
    v = model.input_layer(v)
    for block in model.residual_blocks:
        v = torch.nn.functional.leaky_relu(block(v) + v, negative_slope=0.01)
    v = model.output_layer(v)
    ref = torch.zeros(64, 268, 268, device="cuda")
    ref[:, model.output_mask] = v
    ((ref + ref.transpose(1, 2)) * w).sum().backward()
    for k, p in model.named_parameters():
        assert float((got[k] - p.grad).norm()) <= 1e-3 * float(p.grad.norm()) + 1e-5 * gscale, k
