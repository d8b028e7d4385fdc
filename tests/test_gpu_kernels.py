"""GPU parity tests, kernel level: every libgsr_sm100 entry point against the CPU oracle
(oracle/*.py, pinned to the reference by tests/test_oracle_golden.py) or, for the plain
contractions, against a float64 torch product.  Run on the B200 box with ``-m gpu``.

Tolerances (fp32): bit-exact for the vectoriser, top-k indices and gathers; 1e-6-level
absolute / relative elsewhere, stated at each assert.
"""
import hashlib

import numpy as np
import pytest
import torch

from oracle import gsr_oracle as O
from oracle import vectorizer_oracle as O_vec
from oracle import vectorizer_oracle as VO

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dev(x):
    return torch.as_tensor(x).cuda()


@pytest.fixture(scope="module")
def K(cuda_pkg):
    return cuda_pkg.kernels


@pytest.fixture(params=["simt", "tc"])
def gemm_backend(request, cuda_pkg):
    """Every contraction test runs on the SIMT FP32 kernel and on the tcgen05 3xTF32 kernel."""
    Kmod = cuda_pkg.kernels
    old, old_rows, old_lin = Kmod.GEMM_BACKEND, Kmod.TC_MIN_ROWS, Kmod.LIN_MAX_OUT
    Kmod.set_gemm_backend(request.param)
    Kmod.TC_MIN_ROWS = 0
    if request.param == "tc":
        Kmod.LIN_MAX_OUT = 0          # one-subject products too: the tensor kernels keep their small-shape coverage
    yield request.param
    Kmod.set_gemm_backend(old)
    Kmod.TC_MIN_ROWS, Kmod.LIN_MAX_OUT = old_rows, old_lin


# ------------------------------------------------------------------ vectoriser (bit exact)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 33, 160, 268])
@pytest.mark.parametrize("diag", [False, True])
def test_vectorize_bit_exact(cuda_pkg, n, diag):
    MV = cuda_pkg.MatrixVectorizer
    rng = np.random.default_rng(n)
    for dt in (np.float64, np.float32, np.int64, np.int16, np.uint8):
        m = (rng.random((n, n)) * 100).astype(dt)
        got = MV.vectorize(m, include_diagonal=diag)
        want = VO.vectorize(m, include_diagonal=diag)
        assert got.dtype == want.dtype and got.shape == want.shape, (n, dt)
        assert np.array_equal(got, want), (n, dt)


def test_vectorize_golden_digests(cuda_pkg, golden):
    G = golden("vectorizer")
    MV = cuda_pkg.MatrixVectorizer
    for n in (160, 268):
        m = G[f"vec_n{n}_in"]
        assert sha(MV.vectorize(m)) == str(G[f"vec_n{n}_f32_d0_sha"])
        assert sha(MV.vectorize(m, include_diagonal=True)) == str(G[f"vec_n{n}_f32_d1_sha"])
    for key in G.files:
        if key.startswith("vec_") and key.endswith("_in") and "_d" in key:
            diag = key.split("_d")[1][0] == "1"
            got = MV.vectorize(G[key], include_diagonal=diag)
            assert got.dtype == G[key[:-3] + "_out"].dtype
            assert np.array_equal(got, G[key[:-3] + "_out"]), key


def test_vectorize_batched_device(K):
    x = torch.rand(7, 268, 268, device="cuda")
    for diag in (False, True):
        got = K.vectorize(x, diag).cpu().numpy()
        for b in range(7):
            assert np.array_equal(got[b], VO.vectorize(x[b].cpu().numpy(), diag))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 160, 268])
def test_anti_vectorize_bit_exact(cuda_pkg, golden, n):
    G = golden("vectorizer")
    MV = cuda_pkg.MatrixVectorizer
    v = torch.from_numpy(G[f"anti_n{n}_in"])
    for diag in (False, True):
        for src in (v, v.cuda()):
            got = MV.anti_vectorize(src, n, include_diagonal=diag)
            assert got.dtype == torch.float32 and got.device == src.device
            got = got.cpu().numpy()
            assert np.array_equal(got, VO.anti_vectorize(v.numpy(), n, diag))
            if n <= 8:
                assert np.array_equal(got, G[f"anti_n{n}_d{int(diag)}_out"])
            else:
                assert sha(got) == str(G[f"anti_n{n}_d{int(diag)}_sha"])


def test_anti_vectorize_errors(cuda_pkg):
    MV = cuda_pkg.MatrixVectorizer
    with pytest.raises(AssertionError):
        MV.anti_vectorize(torch.zeros(5), 5)
    with pytest.raises(TypeError):
        MV.anti_vectorize(np.zeros(10), 5)
    with pytest.raises(RuntimeError):
        MV.anti_vectorize(torch.zeros(10, dtype=torch.float64), 5)


def test_vectorizer_not_inverse_and_roundtrip_property(K):
    # the two maps use different orders ...
    v = torch.arange(10, dtype=torch.float32, device="cuda")
    m = K.anti_vectorize(v, 5)
    assert not torch.equal(K.vectorize(m), v)
    # ... but both are permutations of the strict upper triangle: full-size property
    x = torch.rand(64, 35778, device="cuda")
    m = K.anti_vectorize(x, 268)
    assert torch.equal(m, m.transpose(1, 2)) and float(m.diagonal(dim1=1, dim2=2).abs().max()) == 0.0
    back = K.vectorize(m)
    assert torch.equal(back.sort(dim=1).values, x.sort(dim=1).values)


# ------------------------------------------------------------------ normalise
@pytest.mark.parametrize("tag", ["sym12", "sym160", "nonsym9", "zerorow6", "negrow5"])
def test_normalize_golden(cuda_pkg, golden, tag):
    G = golden("ops")
    got = cuda_pkg.normalize_adj_torch(dev(G[f"norm_{tag}_in"])).cpu().numpy()
    want = G[f"norm_{tag}_out"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got, want, rtol=2e-6, atol=1e-8, equal_nan=True)
    if tag == "zerorow6":
        assert np.all(got[2] == 0) and np.all(got[:, 2] == 0)


@pytest.mark.parametrize("n,batch", [(10, 1), (160, 3), (268, 2), (37, 5)])
def test_normalize_fwd_bwd_vs_oracle(cuda_pkg, n, batch):
    g = torch.Generator().manual_seed(n)
    mx = torch.rand(batch, n, n, generator=g)
    T = torch.randn(batch, n, n, generator=g)
    ref_in = mx.clone().requires_grad_(True)
    ref = torch.stack([O.normalize_adj(ref_in[b]) for b in range(batch)])
    (ref * T).sum().backward()
    x = mx.cuda().requires_grad_(True)
    out = cuda_pkg.normalize_adj_torch(x)
    (out * T.cuda()).sum().backward()
    assert torch.allclose(out.cpu(), ref.detach(), rtol=2e-6, atol=1e-9)
    scale = float(ref_in.grad.abs().max())
    assert torch.allclose(x.grad.cpu(), ref_in.grad, rtol=0, atol=2e-6 * scale)


def test_normalize_pre_abs_diag(K):
    g = torch.Generator().manual_seed(3)
    F = torch.randn(2, 268, 268, generator=g)
    T = torch.randn(2, 268, 268, generator=g)
    ref_in = F.clone().requires_grad_(True)
    outs = []
    for b in range(2):
        fd = torch.abs(ref_in[b])
        fd = fd.fill_diagonal_(1)
        outs.append(O.normalize_adj(fd))
    ref = torch.stack(outs)
    (ref * T).sum().backward()
    out, r, m = K.normalize_adj_fwd(F.cuda(), pre=True, want_m=True)
    assert torch.allclose(out.cpu(), ref.detach(), rtol=2e-6, atol=1e-9)
    fd_ref = F.abs()
    fd_ref.diagonal(dim1=1, dim2=2).fill_(1)
    assert torch.equal(m.cpu(), fd_ref)
    assert torch.equal(K.abs_diag1(F.cuda()).cpu(), fd_ref)
    grad = K.normalize_adj_bwd(F.cuda(), r, T.cuda(), pre=True)
    scale = float(ref_in.grad.abs().max())
    assert torch.allclose(grad.cpu(), ref_in.grad, rtol=0, atol=2e-6 * scale)
    assert float(grad.diagonal(dim1=1, dim2=2).abs().max()) == 0.0


# ------------------------------------------------------------------ eigh
def _eigh_checks(A32, w, U, tol_w=2e-6, tol_res=5e-6, tol_orth=1e-5):
    A = A32.double().numpy()
    A = np.triu(A) + np.triu(A, 1).T                      # only the upper triangle is read
    w64 = np.linalg.eigvalsh(A)
    scale = max(1.0, np.abs(w64).max())
    w, U = w.double().numpy(), U.double().numpy()
    assert np.all(np.diff(w) >= 0)
    assert np.abs(w - w64).max() <= tol_w * scale
    assert np.abs(A @ U - U * w).max() <= tol_res * scale
    assert np.abs(U.T @ U - np.eye(A.shape[0])).max() <= tol_orth
    # sign rule: the largest-|.| component of every column is positive
    piv = np.abs(U).argmax(0)
    assert np.all(U[piv, np.arange(U.shape[1])] > 0)


@pytest.mark.parametrize("n", [1, 2, 3, 12, 31, 64, 136, 160, 161, 224])
def test_eigh_properties(cuda_pkg, n):
    F_ = cuda_pkg.functional
    A = torch.stack([O.normalize_adj(O.synth_connectome(n, 50 + s)) if n > 1 else torch.tensor([[0.7]])
                     for s in range(3)])
    w, U, info = F_.eigh(A.cuda())
    assert int(info.min()) > 0
    for b in range(3):
        _eigh_checks(A[b], w[b].cpu(), U[b].cpu())


def test_eigh_hr_target_size_global_workspace(cuda_pkg):
    """n = 268 (utils/gsr_train.py:34) does not fit one CTA's shared memory: global-workspace path."""
    F_ = cuda_pkg.functional
    hrs = torch.stack([O.pad_hr_adj(O.synth_connectome(268, 1500 + s), 0) for s in range(3)])
    w, U, info = F_.eigh(hrs.cuda())
    assert int(info.min()) > 0
    for b in range(3):
        _eigh_checks(hrs[b], w[b].cpu(), U[b].cpu(), tol_w=3e-6, tol_res=1e-5, tol_orth=1e-5)
    # hr_target_basis: the 160 eigenvectors of the smallest eigenvalues, as the oracle picks them
    got = cuda_pkg.hr_target_basis(O.synth_connectome(268, 1500).cuda(), 160).cpu()
    ref = O.hr_target_basis(O.synth_connectome(268, 1500), 160)
    assert got.shape == (268, 160)
    assert float((got.abs() - ref.abs()).abs().mean()) < 1e-3          # loss-level agreement (mean |.|)
    l_got = float((torch.randn(268, 160, generator=torch.Generator().manual_seed(0)) - got).abs().mean())
    l_ref = float((torch.randn(268, 160, generator=torch.Generator().manual_seed(0)) - ref).abs().mean())
    assert abs(l_got - l_ref) < 2e-5
    for n in (225, 256, 288):
        A = O.normalize_adj(O.synth_connectome(n, n))
        w, U, info = F_.eigh(A.cuda())
        _eigh_checks(A, w.cpu(), U.cpu(), tol_w=3e-6, tol_res=1e-5)


def test_eigh_generic_symmetric_and_upper_only(cuda_pkg):
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(0)
    M = torch.randn(4, 96, 96, generator=g)
    A = (M + M.transpose(1, 2)) / 2
    junk = A.clone()
    junk += torch.tril(torch.randn(96, 96, generator=g), -1)       # lower triangle must be ignored
    w, U, _ = F_.eigh(junk.cuda())
    for b in range(4):
        _eigh_checks(A[b], w[b].cpu(), U[b].cpu(), tol_w=3e-6, tol_res=1e-5)


def test_eigh_vs_oracle_vectors(cuda_pkg):
    """Canonicalised eigenvectors against LAPACK's (oracle): compared where the spectrum is
    well separated (gap > 1e-3), since eigenvectors of close pairs are ill conditioned in
    fp32 for BOTH implementations."""
    F_ = cuda_pkg.functional
    A = O.normalize_adj(O.synth_connectome(160, 1000))
    w_ref, U_ref = O.eigh_canonical(A)
    w, U, _ = F_.eigh(A.cuda())
    w, U = w.cpu(), U.cpu()
    assert torch.allclose(w, w_ref, rtol=0, atol=2e-6)
    gaps = torch.minimum(torch.diff(w_ref, prepend=w_ref[:1] - 1), torch.diff(w_ref, append=w_ref[-1:] + 1))
    sep = gaps > 1e-3
    assert int(sep.sum()) >= 10
    assert torch.allclose(U[:, sep], U_ref[:, sep], rtol=0, atol=2e-3)
    # subspace agreement for everything: |U^T U_ref| is a permutation-free near-identity in blocks
    P = (U.double().T @ U_ref.double()).abs()
    assert float((P.max(0).values).min()) > 0.5


def test_eigh_error_contract(cuda_pkg):
    F_ = cuda_pkg.functional
    A = O.normalize_adj(O.synth_connectome(32, 1)).repeat(2, 1, 1)
    A[1, 3, 5] = float("nan")
    with pytest.raises(torch.linalg.LinAlgError):
        F_.eigh(A.cuda())
    w, U, info = F_.eigh(A.cuda(), check_errors=False)
    assert info.tolist()[1] == -1 and info.tolist()[0] > 0
    assert bool(torch.isnan(U[1]).all()) and not bool(torch.isnan(U[0]).any())
    with pytest.raises(cuda_pkg.GsrLibraryError):
        cuda_pkg.kernels.eigh_jacobi(torch.zeros(1, 400, 400, device="cuda"))      # too large: loud, no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 7, 16, 17, 33, 96, 128, 159, 160, 161, 200, 268, 288])
def test_eigh_both_routes_agree(cuda_pkg, n):
    """The tridiagonal-QL kernels (default route; register-resident up to 160 nodes, shared-memory variants up
    to 288) and the Jacobi kernel decompose the same
    matrices; both are held to the float64 properties and their eigenvalues to each other."""
    K = cuda_pkg.kernels
    A = torch.stack([O.normalize_adj(O.synth_connectome(n, 900 + s)) if n > 1 else torch.tensor([[-0.3]])
                     for s in range(5)]).cuda()
    wt, Ut, it = K.eigh_tridiag(A)
    wj, Uj, ij = K.eigh_jacobi(A)
    assert int(it.min()) > 0 and int(ij.min()) > 0
    assert torch.allclose(wt, wj, rtol=0, atol=3e-6)
    for b in range(5):
        _eigh_checks(A[b].cpu(), wt[b].cpu(), Ut[b].cpu(), tol_w=3e-6, tol_res=5e-6, tol_orth=1e-5)
    assert cuda_pkg.kernels.eigh(A)[2].equal(it)                       # "auto" = tridiagonal route up to 288 nodes
    with pytest.raises(ValueError):
        K.eigh_tridiag(torch.zeros(1, 289, 289, device="cuda"))


@pytest.mark.gpu
def test_eigh_ql_failure_falls_back_to_jacobi(cuda_pkg):
    """info = -2 (QL gave up) never reaches a caller: gsr_eigh_tridiag_f32 redoes those graphs with the Jacobi
    kernel on the device, in the same stream, with or without error checks (training steps and CUDA graphs run
    without them).  The failure is provoked by allowing one QL iteration per eigenvalue; both the
    register-resident (n <= 160) and the shared-memory (n = 268, global Jacobi scratch) variants are covered,
    and a batch where only SOME graphs fail (the others must keep their QL result bit for bit)."""
    F_, K = cuda_pkg.functional, cuda_pkg.kernels
    for n, count in ((64, 4), (160, 3), (268, 2)):
        A = torch.stack([O.normalize_adj(O.synth_connectome(n, 40 + s)) for s in range(count)]).cuda()
        w_ok, U_ok, info_ok = K.eigh_tridiag(A)
        assert int(info_ok.min()) > 0
        cuda_pkg.set_tuning("eigh_ql_max_iter", 1)
        try:
            cuda_pkg.set_tuning("eigh_redo_failed", 0)
            w, U, info = K.eigh_tridiag(A)
            assert int(info.max()) == -2 and bool(torch.isnan(U).all())      # raw QL contract, fallback off
            cuda_pkg.set_tuning("eigh_redo_failed", 1)
            w, U, info = K.eigh_tridiag(A)                                    # no sync, no Python-side branch
            w2, U2, info2 = F_.eigh(A, check_errors=False)
        finally:
            cuda_pkg.set_tuning("eigh_ql_max_iter", 30)
            cuda_pkg.set_tuning("eigh_redo_failed", 1)
        assert int(info.min()) > 0 and torch.equal(U, U2) and torch.equal(w, w2)
        for b in range(count):
            # the unpolished global-scratch Jacobi variant (n > 224) is held to the looser bar of its own route
            _eigh_checks(A[b].cpu(), w[b].cpu(), U[b].cpu(), **({} if n <= 160 else dict(tol_w=3e-6, tol_res=2e-5, tol_orth=2e-5)))
        assert float((w - w_ok).abs().max()) < 3e-6
    # a diagonal matrix needs no QL iteration: it survives max_iter = 1 while its neighbours fail and are redone
    A = torch.stack([O.normalize_adj(O.synth_connectome(96, 7)), torch.diag(torch.linspace(-1, 1, 96)),
                     O.normalize_adj(O.synth_connectome(96, 8))]).cuda()
    w_ok, U_ok, _ = K.eigh_tridiag(A)
    cuda_pkg.set_tuning("eigh_ql_max_iter", 1)
    try:
        w, U, info = K.eigh_tridiag(A)
    finally:
        cuda_pkg.set_tuning("eigh_ql_max_iter", 30)
    assert int(info.min()) > 0
    assert torch.equal(U[1], U_ok[1]) and torch.equal(w[1], w_ok[1])
    for b in (0, 2):
        _eigh_checks(A[b].cpu(), w[b].cpu(), U[b].cpu())


@pytest.mark.gpu
def test_eigh_pipeline_matches_one_launch_per_stage(cuda_pkg):
    """The chunked eigensolver pipeline (batches of 384 graphs or more: ONE K1 launch counting finished graphs per
    128-graph chunk, the lane = graph QL kernel resident on its own SMs polling those counts, K3 / K4 per chunk on
    internal streams; csrc/eigh_tridiag.cu) against (a) the same kernels run one launch per stage -- bit identical --
    (b) the one-warp-per-graph QL kernel -- eigenvalues 3e-6 -- and (c) float64 LAPACK on sampled graphs; a ragged
    batch (400 = 3 chunks + 16 graphs) with a NaN graph, a zero matrix, a multiple of the identity and a rank-one
    matrix in different chunks; twice in a row (the per-chunk counters are reset by every call); 1100 graphs = two
    segments; QL failures inside the pipeline are redone by the Jacobi kernel."""
    K = cuda_pkg.kernels
    g = torch.Generator().manual_seed(11)
    B = 400
    x = torch.rand(B, 160, 160, generator=g)
    A = torch.stack([O.normalize_adj((m + m.T) / 2) for m in x])
    A[5, 3, 9] = float("nan")
    A[140] = 0.0
    A[270] = 2.5 * torch.eye(160)
    A[399] = torch.ones(160, 160) / 160
    A = A.cuda()
    try:
        cuda_pkg.set_tuning("eigh_chunks", 8); cuda_pkg.set_tuning("eigh_k2_lanes", 1)
        w_p, U_p, i_p = K.eigh_tridiag(A)
        w_p2, U_p2, i_p2 = K.eigh_tridiag(A)
        # who forms Q = H_{n-1} ... H_1: K3 in front of its rotation replay (0, row form) or a launch of its own behind K1 (1,
        # the default, column form); and K1 with 10 (the whole row), 9, 7 or 6 of a row's ten 16-column blocks in registers instead of 8
        variants = []
        for q_launch, rb in ((0, 8), (0, 10), (1, 9), (1, 7), (0, 6)):
            cuda_pkg.set_tuning("eigh_q_launch", q_launch); cuda_pkg.set_tuning("eigh_k1_regblocks", rb)
            variants.append(K.eigh_tridiag(A))
        cuda_pkg.set_tuning("eigh_q_launch", 1); cuda_pkg.set_tuning("eigh_k1_regblocks", 8)
        cuda_pkg.set_tuning("eigh_chunks", 1); cuda_pkg.set_tuning("eigh_k2_lanes", 2)       # lane = graph, one launch per stage
        w_l, U_l, i_l = K.eigh_tridiag(A)
        cuda_pkg.set_tuning("eigh_k2_lanes", 0)                                              # one warp per graph
        w_w, U_w, i_w = K.eigh_tridiag(A)
    finally:
        cuda_pkg.set_tuning("eigh_chunks", 8); cuda_pkg.set_tuning("eigh_k2_lanes", 1)
        cuda_pkg.set_tuning("eigh_q_launch", 1); cuda_pkg.set_tuning("eigh_k1_regblocks", 8)
    ok = torch.ones(B, dtype=torch.bool); ok[5] = False
    for (q_launch, _), (w_v, U_v, i_v) in zip(((0, 8), (0, 10), (1, 9), (1, 7), (0, 6)), variants):
        assert torch.equal(w_p[ok], w_v[ok]) and torch.equal(i_p, i_v)
        if q_launch:
            assert torch.equal(U_p[ok], U_v[ok])
        else:
            # K3 forms Q in ROW form (2 n^3 flops), the default launch in COLUMN form (4/3 n^3): the same reflectors and the
            # same rotation stream, so the bases differ by the rounding of Q only -- up to the sign rule, which may pick
            # the other of two nearly equal largest components
            sgn = torch.sign((U_p[ok] * U_v[ok]).sum(1, keepdim=True))
            assert float((U_p[ok] - sgn * U_v[ok]).abs().max()) < 2e-5
            assert float((sgn < 0).float().mean()) < 1e-3
    assert int(i_p[5]) == -1 and bool(torch.isnan(w_p[5]).all()) and bool(torch.isnan(U_p[5]).all())
    assert int(i_p[ok].min()) > 0
    assert torch.equal(w_p[ok], w_p2[ok]) and torch.equal(U_p[ok], U_p2[ok]) and torch.equal(i_p, i_p2)
    assert torch.equal(w_p[ok], w_l[ok]) and torch.equal(U_p[ok], U_l[ok]) and torch.equal(i_p, i_l)
    assert float((w_p[ok] - w_w[ok]).abs().max()) < 3e-6
    assert float(w_p[140].abs().max()) == 0.0 and float((w_p[270] - 2.5).abs().max()) < 1e-6
    assert abs(float(w_p[399].max()) - 1.0) < 2e-6 and float(w_p[399][:-1].abs().max()) < 2e-6
    for b in (0, 127, 128, 255, 256, 383, 384, 398):
        _eigh_checks(A[b].cpu(), w_p[b].cpu(), U_p[b].cpu())
    # two segments (1100 > 1024 graphs), QL failures redone inside the pipeline
    A2 = A[ok][:399].repeat(3, 1, 1)[:1100].contiguous()
    w2, U2, i2 = K.eigh_tridiag(A2)
    assert int(i2.min()) > 0 and torch.equal(w2[:399], w_p[ok][:399]) and torch.equal(w2[399:798], w2[:399]) and torch.equal(U2[798:1100], U2[:302])
    cuda_pkg.set_tuning("eigh_ql_max_iter", 1)
    try:
        w3, U3, i3 = K.eigh_tridiag(A2[:384])
    finally:
        cuda_pkg.set_tuning("eigh_ql_max_iter", 30)
    assert int(i3.min()) > 0 and float((w3 - w2[:384]).abs().max()) < 3e-6


def test_eigh_tridiag_special_matrices(cuda_pkg):
    """Zero, identity multiples, rank one, repeated eigenvalues, tiny and huge scales (the kernel
    rescales by an exact power of two), diagonal input, and the NaN contract."""
    K = cuda_pkg.kernels
    dev = "cuda"
    w, U, info = K.eigh_tridiag(torch.zeros(2, 40, 40, device=dev))
    assert float(w.abs().max()) == 0.0 and torch.equal(U, torch.eye(40, device=dev).expand(2, 40, 40)) and int(info.min()) > 0
    w, U, info = K.eigh_tridiag(3.0 * torch.eye(24, device=dev)[None])
    assert torch.equal(w, torch.full((1, 24), 3.0, device=dev))
    ones = torch.ones(1, 64, 64, device=dev)
    w, U, _ = K.eigh_tridiag(ones)
    _eigh_checks(ones[0].cpu(), w[0].cpu(), U[0].cpu(), tol_w=3e-6, tol_res=1e-5)          # eigenvalues 0 (x63) and 64
    g = torch.Generator().manual_seed(5)
    M = torch.randn(3, 80, 80, generator=g)
    A = (M + M.transpose(1, 2)) / 2
    for scale in (1e-20, 1.0, 1e18):
        w, U, info = K.eigh_tridiag((A * scale).cuda())
        assert int(info.min()) > 0
        for b in range(3):
            _eigh_checks(A[b] * scale, w[b].cpu(), U[b].cpu(), tol_w=4e-6, tol_res=2e-5)
    D = torch.diag(torch.tensor([3.0, -1.0, 2.0, 2.0, 0.0, 5.0]))[None].cuda()
    w, U, _ = K.eigh_tridiag(D)
    assert torch.equal(w.cpu(), torch.tensor([[-1.0, 0.0, 2.0, 2.0, 3.0, 5.0]]))
    assert torch.equal(U.abs().cpu()[0], torch.eye(6)[:, [1, 4, 2, 3, 0, 5]])
    bad = O.normalize_adj(O.synth_connectome(48, 3)).repeat(3, 1, 1)
    bad[1, 0, 7] = float("inf")
    w, U, info = K.eigh_tridiag(bad.cuda())
    assert info.tolist()[1] == -1 and info.tolist()[0] > 0 and info.tolist()[2] > 0
    assert bool(torch.isnan(U[1]).all()) and bool(torch.isnan(w[1]).all()) and not bool(torch.isnan(U[[0, 2]]).any())


def test_eigh_full_batch_properties(cuda_pkg):
    """BASELINE size: 1024 normalised 160-node graphs; size-independent properties."""
    F_ = cuda_pkg.functional
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(1024, 160, 160, device="cuda", generator=g)
    lr = (x + x.transpose(1, 2)) / 2
    lr.diagonal(dim1=1, dim2=2).zero_()
    A = cuda_pkg.normalize_adj_torch(lr)
    w, U, info = F_.eigh(A)
    assert int(info.min()) > 0 and int(info.max()) <= 30 * 160        # QL iterations (tridiagonal route, n <= 160)
    Ad, Ud, wd = A.double(), U.double(), w.double()
    As = torch.triu(Ad) + torch.triu(Ad, 1).transpose(1, 2)
    res = (As @ Ud - Ud * wd[:, None, :]).abs().amax()
    orth = (Ud.transpose(1, 2) @ Ud - torch.eye(160, device="cuda", dtype=torch.float64)).abs().amax()
    assert float(res) < 5e-6 and float(orth) < 1e-5
    assert float((wd[:, -1] - 1).abs().max()) < 5e-6          # lambda_max of D^-1/2 A D^-1/2 is 1
    tr = As.diagonal(dim1=1, dim2=2).sum(1)
    assert float((wd.sum(1) - tr).abs().max()) < 2e-5         # trace is preserved


# ------------------------------------------------------------------ GEMMs
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("shape", [(268, 268, 268), (268, 160, 160), (160, 268, 536), (101, 268, 268), (30, 7, 5), (1, 1, 1), (137, 139, 17)])
def test_bgemm_vs_float64(K, gemm_backend, tA, tB, shape):
    M, N, Kd = shape
    g = torch.Generator().manual_seed(M * 7 + N)
    for batch, sharedB in ((1, False), (3, False), (3, True)):
        A = torch.randn((batch, Kd, M) if tA else (batch, M, Kd), generator=g)
        B = torch.randn((Kd, N) if not tB else (N, Kd), generator=g) if sharedB else \
            torch.randn((batch, Kd, N) if not tB else (batch, N, Kd), generator=g)
        opA = A.transpose(-1, -2) if tA else A
        opB = B.transpose(-1, -2) if tB else B
        want = (opA.double() @ opB.double())
        got = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB).cpu()
        tol = 2e-6 * float((opA.abs().double() @ opB.abs().double()).max())
        assert got.shape == want.shape
        assert float((got.double() - want).abs().max()) <= tol, (shape, tA, tB, batch, sharedB)


def test_tcgemm_pair_kernel_equals_one_cta_kernel(cuda_pkg, K):
    """Thread-staged kernels (gsr_set_tuning("tcgemm_tma", 0)): cta_group::2 pairs (M = 256 per MMA) against one CTA per
    128-row tile -- same split, same MMA order, so the results are bit identical; the FP32 ragged rows differ from the
    tensor kernel by rounding only."""
    g = torch.Generator(device="cuda").manual_seed(31)
    A = torch.randn(7, 268, 268, device="cuda", generator=g); B = torch.randn(7, 268, 268, device="cuda", generator=g)
    W = torch.randn(268, 268, device="cuda", generator=g)
    calls = [lambda: K.bgemm(A, B), lambda: K.bgemm(A, B, transA=True), lambda: K.bgemm(A, B, transB=True),
             lambda: K.bgemm(A, W), lambda: K.bgemm(A, W, transB=True), lambda: K.bgemm_reduce(A, B, transA=True)]
    try:
        cuda_pkg.set_tuning("tcgemm_tma", 0)
        cuda_pkg.set_tuning("tcgemm_pairs", 0)
        one = [c() for c in calls]
        cuda_pkg.set_tuning("tcgemm_pairs", 1); cuda_pkg.set_tuning("tcgemm_ragged_rows", 0)
        pairs = [c() for c in calls]
        cuda_pkg.set_tuning("tcgemm_ragged_rows", 1)
        ragged = [c() for c in calls]
    finally:
        cuda_pkg.set_tuning("tcgemm_pairs", 1); cuda_pkg.set_tuning("tcgemm_ragged_rows", 1); cuda_pkg.set_tuning("tcgemm_tma", 1)
    for a, b, c in zip(one, pairs, ragged):
        assert torch.equal(a, b)
        assert torch.equal(a[..., :256, :], c[..., :256, :])
        assert float((a - c).abs().max()) < 4e-6 * float(a.abs().max())


@pytest.mark.parametrize("batch,k2", [(2, 268), (80, 268), (5, 160)])
def test_tcgemm_dual_product(cuda_pkg, K, batch, k2):
    """gsr_tcgemm_dual_f32: C = alpha (op(A1) op(B1) + op(A2) op(B2)) accumulated in tensor memory by ONE launch of the TMA
    kernel (d adj = dH2 S2^T + dH1 S1^T of the two graph convolutions, models/layers.py:85-86 twice) against the float64
    sum within 2e-6 sum|a||b| -- rows 0..255 (tensor kernel) and the ragged rows 256..267 (FP32 kernels, second launch
    adding in place) -- for NT, NN and TN operands, inner sizes that differ, and more tiles than CTA pairs; and against
    the two-launch route (tcgemm_dual = 0: second product with the addend epilogue), which differs by rounding only."""
    g = torch.Generator(device="cuda").manual_seed(500 + batch)
    A1 = torch.randn(batch, 268, 268, device="cuda", generator=g); B1 = torch.randn(batch, 268, 268, device="cuda", generator=g)
    A2 = torch.randn(batch, 268, k2, device="cuda", generator=g); B2 = torch.randn(batch, 268, k2, device="cuda", generator=g)
    d = lambda t: t.double()
    T = lambda t: t.transpose(1, 2)
    cases = [
        (dict(transB=True, alpha=1.0), (A1, B1, A2, B2), d(A1) @ T(d(B1)) + d(A2) @ T(d(B2)), d(A1).abs() @ T(d(B1).abs()) + d(A2).abs() @ T(d(B2).abs())),
        (dict(transB=True, alpha=-0.5), (A1, B1, A2, B2), -0.5 * (d(A1) @ T(d(B1)) + d(A2) @ T(d(B2))), 0.5 * (d(A1).abs() @ T(d(B1).abs()) + d(A2).abs() @ T(d(B2).abs()))),
    ]
    if k2 == 268:
        cases += [
            (dict(alpha=1.0), (A1, B1, A2, B2), d(A1) @ d(B1) + d(A2) @ d(B2), d(A1).abs() @ d(B1).abs() + d(A2).abs() @ d(B2).abs()),
            (dict(transA=True, alpha=2.0), (A1, B1, A2, B2), 2 * (T(d(A1)) @ d(B1) + T(d(A2)) @ d(B2)), 2 * (T(d(A1).abs()) @ d(B1).abs() + T(d(A2).abs()) @ d(B2).abs())),
        ]
    n0 = cuda_pkg.launch_count()
    for kw, ops, ref, bound in cases:
        got = K.bgemm_dual(*ops, **kw)
        err = (got.double() - ref).abs()
        assert float(err[:, :256].max()) <= 2e-6 * float(bound.max()), kw
        assert float(err[:, 256:].max()) <= 2e-6 * float(bound.max()), kw
        try:
            cuda_pkg.set_tuning("tcgemm_dual", 0)
            two = K.bgemm_dual(*ops, **kw)
        finally:
            cuda_pkg.set_tuning("tcgemm_dual", 1)
        assert float((two - got).abs().max()) <= 4e-6 * float(bound.max()), kw
        assert float((two.double() - ref).abs().max()) <= 2e-6 * float(bound.max()), kw
    assert cuda_pkg.launch_count() > n0
    with pytest.raises(RuntimeError):
        K.bgemm_dual(A1, B1, A2[:, :200].contiguous(), B2, transB=True)         # the two products differ in shape


@pytest.mark.parametrize("n,batch", [(268, 7), (160, 3), (64, 2), (12, 5), (300, 2), (320, 1)])
def test_normalize_bwd_cluster_kernel(cuda_pkg, K, n, batch):
    """The one-pass normalise backward (the tile pairs of a subject as a cluster of 16 CTAs exchanging their partial sums
    through distributed shared memory; csrc/elementwise.cu; a switch, measured slower than the default) is BIT identical to the three launches it replaces (partial,
    drho, apply): one and two incoming gradients, pre = 0 / 1 with negative entries, an exact zero (sgn 0) and a zero row sum
    (r = 0) in the data, and against the CPU oracle's autograd."""
    g = torch.Generator(device="cuda").manual_seed(700 + n)
    mx = torch.randn(batch, n, n, device="cuda", generator=g)
    mx[0, 1, 0] = 0.0
    pos = mx.abs() + 0.05
    if n > 4:
        pos[batch - 1, 3, :] = 0.0                                  # a zero row: r = 0 (the reference's inf -> 0 rule)
    g1 = torch.randn(batch, n, n, device="cuda", generator=g); g2 = torch.randn(batch, n, n, device="cuda", generator=g)
    for pre, x in ((True, mx), (False, pos)):
        _, r, _ = K.normalize_adj_fwd(x, pre=pre)
        got = {}
        for cl in (0, 1):
            cuda_pkg.set_tuning("normalize_bwd_cluster", cl)
            try:
                got[cl] = (K.normalize_adj_bwd(x, r, g1, pre=pre), K.normalize_adj_bwd(x, r, g1, pre=pre, grad_out2=g2))
            finally:
                cuda_pkg.set_tuning("normalize_bwd_cluster", 0)
        assert torch.equal(got[0][0], got[1][0]) and torch.equal(got[0][1], got[1][1]), (n, pre)
    if batch > 1:                                                   # subject 0 has no zero row (torch's autograd gives NaN there)
        xr = pos[0].cpu().clone().requires_grad_(True)
        (O.normalize_adj(xr) * g1[0].cpu()).sum().backward()
        _, r, _ = K.normalize_adj_fwd(pos[:1], pre=False)
        assert torch.allclose(K.normalize_adj_bwd(pos[:1], r, g1[:1])[0].cpu(), xr.grad, rtol=1e-5, atol=2e-6)


def test_normalize_bwd_two_gradients(cuda_pkg, K):
    """gsr_normalize_adj_bwd_tiles2: the incoming gradient given as two tensors, summed as the passes read them, is BIT
    identical to the backward of their materialised fp32 sum (pre = 0 and pre = 1; 268 nodes = the tile path, 30 nodes =
    the wrapper adds first and takes the strip kernel)."""
    for n, batch in ((268, 6), (160, 3), (30, 4)):
        g = torch.Generator(device="cuda").manual_seed(900 + n)
        mx = torch.rand(batch, n, n, device="cuda", generator=g) + 0.05      # positive row sums (r = rowsum^-1/2)
        g1 = torch.randn(batch, n, n, device="cuda", generator=g); g2 = torch.randn(batch, n, n, device="cuda", generator=g)
        for pre in (False, True):
            _, r, _ = K.normalize_adj_fwd(mx, pre=pre)
            want = K.normalize_adj_bwd(mx, r, g1 + g2, pre=pre)
            got = K.normalize_adj_bwd(mx, r, g1, pre=pre, grad_out2=g2)
            assert torch.equal(got, want), (n, pre)
    with pytest.raises(ValueError):
        K.normalize_adj_bwd(mx, r, g1, grad_out2=g2[:, :10])


@pytest.mark.parametrize("batch", [1, 3, 80])
def test_tcgemm_tma_kernel(cuda_pkg, K, batch):
    """The TMA-fed persistent CTA-pair kernel (csrc/tcgemm_tma.cu; default route for M >= 256, N <= 272): every operand
    layout, epilogue and the subject reduction against the float64 product within 2e-6 sum|a||b| (its hi term is the
    TRUNCATED tf32 of the operand -- kind::tf32 ignores the low mantissa bits of the raw tile -- so it is not bit
    identical to the thread-staged kernels, which round); its switches (tensor-memory A operand via tcgen05.cp, whole
    32-column blocks for MN-major B, L2 prefetch) change the data path, not one bit of the result; more tiles than CTA
    pairs (batch 80 -> a second round of the persistent loop with the other accumulator)."""
    g = torch.Generator(device="cuda").manual_seed(77 + batch)
    A = torch.randn(batch, 268, 268, device="cuda", generator=g); B = torch.randn(batch, 268, 268, device="cuda", generator=g)
    W = torch.randn(268, 268, device="cuda", generator=g); Wl = torch.randn(268, 160, device="cuda", generator=g)
    U = torch.randn(batch, 160, 160, device="cuda", generator=g); add = torch.randn(batch, 268, 268, device="cuda", generator=g)
    bias = torch.randn(268, device="cuda", generator=g)
    Ad, Bd, Wd = A.double(), B.double(), W.double()
    cases = [
        (lambda: K.bgemm(A, B), Ad @ Bd, Ad.abs() @ Bd.abs()),
        (lambda: K.bgemm(A, B, transA=True), Ad.transpose(1, 2) @ Bd, Ad.abs().transpose(1, 2) @ Bd.abs()),
        (lambda: K.bgemm(A, B, transB=True), Ad @ Bd.transpose(1, 2), Ad.abs() @ Bd.abs().transpose(1, 2)),
        (lambda: K.bgemm(A, B, alpha=2.0, add=add), 2 * (Ad @ Bd) + add.double(), 2 * (Ad.abs() @ Bd.abs())),
        (lambda: K.bgemm(A, W), Ad @ Wd, Ad.abs() @ Wd.abs()),
        (lambda: K.bgemm(A, W, transB=True, bias=bias), Ad @ Wd.T + bias.double(), Ad.abs() @ Wd.abs().T),
        (lambda: K.bgemm(A, Wl), Ad @ Wl.double(), Ad.abs() @ Wl.double().abs()),
        (lambda: K.bgemm(U, Wl, transB=True), U.double() @ Wl.double().T, U.double().abs() @ Wl.double().abs().T),
        (lambda: K.bgemm_reduce(A, B, transA=True), (Ad.transpose(1, 2) @ Bd).sum(0), (Ad.abs().transpose(1, 2) @ Bd.abs()).sum(0)),
        (lambda: K.bgemm_reduce(A, B), (Ad @ Bd).sum(0), (Ad.abs() @ Bd.abs()).sum(0)),
    ]
    n0 = cuda_pkg.launch_count()
    base = []
    for fn, ref, bound in cases:
        out = fn()
        assert float((out.double() - ref).abs().max()) <= 2e-6 * float(bound.max())
        base.append(out)
    assert cuda_pkg.launch_count() > n0
    try:
        for key, val in (("tcgemm_utccp", 1), ("tcgemm_mn_partial", 0), ("tcgemm_prefetch", 4)):
            cuda_pkg.set_tuning(key, val)
            for (fn, _, _), want in zip(cases, base):
                assert torch.equal(fn(), want), key
            cuda_pkg.set_tuning(key, {"tcgemm_utccp": 0, "tcgemm_mn_partial": 1, "tcgemm_prefetch": 0}[key])
    finally:
        cuda_pkg.set_tuning("tcgemm_utccp", 0); cuda_pkg.set_tuning("tcgemm_mn_partial", 1); cuda_pkg.set_tuning("tcgemm_prefetch", 0)


@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("M", [257, 260, 265, 268, 272])
def test_bgemm_ragged_rows_after_the_cta_pairs(K, tA, tB, M):
    """Rows 256.. of a per-subject product run on the FP32 kernels of csrc/tcgemm.cu (1..16 rows, both layouts of
    B, all epilogues); rows 0..255 on the CTA-pair tensor kernel."""
    g = torch.Generator().manual_seed(M)
    batch, N, Kd = 5, 268, 164
    A = torch.randn((batch, Kd, M) if tA else (batch, M, Kd), generator=g)
    B = torch.randn((batch, N, Kd) if tB else (batch, Kd, N), generator=g)
    add, bias = torch.randn(batch, M, N, generator=g), torch.randn(N, generator=g)
    opA = A.transpose(1, 2) if tA else A
    opB = B.transpose(1, 2) if tB else B
    ref = opA.double() @ opB.double()
    tol = 2e-6 * float((opA.abs().double() @ opB.abs().double()).max())
    got = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB).cpu()
    assert float((got.double() - ref).abs().max()) <= tol
    got = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB, alpha=-0.5, add=add.cuda()).cpu()
    assert float((got.double() - (-0.5 * ref + add.double())).abs().max()) <= tol
    got = K.bgemm(A.cuda(), B.cuda(), transA=tA, transB=tB, bias=bias.cuda()).cpu()
    assert float((got.double() - (ref + bias.double())).abs().max()) <= tol


def test_bgemm_epilogues_and_2d(K, gemm_backend):
    g = torch.Generator().manual_seed(5)
    X, W, b = torch.randn(144, 268, generator=g), torch.randn(268, 268, generator=g), torch.randn(268, generator=g)
    want = torch.nn.functional.linear(X.double(), W.double(), b.double())
    got = K.bgemm(X.cuda(), W.cuda(), transB=True, bias=b.cuda()).cpu()
    assert got.dim() == 2 and float((got.double() - want).abs().max()) < 2e-4
    add = torch.randn(4, 268, 268, generator=g)
    A, B = torch.randn(4, 268, 268, generator=g), torch.randn(4, 268, 268, generator=g)
    want = 2.0 * (A.double() @ B.double()) + add.double()
    got = K.bgemm(A.cuda(), B.cuda(), alpha=2.0, add=add.cuda()).cpu()
    assert float((got.double() - want).abs().max()) < 5e-4
    # stacked path with an addend (A 3-D, B shared)
    Wm = torch.randn(268, 160, generator=g)
    add2 = torch.randn(4, 268, 160, generator=g)
    want = A.double() @ Wm.double() + add2.double()
    got = K.bgemm(A.cuda(), Wm.cuda(), add=add2.cuda()).cpu()
    assert float((got.double() - want).abs().max()) < 3e-4


def test_bgemm_variants_agree_bitwise(cuda_pkg, K):
    K.set_gemm_backend("simt")
    g = torch.Generator().manual_seed(9)
    A, B = torch.randn(3, 268, 268, generator=g).cuda(), torch.randn(3, 268, 268, generator=g).cuda()
    outs = []
    for v in (0, 1, 2):
        cuda_pkg.set_tuning("gemm_variant", v)
        outs.append(K.bgemm(A, B, transB=True))
    cuda_pkg.set_tuning("gemm_variant", 0)
    K.set_gemm_backend("tc")
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("M", [160, 268, 256])            # 268 = one CTA pair + the odd tail tile on the one-CTA kernel
@pytest.mark.parametrize("tA,tB", [(False, False), (True, False), (False, True)])
def test_bgemm_reduce(K, gemm_backend, tA, tB, M):
    g = torch.Generator().manual_seed(11)
    batch, N, Kd = 37, 268, 268
    A = torch.randn((batch, Kd, M) if tA else (batch, M, Kd), generator=g)
    B = torch.randn((batch, N, Kd) if tB else (batch, Kd, N), generator=g)
    add = torch.randn(M, N, generator=g)
    opA = A.transpose(-1, -2) if tA else A
    opB = B.transpose(-1, -2) if tB else B
    want = (opA.double() @ opB.double()).sum(0) * 0.5 + add.double()
    got = K.bgemm_reduce(A.cuda(), B.cuda(), transA=tA, transB=tB, alpha=0.5, add=add.cuda()).cpu()
    tol = 2e-6 * float((opA.abs().double() @ opB.abs().double()).sum(0).max())
    assert float((got.double() - want).abs().max()) <= tol
    again = K.bgemm_reduce(A.cuda(), B.cuda(), transA=tA, transB=tB, alpha=0.5, add=add.cuda()).cpu()
    assert torch.equal(got, again)                       # deterministic reduction order


# the Graph U-Net's Linear layers and their gradients at ks = [0.9, 0.7, 0.6, 0.5] (160 -> 144 -> 100 -> 60 -> 30 rows of width
# 268: models/ops.py:60, 92-107), then ragged / minimal / maximal-K shapes
LIN_SHAPES = [(160, 268, 160), (144, 268, 268), (100, 268, 268), (60, 268, 268), (30, 268, 268), (160, 268, 536),
              (160, 536, 268), (268, 268, 144), (268, 268, 30), (268, 536, 160), (160, 320, 320), (1, 4, 4), (33, 36, 4),
              (37, 100, 7), (64, 64, 768), (31, 29, 50)]


@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
def test_bgemm_one_subject_kernel(cuda_pkg, K, tA, tB):
    """csrc/bgemm.cu lin::lin_kernel (one product, whole operand panels in shared memory, 32 x 32 tiles, four k groups):
    against the float64 product for every layout and epilogue, against the 136 x 136 kernel it replaces for these shapes
    (same fp32 class: a few ulp of sum |a||b|), deterministic, and never writing outside C."""
    n0 = cuda_pkg.launch_count()
    for M, N, Kd in LIN_SHAPES:
        g = torch.Generator().manual_seed(1000 * M + 10 * N + Kd)
        # leading dimensions rounded up to 4 floats (the kernel's alignment contract), operands are views into wider buffers
        lda = ((M if tA else Kd) + 3) // 4 * 4
        ldb = ((Kd if tB else N) + 3) // 4 * 4
        Abuf = torch.randn((Kd, lda) if tA else (M, lda), generator=g)
        Bbuf = torch.randn((N, ldb) if tB else (Kd, ldb), generator=g)
        A = Abuf[:, :M] if tA else Abuf[:, :Kd]
        B = Bbuf[:, :Kd] if tB else Bbuf[:, :N]
        opA, opB = (A.T if tA else A), (B.T if tB else B)
        want = opA.double() @ opB.double()
        tol = 2e-6 * float((opA.abs().double() @ opB.abs().double()).max())
        bias, add = torch.randn(N, generator=g), torch.randn(M, N, generator=g)
        Ad, Bd = Abuf.cuda(), Bbuf.cuda()

        def run(alpha=1.0, epi=0, extra=None, ld_extra=0):
            out = torch.full((M + 2, N + 3), 7.0, device="cuda")             # guard rows / columns around C (ldc = N + 3)
            K.check(cuda_pkg.lib().gsr_bgemm_f32(
                Ad.data_ptr(), Bd.data_ptr(), out.data_ptr(), M, N, Kd, lda, ldb, N + 3, 0, 0, 0, int(tA), int(tB), 1,
                float(alpha), epi, extra.data_ptr() if extra is not None else None, ld_extra, 0,
                torch.cuda.current_stream().cuda_stream), "gsr_bgemm_f32")
            assert bool((out[M:] == 7.0).all()) and bool((out[:, N:] == 7.0).all()), (M, N, Kd, "wrote outside C")
            return out[:M, :N].cpu()

        got = run()
        assert float((got.double() - want).abs().max()) <= tol, (M, N, Kd, tA, tB)
        assert torch.equal(got, run())
        got_b = run(alpha=-0.5, epi=K.EPI_BIAS, extra=bias.cuda())
        assert float((got_b.double() - (-0.5 * want + bias.double())).abs().max()) <= tol
        got_a = run(alpha=2.0, epi=K.EPI_ADD, extra=add.cuda(), ld_extra=N)
        assert float((got_a.double() - (2.0 * want + add.double())).abs().max()) <= 2 * tol
        try:
            cuda_pkg.set_tuning("bgemm_lin", 0)
            old = run()
        finally:
            cuda_pkg.set_tuning("bgemm_lin", 1)
        assert float((got - old).abs().max()) <= tol
    assert cuda_pkg.launch_count() >= n0 + 5 * len(LIN_SHAPES)      # every run() above is one launch of this library


def test_colsum_and_wrapper_routing(cuda_pkg, K):
    """gsr_colsum_f32 (32 columns x 8 row phases, fixed fold order) against float64, ragged sizes; kernels.bgemm sends one
    small product to the one-subject kernel whatever its row count and keeps batched / large products where they were."""
    g = torch.Generator().manual_seed(77)
    for rows, cols in ((160, 320), (30, 320), (1, 1), (7, 33), (512, 1024), (268, 268), (9, 31)):
        X = torch.randn(rows, cols, generator=g)
        got = K.colsum(X.cuda()).cpu()
        assert float((got.double() - X.double().sum(0)).abs().max()) <= 2e-6 * float(X.abs().double().sum(0).max())
        assert torch.equal(got, K.colsum(X.cuda()).cpu())
    X, W, b = torch.randn(144, 268, generator=g), torch.randn(268, 268, generator=g), torch.randn(268, generator=g)
    gY = torch.randn(144, 268, generator=g)
    y = K.bgemm(X.cuda(), W.cuda(), transB=True, bias=b.cuda()).cpu()
    assert float((y.double() - (X.double() @ W.double().T + b.double())).abs().max()) < 2e-6 * float((X.abs().double() @ W.abs().double().T).max())
    gW = K.bgemm(gY.cuda(), X.cuda(), transA=True).cpu()                       # 268 rows: the row rule alone would pick a tensor tile
    assert float((gW.double() - gY.double().T @ X.double()).abs().max()) < 2e-6 * float((gY.abs().double().T @ X.abs().double()).max())


@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("shape", [(512, 1024, 2048), (512, 1024, 2090), (300, 268, 1500), (512, 516, 1030)])
def test_bgemm_split_k(cuda_pkg, K, tA, tB, shape):
    """One skinny product with a long inner dimension (the notebook MLP's layers): K slices as the items of one batched tensor
    launch + gsr_splitk_reduce_f32, with a remainder slice (K not a multiple of 4 x splits), every epilogue, ragged rows
    (M = 300) and operands that are row-padded views; against float64 and against the unsplit product."""
    M, N, Kd = shape
    plan = K._splitk_plan(M, N, Kd)
    assert plan is not None and plan[0] >= 2 and plan[1] % 4 == 0 and plan[0] * plan[1] + plan[2] == Kd
    g = torch.Generator().manual_seed(M + N + Kd)
    lda = ((M if tA else Kd) + 3) // 4 * 4 + 4              # padded rows: views of wider buffers
    ldb = ((Kd if tB else N) + 3) // 4 * 4 + 8
    Abuf = torch.randn((Kd, lda) if tA else (M, lda), generator=g).cuda()
    Bbuf = torch.randn((N, ldb) if tB else (Kd, ldb), generator=g).cuda()
    A = Abuf[:, :M] if tA else Abuf[:, :Kd]
    B = Bbuf[:, :Kd] if tB else Bbuf[:, :N]
    opA, opB = (A.T if tA else A).double().cpu(), (B.T if tB else B).double().cpu()
    want = opA @ opB
    tol = 2e-6 * float((opA.abs() @ opB.abs()).max())
    bias, add = torch.randn(N, generator=g).cuda(), torch.randn(M, N, generator=g).cuda()
    n0 = cuda_pkg.launch_count()
    got = K.bgemm(A, B, transA=tA, transB=tB)
    assert cuda_pkg.launch_count() - n0 >= 2                # the sliced product (+ its ragged rows) and the reduction
    assert float((got.double().cpu() - want).abs().max()) <= tol
    assert torch.equal(got, K.bgemm(A, B, transA=tA, transB=tB))
    got = K.bgemm(A, B, transA=tA, transB=tB, alpha=-0.5, bias=bias)
    assert float((got.double().cpu() - (-0.5 * want + bias.double().cpu())).abs().max()) <= tol
    got = K.bgemm(A, B, transA=tA, transB=tB, alpha=2.0, add=add)
    assert float((got.double().cpu() - (2.0 * want + add.double().cpu())).abs().max()) <= 2 * tol
    old = K.SPLITK_MIN_K
    try:
        K.SPLITK_MIN_K = 1 << 30
        unsplit = K.bgemm(A, B, transA=tA, transB=tB)
    finally:
        K.SPLITK_MIN_K = old
    assert float((unsplit - K.bgemm(A, B, transA=tA, transB=tB)).abs().max()) <= 2 * tol     # each is within tol of float64


def test_linear_backward_pads_unaligned_gradient_rows(cuda_pkg, K):
    """LinearFn.backward with a [B, N] gradient whose N is not a multiple of 4 (the notebook MLP's 35 778 outputs): the
    row-padded copy feeds both products; against torch's float64 Linear."""
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(12)
    Bn, fin, fout = 256, 96, 4102
    X = torch.randn(Bn, fin, generator=g); W = torch.randn(fout, fin, generator=g) * 0.1; b = torch.randn(fout, generator=g)
    gY = torch.randn(Bn, fout, generator=g)
    Xc, Wc, bc = (t.cuda().requires_grad_(True) for t in (X, W, b))
    F_.LinearFn.apply(Xc, Wc, bc).backward(gY.cuda())
    Xd, Wd, bd = (t.double().requires_grad_(True) for t in (X, W, b))
    torch.nn.functional.linear(Xd, Wd, bd).backward(gY.double())
    for got, want in ((Xc.grad, Xd.grad), (Wc.grad, Wd.grad), (bc.grad, bd.grad)):
        assert float((got.double().cpu() - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_bgemm_linearity_full_batch(K, gemm_backend):
    """BASELINE size (1024 subjects): C(A1 + A2) = C(A1) + C(A2) up to fp32 rounding, and the
    stacked shared-weight path equals the per-subject path bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(2)
    A1 = torch.rand(1024, 268, 268, device="cuda", generator=g)
    A2 = torch.rand(1024, 268, 268, device="cuda", generator=g)
    W = torch.randn(268, 268, device="cuda", generator=g)
    c12 = K.bgemm(A1 + A2, W)
    c1, c2 = K.bgemm(A1, W), K.bgemm(A2, W)
    assert float((c12 - (c1 + c2)).abs().max()) < 1e-3 * float(c12.abs().max())
    per_subject = K.bgemm(A1, W.expand(1024, 268, 268).contiguous())
    # rows 0..255 of a subject run on the same tensor-core tiles in both paths; the 12 ragged rows of the per-subject
    # path are FP32 FMAs (csrc/tcgemm.cu, tail_kernel) and differ from the 3xTF32 products by rounding only
    assert torch.equal(per_subject[:, :256], c1[:, :256])
    assert float((per_subject[:, 256:] - c1[:, 256:]).abs().max()) < 4e-6 * float(c1.abs().max())
    assert K.colsum(c1[0]).shape == (268,)
    assert torch.allclose(K.colsum(c1[0]), c1[0].sum(0), rtol=1e-5, atol=1e-3)


# ------------------------------------------------------------------ sym / diag / abs
@pytest.mark.parametrize("n,batch", [(268, 2), (5, 1), (33, 3)])
def test_sym_diag_abs(cuda_pkg, n, batch):
    g = torch.Generator().manual_seed(n)
    x = torch.randn(batch, n, n, generator=g)
    T = torch.randn(batch, n, n, generator=g)
    xr = x.clone().requires_grad_(True)
    outs = []
    for b in range(batch):
        z = (xr[b] + xr[b].t()) / 2
        z[torch.eye(n, dtype=torch.bool)] = 1
        outs.append(torch.abs(z))
    ref = torch.stack(outs)
    (ref * T).sum().backward()
    xc = x.cuda().requires_grad_(True)
    out = cuda_pkg.functional.sym_diag_abs(xc)
    (out * T.cuda()).sum().backward()
    assert torch.equal(out.cpu(), ref.detach())
    assert torch.allclose(xc.grad.cpu(), xr.grad, rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,batch", [(268, 3), (64, 2), (12, 1), (200, 2), (4, 5)])
def test_sym_diag_abs_l1_fused(cuda_pkg, K, n, batch):
    """The fused final pass (sym-diag-abs + L1 against the target, prediction never written) against the two launches it
    replaces: loss to rounding (a different summation order), prediction and gradient BIT identical -- also with exact
    ties (pred == y: sgn 0), a zero pre-activation (sgn(z) = 0) and non-uniform upstream gradients -- and the loss against
    float64."""
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(n)
    x = torch.randn(batch, n, n, generator=g)
    y = torch.rand(batch, n, n, generator=g)
    z0 = ((x[0] + x[0].t()) / 2).abs()
    y[0, 0, :] = z0[0, :]                                      # ties in one row of one subject
    if n > 3:
        x[0, 1, 2] = 0.25; x[0, 2, 1] = -0.25                  # z = 0 exactly
    wt = torch.rand(batch, generator=g).cuda()
    xc, yc = x.cuda(), y.cuda()
    x1 = xc.clone().requires_grad_(True)
    per1 = F_.l1_loss_per_subject(F_.sym_diag_abs(x1), yc)
    (per1 * wt).sum().backward()
    x2 = xc.clone().requires_grad_(True)
    assert K.sym_diag_abs_l1_supported(x2, yc)
    per2 = F_.sym_diag_abs_l1_per_subject(x2, yc)
    (per2 * wt).sum().backward()
    assert torch.allclose(per1, per2, rtol=1e-6, atol=0)
    assert torch.equal(x1.grad, x2.grad)
    zd = (x.double() + x.double().transpose(1, 2)) / 2
    pred = zd.abs()
    pred[:, torch.arange(n), torch.arange(n)] = 1.0
    want = (pred.float().double() - y.double()).abs().mean((1, 2))
    assert torch.allclose(per2.cpu().double(), want, rtol=2e-6, atol=0)
    loss, p = K.sym_diag_abs_l1_fwd(xc, yc, want_pred=True)
    assert torch.equal(p, K.sym_diag_abs_fwd(xc)) and torch.equal(loss, per2.detach())


def test_sym_diag_abs_l1_fallback_and_contract(cuda_pkg, K):
    """n % 4 != 0 takes the two-launch route; the C entry point itself refuses (GSR_E_UNSUPPORTED), never guesses."""
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(2, 33, 33, generator=g).cuda(), torch.rand(2, 33, 33, generator=g).cuda()
    assert not K.sym_diag_abs_l1_supported(x, y)
    per = F_.sym_diag_abs_l1_per_subject(x.clone().requires_grad_(True), y)
    assert torch.equal(per, F_.l1_loss_per_subject(F_.sym_diag_abs(x), y))
    with pytest.raises(RuntimeError):
        K.sym_diag_abs_l1_fwd(x, y)
    with pytest.raises(NotImplementedError):
        F_.sym_diag_abs_l1_per_subject(x, y.clone().requires_grad_(True))


# ------------------------------------------------------------------ pooling
def test_topk_bit_exact_with_ties(K, golden):
    G = golden("ops")
    _, idx = K.pool_topk(torch.full((160,), 0.5, device="cuda"), 144)
    assert np.array_equal(idx.cpu().numpy(), G["tie_idx_stable"])
    g = torch.Generator().manual_seed(4)
    for n, k in ((160, 144), (144, 101), (101, 61), (61, 30), (9, 1), (7, 7)):
        s = (torch.randint(0, 12, (n,), generator=g).float() / 16 + 0.25)      # many exact ties
        vals, idx = K.pool_topk(s.cuda(), k)
        rv, ri = O.stable_topk(s, k)
        assert torch.equal(idx.cpu(), ri) and torch.equal(vals.cpu(), rv)
    for i in range(4):
        Gf = golden("gsrnet_full")
        s = torch.from_numpy(Gf[f"pool{i}_scores"])
        k = Gf[f"pool{i}_idx"].shape[0]
        _, idx = K.pool_topk(s.cuda(), k)
        assert np.array_equal(idx.cpu().numpy(), Gf[f"pool{i}_idx"])           # reference scores -> reference idx


def test_pool_unpool_gather_vs_oracle(cuda_pkg, golden):
    G = golden("ops")
    pkg = cuda_pkg
    pool = pkg.GraphPool(0.5, 4).cuda()
    with torch.no_grad():
        pool.proj.weight.copy_(dev(G["pool_w"]))
        pool.proj.bias.copy_(dev(G["pool_b"]))
    A2, X2, idx = pool(dev(G["pool_A"]), dev(G["pool_X"]))
    assert np.array_equal(idx.cpu().numpy(), G["pool_idx"])
    assert np.array_equal(A2.cpu().numpy(), G["pool_A_out"])                  # gather: bit exact
    assert np.allclose(X2.detach().cpu().numpy(), G["pool_X_out"], rtol=1e-6, atol=1e-7)
    unp = pkg.GraphUnpool()
    _, up = unp(torch.zeros(6, 6, device="cuda"), dev(G["unpool_X"]), dev(G["unpool_idx"]))
    assert np.array_equal(up.cpu().numpy(), G["unpool_out"])


def test_pool_backward_vs_oracle(cuda_pkg):
    g = torch.Generator().manual_seed(8)
    n, d, k = 101, 268, 0.6
    X = torch.randn(n, d, generator=g)
    pw, pb = torch.randn(1, d, generator=g) * 0.5, torch.randn(1, generator=g)
    T = torch.randn(O.pool_size(k, n), d, generator=g)
    sd = {"p.proj.weight": pw.clone().requires_grad_(True), "p.proj.bias": pb.clone().requires_grad_(True)}
    Xr = X.clone().requires_grad_(True)
    _, newX, idx, scores = O.graph_pool(sd, "p", k, torch.zeros(n, n), Xr)
    (newX * T).sum().backward()
    pool = cuda_pkg.GraphPool(k, d).cuda()
    with torch.no_grad():
        pool.proj.weight.copy_(pw.cuda())
        pool.proj.bias.copy_(pb.cuda())
    Xc = X.cuda().requires_grad_(True)
    new2, idx2, sc2 = pool.select(Xc)
    (new2 * T.cuda()).sum().backward()
    assert torch.allclose(sc2.cpu(), scores.detach(), rtol=0, atol=2e-7)
    assert torch.equal(idx2.cpu(), idx)
    assert torch.allclose(new2.detach().cpu(), newX.detach(), rtol=1e-6, atol=1e-6)
    assert torch.allclose(Xc.grad.cpu(), Xr.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(pool.proj.weight.grad.cpu(), sd["p.proj.weight"].grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(pool.proj.bias.grad.cpu(), sd["p.proj.bias"].grad, rtol=1e-4, atol=1e-6)


# ------------------------------------------------------------------ L1 / evaluation
def test_l1_loss(cuda_pkg):
    F_ = cuda_pkg.functional
    g = torch.Generator().manual_seed(6)
    x, y = torch.randn(5, 268, 268, generator=g), torch.randn(5, 268, 268, generator=g)
    y[0, :4] = x[0, :4]                                        # exact zeros: sgn(0) = 0
    xr = x.clone().requires_grad_(True)
    ref = torch.nn.L1Loss()(xr, y)
    ref.backward()
    xc = x.cuda().requires_grad_(True)
    got = F_.l1_loss(xc, y.cuda())
    got.backward()
    assert float(got) == pytest.approx(float(ref), rel=2e-6)
    assert torch.allclose(xc.grad.cpu(), xr.grad, rtol=1e-6, atol=0)
    per = F_.l1_loss_per_subject(x.cuda(), y.cuda()).cpu()
    assert torch.allclose(per, (x - y).abs().mean((1, 2)), rtol=2e-6)
    # shared target + gradient into it (criterion(net_outs, start_gcn_outs) has grads on both sides)
    a = torch.randn(160, 268, generator=g)
    b = torch.randn(160, 268, generator=g)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    torch.nn.L1Loss()(ar, br).backward()
    ac, bc = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    F_.l1_loss(ac, bc).backward()
    assert torch.allclose(ac.grad.cpu(), ar.grad) and torch.allclose(bc.grad.cpu(), br.grad)
    # shared y across a batch, gradient reduced over the subjects
    u = torch.randn(4, 268, 160, generator=g)
    w = torch.randn(268, 160, generator=g)
    wr = w.clone().requires_grad_(True)
    torch.stack([torch.nn.L1Loss()(wr, u[i]) for i in range(4)]).mean().backward()
    wc = w.cuda().requires_grad_(True)
    F_.l1_loss_per_subject(u.cuda(), wc).mean().backward()
    assert torch.allclose(wc.grad.cpu(), wr.grad, rtol=1e-5, atol=1e-9)
    # the one-pass kernel for the shared side (x without gradient) equals the general path (x with gradient), with
    # non-uniform upstream gradients and more subjects than batch phases; ties (x == y) contribute 0
    u = torch.randn(70, 33, 7, generator=g); w = torch.randn(33, 7, generator=g); u[3, 2] = w[2]
    wt = torch.rand(70, generator=g)
    w1 = w.cuda().requires_grad_(True)
    (F_.l1_loss_per_subject(u.cuda(), w1) * wt.cuda()).sum().backward()
    w2, u2 = w.cuda().requires_grad_(True), u.cuda().requires_grad_(True)
    (F_.l1_loss_per_subject(u2, w2) * wt.cuda()).sum().backward()
    want = -(torch.sign(u.double() - w.double()) * wt.double()[:, None, None]).sum(0) / (33 * 7)
    assert torch.allclose(w1.grad.cpu().double(), want, rtol=1e-5, atol=1e-8)
    assert torch.allclose(w1.grad, w2.grad, rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize("tag", ["small", "hr"])
def test_eval_reductions_golden(cuda_pkg, golden, tag):
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
    mae, pcc, js = cuda_pkg.evaluation_reductions(np.stack(preds), np.stack(gts))
    assert mae == pytest.approx(float(G[f"{tag}_mae64"]), rel=1e-12)
    assert pcc == pytest.approx(float(G[f"{tag}_pcc64"]), rel=1e-9)
    assert js == pytest.approx(float(G[f"{tag}_js64"]), rel=1e-9)
    # flat (already vectorised) entry agrees
    p1 = np.concatenate([VO.vectorize(m) for m in preds]); g1 = np.concatenate([VO.vectorize(m) for m in gts])
    from importlib import import_module
    ev = import_module(cuda_pkg.__name__ + ".utils.evaluation")
    mae2, pcc2, js2 = ev.evaluation_reductions_flat(p1, g1)
    assert mae2 == pytest.approx(mae, rel=1e-12) and pcc2 == pytest.approx(pcc, rel=1e-10) and js2 == pytest.approx(js, rel=1e-10)


def test_eval_reductions_full_sweep_10k(cuda_pkg):
    """BASELINE configs[4]: 10 000 predictions against ground truth (268 nodes) in one call, checked against
    float64 sums accumulated chunk by chunk with torch (sklearn / scipy definitions of utils/evaluation.py:106-112)."""
    N, n, chunk = 10000, 268, 1000
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand(N, n, n, device="cuda", generator=g)
    gt = (x + x.transpose(1, 2)) / 2
    del x
    noise = torch.randn(N, n, n, device="cuda", generator=g) * 0.1
    pred = torch.clamp(gt + (noise + noise.transpose(1, 2)) / 2, 0, 1)
    del noise
    mae, pcc, js = cuda_pkg.evaluation_reductions(pred, gt)
    iu = torch.triu_indices(n, n, 1, device="cuda")
    S = dict(ad=0.0, p=0.0, g=0.0, pp=0.0, gg=0.0, pg=0.0)
    for s0 in range(0, N, chunk):
        pc = pred[s0:s0 + chunk][:, iu[0], iu[1]].double(); gc = gt[s0:s0 + chunk][:, iu[0], iu[1]].double()
        S["ad"] += float((pc - gc).abs().sum()); S["p"] += float(pc.sum()); S["g"] += float(gc.sum())
        S["pp"] += float((pc * pc).sum()); S["gg"] += float((gc * gc).sum()); S["pg"] += float((pc * gc).sum())
    cnt = N * n * (n - 1) // 2
    assert mae == pytest.approx(S["ad"] / cnt, rel=1e-10)
    cov = S["pg"] - S["p"] * S["g"] / cnt
    want_pcc = cov / np.sqrt((S["pp"] - S["p"] ** 2 / cnt) * (S["gg"] - S["g"] ** 2 / cnt))
    assert pcc == pytest.approx(want_pcc, rel=1e-8)
    kl = 0.0
    for s0 in range(0, N, chunk):
        pc = pred[s0:s0 + chunk][:, iu[0], iu[1]].double() / S["p"]; gc = gt[s0:s0 + chunk][:, iu[0], iu[1]].double() / S["g"]
        m = (pc + gc) / 2
        kl += float(torch.special.xlogy(pc, pc / m).sum()) + float(torch.special.xlogy(gc, gc / m).sum())
    assert js == pytest.approx(np.sqrt(kl / 2), rel=1e-8)


def test_eval_reductions_zero_entries(cuda_pkg):
    # rel_entr conventions: zeros in gt and pred contribute 0, not NaN
    gt = torch.rand(3, 20, 20); pred = torch.rand(3, 20, 20)
    gt[0, 1, 5] = 0; pred[1, 2, 7] = 0; gt[2, 3, 9] = 0; pred[2, 3, 9] = 0
    mae, pcc, js = cuda_pkg.evaluation_reductions(pred.numpy(), gt.numpy())
    m2, p2, j2 = VO.eval_reductions(pred.numpy(), gt.numpy())
    assert mae == pytest.approx(m2, rel=1e-12) and pcc == pytest.approx(p2, rel=1e-9) and js == pytest.approx(j2, rel=1e-9)


# ------------------------------------------------------------------ N4: the on-disk edge
@pytest.mark.parametrize("n", [2, 3, 5, 33, 160, 268])
def test_triu_flatten_is_numpy_triu_indices(cuda_pkg, n):
    """training_gsr.ipynb:112-117: bit exact against matrix[np.triu_indices(n, k=1)], batched."""
    g = torch.Generator().manual_seed(n)
    M = torch.randn(3, n, n, generator=g)
    got = cuda_pkg.kernels.triu_flatten(M.cuda()).cpu().numpy()
    iu = np.triu_indices(n, k=1)
    for b in range(3):
        assert np.array_equal(got[b], M[b].numpy()[iu])
    # inverse of anti_vectorize on symmetric zero-diagonal matrices
    v = torch.randn(2, n * (n - 1) // 2, generator=g).cuda()
    assert torch.equal(cuda_pkg.kernels.triu_flatten(cuda_pkg.kernels.anti_vectorize(v, n)), v)


def test_submission_edge(cuda_pkg, tmp_path):
    """extract_data's NaN rule + anti-vectorise (utils/preprocessing.py:60-75) and both submission orders."""
    from importlib import import_module
    sub = import_module(cuda_pkg.__name__ + ".utils.submission")
    rng = np.random.default_rng(0)
    rows = rng.random((4, 12720))
    rows[1, 5] = np.nan
    mats = sub.extract_matrices(rows, is_hr=False)
    assert mats.shape == (4, 160, 160) and mats.dtype == np.float64
    ref = np.stack([O_vec.anti_vectorize(np.nan_to_num(r, nan=0.0).astype(np.float32), 160) for r in rows])
    assert np.array_equal(mats, ref.astype(np.float64))
    preds = rng.random((3, 268, 268)).astype(np.float32)
    row = sub.submission_vector(preds, "row")
    col = sub.submission_vector(preds, "column")
    iu = np.triu_indices(268, k=1)
    assert np.array_equal(row, np.concatenate([p[iu] for p in preds]))
    assert np.array_equal(col, np.concatenate([O_vec.vectorize(p) for p in preds]))
    path = tmp_path / "submission.csv"
    assert sub.write_submission(str(path), row[:5]) == 5
    lines = path.read_text().splitlines()
    assert lines[0] == "ID,Predicted" and lines[1].startswith("1,") and len(lines) == 6
    assert float(lines[3].split(",")[1]) == float(row[2])


def test_flat_bucket_pack_unpack_roundtrip(K):
    """(e) gradient bucket: 70 tensors (two parameter-table chunks of csrc/api.cu), scale on the way in."""
    g = torch.Generator().manual_seed(70)
    tensors = [torch.randn(int(n), generator=g).cuda() for n in torch.randint(1, 3000, (70,), generator=g)]
    want = torch.cat([t.reshape(-1) for t in tensors]) * 0.25
    bucket = K.FlatBucket(tensors)
    flat = bucket.pack(0.25)
    assert torch.equal(flat, want)
    flat.mul_(2.0)
    bucket.unpack(1.0)
    assert torch.equal(torch.cat([t.reshape(-1) for t in tensors]), want * 2.0)


def test_unpool_rejects_out_of_range_indices(cuda_pkg):
    """caller-supplied indices never become out-of-bounds addresses -- the public module raises like torch's
    indexing does, and the raw kernels skip / zero what is out of range."""
    K = cuda_pkg.kernels
    unp = cuda_pkg.GraphUnpool()
    A, X = torch.zeros(6, 6, device="cuda"), torch.randn(3, 4, device="cuda")
    for bad in ([4, 0, 6], [-1, 0, 2]):
        with pytest.raises(IndexError):
            unp(A, X, torch.tensor(bad, device="cuda"))
    with pytest.raises(IndexError):
        unp(A, X, torch.tensor([4, 0, 2], device="cuda", dtype=torch.int32))
    out = K.unpool_fwd(X, torch.tensor([4, 99, 2], device="cuda"), 6)           # raw kernel: row 99 is dropped
    assert torch.equal(out[4], X[0]) and torch.equal(out[2], X[2]) and float(out[[0, 1, 3, 5]].abs().max()) == 0.0
    G = K.gather_rows_cols(torch.arange(36., device="cuda").reshape(6, 6), torch.tensor([1, 7, 3], device="cuda"))
    assert torch.equal(G[0], torch.tensor([7., 0., 9.], device="cuda")) and float(G[1].abs().max()) == 0.0
