#!/usr/bin/env python
"""Per-kernel device times at the BASELINE batch (1024 subjects): CUDA events, warm-up,
inputs larger than L2.  Prints one JSON object per kernel; used to fill profiles/ and DESIGN.md."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200")
K = pkg.kernels
B = int(os.environ.get("GSR_B", "1024"))
HBM = 6538.9


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, flops=None, bytes_=None, **kw):
    d = {"kernel": name, "ms": round(ms, 4)}
    if flops: d["tflops"] = round(flops / ms / 1e9, 2)
    if bytes_: d["gbs"] = round(bytes_ / ms / 1e6, 1); d["hbm_frac"] = round(bytes_ / ms / 1e6 / HBM, 3)
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(B, 160, 160, device=dev, generator=g); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
    H = torch.rand(B, 268, 268, device=dev, generator=g)
    H2 = torch.rand(B, 268, 268, device=dev, generator=g)
    W = torch.randn(268, 268, device=dev, generator=g)
    nb = 268 * 268 * 4 * B
    A, r, _ = K.normalize_adj_fwd(lr)
    report("normalize_fwd_160", timeit(lambda: K.normalize_adj_fwd(lr)), bytes_=2 * 160 * 160 * 4 * B)
    report("normalize_fwd_268_pre", timeit(lambda: K.normalize_adj_fwd(H, pre=True)), bytes_=2 * nb)
    adj, r2, _ = K.normalize_adj_fwd(H, pre=True)
    report("normalize_bwd_268_pre", timeit(lambda: K.normalize_adj_bwd(H, r2, H2, pre=True)), bytes_=3 * nb)
    report("sym_diag_abs_fwd", timeit(lambda: K.sym_diag_abs_fwd(H)), bytes_=2 * nb)
    report("sym_diag_abs_bwd", timeit(lambda: K.sym_diag_abs_bwd(H, H2)), bytes_=3 * nb)
    report("l1_fwd", timeit(lambda: K.l1_loss_fwd(H, H2, B, False)), bytes_=2 * nb)
    gl = torch.ones(B, device=dev)
    report("l1_bwd", timeit(lambda: K.l1_loss_bwd(H, H2, gl, B, False)), bytes_=3 * nb)
    report("vectorize_268", timeit(lambda: K.vectorize(H)), bytes_=2 * 35778 * 4 * B)
    v = K.vectorize(H)
    report("anti_vectorize_268", timeit(lambda: K.anti_vectorize(v, 268)), bytes_=(35778 + 268 * 268) * 4 * B)
    ms = timeit(lambda: K.eigh_tridiag(A), reps=3, warm=1)
    w, U, info = K.eigh_tridiag(A)
    report("eigh_tridiag_160 (K1-K4, the in-step LR route)", ms, flops=9 * 160 ** 3 * B, us_per_graph=round(ms * 1e3 / B, 2),
           ql_iterations_mean=float(info.float().mean()))
    hr_t = (H + H.transpose(1, 2)) / 2
    hr_t.diagonal(dim1=1, dim2=2).fill_(1.0)
    ms = timeit(lambda: K.eigh_tridiag(hr_t), reps=2, warm=1)
    report("eigh_tridiag_268 (HR target, shared-memory variants)", ms, flops=9 * 268 ** 3 * B, us_per_graph=round(ms * 1e3 / B, 2))
    ms = timeit(lambda: K.eigh_jacobi(hr_t), reps=1, warm=1)
    report("eigh_jacobi_268 (global-workspace variant, for context)", ms, flops=9 * 268 ** 3 * B, us_per_graph=round(ms * 1e3 / B, 2))
    ms = timeit(lambda: K.eigh_jacobi(A), reps=3, warm=1)
    w, U, info = K.eigh_jacobi(A)
    report("eigh_jacobi_160", ms, flops=9 * 160 ** 3 * B, us_per_graph=round(ms * 1e3 / B, 2),
           sweeps_mean=float((info & 0xFF).float().mean()), polish_mean=float((info >> 8).float().mean()))
    ms = timeit(lambda: torch.linalg.eigh(A, UPLO="U"), reps=1, warm=1)
    report("torch.linalg.eigh_160 (cuSOLVER, for context)", ms, us_per_graph=round(ms * 1e3 / B, 2))
    Wl = torch.randn(268, 160, device=dev, generator=g)
    for backend in ("tc", "simt"):
        K.set_gemm_backend(backend)
        tag = "tcgemm (tcgen05 3xTF32)" if backend == "tc" else "bgemm (SIMT FP32)"
        f = 2 * 268 ** 3 * B
        report(f"{tag} NN per-subject 268^3", timeit(lambda: K.bgemm(H, H2)), flops=f)
        report(f"{tag} NT per-subject 268^3", timeit(lambda: K.bgemm(H, H2, transB=True)), flops=f)
        report(f"{tag} TN per-subject 268^3", timeit(lambda: K.bgemm(H, H2, transA=True)), flops=f)
        report(f"{tag} NN stacked shared-W 268^3", timeit(lambda: K.bgemm(H, W)), flops=f)
        report(f"{tag} NT stacked shared-W 268^3", timeit(lambda: K.bgemm(H, W, transB=True)), flops=f)
        report(f"{tag} reduce TN 268^3", timeit(lambda: K.bgemm_reduce(H, H2, transA=True)), flops=f)
        report(f"{tag} NT stacked U.W^T 160x268x160", timeit(lambda: K.bgemm(A, Wl, transB=True)), flops=2 * 160 * 268 * 160 * B)
    K.set_gemm_backend("tc")
    ms = timeit(lambda: torch.bmm(H, H2))
    report("torch.bmm 268^3 fp32 (cuBLAS, for context)", ms, flops=2 * 268 ** 3 * B)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms = timeit(lambda: torch.bmm(H, H2))
    report("torch.bmm 268^3 tf32 (cuBLAS, for context)", ms, flops=2 * 268 ** 3 * B)
    torch.backends.cuda.matmul.allow_tf32 = False
    P = torch.rand(10000 // 8, 268, 268, device=dev, generator=g); Gt = torch.rand(10000 // 8, 268, 268, device=dev, generator=g)
    ms = timeit(lambda: K.eval_reduce(P, Gt, True), reps=3, warm=1)
    report("eval_reduce upper (1250 pairs)", ms, bytes_=2 * 2 * 35778 * 4 * P.shape[0])
    del P, Gt
    # N2: the MLP baseline at the challenge size (12 720 LR features -> 268 x 268), batch 512, hidden 1024 x 2
    Bm, Hm, In, Out = 512, 1024, 12720, 268 * 268
    torch.manual_seed(0)
    mlp = pkg.SuperResMLP(In, Out, Hm, 2).to(dev)
    xm = torch.rand(Bm, In, device=dev, generator=g)
    fwd_flops = 2 * Bm * (In * Hm + Hm * Hm + Hm * Out)
    mlp.eval()
    with torch.no_grad():
        ms = timeit(lambda: mlp(xm), reps=5, warm=2)
    report("SuperResMLP 12720-1024-1024-71824 forward (eval), batch 512", ms, flops=fwd_flops)
    mlp.train()
    def train_step():
        mlp.zero_grad(set_to_none=True)
        mlp(xm).sum().backward()
    ms = timeit(train_step, reps=5, warm=2)
    report("SuperResMLP 12720-1024-1024-71824 forward + backward (train), batch 512", ms, flops=3 * fwd_flops - 2 * Bm * In * Hm)
    xb = torch.randn(4096, 4096, device=dev, generator=g)
    mean, var = K.bn_stats(xb)
    ga, be = torch.ones(4096, device=dev), torch.zeros(4096, device=dev)
    ms = timeit(lambda: K.bn_stats(xb), reps=5, warm=2)
    report("bn_stats 4096 x 4096 (two passes)", ms, bytes_=2 * xb.numel() * 4)
    ms = timeit(lambda: K.bn_act_fwd(xb, mean, var, ga, be, 1e-5, 0.1, 7, True), reps=5, warm=2)
    report("bn_act_fwd 4096 x 4096 (BatchNorm + Dropout 0.1 + ReLU)", ms, bytes_=2 * xb.numel() * 4)
    ms = timeit(lambda: K.bn_act_bwd(xb, mean, var, ga, be, xb, 1e-5, 0.1, 7, True, True), reps=5, warm=2)
    report("bn_act_bwd 4096 x 4096 (reduce + apply)", ms, bytes_=5 * xb.numel() * 4)


if __name__ == "__main__":
    main()
