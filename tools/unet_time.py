"""Graph U-Net (models/ops.py:64-115, fed the identity: once per step) forward + backward as a CUDA graph, and its Linear
products one by one: the one-subject kernel of csrc/bgemm.cu (lin::lin_kernel) against the 136 x 136 kernel
(gsr_set_tuning("bgemm_lin", 0)).  Prints one JSON line."""
import importlib, json, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from importlib import import_module
GraphUnet = import_module("fmri-connectivity-super-resolution_b200.models.ops").GraphUnet

torch.manual_seed(0)
net = GraphUnet([0.9, 0.7, 0.6, 0.5], 160, 268).cuda()
I = torch.eye(160, device="cuda")
A = torch.rand(160, 160, device="cuda")
params = [p for p in net.parameters()]


def fwd_bwd():
    for p in params:
        p.grad = None
    out, start = net(A, I)
    (out.sum() + start.square().sum()).backward()


def graph_us(reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fwd_bwd()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd_bwd()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 1)


def each_us(fn, reps=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps * 1e3, 2)


res = {}
for lin in (1, 0):
    pkg.set_tuning("bgemm_lin", lin)
    tag = "one_subject_kernel" if lin else "kernel_136x136"
    r = {"unet_fwd_bwd_graph_us": graph_us()}
    gen = torch.Generator(device="cuda").manual_seed(1)
    for M in (160, 100, 30):
        X = torch.randn(M, 268, device="cuda", generator=gen); W = torch.randn(268, 268, device="cuda", generator=gen)
        b = torch.randn(268, device="cuda", generator=gen); gY = torch.randn(M, 268, device="cuda", generator=gen)
        r[f"fwd {M}x268x268 (X W^T + b)"] = each_us(lambda: K.bgemm(X, W, transB=True, bias=b))
        r[f"dX {M}x268x268 (gY W)"] = each_us(lambda: K.bgemm(gY, W))
        r[f"dW 268x268x{M} (gY^T X)"] = each_us(lambda: K.bgemm(gY, X, transA=True))
        r[f"colsum {M}x268"] = each_us(lambda: K.colsum(gY))
    X = torch.randn(160, 536, device="cuda", generator=gen); W = torch.randn(268, 536, device="cuda", generator=gen)
    r["fwd 160x268x536 (end_gcn)"] = each_us(lambda: K.bgemm(X, W, transB=True))
    res[tag] = r
pkg.set_tuning("bgemm_lin", 1)
print(json.dumps(res))
