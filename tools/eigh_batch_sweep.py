"""Which K2 / pipeline for which batch: the whole n = 160 decomposition timed for B graphs with (a) one launch per stage and
the one-warp-per-graph QL kernel, (b) one launch per stage and the lane = graph QL kernel, (c) the chunked pipeline."""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from eigh_check import timed
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(2048, 160, 160, device="cuda", generator=gen); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
for B in (64, 128, 256, 384, 512, 768, 1024, 2048):
    row = {"batch": B}
    for tag, lanes, chunks in (("one_warp_per_graph", 0, 1), ("lanes_unchunked", 2, 1), ("default", 1, 8)):
        pkg.set_tuning("eigh_k2_lanes", lanes); pkg.set_tuning("eigh_chunks", chunks)
        _, ms = timed(K.eigh_tridiag, A[:B], reps=5); row[tag + "_ms"] = round(ms, 3)
    print(json.dumps(row), flush=True)
