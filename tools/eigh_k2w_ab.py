"""Chunked eigensolver pipeline: warps (of 32 graphs) per resident K2 CTA -- 2 (16 SMs held for the rotation chains of 1024 graphs),
3 or 4 (8 SMs) -- now that the chains are off the critical path (Q formed by its own launch, tools/eigh_k1_ab.py)."""
import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
from eigh_check import timed
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(1024, 160, 160, device="cuda", generator=gen); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
A = pkg.normalize_adj_torch(lr)
for chunks in (8, 4):
    for warps in (2, 4, 3, 1):
        pkg.set_tuning("eigh_chunks", chunks); pkg.set_tuning("eigh_k2_warps", warps)
        (w, U, info), ms = timed(K.eigh_tridiag, A, reps=7)
        print(json.dumps({"chunks": chunks, "k2_warps": warps, "ms": round(ms, 3), "info_min": int(info.min())}), flush=True)
