#!/usr/bin/env python
"""One training step of the BASELINE batch inside a cudaProfilerStart/Stop window, for
`ncu --profile-from-start off ...`.  Warm-up steps run outside the window."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

pkg = importlib.import_module(bench.PKG_NAME)
for kv in os.environ.get("GSR_TUNE", "").split(","):           # e.g. GSR_TUNE=eigh_chunks=1 under ncu (serialised kernels)
    if kv: pkg.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
B = int(os.environ.get("GSR_B", "1024"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], bench.Args()).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
step = pkg.GSRTrainStep(model)
lr = bench.synth_batch(torch, 160, B, 1000, dev)
hr = bench.synth_batch(torch, 268, B, 5000, dev)
U_hr = torch.zeros(B, 268, 160, device=dev)       # the HR basis is hoisted (see bench.py); its values do not change the kernels


def one():
    opt.zero_grad(set_to_none=True)
    step(lr, hr, U_hr_reduced=U_hr)
    opt.step()


for _ in range(2):
    one()
torch.cuda.synchronize()
torch.cuda.profiler.start()
one()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step, batch", B)
