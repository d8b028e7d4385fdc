"""Kernel-by-kernel device time of one training step of the notebook MLP (bench.mlp_measure's step) with torch.profiler
(CUPTI); prints the top kernels.  Not a bench number (profiler overhead): it says where the step's time is."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pkg = importlib.import_module(bench.PKG_NAME)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = pkg.SuperResMLPNotebook(160, 268, 1024, 2).to(dev); model.train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
x = bench.synth_batch(torch, 160, 512, 77, dev); y = bench.synth_batch(torch, 268, 512, 78, dev)


def one():
    opt.zero_grad(set_to_none=True)
    loss = pkg.functional.l1_loss(model(x), y)
    loss.backward(); opt.step()


for _ in range(3): one()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): one()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time per step {tot / 3 / 1e3:.3f} ms")
for e in rows[:28]:
    print(f"{e.device_time_total / 3:9.1f} us {e.count // 3:3d}x  {e.key[:110]}")
