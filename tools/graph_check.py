"""Eager against CUDA-graphed training step: host enqueue time, device time, and equality of the results."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
pkg = importlib.import_module(bench.PKG_NAME)
B = int(os.environ.get("GSR_B", "1024")); dev = torch.device("cuda", 0)
lr = bench.synth_batch(torch, 160, B, 1000, dev); hr = bench.synth_batch(torch, 268, B, 5000, dev)
U = pkg.hr_target_basis(hr, 160)
def build():
    torch.manual_seed(0)
    m = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], bench.Args()).to(dev)
    return m, torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True)
m1, o1 = build(); eager = pkg.GSRTrainStep(m1)
def eager_step():
    o1.zero_grad(set_to_none=True); out = eager(lr, hr, U_hr_reduced=U); o1.step(); return out
m2, o2 = build(); graphed = pkg.GraphedTrainStep(m2, o2, B)
print(json.dumps({"launches_per_step": graphed.launches_per_step}), flush=True)
for i in range(3):
    l1, _ = eager_step(); l2, _ = graphed(lr, hr, U)
    torch.cuda.synchronize()
    dp = max(float((a - b).abs().max()) for a, b in zip(m1.parameters(), m2.parameters()))
    print(json.dumps({"step": i, "loss_eager": float(l1), "loss_graph": float(l2), "max_param_diff": dp}), flush=True)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return {"host_enqueue_ms": 1e3 * (t1 - t0) / n, "device_ms": a.elapsed_time(b) / n}
print(json.dumps({"eager": timed(eager_step)}), flush=True)
print(json.dumps({"graph": timed(lambda: graphed(lr, hr, U))}), flush=True)
print(json.dumps({"graph_replay_only": timed(graphed.replay)}), flush=True)
