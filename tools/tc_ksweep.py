import importlib, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
K.set_gemm_backend("tc")
g = torch.Generator(device="cuda").manual_seed(1)
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
B = 1024
for Kd in (32, 64, 128, 256, 512, 1024):
    A = torch.rand(B * 268, Kd, device="cuda", generator=g); W = torch.rand(Kd, 268, device="cuda", generator=g)
    ms = timeit(lambda: K.bgemm(A, W))
    tiles = (B * 268 + 127) // 128
    print(json.dumps({"K": Kd, "ms": round(ms, 4), "us_per_tile_per_sm": round(ms * 1e3 * 148 / tiles, 2), "kblocks": Kd // 32}), flush=True)
for N in (32, 128, 256, 268):
    A = torch.rand(B * 268, 256, device="cuda", generator=g); W = torch.rand(256, N, device="cuda", generator=g)
    ms = timeit(lambda: K.bgemm(A, W))
    print(json.dumps({"N": N, "K": 256, "ms": round(ms, 4)}), flush=True)
