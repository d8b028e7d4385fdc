"""gsr_eval_reduce (MAE / PCC / JSD of utils/evaluation.py:106-112) timed on 1250-pair chunks of the 10 000-pair sweep
(BASELINE configs[4]); reports ms per 10 000 pairs and the fraction of the measured HBM peak over its two passes."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200")
g = torch.Generator(device="cuda").manual_seed(0)
N = 1250
gt = torch.rand(N, 268, 268, device="cuda", generator=g); gt = (gt + gt.transpose(1, 2)) / 2
pred = (gt + 0.1 * torch.randn(N, 268, 268, device="cuda", generator=g)).clamp(0, 1)
for _ in range(2): pkg.kernels.eval_reduce(pred, gt, True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(8): r = pkg.kernels.eval_reduce(pred, gt, True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 8
alg = 2 * 2 * N * (268 * 267 // 2) * 4          # two passes x (pred + gt) x upper triangle x 4 B
print(json.dumps({"kernel": "gsr_eval_reduce (eval_pass1 + eval_pass2)", "pairs": N, "ms": ms, "ms_per_10k_pairs": ms * 8,
                  "algorithmic_GB_two_passes": alg / 1e9, "GBps": alg / ms / 1e6, "frac_of_hbm_6538.9": alg / ms / 1e6 / 6538.9,
                  "result": r.tolist()}))
