"""Where does tcgemm_kernel spend its time?  Experiment knobs (gsr_tcgemm_debug key 0, bitmask):
1 = one MMA pass instead of three, 2 = no split arithmetic, 4 = epilogue stores one row, 8 = no global loads."""
import ctypes, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("fmri-connectivity-super-resolution_b200"); K = pkg.kernels
lib = pkg.lib()
dbg = lib.gsr_tcgemm_debug; dbg.argtypes = [ctypes.c_int, ctypes.c_int]; dbg.restype = ctypes.c_int
B = 1024
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g)
W = torch.randn(268, 268, device="cuda", generator=g)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
fl = 2 * 268 ** 3 * B
for mask in (0, 64, 8):
    dbg(0, mask)
    r = {"mask": mask}
    for name, fn in (("NN per-subject", lambda: K.bgemm(X, Y)), ("NT per-subject", lambda: K.bgemm(X, Y, transB=True)),
                     ("TN per-subject", lambda: K.bgemm(X, Y, transA=True)), ("NN stacked W", lambda: K.bgemm(X, W))):
        ms = timeit(fn); r[name] = [round(ms, 4), round(fl / ms / 1e9, 1)]
    print(json.dumps(r), flush=True)
dbg(0, 0)
