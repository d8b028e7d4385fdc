"""Pipeline trace of tcgemm_kernel: builds a private libgsr copy with -DGSR_TC_TRACE, runs one
1024-subject 268^3 product and prints, for one CTA, when each producer group issued its loads, got
its stage back, saw the loads land, finished the split + stores, the proxy fence, and when the MMA
thread saw the stage full / had its MMAs committed (ns relative to the CTA's first event)."""
import ctypes, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
csrc = os.path.join(ROOT, "fmri-connectivity-super-resolution_b200", "csrc")
out = os.path.join(tempfile.mkdtemp(), "libgsr_trace.so")
srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DGSR_TC_TRACE",
                       "-Xcompiler", "-fPIC", "-shared", "-o", out, *srcs, "-lcudart"])
lib = ctypes.CDLL(out)
B = 1024
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g); C = torch.empty_like(X)
vp, ll, i32, f32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
fn = lib.gsr_tcgemm_f32
fn.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, ll, ll, ll, i32, i32, i32, f32, i32, vp, i32, ll, vp]
for _ in range(3):
    rc = fn(X.data_ptr(), Y.data_ptr(), C.data_ptr(), 268, 268, 268, 268, 268, 268, 268 * 268, 268 * 268, 268 * 268, 0, 0, B, 1.0, 0, None, 0, 0,
            torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * 16 * 8))()
lib.gsr_tcgemm_trace_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), i32]
assert lib.gsr_tcgemm_trace_read(buf, 4 * 16 * 8) == 0
t = list(buf)
t0 = min(x for x in t if x)
names = ["issue", "wait_stage", "stage_free", "loads_landed", "stored", "fenced"]
for kb in range(9):
    grp = kb % 2
    ev = [t[(grp * 16 + kb) * 8 + e] for e in range(6)]
    mm = [t[(2 * 16 + kb) * 8 + e] for e in range(3)]
    print(f"kb {kb} grp {grp}: " + " ".join(f"{n}={x - t0 if x else -1:6d}" for n, x in zip(names, ev)) +
          " | mma: " + " ".join(f"{n}={x - t0 if x else -1:6d}" for n, x in zip(["wait", "full", "committed"], mm)))
