"""Pipeline trace of tcgemm3_kernel (csrc/tcgemm_tma.cu): builds a private libgsr copy with -DGSR_T3_TRACE, runs one
1024-subject stacked 268-wide product (or --per-subject) and prints, for both CTAs of cluster 1, when the TMA
producer got its slot back, when the converter saw the bytes land / the lo slot free / arrived, when the MMA thread saw
the stage converted and had its MMAs committed, and the epilogue's drain (ns relative to the first event)."""
import ctypes, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
csrc = os.path.join(ROOT, "fmri-connectivity-super-resolution_b200", "csrc")
out = os.path.join(tempfile.mkdtemp(), "libgsr_trace.so")
srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DGSR_T3_TRACE",
                       "-Xcompiler", "-fPIC", "-shared", "-o", out, *srcs, "-lcudart"])
lib = ctypes.CDLL(out)
B = 1024
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.rand(B, 268, 268, device="cuda", generator=g); Y = torch.rand(B, 268, 268, device="cuda", generator=g); C = torch.empty_like(X)
W = torch.rand(268, 268, device="cuda", generator=g)
vp, ll, i32, f32 = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
fn = lib.gsr_tcgemm_f32
fn.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, ll, ll, ll, i32, i32, i32, f32, i32, vp, i32, ll, vp]
per_subject = "--per-subject" in sys.argv
transB = 1 if "--nt" in sys.argv else 0
with_add = "--add" in sys.argv          # per-subject product with the GSR_EPI_ADD epilogue (the d adj products of the backward pass)
E = torch.rand(B, 268, 268, device="cuda", generator=g) if with_add else None
for _ in range(3):
    if per_subject:
        rc = fn(X.data_ptr(), Y.data_ptr(), C.data_ptr(), 268, 268, 268, 268, 268, 268, 268 * 268, 268 * 268, 268 * 268, 0, transB, B, 1.0,
                2 if with_add else 0, E.data_ptr() if with_add else None, 268 if with_add else 0, 268 * 268 if with_add else 0,
                torch.cuda.current_stream().cuda_stream)
    else:
        rc = fn(X.data_ptr(), W.data_ptr(), C.data_ptr(), B * 268, 268, 268, 268, 268, 268, 0, 0, 0, 0, transB, 1, 1.0, 0, None, 0, 0,
                torch.cuda.current_stream().cuda_stream)
    assert rc == 0
torch.cuda.synchronize()
ITS, EVS = 64, 12
n = 2 * 4 * ITS * EVS
buf = (ctypes.c_ulonglong * n)()
lib.gsr_tcgemm3_trace_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), i32]
assert lib.gsr_tcgemm3_trace_read(buf, n) == 0
t = list(buf)
t0 = min(x for x in t if x)
def ev(rank, role, it, e):
    x = t[((rank * 4 + role) * ITS + it) * EVS + e]
    return x - t0 if x else -1
print("mode:", "per-subject" if per_subject else "stacked", "NT" if transB else "NN", "+ addend" if with_add else "")
for rank in (0, 1):
    print(f"--- CTA rank {rank}")
    for it in range(30):
        line = f"kb {it:2d}: tma wait={ev(rank,0,it,0):6d} got={ev(rank,0,it,1):6d} | conv wait={ev(rank,2,it,0):6d} landed={ev(rank,2,it,1):6d} lo_free={ev(rank,2,it,2):6d} stored={ev(rank,2,it,4):6d} fenced={ev(rank,2,it,5):6d} arrived={ev(rank,2,it,3):6d}"
        if rank == 0:
            line += f" | mma wait={ev(0,1,it,0):6d} full={ev(0,1,it,1):6d} committed={ev(0,1,it,2):6d}"
            if ev(0, 1, it, 3) >= 0:
                line += f" [tile: tmem_empty wait={ev(0,1,it,3):6d} got={ev(0,1,it,4):6d}]"
        print(line)
    for tile in range(4):
        print(f"tile {tile}: epi wait={ev(rank,3,tile,0):6d} tmem_full={ev(rank,3,tile,1):6d} drained={ev(rank,3,tile,2):6d} chunks=" +
              " ".join(f"{ev(rank,3,tile,3+c):6d}" for c in range(9)))
