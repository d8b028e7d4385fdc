"""Timeline of the chunked n <= 160 eigendecomposition (csrc/eigh_tridiag.cu): builds a private libgsr copy with
-DGSR_EIGH_TRACE (a timing event after every stage of every chunk on its internal stream) and prints, per chunk, when
K1 / K2 / K3 / K4 started and ended (ms since the fork), for 1024 graphs of 160 nodes."""
import ctypes, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import torch
csrc = os.path.join(ROOT, "fmri-connectivity-super-resolution_b200", "csrc")
out = os.path.join(tempfile.mkdtemp(), "libgsr_trace.so")
srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith(".cu")]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DGSR_EIGH_TRACE",
                       "-Xcompiler", "-fPIC", "-shared", "-o", out, *srcs, "-lcudart"])
lib = ctypes.CDLL(out)
B, n = 1024, 160
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(B, n, n, device="cuda", generator=g); lr = (x + x.transpose(1, 2)) / 2; lr.diagonal(dim1=1, dim2=2).zero_()
d = lr.sum(2).pow(-0.5); A = (lr * d[:, :, None] * d[:, None, :]).contiguous()
vp, i32 = ctypes.c_void_p, ctypes.c_int
lib.gsr_eigh_tridiag_workspace.restype = ctypes.c_size_t
lib.gsr_eigh_tridiag_workspace.argtypes = [i32, i32]
ws_bytes = lib.gsr_eigh_tridiag_workspace(B, n)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
w = torch.empty(B, n, device="cuda"); U = torch.empty(B, n, n, device="cuda"); info = torch.empty(B, dtype=torch.int32, device="cuda")
fn = lib.gsr_eigh_tridiag_f32
fn.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, i32, i32, i32, vp]
lib.gsr_set_tuning.argtypes = [ctypes.c_char_p, i32]
lib.gsr_eigh_timeline_read.argtypes = [ctypes.POINTER(ctypes.c_float), i32]
for chunks in [int(a) for a in sys.argv[1:]] or [8]:
    lib.gsr_set_tuning(b"eigh_chunks", chunks)
    for _ in range(3):
        rc = fn(A.data_ptr(), w.data_ptr(), U.data_ptr(), info.data_ptr(), ws.data_ptr(), ws_bytes, B, n, 1, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
    buf = (ctypes.c_float * 40)()
    lib.gsr_eigh_timeline_read(buf, max(chunks, 1))
    print(f"chunks = {chunks}: ms since the fork     K1 launch (chunk 0 row only when chunked: ONE K1 launch)    K2 start (chunked: resident, polling)    K2 end=K3 start    K3 end=K4 start    K4 end")
    for c in range(max(chunks, 1)):
        print(f"  chunk {c}: " + "  ".join(f"{buf[c * 5 + k]:8.3f}" for k in range(5)))
