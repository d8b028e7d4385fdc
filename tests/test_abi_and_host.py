"""CPU-only checks: the C-ABI library loads and exports every symbol include/gsr_sm100.h
declares (no compute calls -- there is no GPU here), the ctypes table covers the header, the
host-side mirror keeps the reference's module / parameter names and seeded initialisation, the
product package never touches oracle/, and the subject-parallel host logic works over gloo."""
import ast
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG_NAME, ROOT

from oracle import gsr_oracle as O
from oracle import ref_shim

PKG_DIR = os.path.join(ROOT, PKG_NAME)
HEADER = os.path.join(ROOT, "include", "gsr_sm100.h")


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^GSR_API\s+[\w\s\*]+?\b(gsr_[a-z0-9_]+)\s*\(", text, flags=re.M)))


def test_header_declares_the_path():
    syms = header_symbols()
    assert len(syms) >= 30
    for must in ("gsr_eigh_jacobi_f32", "gsr_eigh_tridiag_f32", "gsr_tcgemm_f32", "gsr_bgemm_f32", "gsr_normalize_adj_fwd", "gsr_vectorize",
                 "gsr_anti_vectorize", "gsr_pool_topk", "gsr_l1_loss_fwd", "gsr_eval_reduce", "gsr_flat_pack"):
        assert must in syms


def test_library_loads_and_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "libgsr_sm100.so not built (python __graft_entry__.py build)"
    handle = ctypes.CDLL(pkg.LIB_PATH)
    for name in header_symbols():
        assert hasattr(handle, name), f"{name} declared in gsr_sm100.h but not exported"
    lib = pkg.lib()
    assert lib.gsr_abi_version() == 1
    assert lib.gsr_vectorize_len(268, 0) == 35778 and lib.gsr_vectorize_len(160, 0) == 12720
    assert lib.gsr_vectorize_len(268, 1) == 35778 + 267
    assert lib.gsr_l1_partials(268 * 268) >= 1
    assert lib.gsr_eval_workspace() > 0
    assert lib.gsr_eigh_workspace(4, 160) == 0 and lib.gsr_eigh_workspace(4, 268) == 4 * 268 * 292
    # tridiagonal route: n <= 288, workspace grows linearly with the batch
    assert lib.gsr_eigh_tridiag_workspace(4, 289) == 0 and lib.gsr_eigh_tridiag_workspace(0, 160) == 0
    assert lib.gsr_eigh_tridiag_workspace(2, 268) > 2 * 2 * 268 * 268 * 8
    w1, w8 = lib.gsr_eigh_tridiag_workspace(1, 160), lib.gsr_eigh_tridiag_workspace(8, 160)
    assert w1 > 2 * 160 * 160 * 8 and 7 * w1 < w8 <= 8 * w1
    assert lib.gsr_launch_count() == 0           # nothing may launch on a CPU-only box


def test_ctypes_table_matches_header(pkg):
    from importlib import import_module
    table = import_module(PKG_NAME + "._lib").PROTOTYPES
    assert sorted(table) == header_symbols()


def test_ctypes_argument_counts_match_header(pkg):
    """Every ctypes prototype has as many arguments as the C declaration (no silent ABI drift)."""
    from importlib import import_module
    table = import_module(PKG_NAME + "._lib").PROTOTYPES
    text = re.sub(r"/\*.*?\*/", " ", open(HEADER).read(), flags=re.S)
    decls = dict(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    assert sorted(decls) == header_symbols()
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(table[name][1]), (name, n, len(table[name][1]))


def test_bad_arguments_are_rejected_without_a_gpu(pkg):
    lib = pkg.lib()
    assert lib.gsr_normalize_adj_fwd(None, None, None, None, 1, 4, 0, None) == -1
    assert b"normalize_adj_fwd" in lib.gsr_last_error()
    assert lib.gsr_eigh_jacobi_f32(None, None, None, None, None, 1, 4, 1, 0.0, 0, None) == -1
    assert lib.gsr_eigh_tridiag_f32(None, None, None, None, None, 0, 1, 4, 1, None) == -1
    assert b"eigh_tridiag" in lib.gsr_last_error()
    assert lib.gsr_set_tuning(b"no_such_knob", 1) == -1
    for knob in (b"tcgemm_pairs", b"tcgemm_ragged_rows", b"gemm_variant"):
        assert lib.gsr_set_tuning(knob, 1 if knob != b"gemm_variant" else 0) == 0
    # round-1 additions: MLP passes, shared-target L1 backward, flat bucket, ragged-row / pair GEMM entry
    assert lib.gsr_scale_div_f32(None, None, None, 4, None) == -1 and b"scale_div" in lib.gsr_last_error()
    assert lib.gsr_spectral_bwd_f32(None, None, None, None, None, None, 4, 4, None, None) == -1
    assert lib.gsr_spectral_bwd_workspace() >= 2
    assert lib.gsr_splitk_reduce_f32(None, None, 4, 4, 4, 2, 1.0, 0, None, 0, None) == -1
    assert b"splitk_reduce" in lib.gsr_last_error()
    assert lib.gsr_bn_stats_f32(None, 4, 4, None, None, None, None, 0.1, None) == -1
    assert lib.gsr_bn_act_fwd_f32(None, None, None, None, None, 1e-5, 0.1, 0, 1, None, 4, 4, None) == -1
    assert lib.gsr_bn_act_bwd_f32(None, None, None, None, None, None, 1e-5, 0.1, 0, 1, 1, None, None, None, 4, 4, None) == -1
    assert lib.gsr_l1_loss_bwd_shared_y(None, None, None, None, 4, 16, None) == -1
    assert lib.gsr_flat_pack(None, None, 3, None, 10, 1.0, None) == -1
    assert lib.gsr_tcgemm_f32(None, None, None, 268, 268, 268, 268, 268, 268, 0, 0, 0, 0, 0, 1, 1.0, 0, None, 0, 0, None) == -1
    assert lib.gsr_launch_count() == 0


def test_gemm_dispatch_host_rules(pkg):
    """Host-side dispatch of kernels.bgemm (no GPU): the K-split plan of a skinny product covers K exactly with 16-byte aligned
    slices, leaves the step's shapes alone, and the one-subject rule is about operands without a subject dimension."""
    K = pkg.kernels
    for M, N, Kd in ((512, 1024, 12720), (512, 1024, 35778), (512, 1024, 1024), (300, 268, 1500), (512, 516, 1030)):
        splits, kc, krem = K._splitk_plan(M, N, Kd)
        assert splits >= 2 and kc % 4 == 0 and kc >= K.SPLITK_MIN_SLICE - 3 and 0 <= krem < 4 * splits
        assert splits * kc + krem == Kd
    for shape in ((1024 * 268, 268, 268), (268, 268, 268), (512, 35778, 1024), (35778, 1024, 512), (160, 268, 536), (128, 1024, 4096)):
        assert K._splitk_plan(*shape) is None, shape         # stacked / per-subject / wide / short-K / small-M products: no split
    assert K.LIN_MAX_OUT >= 268 * 536 and K.LIN_MAX_K >= 536   # every Linear product of the Graph U-Net (models/ops.py:92-107)


def test_product_package_never_imports_the_oracle():
    offenders = []
    for dirpath, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    if any(n == "oracle" or n.startswith("oracle.") for n in names):
                        offenders.append(os.path.join(dirpath, f))
    assert not offenders
    for dirpath, _, files in os.walk(os.path.join(PKG_DIR, "csrc")):
        for f in files:
            assert "oracle" not in open(os.path.join(dirpath, f)).read()


def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError):
        pkg.normalize_adj_torch(torch.rand(4, 4)) if not torch.cuda.is_available() else (_ for _ in ()).throw(RuntimeError())
    with pytest.raises(RuntimeError):
        pkg.kernels.bgemm(torch.rand(4, 4), torch.rand(4, 4))


class Args:
    lr_dim, hr_dim, hidden_dim, padding, lmbda, epochs = 160, 268, 268, 0, 16, 1


def test_module_mirror_names_and_shapes(pkg):
    model = pkg.GSRNet(list(O.DEFAULT_KS), Args())
    sd = model.state_dict()
    assert list(sd.keys()) == O.state_dict_keys()
    ref = O.make_state_dict(0)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.items()}
    assert sum(v.numel() for v in sd.values()) == 1023496
    for attr in ("layer", "net", "gc1", "gc2", "lr_dim", "hr_dim", "hidden_dim"):
        assert hasattr(model, attr)
    for attr in ("start_gcn", "bottom_gcn", "end_gcn", "down_gcns", "up_gcns", "pools", "unpools", "ks", "l_n"):
        assert hasattr(model.net, attr)
    assert [p.pool_size(n) for p, n in zip(model.net.pools, (160, 144, 101, 61))] == [144, 101, 61, 30]
    assert model.layer.weights.shape == (268, 160) and model.layer.lr_dim == 160


@pytest.mark.skipif(not ref_shim.available(), reason="reference not mounted (GPU box)")
def test_seeded_initialisation_equals_the_reference(pkg):
    ref = ref_shim.load()
    torch.manual_seed(0)
    theirs = ref.GSRNet(list(O.DEFAULT_KS), Args()).state_dict()
    torch.manual_seed(0)
    ours = pkg.GSRNet(list(O.DEFAULT_KS), Args()).state_dict()
    assert list(theirs.keys()) == list(ours.keys())
    for k in theirs:
        assert torch.equal(theirs[k], ours[k]), k


def test_vectorizer_host_side_contract(pkg):
    MV = pkg.MatrixVectorizer
    assert MV.vectorize(np.zeros((1, 1))).shape == (0,)
    with pytest.raises(AssertionError):
        MV.anti_vectorize(torch.zeros(5), 5)
    with pytest.raises(TypeError):
        MV.anti_vectorize(np.zeros(10), 5)
    with pytest.raises(RuntimeError):
        MV.anti_vectorize(torch.zeros(10, dtype=torch.float64), 5)


def test_unpad_and_shards(pkg):
    x = torch.arange(36.).reshape(6, 6)
    assert torch.equal(pkg.unpad(x, 2), x[2:4, 2:4]) and torch.equal(pkg.unpad(x, 0), x)
    for n, w in ((1024, 8), (167, 4), (5, 8), (10, 3)):
        b = [pkg.shard_bounds(n, r, w) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1
    assert pkg.fold_assignment(8, 3) == [[0, 1, 2], [3, 4, 5], [6, 7]]
    assert pkg.fold_assignment(4, 3) == [[0, 1], [2], [3]]
    assert pkg.fold_assignment(2, 3) == [[0], [1]]


_GLOO_WORKER = r"""
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
pkg = importlib.import_module({pkg!r})
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(1234 + rank)                       # deliberately different weights per rank
model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
sp = pkg.SubjectParallel(model, broadcast_from=0)    # rank 0's weights everywhere
w0 = [p.detach().clone() for p in model.parameters()]
gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
dist.all_gather(gathered, w0[0])
assert all(torch.equal(g, gathered[0]) for g in gathered)
# the subjects of one step are sharded over the ranks; mean loss per shard
g = torch.Generator().manual_seed(7)
X, Y = torch.randn(10, 6, generator=g), torch.randn(10, 3, generator=g)
lo, hi = pkg.shard_bounds(10, rank, world)
assert hi - lo == 5
loss = (model(X[lo:hi]) - Y[lo:hi]).abs().mean()
loss.backward()
sp.sync_gradients()
# reference: the whole batch on one rank
ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
ref.load_state_dict(model.state_dict())
(ref(X) - Y).abs().mean().backward()
for p, q in zip(model.parameters(), ref.parameters()):
    assert torch.allclose(p.grad, q.grad, atol=1e-7), (p.grad - q.grad).abs().max()
# fold groups: 2 ranks, 3 folds -> one single-rank group each, folds dealt round-robin
folds, group, ranks = pkg.fold_groups(3)
assert ranks == [rank] and folds == [f for f in range(3) if f % world == rank]
dist.barrier()
dist.destroy_process_group()
print("worker", rank, "ok")
"""


def test_subject_parallel_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, pkg=PKG_NAME))
    port = 29500 + (os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"worker {r} ok" in o


_FOLD_WORKER = r"""
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
pkg = importlib.import_module({pkg!r})
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
assert world == 4
folds, group, ranks = pkg.fold_groups(3)               # 4 ranks, 3 folds -> {{0, 1}}, {{2}}, {{3}}
assert ranks == [[0, 1], [0, 1], [2], [3]][rank] and folds == [[0], [0], [1], [2]][rank]
fold = folds[0]
torch.manual_seed(100 + rank)                          # different weights on every rank before the broadcast
model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
sp = pkg.SubjectParallel(model, group=group, broadcast_from=0)     # the fold's first rank seeds its group
assert sp.world == len(ranks)
# each fold trains on its own subjects; the ranks of a fold shard them
g = torch.Generator().manual_seed(50 + fold)
X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
lo, hi = pkg.shard_bounds(8, ranks.index(rank), len(ranks))
loss = (model(X[lo:hi]) - Y[lo:hi]).abs().mean()
loss.backward()
local = [p.grad.clone() for p in model.parameters()]
sp.sync_gradients()
ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
ref.load_state_dict(model.state_dict())
ref_loss = (ref(X) - Y).abs().mean()
ref_loss.backward()
for p, q in zip(model.parameters(), ref.parameters()):            # == the fold's whole batch on one rank
    assert torch.allclose(p.grad, q.grad, atol=1e-7), (rank, (p.grad - q.grad).abs().max())
if len(ranks) == 2:                                               # the multi-rank group really reduced
    assert any(not torch.equal(a, p.grad) for a, p in zip(local, model.parameters()))
else:                                                             # single-rank folds: nothing touched the gradient
    assert all(torch.equal(a, p.grad) for a, p in zip(local, model.parameters()))
# no traffic between folds: every fold has its own weights and its own loss; equal inside a fold
flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()] + [ref_loss.detach().reshape(1)])
everyone = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(everyone, flat)
assert torch.equal(everyone[0], everyone[1])
assert not torch.equal(everyone[0], everyone[2]) and not torch.equal(everyone[2], everyone[3])
assert len({{float(e[-1]) for e in everyone}}) == 3
dist.barrier()
dist.destroy_process_group()
print("worker", rank, "ok")
"""


def test_fold_groups_gloo_world4(tmp_path):
    """BASELINE configs[2] host logic: 3 cross-validation folds on 4 ranks -> one two-rank fold group that
    all-reduces its shards' gradients and two single-rank folds; nothing crosses fold boundaries."""
    script = tmp_path / "fold_worker.py"
    script.write_text(_FOLD_WORKER.format(root=ROOT, pkg=PKG_NAME))
    port = 31500 + (os.getpid() % 2000)
    procs = []
    for r in range(4):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="4", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"worker {r} ok" in o
