#!/usr/bin/env python
"""bench.py -- GSR-Net fwd+bwd subjects/s (LR 160 -> HR 268) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
                    [--scaling weak|strong] [--folds F] [--e2e-format vector|matrix]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one training step of the reference's loop (utils/gsr_train.py:23-46) over a batch
of B synthetic subjects per GPU: adjacency normalisation, LR eigendecomposition, Graph U-Net,
GSR layer, two graph convolutions, the three-term L1 loss, the full backward, the gradient
all-reduce (N > 1) and the Adam update, captured once in a CUDA graph (GraphedTrainStep) and
replayed; --no-graph launches every kernel from Python.  `value` = subjects of all ranks /
max-over-ranks device time, inputs resident in HBM.  `e2e` = the same step through the public API with the
LR / HR batches coming from pinned host memory -- as vectorised upper triangles, the dataset's on-disk form, rebuilt on
the device by gsr_anti_vectorize (utils/preprocessing.py:72-73) -- and the loss read back, every step.
--scaling strong keeps the GLOBAL batch at --batch (1024 subjects over all ranks); --folds F trains F independent
cross-validation models on disjoint rank groups (BASELINE configs[2]; training_gsr.ipynb:88-96).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG_NAME = "fmri-connectivity-super-resolution_b200"

LR_DIM, HR_DIM = 160, 268
# algorithmic work of the subject-dependent contractions, per subject
GEMM_FLOP_PER_SUBJECT_FWD_BWD = 635.4e6
METRIC = "GSR-Net fwd+bwd subjects/sec (160->268 nodes)"


class Args:
    lr_dim, hr_dim, hidden_dim, padding, lmbda, epochs = LR_DIM, HR_DIM, HR_DIM, 0, 16, 1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=1024, help="subjects per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm-variant", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="torch", choices=["torch", "flat"],
                    help="torch: torch.optim.Adam(fused=True) (default); flat: the library's gsr_adam_step (pkg.FlatAdam)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--cpu-sample", type=int, default=None, help="subjects in the CPU baseline sample")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch subjects per GPU (default); strong: --batch subjects over all GPUs")
    ap.add_argument("--folds", type=int, default=0, help="F > 0: fold-parallel cross-validation, F independent models on disjoint rank groups")
    ap.add_argument("--e2e-format", default="vector", choices=["vector", "matrix"],
                    help="what the end-to-end loop ships per step: vectorised upper triangles (default) or full matrices")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--workload", default="gsrnet", choices=["gsrnet", "mlp"],
                    help="mlp: only the notebook-MLP line of BASELINE configs[3] (also reported inside the default line at N = 1)")
    ap.add_argument("--no-mlp", action="store_true")
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=VALUE",
                    help="gsr_set_tuning(KEY, VALUE) before the run (A/B switches of the kernels, e.g. tcgemm_tma=0)")
    return ap.parse_args()


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  nvidia-smi is started before the
    warm-up steps (its start-up attaches to the driver and can stall launches for a second on a fresh box, which
    must not fall into the timed region) and every row is stamped on arrival; `stop()` keeps the rows that arrived
    between `begin()` and `end()` (plus one sampling period)."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 50

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.thread = [], None, None
        self.gpu_index = gpu_index
        self.t_begin = self.t_end = None

    def start(self, wait_s: float = 10.0):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.monotonic(), line.strip()))
        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()
        t0 = time.monotonic()
        while not self.rows and time.monotonic() - t0 < wait_s and self.proc.poll() is None:
            time.sleep(0.02)

    def begin(self):
        self.t_begin = time.monotonic()

    def end(self):
        self.t_end = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(1.5 * self.PERIOD_MS / 1e3)          # the row that covers the end of the region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        lo = self.t_begin if self.t_begin is not None else 0.0
        hi = (self.t_end if self.t_end is not None else time.monotonic()) + 1.5 * self.PERIOD_MS / 1e3
        rows = [r for t, r in self.rows if lo <= t <= hi]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "period_ms": self.PERIOD_MS}


# ------------------------------------------------------------------------------ CPU reference arm
def oracle_step_seconds(n_subjects: int, threads: int, with_adam: bool = True):
    """The reference's per-subject training step (fwd, HR-target eigh, 3-term L1 loss, bwd, Adam)
    restated by oracle/gsr_oracle.py, timed on the host cores.  Returns seconds for n_subjects."""
    import torch
    from collections import OrderedDict
    from oracle import gsr_oracle as O
    torch.set_num_threads(threads)
    sd = O.make_state_dict(0)
    params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items())
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    lrs = [O.synth_connectome(LR_DIM, 1000 + s) for s in range(min(n_subjects, 16))]
    hrs = [O.synth_connectome(HR_DIM, 1500 + s) for s in range(min(n_subjects, 16))]

    def one(s):
        opt.zero_grad()
        loss, err, _ = O.train_loss(params, lrs[s % len(lrs)], hrs[s % len(hrs)])
        loss.backward()
        if with_adam:
            opt.step()
        return float(loss)

    one(0); one(1)                      # warm-up
    t0 = time.perf_counter()
    for s in range(n_subjects):
        one(s)
    return time.perf_counter() - t0


def oracle_gpu_eager_seconds(n_subjects: int):
    """The same per-subject training step (oracle port, op for op the reference's eager PyTorch graph: padded mat-muls,
    diag-matmul normalisation, torch.linalg.eigh per subject, torch.optim.Adam) with every tensor on the B200: the
    reference as it runs today with `device = cuda` (utils/gsr_train.py:12-13) -- cuBLAS / cuSOLVER / ATen kernels, none
    of this library's.  A BASELINE leg like cpu_baseline; the product never routes through it."""
    import torch
    from collections import OrderedDict
    from oracle import gsr_oracle as O
    dev = torch.device("cuda", torch.cuda.current_device())
    sd = O.make_state_dict(0)
    params = OrderedDict((k, v.clone().to(dev).requires_grad_(True)) for k, v in sd.items())
    opt = torch.optim.Adam(list(params.values()), lr=1e-4)
    lrs = [O.synth_connectome(LR_DIM, 1000 + s).to(dev) for s in range(min(n_subjects, 16))]
    hrs = [O.synth_connectome(HR_DIM, 1500 + s).to(dev) for s in range(min(n_subjects, 16))]

    def one(s):
        opt.zero_grad()
        loss, err, _ = O.train_loss(params, lrs[s % len(lrs)], hrs[s % len(hrs)])
        loss.backward()
        opt.step()
        return float(loss.detach()) + float(err.detach())          # the two .item() reads of utils/gsr_train.py:48-49

    with torch.device(dev):                 # torch.eye / zeros inside the oracle land on the GPU, as the reference's `.to(device)` calls do
        one(0); one(1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(n_subjects):
            one(s)
        torch.cuda.synchronize()
        return time.perf_counter() - t0


def bind_to_gpu_numa_node(local_gpu: int):
    """Pin this process to the CPUs NVML reports as local to its GPU BEFORE it allocates pinned host buffers, so that the
    staging memory of every rank lives on the NUMA node its GPU's PCIe root hangs off.  Returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_gpu)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"process bound to the {len(cpus)} CPUs local to GPU {local_gpu} (NVML) before the pinned buffers were allocated"
    except Exception as e:          # noqa: BLE001 -- affinity is an optimisation, never a failure
        return f"no NUMA binding ({type(e).__name__})"
    return "no NUMA binding (empty affinity mask)"


def ncu_traffic(summary_file: str, kernel_substr: str):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes) of the first launch of `kernel_substr` in a committed
    profiles/*.txt summary (tools/ncu_summary.py output), or None."""
    try:
        lines = open(os.path.join(ROOT, "profiles", summary_file)).read().splitlines()
    except OSError:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    inside, got = False, {}
    for ln in lines:
        if ln.startswith("== "):
            if inside and len(got) == 2:
                break
            inside, got = kernel_substr in ln, {}
        elif inside and ln.startswith("dram__bytes_"):
            key, rest = ln.split(" = ")
            val, u = rest.split()[:2]
            got[key.strip()] = float(val.replace(",", "")) * unit.get(u, 1.0)
    return sum(got.values()) if len(got) == 2 else None


def ncu_traffic_all(summary_file: str):
    """{kernel name: dram bytes read + written per launch} for every launch section of a committed profiles/*.txt summary
    (first launch of each kernel), or None."""
    try:
        lines = open(os.path.join(ROOT, "profiles", summary_file)).read().splitlines()
    except OSError:
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out, name = {}, None
    for ln in lines:
        if ln.startswith("== "):
            name = ln[3:].split("(")[0].replace("void ", "").strip()
            if name in out:
                name = None
            else:
                out[name] = 0.0
        elif name and ln.startswith("dram__bytes_"):
            val, u = ln.split(" = ")[1].split()[:2]
            out[name] += float(val.replace(",", "")) * unit.get(u, 1.0)
    return out or None


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port --
    /root/reference does not exist on the GPU box), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    per_step = args.cpu_sample or 24
    oracle_step_seconds(2, threads)      # import + warm-up
    times = []
    for _ in range(args.warmup):
        oracle_step_seconds(2, threads)
    for _ in range(args.steps):
        times.append(oracle_step_seconds(per_step, threads))
    total = sum(times)
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "subjects/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"GSR-Net train step, batch 1 per optimiser step as utils/gsr_train.py:23-46, "
                               f"LR {LR_DIM} -> HR {HR_DIM}; each bench step = {per_step} subjects (bounded sample)",
                   "hr_target_eigh": "in step (as the reference)", "optimizer": "Adam lr=1e-4"},
        "cpu_baseline": {"value": value, "unit": "subjects/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step * args.steps} subjects, per-subject steps, torch-CPU fp32 (MKL), {threads} threads"},
        "e2e": {"value": value, "unit": "subjects/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def synth_batch(torch, n, batch, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.rand(batch, n, n, device=device, generator=g, dtype=torch.float32)
    m = (x + x.transpose(1, 2)) / 2
    m.diagonal(dim1=1, dim2=2).zero_()
    return m.contiguous()


def mlp_measure(pkg, torch, dev, steps=5, batch=512, hidden=1024, n_layers=2):
    """BASELINE configs[3]: the notebook MLP (training_template.ipynb:100-216) on vectorised upper triangles, 160-node LR ->
    268-node HR = 12 720 -> hidden -> 35 778 features; one training step = forward + backward + fused Adam on dense
    [batch, 160, 160] inputs already resident in HBM.  Secondary line (the headline metric is the GSR-Net step)."""
    torch.manual_seed(0)
    model = pkg.SuperResMLPNotebook(LR_DIM, HR_DIM, hidden, n_layers).to(dev)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    x = synth_batch(torch, LR_DIM, batch, 77, dev)
    y = synth_batch(torch, HR_DIM, batch, 78, dev)
    crit = pkg.functional.l1_loss

    def one():
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        return loss

    n0 = pkg.launch_count()
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = one()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    n_in, n_out = LR_DIM * (LR_DIM - 1) // 2, HR_DIM * (HR_DIM - 1) // 2
    flop = 3 * 2.0 * batch * (n_in * hidden + n_layers * hidden * hidden + hidden * n_out)        # fwd + dX + dW of every Linear
    return {"workload": f"BASELINE configs[3]: notebook MLP {n_in} -> {hidden} x {n_layers} residual blocks -> {n_out}, batch {batch}, "
                        f"fwd + bwd + Adam (spectral-norm Linear, BatchNorm1d + Dropout + LeakyReLU, residual adds, Sigmoid, "
                        f"triangle scatter: all on libgsr_sm100 kernels)",
            "samples_per_s": batch / (ms * 1e-3), "ms_per_step": ms, "gemm_tflops": flop / (ms * 1e-3) / 1e12,
            "loss_finite": bool(torch.isfinite(loss)), "gpu_launches_per_step": (pkg.launch_count() - n0) // (steps + 3),
            "note": "M = batch = 512 rows against K = 12 720 / 35 778 gives 8 tiles of 256 x 256: the long inner dimensions run as K "
                    "slices of one batched tensor launch + gsr_splitk_reduce_f32 (kernels.bgemm), the [512, 35 778] gradient is padded "
                    "to 16-byte aligned rows for its two products (functional.LinearFn), W / sigma of spectral_norm and its gradient run "
                    "on gsr_scale_div_f32 / gsr_spectral_bwd_f32 (the power iteration's three mat-vecs per layer are torch's, as in its "
                    "hook); gemm_tflops is the WHOLE step (BatchNorm, Sigmoid, triangle scatter and Adam over 51.7 M parameters included) divided into the "
                    "Linear flops; the forward of the output layer (N = 35 778, ldc not a multiple of 4) stays on the thread-staged "
                    "tensor kernel"}


def run_folds(args):
    """--folds F: fold-parallel cross-validation (BASELINE configs[2]; the KFold loop of training_gsr.ipynb:88-96 with one
    model per fold, unet-transformer.py:484 for the 3 folds).  The ranks are split into F disjoint groups
    (subject_parallel.fold_groups: 8 GPUs, 3 folds -> {0,1,2} {3,4,5} {6,7}); every group trains ITS OWN model on ITS OWN
    subjects, subject-data-parallel inside the group with one NCCL all-reduce of the flat gradient bucket per step, and no
    traffic between groups.  With fewer ranks than folds every rank trains its folds one after the other."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29513")
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG_NAME)
    pkg.lib()
    F, B = args.folds, args.batch
    my_folds, group, ranks = pkg.fold_groups(F)
    results = []
    for fold in my_folds:
        torch.manual_seed(100 + fold)                                     # one model per fold
        model = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], Args()).to(dev)
        gsize = len(ranks)
        sp = pkg.SubjectParallel(model, group=group) if gsize > 1 else None
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
        gr = ranks.index(rank)
        lr = synth_batch(torch, LR_DIM, B, 7000 + 97 * fold + gr, dev)    # the fold's own subjects, sharded over its ranks
        hr = synth_batch(torch, HR_DIM, B, 9000 + 97 * fold + gr, dev)
        U_hr = pkg.hr_target_basis(hr, LR_DIM, 0, check_errors=True)
        graphed = pkg.GraphedTrainStep(model, opt, B, lmbda=Args.lmbda, padding=Args.padding,
                                       grad_sync=sp.sync_gradients if sp is not None else None, warmup=max(args.warmup, 3))
        graphed.load(lr, hr, U_hr)
        for _ in range(max(args.warmup, 3)):
            graphed.replay()
        torch.cuda.synchronize()
        if gsize > 1:
            dist.barrier(group=group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss, err = graphed.replay()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if gsize > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        w0 = float(model.layer.weights.detach().double().sum())           # identical on the ranks of a fold, different across folds
        results.append({"fold": fold, "ranks": ranks, "subjects_per_step": gsize * B, "ms_per_step": float(t) / args.steps,
                        "subjects_per_s": gsize * B * args.steps / (float(t) / 1e3), "rank": rank,
                        "loss_local_shard": float(loss), "weights_checksum": w0})
        del graphed
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        per_fold = {}
        for rs in gathered:
            for r in rs:
                per_fold.setdefault(r["fold"], []).append(r)
        folds_out, total = [], 0.0
        for fold in sorted(per_fold):
            rs = per_fold[fold]
            same = all(abs(r["weights_checksum"] - rs[0]["weights_checksum"]) == 0.0 for r in rs)
            folds_out.append({"fold": fold, "ranks": rs[0]["ranks"], "subjects_per_step": rs[0]["subjects_per_step"],
                              "ms_per_step": max(r["ms_per_step"] for r in rs), "subjects_per_s": min(r["subjects_per_s"] for r in rs),
                              "weights_identical_inside_fold": same, "weights_checksum": rs[0]["weights_checksum"],
                              "loss_per_rank": [r["loss_local_shard"] for r in rs]})
            total += folds_out[-1]["subjects_per_s"]
        distinct = len({f["weights_checksum"] for f in folds_out}) == len(folds_out)
        concurrent = world >= F
        # whole-job rate: all subjects of all folds / the busiest rank's time (a rank that trains several folds runs them in turn)
        busy_ms = max(sum(r["ms_per_step"] for r in rs) for rs in gathered if rs)
        total_subjects_per_step = sum(f["subjects_per_step"] for f in folds_out)
        line = {"metric": METRIC, "value": total_subjects_per_step / (busy_ms / 1e3),
                "unit": "subjects/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": max(f["ms_per_step"] for f in folds_out), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[2]: {F}-fold cross-validation, fold-parallel x subject-data-parallel, "
                                       f"{B} subjects per GPU per step, LR {LR_DIM} -> HR {HR_DIM}",
                           "fold_groups": [f["ranks"] for f in folds_out],
                           "folds_run": "concurrently on disjoint rank groups" if concurrent else "round-robin on the ranks, one after the other",
                           "collectives": "one NCCL all-reduce of the 4.09 MB gradient bucket per step INSIDE each fold group; none across folds"},
                "folds": folds_out, "weights_differ_across_folds": distinct,
                "gpu_launches": None}
        print(json.dumps(line), flush=True)
    sys.stdout.flush()
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GSR-Net path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG_NAME)
    pkg.lib()
    K = pkg.kernels
    if args.gemm_variant is not None:
        pkg.set_tuning("gemm_variant", args.gemm_variant)
    for kv in args.tune:
        key, val = kv.split("=")
        pkg.set_tuning(key, int(val))

    numa_note = bind_to_gpu_numa_node(local)
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"bench.py --scaling strong: --batch {args.batch} is not divisible by {world} GPUs")
        B = args.batch // world            # the GLOBAL batch is fixed; every rank takes its shard (shard_bounds)
    else:
        B = args.batch
    torch.manual_seed(0)
    model = pkg.GSRNet([0.9, 0.7, 0.6, 0.5], Args()).to(dev)
    sp = pkg.SubjectParallel(model) if world > 1 else None
    use_graph = not args.no_graph
    if args.optimizer == "flat":
        opt = pkg.FlatAdam(model.parameters(), lr=1e-4)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=use_graph)
    step = pkg.GSRTrainStep(model, lmbda=Args.lmbda, padding=Args.padding)
    lr = synth_batch(torch, LR_DIM, B, 1000 + rank, dev)
    hr = synth_batch(torch, HR_DIM, B, 5000 + rank, dev)
    # HR-target basis (utils/gsr_train.py:33-36): data-only (the weights do not enter it), so the
    # trainer computes it once per dataset.  Default here: hoisted, computed with the library's own
    # 268-node eigh kernels outside the timed region; GSR_BENCH_HR_BASIS=in_step recomputes it
    # every step as the reference does.
    in_step_hr = os.environ.get("GSR_BENCH_HR_BASIS", "hoisted") == "in_step"
    if in_step_hr:
        U_hr = None
        hr_note = "in step: recomputed every step with the library's 268-node tridiagonal-QL kernels (csrc/eigh_tridiag.cu, n = 161..288 variants)"
    else:
        U_hr = pkg.hr_target_basis(hr, LR_DIM, 0, check_errors=True)
        hr_note = ("hoisted: data-only, computed once outside the timed region with the library's own 268-node "
                   "tridiagonal-QL kernels (SURVEY 8 a12); `with_hr_target_eigh_in_step` below is the same step with it "
                   "recomputed every step (GSR_BENCH_HR_BASIS=in_step makes that the headline)")

    def one_step(lr_b, hr_b):
        opt.zero_grad(set_to_none=True)
        loss, err = step(lr_b, hr_b, U_hr_reduced=U_hr)
        if sp is not None:
            sp.sync_gradients()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                 # before the warm-up: nvidia-smi's start-up must not fall into the timed region
    # The whole optimiser step (forward, backward, gradient all-reduce, Adam) is captured once in a CUDA graph
    # (pkg.GraphedTrainStep, SURVEY 8 f row N3) and the timed region replays it: the ~180 launches of a step keep the
    # host busy for 10 of the 15 ms when issued from Python, so an eager loop is only just device-bound.
    graphed = None
    if use_graph:
        graphed = pkg.GraphedTrainStep(model, opt, B, lmbda=Args.lmbda, padding=Args.padding, hr_basis_in_graph=in_step_hr,
                                       grad_sync=sp.sync_gradients if sp is not None else None, warmup=max(args.warmup, 3),
                                       slots=2)
        graphed.load(lr, hr, U_hr)      # the batch is resident in the graph's input buffers when the timed region starts
        timed_step = graphed.replay
    else:
        timed_step = lambda: one_step(lr, hr)
    for _ in range(max(args.warmup, 3)):
        timed_step()
    barrier()

    # ---- timed region: K steps, device time, max over ranks
    launches0 = pkg.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    for _ in range(args.steps):
        timed_step()
    e1.record()
    barrier()
    sampler.end()
    elapsed_ms = e0.elapsed_time(e1)
    launches = graphed.launches_per_step * args.steps if use_graph else pkg.launch_count() - launches0

    # ---- per-kernel durations for the rooflines: the same K steps launched eagerly with CUDA events around every
    #      library call (events cannot be placed inside a graph replay); shares are relative to the timed step
    K.GEMM_TIMER, K.EIGH_TIMER, K.EW_TIMER = K.KernelTimer(), K.KernelTimer(), K.KernelTimer()
    e0i, e1i = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0i.record()
    for _ in range(args.steps):
        one_step(lr, hr)
    e1i.record()
    barrier()
    instrumented_ms = e0i.elapsed_time(e1i)
    # contractions on the tensor-core kernel (per-subject and stacked 268-wide products) and, separately, the small
    # Graph U-Net products on the SIMT kernel (they run on a side stream under the eigendecomposition)
    gemm_ms, gemm_flops, gemm_launches = K.GEMM_TIMER.totals("tc" if K.GEMM_BACKEND == "tc" else None)
    simt_ms, simt_flops, simt_launches = K.GEMM_TIMER.totals("simt") if K.GEMM_BACKEND == "tc" else (0.0, 0.0, 0)
    eigh_ms, eigh_graphs, eigh_launches = K.EIGH_TIMER.totals()
    ew_ms, ew_bytes, ew_launches = K.EW_TIMER.totals()
    eigh_info = getattr(K.EIGH_TIMER, "last_info", None)
    eigh_timer = K.EIGH_TIMER
    K.GEMM_TIMER = K.EIGH_TIMER = K.EW_TIMER = None
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t)
    value = world * B * args.steps / (elapsed_ms / 1e3)

    # ---- end to end: pinned host buffers -> device -> step -> loss back, every step.
    # Double-buffered as a training loop over a host-resident dataset would be: the H2D copy of step i+1 runs on a copy
    # stream while step i computes; every step's inputs are copied inside the timed region and every step's loss is read
    # back to pinned host memory.  Default format: the vectorised upper triangles the challenge data comes as
    # (12 720 + 35 778 floats per subject, half the bytes of the matrices); gsr_anti_vectorize rebuilds the symmetric
    # zero-diagonal matrices on the device, exactly what the reference's loader does per CSV row
    # (utils/preprocessing.py:72-73), straight into the input buffers of the captured step.
    vec_fmt = args.e2e_format == "vector"
    if vec_fmt:
        lr_src, hr_src = K.triu_flatten(lr), K.triu_flatten(hr)             # [B, 12720], [B, 35778]
        assert torch.equal(K.anti_vectorize(lr_src, LR_DIM), lr) and torch.equal(K.anti_vectorize(hr_src, HR_DIM), hr)
    else:
        lr_src, hr_src = lr, hr
    lr_host = lr_src.cpu().pin_memory()
    hr_host = hr_src.cpu().pin_memory()
    if use_graph:
        # the two input-buffer sets of the graphed step are the double buffer: set (i + 1) % 2 is filled from pinned host
        # memory while the graph captured on set i % 2 replays
        lr_dev = [graphed.inputs(0)[0], graphed.inputs(1)[0]]
        hr_dev = [graphed.inputs(0)[1], graphed.inputs(1)[1]]
        if U_hr is not None:
            graphed.inputs(1)[2].copy_(U_hr)            # the hoisted HR basis is resident in both sets (set 0 was loaded above)
    else:
        lr_dev = [torch.empty_like(lr), torch.empty_like(lr)]
        hr_dev = [torch.empty_like(hr), torch.empty_like(hr)]
    lr_stage = [torch.empty_like(lr_src), torch.empty_like(lr_src)] if vec_fmt else lr_dev
    hr_stage = [torch.empty_like(hr_src), torch.empty_like(hr_src)] if vec_fmt else hr_dev
    U_host = U_hr.cpu().pin_memory() if U_hr is not None else None
    copy_stream = torch.cuda.Stream()

    def e2e_run(nsteps, copy_basis=False):
        loss_host = torch.empty((nsteps,), dtype=torch.float32).pin_memory()
        ev_copy = [torch.cuda.Event(), torch.cuda.Event()]
        ev_done = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream()

        def issue_copy(i):
            buf = i % 2
            with torch.cuda.stream(copy_stream):
                if i >= 2:
                    copy_stream.wait_event(ev_done[buf])          # step i-2 has consumed this buffer
                lr_stage[buf].copy_(lr_host, non_blocking=True)
                hr_stage[buf].copy_(hr_host, non_blocking=True)
                if copy_basis and use_graph and U_host is not None:
                    graphed.inputs(buf)[2].copy_(U_host, non_blocking=True)
                ev_copy[buf].record(copy_stream)

        issue_copy(0)
        for i in range(nsteps):
            buf = i % 2
            if i + 1 < nsteps:
                issue_copy(i + 1)
            main.wait_event(ev_copy[buf])
            if vec_fmt:                                           # two launches: vectors -> the step's input matrices
                K.anti_vectorize(lr_stage[buf], LR_DIM, out=lr_dev[buf])
                K.anti_vectorize(hr_stage[buf], HR_DIM, out=hr_dev[buf])
            if use_graph:
                loss, _ = graphed.replay(buf)
                ev_done[buf].record(main)
            else:
                loss = one_step(lr_dev[buf], hr_dev[buf])
                ev_done[buf].record(main)
            loss_host[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        main.synchronize()
        return loss_host

    def e2e_timed(copy_basis):
        e2e_run(2, copy_basis)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        losses = e2e_run(args.steps, copy_basis)
        e3.record()
        barrier()
        assert bool(torch.isfinite(losses).all())
        t = torch.tensor([e2.elapsed_time(e3)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * B * args.steps / (float(t) / 1e3)

    e2e_value = e2e_timed(False)
    # a dataset of more than one batch: the data-only HR-target basis is either cached on the device (172 KB per subject;
    # the default above), recomputed in the step (`with_hr_target_eigh_in_step`), or shipped with the batch -- this variant:
    e2e_basis_value = e2e_timed(True) if (use_graph and U_host is not None) else None
    h2d_per_rank = (lr_host.numel() + hr_host.numel()) * 4

    # ---- the same step with the HR-target eigendecomposition recomputed inside it, as utils/gsr_train.py:33-36
    #      literally does (2 steps; reported next to the headline, which hoists that data-only computation)
    hr_in_step = None
    if not in_step_hr:
        def one_step_hr(lr_b, hr_b):
            opt.zero_grad(set_to_none=True)
            loss, err = step(lr_b, hr_b, U_hr_reduced=None)
            if sp is not None:
                sp.sync_gradients()
            opt.step()
            return loss
        one_step_hr(lr, hr)
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for _ in range(2):
            one_step_hr(lr, hr)
        e5.record()
        barrier()
        t = torch.tensor([e4.elapsed_time(e5)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        hr_in_step = {"value": world * B * 2 / (float(t) / 1e3), "unit": "subjects/s", "ms_per_step": float(t) / 2,
                      "note": "HR-target eigh (268 nodes, tridiagonal-QL kernels with the packed triangle / the rows of Z in shared "
                              "memory: row N1 of SURVEY 8(f)) recomputed every step, as utils/gsr_train.py:33-36 literally does"}

    # ---- FP32 FFMA peak of this GPU (denominator for the SIMT contraction kernel)
    scratch = torch.zeros(4, device=dev)
    ffma_best = 0.0
    for it in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flops = pkg.lib().gsr_ffma_probe(1 << 15, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        b.record()
        torch.cuda.synchronize()
        if it:
            ffma_best = max(ffma_best, flops / (a.elapsed_time(b) * 1e-3) / 1e12)

    # ---- dense TF32 rate of the tensor pipe on this GPU (resident operands, no traffic): the measured denominator of the
    #      3xTF32 contraction kernels (three MMAs per algorithmic product -> their ceiling is a third of it)
    tf32_best = 0.0
    for it in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        flops = pkg.lib().gsr_tf32_probe(8192, torch.cuda.current_stream().cuda_stream)
        b.record()
        torch.cuda.synchronize()
        if it:
            tf32_best = max(tf32_best, flops / (a.elapsed_time(b) * 1e-3) / 1e12)

    # ---- the reference's eager PyTorch graph on this same B200 (cuBLAS / cuSOLVER / ATen; per-subject loop): the
    #      "existing kernels" bar next to the CPU arm
    gpu_eager = None
    if world == 1 and rank == 0 and not args.no_gpu_eager_baseline:
        n_eager = 48
        secs = oracle_gpu_eager_seconds(n_eager)
        gpu_eager = {"value": n_eager / secs, "unit": "subjects/s", "kind": "port on cuda",
                     "sample": f"{n_eager} subjects, per-subject fwd+bwd+Adam incl. both torch.linalg.eigh calls "
                               f"(utils/gsr_train.py:23-46 restated in oracle/, every tensor on the GPU: eager PyTorch / cuBLAS / "
                               f"cuSOLVER kernels, none of this library's)"}

    mlp_line = None
    if world == 1 and rank == 0 and not args.no_mlp:
        del lr_stage, hr_stage
        torch.cuda.empty_cache()
        mlp_line = mlp_measure(pkg, torch, dev)

    def finish():
        # N > 1: the captured graph holds the communicator's all-reduce; tearing the process group down with it
        # alive was observed to block at exit.  Release the graph, meet the other ranks and leave without the
        # NCCL / interpreter teardown.
        nonlocal graphed, timed_step
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            graphed = timed_step = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    if rank != 0:
        finish()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    backend = K.GEMM_BACKEND
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    bf16_peak = peaks.get("bf16_tflops_sustained") or 1400.0
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # --- contraction kernel: tensor-bound.  tcgen05 kind::tf32 runs at half the bf16 rate and the 3xTF32 split
    #     issues three MMAs per algorithmic product, both stated so that `frac` is against the MEASURED bf16 peak.
    tf32_peak = tf32_best if tf32_best > 0 else bf16_peak / 2.0
    tma_on = backend == "tc"
    traffic = ncu_traffic("r2d_ncu_summary_tcgemm3.txt", "tcgemm3_kernel") if backend == "tc" else None
    gemm_roof = {
        "kernel": ("gsr::tc::t3::tcgemm3_kernel (persistent CTA pairs, tcgen05.mma.cta_group::2 kind::tf32, 3xTF32: operands by "
                   "cp.async.bulk.tensor into SWIZZLE_128B / 128B_ATOM_32B tiles, the raw fp32 tile is the hi term, converter warps "
                   "write lo, two overlapping accumulators in TMEM, TMA-store epilogue; 256 x 272 x 32 tiles) + the FP32 "
                   "ragged-row kernels for rows 256..267 and the one-CTA gsr::tc::tcgemm_kernel for the few remaining M < 256 tiles"
                   if backend == "tc" else "gsr::bgemm_kernel (136x136x16 SIMT FP32 tile)"),
        "bound": "tensor" if backend == "tc" else "fp32",
        "achieved": gemm_tflops, "unit": "TFLOP/s (algorithmic fp32 flops of the contractions)",
        "peak": bf16_peak if backend == "tc" else ffma_best,
        "frac": (gemm_tflops / (bf16_peak if backend == "tc" else ffma_best)) if gemm_tflops else None,
        "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside the step).  kind::tf32 runs at half the "
                        "bf16 rate and the fp32-accurate 3xTF32 split issues three MMAs per product: see tf32_measured / "
                        "frac_of_3xtf32_ceiling for the denominators this kernel can actually reach"
                        if backend == "tc" else "FP32 FFMA probe kernel measured in this run"),
        "tf32_measured_tflops": tf32_best if tma_on else None,
        "tf32_measured_how": "gsr_tf32_probe in this run: 74 CTA pairs x 32768 tcgen05.mma.cta_group::2 kind::tf32 (M256 N256 K8) on "
                             "resident zero operands (no operand traffic, so an upper bound under the power cap), best of 4",
        "frac_of_measured_tf32": (gemm_tflops / tf32_peak) if (gemm_tflops and tma_on) else None,
        "frac_of_3xtf32_ceiling": (gemm_tflops / (tf32_peak / 3.0)) if (gemm_tflops and tma_on) else None,
        "frac_of_3xtf32_ceiling_bf16_derived": (gemm_tflops / (bf16_peak / 6.0)) if (gemm_tflops and tma_on) else None,
        "share_of_step": gemm_ms / elapsed_ms if elapsed_ms else None, "launches": gemm_launches,
        "algorithmic_flop_per_subject": GEMM_FLOP_PER_SUBJECT_FWD_BWD,
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (1024 per-subject 256 x 268 x 268 products) from the committed
        # ncu --set full capture; the algorithmic bytes of that launch are 1024 x (256 x 268 + 268 x 268 + 256 x 268) x 4 B = 856 MB
        "traffic": traffic,
        "traffic_source": "profiles/r2d_ncu_summary_tcgemm3.txt (tools/ncu_summary.py of the ncu --set full capture), first tcgemm3_kernel launch (1024 x 268^3 stacked product: 856 MB algorithmic)",
        "unet_simt_products": {"launches": simt_launches, "gflop_per_step": simt_flops / max(args.steps, 1) / 1e9,
                               "note": "Graph U-Net linears and their gradients (once per step, subject independent) on "
                                       "gsr::lin::lin_kernel (one product: whole operand panels in shared memory, 32 x 32 tiles, 4-6 us "
                                       "per launch), side stream under the eigendecomposition; not part of `achieved`"},
    }
    # --- LR eigendecomposition
    eigh_roof = None
    eigh_method = getattr(eigh_timer, "last_method", None)
    if eigh_ms > 0 and eigh_info is not None and eigh_method == "tridiag":
        # tridiagonal route (csrc/eigh_tridiag.cu): matrices in registers, bound by the FP32 FMA pipe.
        # Executed flops per graph: Householder reduction 2 n^3 (n^3 FMAs), Q accumulation in column form (a warp owns 32
        # columns and starts at reflector c0 + 1; a reflector k touches 16 ceil(k / 16) coordinates, 4 flops each: 6.6 MFLOP
        # at n = 160 = 0.81 of the 2 n^3 the row form executed), and 6 n per replayed plane rotation (2 FMUL + 2 FFMA on each
        # of the n rows); the rotation count (identity padding included, it is executed) is read back from one extra, untimed call.
        n = LR_DIM
        A_lr = pkg.normalize_adj_torch(lr)
        _, _, _, ws = K.eigh_tridiag(A_lr, _return_workspace=True)
        torch.cuda.synchronize()
        # The decomposition timed ALONE on the device (5 calls, CUDA events, nothing on other streams): inside a step it runs
        # as a pipeline over internal streams next to the Graph U-Net's side stream, so the events the instrumented pass
        # puts around the call on the caller's stream also see whatever shares the SMs with it (kept as `in_step_pass_ms`).
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            K.eigh_tridiag(A_lr)
        ev0.record()
        for _ in range(5):
            K.eigh_tridiag(A_lr)
        ev1.record()
        torch.cuda.synchronize()
        eigh_pass_ms = eigh_ms / max(eigh_launches, 1)
        eigh_alone_ms = ev0.elapsed_time(ev1) / 5
        eigh_ms, eigh_graphs_alone = eigh_alone_ms * max(eigh_launches, 1), eigh_graphs      # totals over the timed steps, as before
        cap = ((2 * n * n + 512 + 63) // 64) * 64 + 64
        al = lambda x: (x + 255) // 256 * 256
        off = al(B * cap * 8) + al(B * n * 164 * 4) + al(B * n * n * 4) + 2 * al(B * n * 4) + al(B * 4)
        counts = ws[off:off + B * 16].view(torch.int32).view(B, 4)
        rot_mean = float(counts[:, 1].double().mean())
        q_flops = float(sum(32 * sum(4 * 16 * ((k + 15) // 16) for k in range(c0 + 1, n)) for c0 in range(0, n, 32)))
        flops_per_graph = 2.0 * n ** 3 + q_flops + 6.0 * n * rot_mean
        tflops = flops_per_graph * eigh_graphs / (eigh_ms * 1e-3) / 1e12
        eigh_roof = {
            "kernel": "gsr::tq::{tridiag,form_q,tql_scalar_lanes,tql_apply,tql_finish}_kernel (Householder + implicit QL, one thread per "
                      "matrix row, rows resident in registers (K1: 128 of 160 columns, the rest in shared memory); QL iteration with lane = "
                      "graph; run as a pipeline: K1 in one launch, Q = H_{n-1} ... H_1 as a launch of its own behind it (column form: thread = column of Q), K2 CTAs resident "
                      "on their own SMs polling K1's per-chunk counts, K3 (rotation replay) / K4 per 128-graph chunk on internal streams)",
            "timed": "alone on the device, mean of 5 calls of 1024 graphs", "in_step_pass_ms": eigh_pass_ms, "alone_ms": eigh_alone_ms,
            "bound": "fp32", "achieved": tflops, "unit": "TFLOP/s (executed fp32 flops: 2 n^3 + 1.6 n^3 (Q in column form) + 6 n per plane rotation)",
            "peak": ffma_best, "frac": tflops / ffma_best if ffma_best else None,
            "peak_source": "FP32 FFMA probe kernel measured in this run",
            "share_of_step": eigh_ms / elapsed_ms, "launches": eigh_launches,
            "us_per_graph": eigh_ms * 1e3 / eigh_graphs if eigh_graphs else None,
            "ql_iterations_mean": float(eigh_info.double().mean()), "rotations_mean": rot_mean,
            "conventional_9n3_tflops": 9.0 * n ** 3 * eigh_graphs / (eigh_ms * 1e-3) / 1e12,
            "traffic": ncu_traffic_all("r2d_ncu_summary_eigh.txt"),
            "traffic_source": "profiles/r2d_ncu_summary_eigh.txt: dram bytes per launch of each stage (1024 graphs, one launch per stage, "
                              "the one-warp QL kernel in place of the lane = graph one)",
        }
    elif eigh_ms > 0 and eigh_info is not None:
        # Jacobi route: resident in shared memory, bound by the LSU / shared-memory data pipe
        sw = (eigh_info & 0xFF).double().sum().item()
        ps = (eigh_info >> 8).double().sum().item()
        n = LR_DIM
        # a one-sided sweep is n steps in which every slot loads one column and stores one (the other stays
        # in registers): n * n * 4 B read + as much written per step; a polish round reads 2n columns (U and
        # A.U) and writes back only what it rotates (counted as read only: the algorithmic minimum)
        bytes_per_launch = (sw * n * 2 * n * n * 4.0 / 2.0) + (ps * (n - 1) * 2 * n * n * 4.0)
        launches_per_timer = max(eigh_launches, 1)
        gbs = bytes_per_launch / (eigh_ms / launches_per_timer * 1e-3) / 1e9
        smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        eigh_roof = {
            "kernel": "gsr::eigh_jacobi_kernel<5, polish> (one CTA per 160-node graph, 212 KB shared memory)",
            "bound": "smem", "achieved": gbs, "unit": "GB/s (shared-memory bytes the Jacobi rounds must move)",
            "peak": smem_peak, "frac": gbs / smem_peak,
            "peak_source": f"148 SMs x 128 B/clk x {sm_mhz:.0f} MHz (clock sampled in this run); ncu: lsu data-pipe 66 % busy",
            "share_of_step": eigh_ms / elapsed_ms, "launches": eigh_launches,
            "us_per_graph": eigh_ms * 1e3 / eigh_graphs if eigh_graphs else None,
            "sweeps_mean": sw / eigh_info.numel(), "polish_sweeps_mean": ps / eigh_info.numel(), "traffic": None,
        }
    ew_roof = None
    if ew_ms > 0:
        gbs = ew_bytes / (ew_ms * 1e-3) / 1e9
        ew_roof = {"kernel": "normalize_{rowsum,scale_tiles,bwd_partial_tiles,bwd_apply_tiles}_kernel + sym_diag_abs_tiles_kernel<fwd,bwd> + "
                             "sym_diag_abs_l1_tiles_kernel<fwd,bwd> (the final sym-diag-abs and the L1 term against the HR target in one "
                             "pass): the memory-bound n x n kernels",
                   "bound": "hbm", "achieved": gbs, "unit": "GB/s", "peak": peaks.get("hbm_gbs"),
                   "frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs", "share_of_step": ew_ms / elapsed_ms,
                   "launches": ew_launches, "traffic": ncu_traffic_all("r2d_ncu_summary_nxn.txt"),
                   "traffic_source": "profiles/r2d_ncu_summary_nxn.txt: dram bytes per launch (1024 x 268 x 268; algorithmic: 2 tensors = 588 MB "
                                     "per forward kernel, 3 = 882 MB per backward pass, 4 with two incoming gradients)"}
    by_kernel = [r for r in (eigh_roof, gemm_roof, ew_roof) if r]
    roofline = max(by_kernel, key=lambda r: r["share_of_step"] or 0.0)          # the dominant kernel of the step
    others = [r for r in by_kernel if r is not roofline]
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:        # reported on rank 0 at N = 1 only
        threads = os.cpu_count() or 1
        n_cpu = args.cpu_sample or 96
        secs = oracle_step_seconds(n_cpu, threads)
        cpu_baseline = {"value": n_cpu / secs, "unit": "subjects/s", "cores": threads, "kind": "port",
                        "sample": f"{n_cpu} subjects, per-subject fwd+bwd+Adam incl. HR-target eigh "
                                  f"(utils/gsr_train.py:23-46 restated in oracle/), torch-CPU fp32, {threads} threads"}
    line = {
        "metric": METRIC, "value": value, "unit": "subjects/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: GSR-Net fwd+bwd (+grad all-reduce, +Adam) batched over {B} "
                               f"synthetic subjects per GPU, LR {LR_DIM} -> HR {HR_DIM}"
                               + (f" (strong scaling: {world * B} subjects over {world} GPUs)" if args.scaling == "strong" else ""),
                   "subjects_per_gpu": B, "global_subjects_per_step": world * B,
                   "parallelism": f"subject-data-parallel x{world}, one NCCL all-reduce of the flat 4.09 MB gradient bucket",
                   "lr_eigh": f"in step ({eigh_method or K.EIGH_METHOD} kernels)", "hr_target_eigh": hr_note,
                   "with_hr_target_eigh_in_step": hr_in_step,
                   "gemm_backend": K.GEMM_BACKEND, "tuning": args.tune or "defaults",
                   "l2": "inputs of one step are 400 MB per GPU (> 126 MB L2); no flush needed",
                   "optimizer": ("pkg.FlatAdam (gsr_adam_step), in the timed region" if args.optimizer == "flat"
                                 else "torch.optim.Adam(fused=True, capturable=True), in the timed region"),
                   "cuda_graph": ("the timed region replays one captured graph per step (forward, backward, gradient all-reduce, Adam; "
                                  "GraphedTrainStep); --no-graph launches every kernel from Python" if use_graph else "off (--no-graph)"),
                   "kernel_timing": (f"roofline kernel durations come from an instrumented eager pass of the same {args.steps} steps right "
                                     f"after the timed region (CUDA events around every library call; that pass is host bound, "
                                     f"{instrumented_ms / args.steps:.3f} ms per step); share_of_step = those device-side "
                                     f"durations / the timed (graph-replayed) step")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "subjects/s",
                "h2d_bytes_per_step": world * h2d_per_rank,
                "d2h_bytes_per_step": world * 4,
                "format": ("vectorised upper triangles (12 720 + 35 778 fp32 per subject, the dataset's on-disk form); "
                           "gsr_anti_vectorize rebuilds the matrices on the device inside the timed region "
                           "(utils/preprocessing.py:72-73)" if vec_fmt else "full fp32 matrices"),
                "hr_target_basis": "cached on the device across steps (data-only, 172 KB per subject): a dataset of more than one "
                                   "batch keeps it resident (1.7 GB per 10 000 subjects of the 180 GB), ships it with the batch "
                                   "(`with_basis_copied`) or recomputes it (`config.with_hr_target_eigh_in_step`)",
                "with_basis_copied": ({"value": e2e_basis_value, "unit": "subjects/s",
                                       "h2d_bytes_per_step": world * (h2d_per_rank + B * HR_DIM * LR_DIM * 4)}
                                      if e2e_basis_value else None),
                "host_buffers": numa_note,
                "pipeline": "H2D of step i+1 overlaps the compute of step i (two device buffers, copy stream)"
                            + ("; the copies land in the input buffers of the graph captured for that buffer set, which is then replayed" if use_graph else "")},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_other_kernels": others,
        "cpu_baseline": cpu_baseline,
        "gpu_eager_baseline": gpu_eager,
        "mlp_baseline": mlp_line,
    }
    print(json.dumps(line), flush=True)
    finish()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "mlp":
        import torch
        pkg = importlib.import_module(PKG_NAME)
        pkg.lib()
        torch.cuda.set_device(0)
        print(json.dumps(mlp_measure(pkg, torch, torch.device("cuda", 0), steps=max(args.steps, 3))), flush=True)
    elif args.folds > 0:
        run_folds(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
