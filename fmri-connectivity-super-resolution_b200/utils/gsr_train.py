"""train / test -- drop-ins for the reference's utils/gsr_train.py:11-85, plus the batched
step the B200 path is built around.

``train`` and ``test`` keep the reference's per-subject semantics (one optimiser step per
subject, numpy inputs, print-only, return None).  ``GSRTrainStep`` / ``train_batched`` run B
subjects per step with frozen weights inside the step: the loss is the MEAN over the subjects of
the reference's per-subject loss, so the gradients equal the average of the reference's
per-subject gradients.  Importing this module has no side effects (the reference
loads its CSVs at import time, utils/gsr_train.py:88 -- deliberately not reproduced).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import functional as F_
from .preprocessing import pad_HR_adj, unpad


def criterion(x, y):
    """nn.L1Loss() of utils/gsr_train.py:9 on the sm_100a reduction kernels."""
    return F_.l1_loss(x, y)


def hr_target_basis(hr, lr_dim, padding=0, check_errors=True):
    """utils/gsr_train.py:33-36: eigenvectors of the ``lr_dim`` SMALLEST eigenvalues of the
    padded, unit-diagonal HR target.  Data-only: callers may cache it across epochs."""
    padded_hr = pad_HR_adj(hr, padding)
    _, U_hr, _ = F_.eigh(padded_hr, "U", check_errors=check_errors)
    return U_hr[..., :, :lr_dim].contiguous()


def _as_device_matrix(a, device):
    """numpy (any float dtype) or torch -> float32 on the device, as `.type(torch.FloatTensor).to(device)` does."""
    t = torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else torch.as_tensor(a)
    return t.to(device=device, dtype=torch.float32)


def train(model, optimizer, subjects_adj, subjects_labels, args):
    """Drop-in for utils/gsr_train.py:11-53: ``args.epochs`` passes over the subjects, ONE optimiser step per subject
    (the reference is strictly batch 1), the three-term loss of :39-41, and the reference's only observable output --
    one ``Epoch: i, Loss: ..., Error (MAE): ...`` line per epoch.  Returns None.  Each step is a ``GSRTrainStep`` on a
    one-subject batch (forward, HR-target eigendecomposition, loss, backward on the library's kernels); ``loss`` /
    ``error`` are read back per step as the reference's ``.item()`` calls do."""
    device = torch.device("cuda")
    model.to(device)
    step = GSRTrainStep(model, lmbda=args.lmbda, padding=args.padding, check_errors=True)
    for epoch in range(args.epochs):
        losses, errors = [], []
        for lr, hr in zip(subjects_adj, subjects_labels):
            model.train()
            optimizer.zero_grad()
            loss, error = step(_as_device_matrix(lr, device)[None], _as_device_matrix(hr, device)[None])
            optimizer.step()
            losses.append(loss.item())
            errors.append(error.item())
        print(f"Epoch: {epoch + 1}, Loss: {np.mean(losses)}, Error (MAE): {np.mean(errors)}")


def test(model, test_adj, test_labels, args):
    """Drop-in for utils/gsr_train.py:55-85: per-subject MAE of the prediction against the HR target with a unit diagonal
    (set IN PLACE on the caller's array, :70), all-zero samples skipped (:65-68), one ``MAE:`` line per evaluated
    subject and the mean at the end.  Returns None."""
    device = torch.device("cuda")
    model.to(device)
    test_error = []
    for lr, hr in zip(test_adj, test_labels):
        if not np.any(lr) or not np.any(hr):
            continue
        np.fill_diagonal(hr, 1)
        with torch.no_grad():           # the reference builds and discards an autograd graph here
            preds = unpad(model(_as_device_matrix(lr, device))[0], args.padding)
            error = criterion(preds, _as_device_matrix(hr, device)).item()
        print(f"MAE: {error}")
        test_error.append(error)
    print(f"Test error MAE: {np.mean(test_error)}")


# ----------------------------------------------------------------------------- batched step
class GSRTrainStep:
    """One forward + backward of the reference's loss (utils/gsr_train.py:30-45) over a batch
    of subjects resident on the device.

        step = GSRTrainStep(model, lmbda=16, padding=0)
        loss, error = step(lr_batch, hr_batch)        # [B,lr,lr], [B,hr,hr]; gradients in .grad

    ``loss`` / ``error`` are device scalars (means over the subjects); nothing synchronises
    unless ``check_errors`` asks for torch.linalg.eigh's raise-on-failure contract.
    ``U_hr_reduced`` may be passed when the caller caches the data-only HR target basis."""

    def __init__(self, model, lmbda=16, padding=0, check_errors=False, fuse_final=True):
        self.model, self.lmbda, self.padding, self.check_errors = model, lmbda, padding, check_errors
        self.fuse_final = fuse_final      # False: the prediction is materialised (model(lr)) and compared by l1_loss

    def forward_loss(self, lr, hr, U_hr_reduced=None, U_lr=None):
        model = self.model
        # check_errors=True raises inside the forward pass, before any product consumes a failed basis; the
        # caller's model keeps its own ``layer.check_eigh`` either way
        fused = self.fuse_final and self.padding == 0 and lr.dim() == 3 and hasattr(model, "forward_hidden")
        if fused:
            # :32-37 of the model and the L1 term against the HR target in one pass over hidden2 and hr
            hidden2, net_outs, start_gcn_outs, _ = model.forward_hidden(lr, U=U_lr, check_eigh=self.check_errors)
        else:
            out, net_outs, start_gcn_outs, _ = model(lr, U=U_lr, check_eigh=self.check_errors)
            out = unpad(out, self.padding)
        if U_hr_reduced is None:
            U_hr_reduced = hr_target_basis(hr, model.layer.lr_dim, self.padding, check_errors=False)
        err_b = F_.sym_diag_abs_l1_per_subject(hidden2, hr) if fused else F_.l1_loss_per_subject(out, hr)      # [B]
        w_b = F_.l1_loss_per_subject(U_hr_reduced, model.layer.weights)          # |W - U_hr| per subject
        reg = F_.l1_loss(net_outs, start_gcn_outs)
        error = err_b.mean()
        loss = self.lmbda * reg + w_b.mean() + error
        return loss, error

    def __call__(self, lr, hr, U_hr_reduced=None, U_lr=None):
        loss, error = self.forward_loss(lr, hr, U_hr_reduced, U_lr)
        loss.backward()
        return loss.detach(), error.detach()


class GraphedTrainStep:
    """The whole optimiser step -- forward, backward, (gradient all-reduce,) optimiser update -- captured once
    in a CUDA graph and replayed (SURVEY 8 f, row N3).  A step is ~180 kernel launches of 10-300 us; replaying
    them removes the Python / autograd / ctypes launch path, which otherwise bounds the step whenever the host is
    slower than the device (profiles/README.md).

        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)     # device-side step counter
        step = GraphedTrainStep(model, opt, batch=1024)
        loss, error = step(lr_batch, hr_batch, U_hr_reduced)      # device scalars, valid until the next call

    The inputs are copied into the graph's static buffers (device to device, or from pinned host memory);
    nothing synchronises.  ``U_hr_reduced=None`` captures the HR-target eigendecomposition inside the graph,
    as the reference's loop computes it every step.  Capture runs ``warmup`` real steps first (lazy
    allocations, cudaFuncSetAttribute); parameters and optimiser state are restored afterwards, so the first
    replay is the first step of the run.  ``grad_sync`` (e.g. ``SubjectParallel.sync_gradients``) is captured
    between backward and the optimiser update.

    ``slots=2`` captures the step once per set of input buffers (the graphs share one memory pool): a loader
    fills ``inputs(1)`` -- e.g. host-to-device copies on another stream -- while ``replay(0)`` runs, and no
    device-to-device staging copy is needed."""

    def __init__(self, model, optimizer, batch, lmbda=16, padding=0, hr_basis_in_graph=False, grad_sync=None,
                 warmup=3, slots=1):
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("GraphedTrainStep: build the optimiser with capturable=True (its step counter must "
                                 "live on the device)")
        layer = model.layer
        dev = layer.weights.device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep: move the model to the GPU first; there is no CPU path")
        self.model, self.optimizer = model, optimizer
        self.eager = GSRTrainStep(model, lmbda, padding)
        hr_dim = layer.hr_dim - 2 * padding
        self.slots = max(int(slots), 1)
        self._in = []
        for _ in range(self.slots):
            lr = torch.zeros(batch, layer.lr_dim, layer.lr_dim, device=dev)
            hr = torch.zeros(batch, hr_dim, hr_dim, device=dev)
            # unit-diagonal placeholders keep the warm-up eigendecompositions well defined
            lr.diagonal(dim1=1, dim2=2).fill_(1.0)
            hr.diagonal(dim1=1, dim2=2).fill_(1.0)
            U = None if hr_basis_in_graph else torch.zeros(batch, layer.hr_dim, layer.lr_dim, device=dev)
            self._in.append((lr, hr, U))
        self.lr, self.hr, self.U_hr = self._in[0]

        def one(slot=0):
            lr, hr, U = self._in[slot]
            optimizer.zero_grad(set_to_none=True)
            out = self.eager(lr, hr, U_hr_reduced=U)
            if grad_sync is not None:
                grad_sync()
            optimizer.step()
            return out

        params = [p for g in optimizer.param_groups for p in g["params"]]
        saved_params = [p.detach().clone() for p in params]
        saved_state = {id(p): {k: (v.clone() if torch.is_tensor(v) else v) for k, v in optimizer.state.get(p, {}).items()}
                       for p in params}
        from .. import launch_count
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 1)):
                one()
        torch.cuda.current_stream().wait_stream(side)
        self._graphs, self._out = [], []
        for slot in range(self.slots):
            optimizer.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            n0 = launch_count()
            with torch.cuda.graph(graph, pool=self._graphs[0].pool() if self._graphs else None):
                out = one(slot)
            self.launches_per_step = launch_count() - n0    # launches of this library inside one replay
            self._graphs.append(graph)
            self._out.append(out)
        self.graph = self._graphs[0]
        self.loss, self.error = self._out[0]
        with torch.no_grad():
            for p, q in zip(params, saved_params):
                p.copy_(q)
            for p in params:
                old = saved_state[id(p)]
                for k, v in optimizer.state.get(p, {}).items():
                    if torch.is_tensor(v):
                        v.copy_(old[k]) if k in old else v.zero_()

    def inputs(self, slot=0):
        """(lr, hr, U_hr_reduced or None): the input buffers graph ``slot`` reads; fill them before ``replay(slot)``."""
        return self._in[slot]

    def load(self, lr, hr, U_hr_reduced=None, non_blocking=True, slot=0):
        """Copy one batch into the input buffers of ``slot`` (on the current stream)."""
        lr_s, hr_s, U_s = self._in[slot]
        lr_s.copy_(lr, non_blocking=non_blocking)
        hr_s.copy_(hr, non_blocking=non_blocking)
        if U_s is not None:
            if U_hr_reduced is None:
                raise ValueError("GraphedTrainStep: captured with a cached HR basis; pass U_hr_reduced")
            U_s.copy_(U_hr_reduced, non_blocking=non_blocking)

    def replay(self, slot=0):
        """Returns (loss, error) device scalars of this slot's graph; valid until its next replay."""
        self._graphs[slot].replay()
        return self._out[slot]

    def __call__(self, lr, hr, U_hr_reduced=None, slot=0):
        self.load(lr, hr, U_hr_reduced, slot=slot)
        return self.replay(slot)


def train_batched(model, optimizer, subjects_adj, subjects_labels, args, batch_size=64,
                  cache_eigenbases=True, grad_sync=None, verbose=True):
    """Batched counterpart of ``train``: one optimiser step per ``batch_size`` subjects.
    The LR / HR eigenbases depend on the data only and are computed once (first epoch) when
    ``cache_eigenbases`` is set.  ``grad_sync`` (e.g. SubjectParallel.sync_gradients) is called
    between backward and the optimiser step.  Returns the list of epoch mean losses."""
    device = torch.device("cuda")
    model.to(device)
    lr_all = torch.as_tensor(np.asarray(subjects_adj), dtype=torch.float32).to(device)
    hr_all = torch.as_tensor(np.asarray(subjects_labels), dtype=torch.float32).to(device)
    step = GSRTrainStep(model, args.lmbda, args.padding)
    n = lr_all.shape[0]
    U_lr_cache, U_hr_cache = {}, {}
    history = []
    for epoch in range(args.epochs):
        losses, errors = [], []
        for s in range(0, n, batch_size):
            lr, hr = lr_all[s:s + batch_size], hr_all[s:s + batch_size]
            model.train()
            optimizer.zero_grad(set_to_none=True)
            U_lr = U_hr = None
            if cache_eigenbases:
                if s not in U_hr_cache:
                    U_hr_cache[s] = hr_target_basis(hr, model.layer.lr_dim, args.padding)
                    from .preprocessing import normalize_adj_torch
                    U_lr_cache[s] = F_.eigh(normalize_adj_torch(lr), "U")[1]
                U_lr, U_hr = U_lr_cache[s], U_hr_cache[s]
            loss, error = step(lr, hr, U_hr_reduced=U_hr, U_lr=U_lr)
            if grad_sync is not None:
                grad_sync()
            optimizer.step()
            losses.append(loss)
            errors.append(error)
        ml = float(torch.stack(losses).mean())
        me = float(torch.stack(errors).mean())
        history.append(ml)
        if verbose:
            print(f"Epoch: {epoch + 1}, Loss: {ml}, Error (MAE): {me}")
    return history
