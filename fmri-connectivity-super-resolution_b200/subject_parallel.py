"""Subject-data-parallel training over the GPUs of one box.

Subjects are independent given the weights, so rank g of G takes the subjects
[g*B/G, (g+1)*B/G); every rank runs the (subject independent) Graph U-Net redundantly; the 31
partial gradients are packed into ONE flat fp32 bucket (1,023,496 floats = 4.09 MB), reduced
with a single all-reduce (NCCL over NVLink / NVSwitch on the GPUs, gloo in the CPU tests of the
host logic) and unpacked scaled by 1/G, so that every rank applies the identical optimiser step.
Cross-validation folds are independent models: ``fold_groups`` splits the world into one
process group per fold with no traffic between folds.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_subjects: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of ``n_subjects`` for ``rank``."""
    base, rem = divmod(n_subjects, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def fold_assignment(world: int, n_folds: int) -> List[List[int]]:
    """Ranks of each fold group: folds get contiguous, near-equal GPU sets (8 GPUs, 3 folds ->
    [0,1,2], [3,4,5], [6,7]).  With fewer GPUs than folds every rank runs folds round-robin and
    the list holds one single-rank group per rank."""
    if world >= n_folds:
        out, lo = [], 0
        for f in range(n_folds):
            size = world // n_folds + (1 if f < world % n_folds else 0)
            out.append(list(range(lo, lo + size)))
            lo += size
        return out
    return [[r] for r in range(world)]


def fold_groups(n_folds: int):
    """Create one process group per fold (collective: every rank must call it).  Returns
    (my_fold_indices, my_group, ranks_of_my_group)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    assign = fold_assignment(world, n_folds)
    mine, my_ranks = None, None
    for ranks in assign:
        g = dist.new_group(ranks=ranks)
        if rank in ranks:
            mine, my_ranks = g, ranks
    if world >= n_folds:
        my_folds = [assign.index(my_ranks)]
    else:
        my_folds = [f for f in range(n_folds) if f % world == rank]
    return my_folds, mine, my_ranks


class SubjectParallel:
    """Gradient averaging for a model replicated over the ranks of ``group``."""

    def __init__(self, model: torch.nn.Module, group=None, broadcast_from: Optional[int] = 0):
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        self._bucket = None
        if broadcast_from is not None and self.world > 1:
            src = broadcast_from if group is None else dist.get_global_rank(group, broadcast_from)
            for p in model.state_dict().values():
                dist.broadcast(p, src=src, group=group)

    def _grads(self):
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        return [p.grad for p in self.params]

    def sync_gradients(self) -> None:
        """Average .grad over the ranks (call between backward and optimizer.step)."""
        if self.world == 1:
            return
        grads = self._grads()
        if grads[0].is_cuda:
            from . import kernels as K
            ptrs = tuple(g.data_ptr() for g in grads)
            # one bucket per set of gradient buffers, all kept: a captured CUDA graph (GraphedTrainStep) keeps
            # replaying into the bucket it was captured with while eager steps use another
            if self._bucket is None:
                self._bucket = {}
            bucket = self._bucket.get(ptrs)
            if bucket is None:
                if len(self._bucket) >= 4:
                    self._bucket.pop(next(iter(self._bucket)))
                bucket = self._bucket[ptrs] = K.FlatBucket(grads)
            flat = bucket.pack(1.0 / self.world)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            bucket.unpack(1.0)
        else:   # host-logic tests (gloo): same bucket layout through torch ops
            flat = torch.cat([g.reshape(-1) for g in grads]) / self.world
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
