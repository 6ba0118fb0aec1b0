"""Data-parallel plumbing for the head (one process per GPU, `torch.distributed`).

The path shards over the batch: every quantity is per-sample except the tanh loss, a sum
over the LOCAL batch -- the reference under DDP computes it per rank too (`main_dist.py:330`), so per-rank
losses with mean-reduced gradients is the parity target.  The only exchange step is the gradient all-reduce:
  * prototype kernels: ONE flat [P, C] fp32 buffer, reduced on a side stream right after the dW GEMM so it
    overlaps the dX GEMM and the backbone backward (`ops.GRAD_ALLREDUCE_GROUP`, set by `enable_overlapped_allreduce`);
  * classifier weights: a few KB, one flat bucket after backward (`flat_allreduce_mean_`).
When the model is wrapped in `DistributedDataParallel` (as `main_dist.py` does) none of this is needed: DDP's
reducer sees every per-node parameter as an input of the flat gather and all-reduces them in its first bucket.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def flat_allreduce_mean_(tensors: Iterable[torch.Tensor], group=None) -> None:
    """In-place mean all-reduce of many small tensors through one flat bucket (works on NCCL and gloo)."""
    ts: List[torch.Tensor] = [t for t in tensors if t is not None]
    if not ts or not dist.is_initialized():
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    flat = torch.cat([t.reshape(-1) for t in ts])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for t in ts:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


def enable_overlapped_allreduce(group=None, fresh_grads: bool = True) -> None:
    """Data parallelism handled by the head itself (instead of DistributedDataParallel): mean all-reduce of its gradients
    (prototype kernels, classifiers, presence logits) inside the backward, on a side stream over NCCL.
    fresh_grads=True: all producers write into ONE flat bucket, ONE all-reduce is issued right after the dW GEMM and is
    joined at the end of the backward pass, so it overlaps the dX GEMM and the backbone backward; `param.grad` is set
    (or accumulated into) by the head, not by autograd.  False: gradients travel through autograd and every all-reduce
    is joined before its tensor is returned (dW still overlaps the dX GEMM).
    Do not combine with DistributedDataParallel on the same parameters (they would be reduced twice)."""
    from . import ops
    ops.GRAD_ALLREDUCE_GROUP = group if group is not None else dist.group.WORLD
    ops.ASYNC_GRAD_ALLREDUCE = bool(fresh_grads)


def disable_overlapped_allreduce() -> None:
    from . import ops
    ops.GRAD_ALLREDUCE_GROUP = None


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n units for `rank` (units = images; remainder goes to the first ranks)."""
    q, r = divmod(n, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)
