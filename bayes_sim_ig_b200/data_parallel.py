"""Data-parallel MDNN / MDRFF training across the GPUs of one node.

The reference is single-process; this module adds the one
exchange step data parallelism needs.  One process per GPU
(``torch.distributed``, backend nccl); every rank

  * owns a contiguous shard of the trajectories (``shard_rows``) and summarizes
    it locally -- summaries need no collective;
  * holds a full replica of the model (``enable`` broadcasts rank 0's flat
    parameter buffer once);
  * draws its own minibatches from its shard, and after the backward kernels
    sum-all-reduces the ONE flat fp32 gradient buffer over NVLink (NCCL,
    recorded inside the same CUDA graph as the step), scales by 1/world inside
    the Adam kernel and applies the identical update everywhere.

Semantics: G ranks x minibatch B == one process with minibatch G*B whose rows
are the concatenation of the ranks' rows, except that (1) the eps-noise scale
``1e-5 * mean(L_d)`` (mdnn.py:115) is the mean over the local minibatch, and
(2) reported losses are rank-local; both differences are O(1e-5) relative.
"""
import torch
import torch.distributed as dist


def shard_rows(n_total, rank, world):
    """Contiguous, balanced [start, stop) row range of ``rank`` (first n % world
    ranks get one extra row)."""
    base, extra = divmod(int(n_total), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def enable(model, group=None):
    """Turn on gradient all-reduce for ``model.run_training`` and make all
    replicas start from rank 0's parameters."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        model._dp_group = None
        model._dp_world = 1
        return model
    model._ensure_flat()
    dist.broadcast(model.flat_params, src=dist.get_global_rank(group, 0) if group else 0,
                   group=group)
    rff = getattr(model, 'rff', None)
    if rff is not None:
        dist.broadcast(rff.freqs, src=0, group=group)
    model._dp_group = group
    model._dp_world = dist.get_world_size(group)
    model._plans = {}
    return model


def world_of(model):
    return int(getattr(model, '_dp_world', 1) or 1)


def allreduce_gradients(model, flat_grads):
    """Sum the flat gradient buffer over the replicas (on the current stream, so
    it is captured into the step's CUDA graph)."""
    if world_of(model) > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=model._dp_group)
