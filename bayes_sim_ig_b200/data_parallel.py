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
import os

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
    world = dist.get_world_size(group)
    if model.flat_params.numel() % (4 * world) != 0:
        # equal 16-byte aligned shards for the sharded exchange of large models (train_engine)
        model._flatten_parameters(model.flat_params.device, pad_multiple=4 * world)
    src = dist.get_global_rank(group, 0) if group else 0
    dist.broadcast(model.flat_params, src=src, group=group)
    rff = getattr(model, 'rff', None)
    if rff is not None:
        dist.broadcast(rff.freqs, src=src, group=group)
        rff._coeff_key = None       # in-place collectives may not bump the tensor version
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


# --------------------------------------------------------------------------- P2P exchange
P2P_MAX_PARAMS = 1 << 20      # one-shot all-reduce reads (world-1) x the buffer: small models only


class P2PComm(object):
    """Peer-mapped gradient buffers for the fused all-reduce + Adam kernel
    (csrc/p2p.cu): two gradient buffers (update parity), one flag array and one control
    block per rank, allocated with cudaMalloc and exchanged as CUDA IPC handles."""

    def __init__(self, n_floats, group=None):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self.n_floats = int(n_floats)
        lib = _lib.load()
        sizes = [4 * self.n_floats, 4 * self.n_floats, 256, 256]   # grads[0], grads[1], flags, ctrl
        self.local, handles = [], []
        for nbytes in sizes:
            ptr = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(64)
            rc = lib.bsig_p2p_alloc(ctypes.byref(ptr), nbytes, buf)
            if rc != 0:
                raise _lib.BsigError('bsig_p2p_alloc failed: ' + _lib.last_error())
            self.local.append(ptr.value)
            handles.append(buf.raw)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=group)
        # ptrs[kind][rank]
        self.ptrs = [[None] * self.world for _ in sizes]
        self._opened = []
        for r in range(self.world):
            for kind in range(len(sizes)):
                if r == self.rank:
                    self.ptrs[kind][r] = self.local[kind]
                else:
                    ptr = ctypes.c_void_p()
                    rc = lib.bsig_p2p_open(gathered[r][kind], ctypes.byref(ptr))
                    if rc != 0:
                        raise _lib.BsigError('bsig_p2p_open failed: ' + _lib.last_error())
                    self.ptrs[kind][r] = ptr.value
                    self._opened.append(ptr.value)
        arr = ctypes.c_void_p * self.world
        self.grad_arrays = [arr(*self.ptrs[0]), arr(*self.ptrs[1])]
        self.flag_array = arr(*self.ptrs[2])
        self.ctrl = self.local[3]
        dist.barrier(group=group)

    def last_launch_stamps(self):
        """Nanosecond timestamps of block 0 in the most recent exchange kernel:
        (start, flags published, all peers arrived, reduction + Adam done)."""
        import ctypes
        buf = (ctypes.c_uint64 * 4)()
        self._lib.call('bsig_p2p_read', self.ctrl + 32, ctypes.addressof(buf), 32)
        return [int(v) for v in buf]

    def peer_timeout(self):
        """0 if every peer arrived in time, else 1 + index of the first peer the exchange
        kernel gave up waiting for (sticky; see the watchdog in csrc/p2p.cu)."""
        import ctypes
        word = ctypes.c_uint32(0)
        self._lib.call('bsig_p2p_read', self.ctrl + 8, ctypes.addressof(word), 4)
        return int(word.value)

    def local_grads(self, parity):
        return self.local[parity & 1]

    def adam_allreduce(self, model, exp_avg, exp_avg_sq, parity, step, stream):
        self._lib.call('bsig_adam_allreduce_step', model.flat_params.data_ptr(),
                       self.grad_arrays[parity & 1], self.flag_array, self.ctrl, self.rank,
                       self.world, exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                       model.flat_params.numel(), step, float(model.lr), 0.9, 0.999, 1e-8, stream)


def p2p_comm_for(model):
    """The model's P2P communicator (created on first use), or None when the exchange
    should go through NCCL (single rank, > 8 ranks, or a model too large for one-shot)."""
    world = world_of(model)
    if world <= 1 or world > 8 or model.flat_params.numel() > P2P_MAX_PARAMS:
        return None
    if os.environ.get('BSIG_DP_EXCHANGE', 'p2p') == 'nccl':
        return None
    comm = getattr(model, '_p2p_comm', None)
    if comm is None or comm.n_floats != model.flat_params.numel():
        comm = P2PComm(model.flat_params.numel(), getattr(model, '_dp_group', None))
        model._p2p_comm = comm
    return comm
