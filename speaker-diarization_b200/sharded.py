"""Sharded agglomerative clustering of ONE very long recording (BASELINE config 5).

The pair matrix of ``spk-clustering.py``'s ``spk_cluster_hi`` (CL1:178-260) is dealt
out over the ranks of a ``torch.distributed`` job - pair (r, c) lives on rank
``(r + c) % world`` - the cluster statistics are replicated, and per merge the ranks
exchange their best candidate (16 bytes each: fp64 distance, int64 flat index) and pick
the global minimum with ``ndarray.argmin``'s order.  The merge sequence is identical to
the single-GPU run for any number of ranks (tests/test_gpu_sharded.py,
tests/test_distributed.py).

Exchange functions (``mine: bytes[16] -> bytes[16 * world]``):
  * ``dist_exchange(group)``    all_gather over torch.distributed (nccl: over NVLink;
                                gloo: the CPU tests)
  * ``ThreadExchange(n)``       ranks as threads of one process (tests, one GPU)
"""

import threading

import numpy as np


def dist_exchange(group=None, device=None):
    """An all-gather of 16-byte candidates over ``torch.distributed``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = device if device is not None else (torch.device('cuda', torch.cuda.current_device())
                                             if backend == 'nccl' else torch.device('cpu'))
    mine_t = torch.empty(16, dtype=torch.uint8, device=dev)
    all_t = torch.empty(16 * world, dtype=torch.uint8, device=dev)
    stage = torch.empty(16, dtype=torch.uint8).pin_memory() if dev.type == 'cuda' else None

    def exchange(mine):
        src = torch.frombuffer(bytearray(mine), dtype=torch.uint8)
        if stage is not None:
            stage.copy_(src)
            mine_t.copy_(stage, non_blocking=True)
        else:
            mine_t.copy_(src)
        dist.all_gather_into_tensor(all_t, mine_t, group=group)
        return all_t.cpu().numpy().tobytes()
    return exchange


class ThreadExchange(object):
    """Ranks are threads of one process: a barrier-based all-gather."""

    def __init__(self, nranks):
        self.n = nranks
        self.slots = [b''] * nranks
        self.barrier = threading.Barrier(nranks)

    def for_rank(self, rank):
        def exchange(mine):
            self.slots[rank] = mine
            self.barrier.wait()
            got = b''.join(self.slots)
            self.barrier.wait()
            return got
        return exchange


def pick_global(candidates):
    """The global minimum of the ranks' (value, flat index) candidates in
    ``ndarray.argmin`` order: NaN first, then value, then index (host twin of
    ``cl_before`` in csrc/cluster.cuh; the C loop does the same)."""
    arr = np.frombuffer(candidates, dtype=np.dtype([('v', '<f8'), ('i', '<i8')]))
    best = None
    for v, i in arr:
        if best is None:
            best = (v, i)
            continue
        bv, bi = best
        an, bn = v != v, bv != bv
        if an or bn:
            before = an and (not bn or i < bi)
        else:
            before = v < bv or (v == bv and i < bi)
        if before:
            best = (v, i)
    return float(best[0]), int(best[1])


def broadcast_nccl_id(group=None):
    """A fresh NCCL unique id made on rank 0 and sent to all ranks of a torch.distributed job.
    An id makes ONE communicator: call this once per ``cluster_sharded`` run."""
    import torch.distributed as dist
    from . import _abi
    box = [_abi.nccl_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


class Mailboxes(object):
    """The peer-memory mailboxes of a torch.distributed job (one process per GPU): every rank
    creates its own, the CUDA IPC handles are all-gathered, every rank maps the others'.  Use one
    object for many runs (``next_base`` hands out the sequence-number bases)."""

    def __init__(self, ctx, group=None):
        import ctypes as C
        import torch.distributed as dist
        self.ctx = ctx
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        local = C.c_void_p()
        handle = C.create_string_buffer(64)
        ctx._check(ctx.lib.spkdiar_mailbox_create(ctx.h, self.world, C.byref(local), handle))
        self.local = local.value
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local)
            else:
                p = C.c_void_p()
                ctx._check(ctx.lib.spkdiar_mailbox_open(ctx.h, C.create_string_buffer(h, 64), C.byref(p)))
                self.ptrs.append(p.value)
        self._base = 0
        dist.barrier(group)

    def next_base(self, nseg):
        base = self._base
        self._base += int(nseg) + 4
        return base

    def close(self):
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                self.ctx.lib.spkdiar_mailbox_close(self.ctx.h, p)
        self.ctx.lib.spkdiar_mailbox_free(self.ctx.h, self.local)
        self.ptrs = []


def cluster_sharded(ctx, frames_or_feat, seg_a, seg_b, metric, lambdac, threshold, max_spk,
                    rank, nranks, exchange=None, nccl_id=None, device_loop=False, mailboxes=None):
    """Run this rank's share; -> (merges, stats) identical on every rank.  With ``nccl_id``
    (see ``broadcast_nccl_id``) the library exchanges the candidates itself over NCCL, on its
    stream (``device_loop=True`` selects that loop for a single rank, where nothing is exchanged);
    with ``mailboxes`` (a ``Mailboxes``) the exchange happens inside one persistent kernel per rank
    through peer memory; else ``exchange`` is called once per merge."""
    from . import _abi
    own = not isinstance(frames_or_feat, _abi.Features)
    # clustering only: frames without window statistics, cluster records straight from the frames
    feat = ctx.upload_frames(frames_or_feat) if own else frames_or_feat
    try:
        with feat.cluster(seg_a, seg_b, metric, lambdac) as cl:
            if mailboxes is not None:
                return cl.run_sharded_p2p(threshold, max_spk, rank, nranks, mailboxes.ptrs,
                                          mailboxes.next_base(len(seg_a)))
            if nccl_id is not None or device_loop:
                return cl.run_sharded_nccl(threshold, max_spk, rank, nranks, nccl_id)
            return cl.run_sharded(threshold, max_spk, rank, nranks, exchange)
    finally:
        if own:
            feat.close()
