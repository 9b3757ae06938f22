"""Sharded clustering (BASELINE config 5's path) on ONE GPU: the ranks are threads of
this process, each with its own context, features and cluster handle, exchanging their
candidates through a barrier.  The merge sequence, the distances and the statistics must
be identical to the resident single-GPU engine for every number of ranks."""

import threading

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi, synth, sharded

pytestmark = pytest.mark.gpu


def _single(rec, a, b, metric, threshold, max_spk):
    with _abi.Context(0) as ctx:
        f = ctx.upload_frames(rec.frames)           # as sharded.cluster_sharded uploads them
        with f.cluster(a, b, metric, 1.3) as cl:
            m, st = cl.run(threshold, max_spk, 1)
        f.close()
    return m, st


def _sharded(rec, a, b, metric, threshold, max_spk, nranks):
    ex = sharded.ThreadExchange(nranks)
    out = [None] * nranks
    err = []

    def work(rank):
        try:
            with _abi.Context(0) as ctx:
                out[rank] = sharded.cluster_sharded(ctx, rec.frames, a, b, metric, 1.3, threshold, max_spk,
                                                    rank, nranks, ex.for_rank(rank))
        except Exception as e:                       # pragma: no cover
            err.append(e)
            ex.barrier.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    return out


@pytest.mark.parametrize('metric', [_abi.BIC, _abi.GLR])
@pytest.mark.parametrize('nranks', [1, 2, 3, 8])
def test_sharded_equals_resident(metric, nranks):
    rec = synth.make_recording(77, 24000, 5, turn_lo=1, turn_hi=4)
    a = [t[0] for t in rec.turns]
    b = [t[1] for t in rec.turns]
    thr = 0.0 if metric == _abi.BIC else 1.0e9
    max_spk = 0 if metric == _abi.BIC else 4
    ref_m, ref_st = _single(rec, a, b, metric, thr, max_spk)
    assert len(ref_m) > 10
    res = _sharded(rec, a, b, metric, thr, max_spk, nranks)
    for m, st in res:
        assert m.tobytes() == ref_m.tobytes()
        assert np.array_equal(st, ref_st)


def test_device_decided_loop_equals_resident():
    """The NCCL variant's loop (argmin kernel -> gather -> deciding + rescoring kernel, merges queued in
    batches, stop flag on the device) with a single rank: no exchange, same merges."""
    rec = synth.make_recording(78, 30000, 6, turn_lo=1, turn_hi=4)
    a = [t[0] for t in rec.turns]
    b = [t[1] for t in rec.turns]
    ref_m, ref_st = _single(rec, a, b, _abi.BIC, 0.0, 0)
    with _abi.Context(0) as ctx:
        m, st = sharded.cluster_sharded(ctx, rec.frames, a, b, _abi.BIC, 1.3, 0.0, 0, 0, 1, device_loop=True)
    assert len(ref_m) > 20 and m.tobytes() == ref_m.tobytes()
    assert np.array_equal(st, ref_st)


def _p2p_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    rec = synth.make_recording(79, 40000, 6, turn_lo=1, turn_hi=4)
    a = [t[0] for t in rec.turns]
    b = [t[1] for t in rec.turns]
    with _abi.Context(rank) as ctx:
        mbx = sharded.Mailboxes(ctx)
        out = []
        for _ in range(2):                                   # two runs on the same mailboxes
            m, st = sharded.cluster_sharded(ctx, rec.frames, a, b, _abi.BIC, 1.3, 0.0, 0, rank, world, mailboxes=mbx)
            out.append((m.tobytes(), st.tobytes()))
        dist.barrier()
        mbx.close()
    q.put((rank, out))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_peer_memory_exchange_two_gpus():
    """One persistent kernel per GPU, candidates exchanged through peer-memory mailboxes: same
    merges and statistics as the resident single-GPU engine.  Needs two GPUs."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    rec = synth.make_recording(79, 40000, 6, turn_lo=1, turn_hi=4)
    a = [t[0] for t in rec.turns]
    b = [t[1] for t in rec.turns]
    ref_m, ref_st = _single(rec, a, b, _abi.BIC, 0.0, 0)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        for mb, sb in got[r]:
            assert mb == ref_m.tobytes() and sb == ref_st.tobytes()
