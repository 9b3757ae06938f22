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
        f = ctx.upload(rec.frames)
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
