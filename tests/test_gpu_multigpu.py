"""The two paths that shard, on REAL GPUs (one process per GPU, torch.distributed over NCCL): needs at least
two devices, so it is skipped on a one-GPU box and meant for `gpurun --gpus 2|4|8 -- python -m pytest
tests/test_gpu_multigpu.py -m gpu` (and for the scaling lease).

* sharded clustering of one long recording (BASELINE config 5; spk-clustering.py:201-237) through all three
  exchanges - peer-memory mailboxes inside the persistent kernel (the product), NCCL on the library's
  stream, and the host callback over torch.distributed - merge sequence, distances and statistics
  byte-identical to the single-GPU resident engine, twice in a row on the same mailboxes;
* a corpus sharded by file (config 4; spk-diarization2.py:122-128 per file): every rank's recipes equal the
  one-process run."""

import hashlib
import os
import socket

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi, synth, sharded, corpus

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _long_recording():
    rec = synth.make_recording(79, 60000, 6, turn_lo=1, turn_hi=4)
    return rec, [t[0] for t in rec.turns], [t[1] for t in rec.turns]


def _cluster_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    rec, a, b = _long_recording()
    out = {}
    with _abi.Context(rank) as ctx:
        mbx = sharded.Mailboxes(ctx)
        with ctx.upload(rec.frames) as feat:
            runs = []
            for _ in range(3):                               # several runs on the same mailboxes, no barrier between
                m, st = sharded.cluster_sharded(ctx, feat, a, b, _abi.BIC, 1.3, 0.0, 0, rank, world, mailboxes=mbx)
                runs.append((m.tobytes(), st.tobytes()))
            out['p2p'] = runs
            dist.barrier()
            m, st = sharded.cluster_sharded(ctx, feat, a, b, _abi.BIC, 1.3, 0.0, 0, rank, world,
                                            nccl_id=sharded.broadcast_nccl_id())
            out['nccl'] = [(m.tobytes(), st.tobytes())]
            m, st = sharded.cluster_sharded(ctx, feat, a, b, _abi.BIC, 1.3, 0.0, 0, rank, world,
                                            exchange=sharded.dist_exchange())
            out['host'] = [(m.tobytes(), st.tobytes())]
        dist.barrier()
        mbx.close()
    q.put((rank, out))
    dist.destroy_process_group()


def _spawn(target, world, extra=()):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + tuple(extra)) for r in range(world)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return got


@pytest.mark.timeout(900)
@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_clustering_on_real_gpus_equals_resident_engine(world):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    rec, a, b = _long_recording()
    with _abi.Context(0) as ctx, ctx.upload(rec.frames) as f, f.cluster(a, b, _abi.BIC, 1.3) as cl:
        ref_m, ref_st = cl.run(0.0, 0, 1)
    assert len(ref_m) > 100
    got = _spawn(_cluster_worker, world)
    for r in range(world):
        for kind in ('p2p', 'nccl', 'host'):
            for mb, sb in got[r][kind]:
                assert mb == ref_m.tobytes(), (r, kind)
                assert sb == ref_st.tobytes(), (r, kind)


def _corpus_items(n=12):
    items = []
    for k in range(n):
        rec = synth.make_recording(5000 + k, 9000 + 500 * (k % 4), 2 + k % 4, turn_lo=3, turn_hi=8)
        items.append(('r%02d' % k, synth.one_line_recipe('/syn/r%02d.wav' % k, rec), rec.frames))
    return items


def _digest(outdir):
    h = {}
    for name in sorted(os.listdir(outdir)):
        h[name] = hashlib.sha256(open(os.path.join(outdir, name), 'rb').read()).hexdigest()
    return h


def _corpus_worker(rank, world, port, q, outdir):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))

    def gather(obj):
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts
    merged = corpus.run_corpus(_corpus_items(), rank, world, device=rank, outdir=outdir, frame_rate=100,
                               gather=gather, batch=4, overlap=True)
    dist.barrier()
    q.put((rank, merged))
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('world', [2, 8])
def test_corpus_sharded_by_file_on_real_gpus_equals_one_process(world, tmp_path):
    if _ngpu() < world:
        pytest.skip('needs %d GPUs' % world)
    one = str(tmp_path / 'one')
    single = corpus.run_corpus(_corpus_items(), 0, 1, device=0, outdir=one, frame_rate=100, batch=4, overlap=True)
    many = str(tmp_path / 'many')
    got = _spawn(_corpus_worker, world, (many,))
    for r in range(world):
        assert got[r] == single
    assert _digest(many) == _digest(one) and len(_digest(one)) == 2 * len(single)
