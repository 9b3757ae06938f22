"""Multi-process (N > 1) path of the corpus driver on CPU: two ``gloo`` ranks
shard a small corpus by file, diarize their shares and gather the summaries.
There is no GPU here, so the per-recording runner is the CPU oracle (tests may
use it as the checker); what is under test is the host logic of
``spkdiar.corpus`` - sharding, per-rank output files, the gather, and that the
result does not depend on the number of ranks."""

import io
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import spkdiar                                   # noqa: E402,F401
from spkdiar import corpus, synth                # noqa: E402


def _items(n=5):
    out = []
    for k in range(n):
        rec = synth.make_recording(500 + k, 2400 + 300 * k, 2 + k % 2, rate=100, turn_lo=4, turn_hi=8)
        out.append(('rec%d' % k, synth.one_line_recipe('/syn/rec%d.wav' % k, rec), rec.frames))
    return out


def _oracle_runner(name, lines, frames):
    """spk-diarization2.py:122-128 on the CPU oracle."""
    import warnings
    warnings.simplefilter('ignore')
    from oracle import change_detection as ocd, clustering as ocl
    recipe = ocd.parse_recipe(lines, lambda *a: None)
    cd = ocd.ChangeDetection(100, 'gw', 'BIC', 1.0, 3.0, 0.1, 0.0, 1.0)
    seg = io.StringIO()
    cd.detect_changes(recipe, seg, loader=lambda rl: (39, frames))
    cl = ocl.Clustering(100, 1, 'hi', 'BIC', 0.0, 0, 1.3)
    out = io.StringIO()
    cl.process_recipe(ocd.parse_recipe(seg.getvalue().splitlines(True), lambda *a: None), out,
                      loader=lambda rl: (39, frames))
    return seg.getvalue(), out.getvalue(), dict(turns=seg.getvalue().count('\n'), speakers=len(cl.speakers))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, outdir, result_q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)

    def gather(obj):
        parts = [None] * world
        dist.all_gather_object(parts, obj)
        return parts
    merged = corpus.run_corpus(_items(), rank, world, outdir=outdir, frame_rate=100,
                               runner=_oracle_runner, gather=gather)
    dist.barrier()
    result_q.put((rank, merged))
    dist.destroy_process_group()


def test_shard_is_a_partition():
    for n in (0, 1, 5, 17):
        for world in (1, 2, 4, 8):
            parts = [corpus.shard(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


@pytest.mark.timeout(600)
def test_two_gloo_ranks_equal_one_process(tmp_path):
    single = corpus.run_corpus(_items(), 0, 1, outdir=str(tmp_path / 'one'), frame_rate=100,
                               runner=_oracle_runner)
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path / 'two'), q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=540) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0] == got[1] == single                       # every rank holds the whole corpus' summaries
    for name in single:
        for ext in ('.spkc.recipe', '.recipe'):
            a = open(str(tmp_path / 'one' / (name + ext))).read()
            b = open(str(tmp_path / 'two' / (name + ext))).read()
            assert a == b and a.count('\n') >= 1


# ---- sharded clustering: exchange + global pick over two gloo ranks ---------------------------

def _sharded_cpu_loop(rank, world, exchange, frames, seg, threshold):
    """The host logic of the sharded merge loop (spkdiar_cluster_run_sharded) with the oracle's
    BIC as the distance: this rank keeps the pairs (r, c) with (r + c) % world == rank, proposes
    its best one, the ranks exchange 16-byte candidates and apply the global minimum."""
    import struct
    import warnings
    import numpy as np
    from oracle import distances as OD
    from spkdiar import sharded
    warnings.simplefilter('ignore')
    n = len(seg)
    members = [[s] for s in seg]
    alive = [True] * n

    def feats(k):
        return np.concatenate([frames[a:b] for a, b in members[k]])

    def d(i, j):
        return float(OD.bic_cl(feats(i), feats(j), 1.3))
    M = {}
    for i in range(n):
        for j in range(i + 1, n):
            if (i + j) % world == rank:
                M[(i, j)] = d(i, j)
    merges = []
    while True:
        best = (float('inf'), 2 ** 63 - 1)
        for (i, j), v in M.items():
            if alive[i] and alive[j]:
                cand = (v, i * n + j)
                if cand[0] < best[0] or (cand[0] == best[0] and cand[1] < best[1]):
                    best = cand
        got = exchange(struct.pack('<dq', best[0], best[1]))
        v, idx = sharded.pick_global(got)
        if idx == 2 ** 63 - 1 or not v <= threshold:
            break
        a, b = divmod(idx, n)
        merges.append((a, b, v))
        members[a] += members[b]
        alive[b] = False
        for k in range(n):
            if alive[k] and k != a and (a + k) % world == rank:
                M[(min(a, k), max(a, k))] = d(a, k)
    return merges


def _sharded_worker(rank, world, port, result_q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from spkdiar import sharded
    rec = synth.make_recording(321, 6000, 3, turn_lo=2, turn_hi=4)
    seg = [(t[0], t[1]) for t in rec.turns]
    merges = _sharded_cpu_loop(rank, world, sharded.dist_exchange(), rec.frames, seg, 0.0)
    dist.barrier()
    result_q.put((rank, merges))
    dist.destroy_process_group()


def test_pick_global_order():
    import struct
    from spkdiar import sharded
    nan = float('nan')
    pk = lambda *c: b''.join(struct.pack('<dq', v, i) for v, i in c)
    assert sharded.pick_global(pk((1.0, 5), (1.0, 3), (2.0, 1))) == (1.0, 3)          # ties: smaller flat index
    v, i = sharded.pick_global(pk((1.0, 5), (nan, 9), (nan, 7), (-5.0, 1)))           # NaN first, first NaN wins
    assert v != v and i == 7
    assert sharded.pick_global(pk((float('inf'), 2 ** 63 - 1), (-float('inf'), 4))) == (-float('inf'), 4)


@pytest.mark.timeout(600)
def test_sharded_loop_two_gloo_ranks_equal_one():
    from spkdiar import sharded
    rec = synth.make_recording(321, 6000, 3, turn_lo=2, turn_hi=4)
    seg = [(t[0], t[1]) for t in rec.turns]
    single = _sharded_cpu_loop(0, 1, lambda mine: mine, rec.frames, seg, 0.0)
    assert len(single) >= 5
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=540) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0] == got[1] == single
