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
