"""Host logic of the batched / asynchronous paths without a GPU: the offsetting views over a
packed batch of recordings, the per-recording split of one growing-window launch, the
serialising proxy that lets a worker thread queue device work, and the corpus sharding with
device batches (fake device objects record what they are asked to do)."""

import threading
import time

import numpy as np

import spkdiar                              # noqa: F401
from spkdiar import _abi, corpus


class FakePack(_abi.FeaturePack):
    """FeaturePack without a library: records the calls that reach the packed handle."""

    def __init__(self, base, lengths):
        self.ctx = None
        self.h = None
        self.base, self.lengths = list(base), list(lengths)
        self.n = base[-1] + (lengths[-1] // 128 + 1) * 128
        self.calls = []

    def gw_run(self, seg_a, seg_b, *args, **kw):
        self.calls.append(('gw_run', list(map(int, seg_a)), list(map(int, seg_b))))
        # two records per chain: seq 0 negative, seq 1 positive
        n = len(seg_a)
        win = np.zeros(2 * n, dtype=_abi.GW_WINDOW_DTYPE)
        win['chain'] = np.repeat(np.arange(n), 2)
        win['seq'] = np.tile([0, 1], n)
        win['start'] = np.repeat(np.asarray(seg_a, dtype=float), 2)          # marks whose chain it was
        first = np.arange(0, 2 * n + 1, 2, dtype=np.int64)
        return win, first

    def score_windows(self, a, m, b, metric, lambdac=1.3, terms=False):
        self.calls.append(('score_windows', list(map(int, a)), list(map(int, m)), list(map(int, b))))
        return np.zeros(len(a))

    def score_sets(self, s1, s2, metric, lambdac=1.3, terms=False):
        self.calls.append(('score_sets', s1, s2))
        return np.zeros(len(s1))

    def stats_window(self, a, b):
        self.calls.append(('stats_window', a, b))
        return None


def test_views_offset_every_position():
    pack = FakePack([0, 6144, 12416], [6000, 6200, 100])
    v = pack.view(1)
    assert (v.n, v.off) == (6200, 6144)
    v.score_windows([0, 10], [5, 20], [9, 30], _abi.BIC)
    v.score_sets([[(1, 2), (3, 4)]], [[(5, 6)]], _abi.GLR)
    v.stats_window(7, 9)
    v.gw_run([0, 100], [50, 6200], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
    assert pack.calls == [
        ('score_windows', [6144, 6154], [6149, 6164], [6153, 6174]),
        ('score_sets', [[(6145, 6146), (6147, 6148)]], [[(6149, 6150)]]),
        ('stats_window', 6151, 6153),
        ('gw_run', [6144, 6244], [6194, 12344]),
    ]
    v.close()                                   # a view owns nothing


def test_one_launch_is_split_back_per_recording():
    pack = FakePack([0, 6144, 12416], [6000, 6200, 100])
    chains = [[(0, 3000), (3000, 6000)], [], [(10, 100)]]
    out = pack.gw_run_batch(chains, 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
    assert pack.calls == [('gw_run', [0, 3000, 12426], [3000, 6000, 12516])]
    assert [len(w) for w, _ in out] == [4, 0, 2]
    assert out[0][0]['chain'].tolist() == [0, 0, 1, 1] and out[0][1].tolist() == [0, 2, 4]
    assert out[1][1].tolist() == [0]
    assert out[2][0]['chain'].tolist() == [0, 0] and out[2][1].tolist() == [0, 2]
    assert out[2][0]['start'].tolist() == [12426.0, 12426.0]


def test_serialised_proxy_never_overlaps_calls():
    class Lib(object):
        def __init__(self):
            self.inside = 0
            self.worst = 0

        def spkdiar_x(self, v):
            self.inside += 1
            self.worst = max(self.worst, self.inside)
            time.sleep(0.002)
            self.inside -= 1
            return v + 1
    lib = Lib()
    prox = _abi._Serialised(lib)
    got = []
    ts = [threading.Thread(target=lambda k=k: got.append(prox.spkdiar_x(k))) for k in range(16)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert sorted(got) == list(range(1, 17)) and lib.worst == 1


def test_shard_and_batches_cover_the_corpus_once():
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            mine = corpus.shard(37, rank, world)
            parts = [mine[b0:b0 + 5] for b0 in range(0, len(mine), 5)]
            assert all(0 < len(p) <= 5 for p in parts)
            seen += [k for p in parts for k in p]
        assert sorted(seen) == list(range(37))


def test_items_from_recipes_reads_lazily(tmp_path):
    from spkdiar import synth
    fea = tmp_path / 'fea'
    fea.mkdir()
    paths = []
    for k in range(3):
        rec = synth.make_recording(40 + k, 500 + 10 * k, 2)
        lines = synth.one_line_recipe('/media/r%d.wav' % k, rec)
        rp, _ = synth.write_case(str(tmp_path), 'r%d' % k, rec, lines)
        paths.append(rp)
    items = corpus.items_from_recipes(paths, str(fea))
    assert [it[0] for it in items] == ['r0', 'r1', 'r2']
    assert callable(items[1][2]) and items[1][2]().shape == (510, 39)
    (tmp_path / 'bad.recipe').write_text('audio=/a.wav lna=a_1 start-time=0.0 end-time=1.0\n'
                                         'audio=/b.wav lna=b_1 start-time=0.0 end-time=1.0\n')
    import pytest
    with pytest.raises(ValueError):
        corpus.items_from_recipes([str(tmp_path / 'bad.recipe')], str(fea))
    # two recipes with the same basename would silently overwrite each other's output files
    sub = tmp_path / 'other'
    sub.mkdir()
    (sub / 'r0.recipe').write_text(open(paths[0]).read())
    with pytest.raises(ValueError, match='unique'):
        corpus.items_from_recipes([paths[0], str(sub / 'r0.recipe')], str(fea))
