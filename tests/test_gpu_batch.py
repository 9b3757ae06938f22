"""GPU parity of the batched corpus path (BASELINE config 4): a packed batch of
recordings must give, recording by recording, exactly what the one-at-a-time
path gives - statistics bit for bit, growing-window records, merge sequences,
recipes byte for byte - and the one-at-a-time path is itself pinned to the
oracle / the reference's golden recipes by the other GPU tests."""

import io

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi, synth, corpus
from oracle import change_detection as ocd, clustering as ocl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    c = _abi.Context(0)
    yield c
    c.close()


# lengths chosen to hit the layout's edge cases: a multiple of 128, one frame more, one less,
# shorter than a block, and an empty recording
LENGTHS = [6000, 6400, 6401, 6399, 100, 0, 5000, 128]


def _recordings():
    out = []
    for k, n in enumerate(LENGTHS):
        if n == 0:
            out.append(np.zeros((0, 39), dtype=np.float32))
        else:
            out.append(synth.make_recording(700 + k, n, 3, turn_lo=3, turn_hi=8).frames)
    return out


def test_packed_statistics_are_bit_identical(ctx):
    recs = _recordings()
    pack = ctx.upload_batch(recs)
    try:
        assert [b % 128 for b in pack.base] == [0] * len(recs)
        rng = np.random.default_rng(5)
        for r, x in enumerate(recs):
            n = x.shape[0]
            v = pack.view(r)
            assert v.n == n
            with ctx.upload(x) as alone:
                wins = [(0, n)] + [tuple(sorted(rng.integers(0, n + 1, 2).tolist())) for _ in range(6)]
                for a, b in wins:
                    s1, m1, sh1 = alone.stats_window(a, b)
                    s2, m2, _ = v.stats_window(a, b)
                    assert np.array_equal(s1, s2) and np.array_equal(m1, m2), (r, a, b)
    finally:
        pack.close()


@pytest.mark.parametrize('metric,thr', [(_abi.BIC, 0.0), (_abi.GLR, 1000.0), (_abi.KL2, 400.0)])
def test_batched_growing_window_equals_single(ctx, metric, thr):
    recs = [x for x in _recordings() if x.shape[0] >= 100]
    pack = ctx.upload_batch(recs)
    try:
        # two chains for the first recording (two recipe lines), one for the others
        chains = [[(0, 3000), (3000, 6000)]] + [[(0, x.shape[0])] for x in recs[1:]]
        got = pack.gw_run_batch(chains, 100.0, 100.0, 300.0, 10.0, thr, 1.0, metric)
        for r, x in enumerate(recs):
            with ctx.upload(x) as alone:
                win, first = alone.gw_run([c[0] for c in chains[r]], [c[1] for c in chains[r]],
                                          100.0, 100.0, 300.0, 10.0, thr, 1.0, metric)
            bw, bf = got[r]
            assert np.array_equal(first, bf), r
            assert len(win) == len(bw)
            for name in win.dtype.names:
                if name == 'pad':
                    continue
                assert np.array_equal(win[name], bw[name]), (r, name)
    finally:
        pack.close()


@pytest.mark.parametrize('variant', [1, 2])
@pytest.mark.parametrize('metric', [_abi.BIC, _abi.GLR])
def test_batched_clustering_equals_resident_engine(ctx, variant, metric):
    recs, problems = [], []
    for k, n in enumerate([9000, 20000, 6000, 300, 12000]):
        rec = synth.make_recording(800 + k, n, 4, turn_lo=2, turn_hi=6)
        recs.append(rec.frames)
        problems.append([(t[0], t[1]) for t in rec.turns])
    problems[3] = [(0, 300)]                       # a single segment: nothing to merge
    pack = ctx.upload_batch(recs)
    try:
        thr = 0.0 if metric == _abi.BIC else 500.0
        got = pack.cluster_batch(problems, metric, 1.3, thr, 0, variant)
        for r, x in enumerate(recs):
            with ctx.upload(x) as alone:
                with alone.cluster([p[0] for p in problems[r]], [p[1] for p in problems[r]], metric, 1.3) as cl:
                    merges, stats = cl.run(thr, 0, variant)
            bm, bs = got[r]
            assert len(bm) == len(merges), r
            assert np.array_equal(bm['a'], merges['a']) and np.array_equal(bm['b'], merges['b']), r
            assert np.array_equal(bm['d'], merges['d']), r
            assert np.array_equal(np.asarray(bs), np.asarray(stats), equal_nan=True), (r, bs, stats)
        # the max-speakers stopping rule (spk-clustering.py:207)
        got = pack.cluster_batch(problems[:2], metric, 1.3, -1e300, 3, variant)
        for r in range(2):
            with ctx.upload(recs[r]) as alone:
                with alone.cluster([p[0] for p in problems[r]], [p[1] for p in problems[r]], metric, 1.3) as cl:
                    merges, stats = cl.run(-1e300, 3, variant)
            assert np.array_equal(got[r][0], merges), r
            assert len(problems[r]) - len(merges) == 3
    finally:
        pack.close()


def test_diarize_batch_equals_one_at_a_time_and_oracle(ctx):
    items = []
    for k in range(5):
        rec = synth.make_recording(900 + k, 5000 + 700 * k, 2 + k % 3, turn_lo=3, turn_hi=9)
        lines = synth.one_line_recipe('/syn/b%d.wav' % k, rec) if k != 2 else \
            ['audio=/syn/b2.wav lna=a_1 start-time=0.0 end-time=30.0\n',
             'audio=/syn/b2.wav lna=a_2 start-time=31.5 end-time=%s\n' % (rec.frames.shape[0] / 100.0)]
        items.append((lines, rec.frames))
    batched = corpus.diarize_batch(ctx, items, 100)
    for k, (lines, frames) in enumerate(items):
        with ctx.upload(frames) as feat:
            seg, clu, summary = corpus.diarize_recording(ctx, lines, lambda l: feat, 100)
        assert batched[k][0] == seg, k
        assert batched[k][1] == clu, k
        assert batched[k][2] == summary, k
    # and against the CPU oracle (recipes byte for byte) for one of them
    lines, frames = items[1]
    from spkdiar import recipe as recipe_mod
    cd = ocd.ChangeDetection(100, 'gw', 'BIC', 1.0, 3.0, 0.1, 0.0, 1.0)
    seg = io.StringIO()
    cd.detect_changes(recipe_mod.parse(lines), seg, loader=lambda l: (39, frames))
    assert seg.getvalue() == batched[1][0]
    oc = ocl.Clustering(100, 1, 'hi', 'BIC', 0.0, 0, 1.3)
    out = io.StringIO()
    oc.process_recipe(recipe_mod.parse(seg.getvalue().splitlines(True)), out, loader=lambda l: (39, frames))
    assert out.getvalue() == batched[1][1]


def test_large_problems_of_a_batch_take_the_resident_engine(ctx, monkeypatch):
    """A recording with more turns than the one-CTA engine should take is clustered by the resident engine
    inside the same batch; the recipes do not change (the two engines are bit-identical)."""
    items = []
    for k in range(4):
        rec = synth.make_recording(930 + k, 6000 + 2000 * k, 3, turn_lo=3, turn_hi=6)
        items.append((synth.one_line_recipe('/syn/r%d.wav' % k, rec), rec.frames))
    want = corpus.diarize_batch(ctx, items, 100)
    sizes = sorted(r[2]['turns'] for r in want)
    monkeypatch.setattr(corpus, 'BATCH_MAX_SEGMENTS', sizes[1])        # the two largest leave the batch
    l0 = ctx.launches
    got = corpus.diarize_batch(ctx, items, 100)
    assert got == want
    monkeypatch.setattr(corpus, 'BATCH_MAX_SEGMENTS', 0)               # all of them
    assert corpus.diarize_batch(ctx, items, 100) == want
    assert ctx.launches > l0


def test_overlapped_batches_equal_batches(ctx):
    """The pipelined driver (device stages on a worker thread, host replay on this one) returns
    exactly what the synchronous one returns, batch by batch."""
    batches = []
    for b in range(4):
        part = []
        for k in range(3 + b % 2):
            rec = synth.make_recording(950 + 10 * b + k, 4000 + 500 * k, 2 + k % 3, turn_lo=3, turn_hi=9)
            part.append((synth.one_line_recipe('/syn/o%d_%d.wav' % (b, k), rec), rec.frames))
        batches.append(part)
    want = [corpus.diarize_batch(ctx, part, 100) for part in batches]
    got = list(corpus.diarize_batches(ctx, batches, 100))
    assert got == want
    assert list(corpus.diarize_batches(ctx, [], 100)) == []
    assert list(corpus.diarize_batches(ctx, batches[:1], 100)) == want[:1]


def test_run_corpus_batched_equals_one_at_a_time(tmp_path):
    """The corpus driver: device batches, with and without the overlapped pipeline, write the
    files the one-recording-at-a-time driver writes."""
    items = []
    for k in range(7):
        rec = synth.make_recording(980 + k, 4000 + 300 * k, 2 + k % 3, turn_lo=3, turn_hi=9)
        items.append(('r%d' % k, synth.one_line_recipe('/syn/r%d.wav' % k, rec), rec.frames))
    outs = {}
    for name, kw in (('single', {}), ('batch', dict(batch=3)), ('overlap', dict(batch=3, overlap=True))):
        d = tmp_path / name
        outs[name] = corpus.run_corpus(items, outdir=str(d), frame_rate=100, device=0, **kw)
    assert outs['single'] == outs['batch'] == outs['overlap'] and len(outs['single']) == 7
    for k in range(7):
        for ext in ('.recipe', '.spkc.recipe'):
            want = (tmp_path / 'single' / ('r%d%s' % (k, ext))).read_text()
            assert (tmp_path / 'batch' / ('r%d%s' % (k, ext))).read_text() == want
            assert (tmp_path / 'overlap' / ('r%d%s' % (k, ext))).read_text() == want


def test_argument_errors_of_the_batch_and_async_entry_points(ctx):
    rec = synth.make_recording(990, 3000, 2)
    with pytest.raises(_abi.SpkdiarError) as e:
        ctx.upload_batch([np.zeros((10, 40), dtype=np.float32)])              # more than 39 dimensions
    assert e.value.code == -5
    with pytest.raises(_abi.SpkdiarError):
        ctx.upload_batch([])                                                  # an empty batch
    pack = ctx.upload_batch([rec.frames, rec.frames[:500]])
    with pytest.raises(_abi.SpkdiarError) as e:
        pack.cluster_batch([[(0, 100)], []], _abi.BIC)                        # a problem without segments
    assert e.value.code == -2
    with pytest.raises(_abi.SpkdiarError) as e:
        pack.cluster_batch([[(0, 100), (100, 200)]], _abi.KL2)                # KL2 is not in the engines
    assert e.value.code == -5
    with pytest.raises(_abi.SpkdiarError) as e:
        pack.cluster_batch([[(0, 10 ** 9)]], _abi.BIC)                        # outside the packed matrix
    assert e.value.code == -2
    runs = [dict(rate=100.0, winsize=100.0, winstep=300.0, deltaws=10.0, threshold=0.0, lambdac=1.0, metric=_abi.BIC)]
    h = pack.gw_multi_begin([0], [3000], runs)
    try:
        with pytest.raises(_abi.SpkdiarError):
            pack.gw_multi_begin([0], [3000], runs)                            # one asynchronous object per context
        with pytest.raises(_abi.SpkdiarError):
            h.wait(1)                                                         # no such search
        win, first = h.wait(0)
        assert len(win) > 0
    finally:
        h.close()
    pack.gw_multi_begin([0], [3000], runs).close()                            # and it can be opened again
    pack.close()


def test_corpus_cli_equals_the_two_scripts(tmp_path, capsys):
    """scripts/spk-diarization-corpus.py writes, per recording, the files the two drop-in scripts
    write when spk-diarization2.py calls them one after the other (D2:122-128)."""
    from conftest import run_product
    from cases import D2_GW
    paths = []
    for k in range(3):
        rec = synth.make_recording(60 + k, 5000 + 400 * k, 2 + k, turn_lo=3, turn_hi=8)
        lines = synth.one_line_recipe('/media/c%d.wav' % k, rec)
        rp, feadir = synth.write_case(str(tmp_path), 'c%d' % k, rec, lines)
        paths.append(rp)
    outdir = tmp_path / 'out'
    done = corpus.main(paths + [feadir, '-o', str(outdir), '-f', '100', '--batch', '2'])
    assert sorted(done) == ['c0', 'c1', 'c2'] and 'c1: ' in capsys.readouterr().out
    for k, rp in enumerate(paths):
        seg, clu = str(tmp_path / ('s%d.recipe' % k)), str(tmp_path / ('k%d.recipe' % k))
        run_product('cd', 0, [rp, feadir, '-o', seg, '-f', '100'] + D2_GW)
        run_product('cl', 1, [seg, feadir + '/', '-o', clu, '-f', '100', '-m', 'hi', '-l', '1.3'])
        assert (outdir / ('c%d.spkc.recipe' % k)).read_text() == open(seg).read()
        assert (outdir / ('c%d.recipe' % k)).read_text() == open(clu).read()
