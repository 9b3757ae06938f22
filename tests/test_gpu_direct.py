"""Frames-only feature handles (``spkdiar_features_upload_frames``): cluster records accumulated
straight from the frames (K5, what get_spk_features + np.cov read, spk-clustering.py:46-52, 91-96)
against the records taken as differences of the window statistics, against the oracle, and the
window statistics built on first use."""

import io

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    c = _abi.Context(0)
    yield c
    c.close()


def _oracle_merges(rec, sa, sb, variant, metric='BIC', lam=1.3):
    from oracle import clustering as ocl
    trace = []
    oc = ocl.Clustering(100, variant, 'hi', metric, 0.0, 0, lam, trace=trace)
    recipe = [('/x.wav', 'a_%d' % (k + 1), a / 100.0, b / 100.0) for k, (a, b) in enumerate(zip(sa, sb))]
    oc.process_recipe(recipe, io.StringIO(), loader=lambda rl: (39, rec.frames))
    return trace


@pytest.mark.parametrize('variant', [1, 2])
def test_direct_records_give_the_merge_sequence_of_the_oracle(ctx, variant):
    rec = synth.make_recording(77, 24000, 4, turn_lo=2, turn_hi=6)
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    want = _oracle_merges(rec, sa, sb, variant)
    with ctx.upload_frames(rec.frames) as feat, feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
        got, stats = cl.run(0.0, 0, variant)
    assert [(int(m['a']), int(m['b'])) for m in got] == [(m[0], m[1]) for m in want]
    for g, w in zip(got, want):
        if np.isfinite(w[2]):
            assert abs(g['d'] - w[2]) <= 1e-9 * max(abs(w[2]), 1.0) * 50, (g, w)


def test_direct_and_prefix_records_agree(ctx):
    """Same merges, distances to rounding (the two sums associate differently), statistics alike; a range
    longer than one task (4,096 frames) goes through partial records."""
    rec = synth.make_recording(78, 60000, 5, turn_lo=3, turn_hi=70)       # turns of up to 70 s = 7,000 frames
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    assert max(b - a for a, b in zip(sa, sb)) > 4096
    with ctx.upload(rec.frames) as f1, f1.cluster(sa, sb, _abi.BIC, 1.3) as c1:
        m1, s1 = c1.run(0.0, 0, 1)
    with ctx.upload_frames(rec.frames) as f2, f2.cluster(sa, sb, _abi.BIC, 1.3) as c2:
        m2, s2 = c2.run(0.0, 0, 1)
        # a second run on the same handle starts from the initial records again
        m3, _ = c2.run(0.0, 0, 1)
    assert m1['a'].tolist() == m2['a'].tolist() and m1['b'].tolist() == m2['b'].tolist()
    assert np.allclose(m1['d'], m2['d'], rtol=1e-11, atol=1e-7)
    assert np.allclose(s1, s2, rtol=1e-11, atol=1e-7)
    assert m2.tobytes() == m3.tobytes()


def test_window_statistics_are_built_on_first_use(ctx):
    rec = synth.make_recording(79, 12000, 3, turn_lo=3, turn_hi=8)
    n = rec.frames.shape[0]
    with ctx.upload(rec.frames) as eager, ctx.upload_frames(rec.frames) as lazy:
        w1, _ = eager.gw_run([0], [n], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
        w2, _ = lazy.gw_run([0], [n], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
        assert w1.tobytes() == w2.tobytes()
        a, m, b = [0, 500], [300, 900], [700, 1500]
        assert eager.score_windows(a, m, b, _abi.GLR).tobytes() == lazy.score_windows(a, m, b, _abi.GLR).tobytes()
        # from here on the handle has window statistics: its cluster records are their differences, bit for bit
        sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
        with eager.cluster(sa, sb, _abi.BIC, 1.3) as c1, lazy.cluster(sa, sb, _abi.BIC, 1.3) as c2:
            assert c1.run(0.0, 0, 1)[0].tobytes() == c2.run(0.0, 0, 1)[0].tobytes()


def test_frames_only_sharded_run_equals_resident(ctx):
    """The sharded engine on frames-only handles (what bench.py's config5 uses), two thread-emulated ranks."""
    import threading
    from spkdiar import sharded
    rec = synth.make_recording(80, 30000, 4, turn_lo=1, turn_hi=3)
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    with ctx.upload_frames(rec.frames) as feat, feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
        want, _ = cl.run(0.0, 0, 1)
    ex = sharded.ThreadExchange(2)
    out, err = [None, None], []

    def work(rank):
        try:
            with _abi.Context(0) as c2, c2.upload_frames(rec.frames) as f2:
                out[rank] = sharded.cluster_sharded(c2, f2, sa, sb, _abi.BIC, 1.3, 0.0, 0, rank, 2, ex.for_rank(rank))
        except Exception as e:                       # pragma: no cover
            err.append(e)
            ex.barrier.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not err, err
    for m, _ in out:
        assert m.tobytes() == want.tobytes()


def test_cluster_of_fewer_frames_than_dimensions_stays_alone(ctx):
    """A turn of 30 frames in 39 dimensions has a singular covariance: the reference's np.log(det(S)) is -inf
    (the determinant underflows) or rounding noise; the ruling here is -inf, i.e. a pair distance of +inf that the
    agglomeration ignores - the other clusters merge as if the short turn were not there (a NaN instead would
    stop the loop at once: ndarray.argmin returns the first NaN, spk-clustering.py:203-208)."""
    rec = synth.make_recording(91, 20000, 3, turn_lo=3, turn_hi=6)
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    short = len(sa) // 2
    sa2 = sa[:short] + [sa[short]] + sa[short + 1:]
    sb2 = sb[:short] + [sa[short] + 30] + sb[short + 1:]
    for upload in (ctx.upload, ctx.upload_frames):
        with upload(rec.frames) as feat:
            with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
                full, _ = cl.run(0.0, 0, 1)
            with feat.cluster(sa2, sb2, _abi.BIC, 1.3) as cl:
                got, _ = cl.run(0.0, 0, 1)
                M, alive = cl.matrix()
        assert len(full) > 5
        assert len(got) >= len(full) - 2 and np.all(np.isfinite(got['d']))
        # nobody merged with the short turn; every distance to it is +inf
        members = [[k] for k in range(len(sa2))]
        for m in got:
            members[m['a']].extend(members[m['b']])
            members.pop(m['b'])
        assert [short] in members
        row = M[short][alive.astype(bool)]
        assert np.all(np.isposinf(np.delete(row, np.flatnonzero(np.flatnonzero(alive) == short))))
