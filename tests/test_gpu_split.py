"""GPU parity of the speculative chain splitting (abi_gw.inc): the growing-window search of one
long recipe line, cut into sub-chains that run side by side and stitched at exactly coinciding
changes, must return the records of the sequential search bit for bit - also when sub-chains
give up and are continued in further rounds."""

import os

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


class env(object):
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kw}
        os.environ.update({k: str(v) for k, v in self.kw.items()})

    def __exit__(self, *exc):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _same(x, y):
    (w1, f1), (w2, f2) = x, y
    assert np.array_equal(f1, f2)
    assert len(w1) == len(w2)
    for name in w1.dtype.names:
        if name != 'pad':
            assert np.array_equal(w1[name], w2[name]), name


# the last case detects NO change at all: no sub-chain ever meets another one, every one of them
# gives up and is continued round after round - the worst case of the scheme, still exact
CASES = [(_abi.BIC, 0.0, 100), (_abi.GLR, 1500.0, 100), (_abi.KL2, 4000.0, 100), (_abi.BIC, 0.0, 125),
         (_abi.KL2, 800.0, 125), (_abi.KL2, 3000.0, 125)]


@pytest.mark.parametrize('metric,thr,rate', CASES)
@pytest.mark.parametrize('limit', [3.0, 0.3])
def test_split_equals_sequential(ctx, metric, thr, rate, limit):
    rec = synth.make_recording(1200 + rate, 45000, 6, rate=rate)
    n = rec.frames.shape[0]
    # three recipe lines: a long one, a short one (never split) and a medium one; the third starts
    # on an odd frame
    seg_a, seg_b = [0, 30000, 30801], [30000, 30800, n]
    args = (float(rate), float(rate), 3.0 * rate, float(rate // 10), thr, 1.0, metric)
    with ctx.upload(rec.frames) as feat:
        with env(SPKDIAR_GW_NOSPLIT=1):
            plain = feat.gw_run(seg_a, seg_b, *args)
        with env(SPKDIAR_GW_MINLEN=4, SPKDIAR_GW_LIMIT=limit, SPKDIAR_GW_SPLIT_KL2=1):
            split = feat.gw_run(seg_a, seg_b, *args)
        one = feat.gw_run(seg_a, seg_b, *args, max_groups=1)
    assert len(plain[0]) > 100 and (plain[0]['positive'].sum() > 5 or thr == 3000.0)
    _same(plain, split)
    _same(plain, one)


def test_split_multi_equals_separate(ctx):
    rec = synth.make_recording(1300, 60000, 6)
    runs = [dict(rate=100.0, winsize=100.0, winstep=300.0, deltaws=10.0, threshold=t, lambdac=1.0, metric=m)
            for m, t in ((_abi.BIC, 0.0), (_abi.GLR, 1500.0), (_abi.KL2, 4000.0))]
    with ctx.upload(rec.frames) as feat:
        multi = feat.gw_run_multi([0], [60000], runs)
        for r, got in zip(runs, multi):
            with env(SPKDIAR_GW_NOSPLIT=1):
                want = feat.gw_run([0], [60000], r['rate'], r['winsize'], r['winstep'], r['deltaws'],
                                   r['threshold'], r['lambdac'], r['metric'])
            _same(want, got)


def test_split_declines_inexact_rates(ctx):
    """rate / 10 is not a multiple of 1/8 frame at 48 fps: positions are not translation-exact
    there, so the search must run unsplit - and still be right (host-driven loop as the check)."""
    rec = synth.make_recording(1400, 20000, 4, rate=48)
    args = (48.0, 48.0, 144.0, 4.0, 0.0, 1.0, _abi.BIC)
    with ctx.upload(rec.frames) as feat:
        with env(SPKDIAR_GW_NOSPLIT=1):
            plain = feat.gw_run([0], [20000], *args)
        with env(SPKDIAR_GW_MINLEN=4):
            auto = feat.gw_run([0], [20000], *args)
    _same(plain, auto)


def test_async_searches_with_clustering_queued_behind_the_first(ctx):
    """begin / wait / exec_on: the BIC search is collected first and its turns are clustered on the
    stream and SMs it ran on while the KL2 chain is still running; everything equals the
    synchronous calls."""
    rec = synth.make_recording(1500, 60000, 6)
    runs = [dict(rate=100.0, winsize=100.0, winstep=300.0, deltaws=10.0, threshold=t, lambdac=1.0, metric=m)
            for m, t in ((_abi.BIC, 0.0), (_abi.GLR, 1500.0), (_abi.KL2, 4000.0))]

    def turns(win):
        cuts = [0] + [int(r['start'] + r['maxi_fine']) for r in win if r['positive']] + [60000]
        return cuts[:-1], cuts[1:]
    with ctx.upload(rec.frames) as feat:
        want = feat.gw_run_multi([0], [60000], runs)
        sa, sb = turns(want[0][0])
        with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
            want_merges, want_stats = cl.run(0.0, 0, 1)
        with feat.gw_multi_begin([0], [60000], runs) as h:
            bic = h.wait(0)
            stream, sms = h.where(0)
            assert stream and 0 < sms < ctx.sm_count
            ctx.exec_on(stream, sms)
            try:
                with feat.cluster(*turns(bic[0]), _abi.BIC, 1.3) as cl:
                    merges, stats = cl.run(0.0, 0, 1)
            finally:
                ctx.exec_on()
            kl2 = h.wait(2)                      # out of order on purpose
            glr = h.wait(1)
            with pytest.raises(_abi.SpkdiarError):
                h.wait(1)                        # collected already
        for w, g in zip(want, (bic, glr, kl2)):
            _same(w, g)
        assert np.array_equal(merges, want_merges) and np.array_equal(stats, want_stats)
        # an object that is closed without being collected, and a second one right after it
        feat.gw_multi_begin([0], [60000], runs[:2]).close()
        with feat.gw_multi_begin([0], [60000], runs[:1]) as h:
            _same(want[0], h.wait(0))


@pytest.mark.parametrize('rate,thr', [(100, 4000.0), (100, 400.0), (125, 60.0), (125, 1.0), (100, 1e9)])
def test_kl2_fused_first_wave_equals_separate_fine_wave(ctx, rate, thr):
    """KL2: the first wave after a change also scores every frame a fine tune of window 0 could look
    at, so that a change found in window 0 needs no fine wave.  The records must be those of the
    search with separate fine waves (SPKDIAR_GW_NOFUSE), bit for bit - also at 125 fps (half-frame
    positions), with low thresholds (nearly every window 0 fires) and with none firing at all."""
    rec = synth.make_recording(1600 + rate, 30000, 5, rate=rate)
    n = rec.frames.shape[0]
    seg_a, seg_b = [0, 20001], [20000, n]
    args = (float(rate), float(rate), 3.0 * rate, float(rate // 10), thr, 1.0, _abi.KL2)
    with ctx.upload(rec.frames) as feat:
        fused = feat.gw_run(seg_a, seg_b, *args)
        with env(SPKDIAR_GW_NOFUSE=1):
            plain = feat.gw_run(seg_a, seg_b, *args)
        one_cta = feat.gw_run(seg_a * 80, seg_b * 80, *args)         # 160 chains: one CTA each, no speculation
    _same(plain, fused)
    if thr <= 400.0:
        assert plain[0]['positive'].sum() > 5
    k = len(fused[0])
    assert len(one_cta[0]) == 80 * k
    for name in ('start', 'end', 'maxi', 'maxd', 'maxi_fine', 'maxd_fine', 'positive', 'ncand', 'ninf', 'seq'):
        assert np.array_equal(one_cta[0][name][:k], fused[0][name]), name
