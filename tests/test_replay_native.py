"""The native record -> recipe replay (``spkdiar_replay_*``, csrc/spkdiar_replay.cu) against the
Python replay (``Detector`` / ``Clusterer``, which the golden tests pin to the reference's own
scripts): same chains, same segmentation recipe, same initial clusters, same clustered recipe,
byte for byte, on random recipes, window records and merge sequences.  Host only - no GPU."""

import io
import random

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from spkdiar import _abi, corpus, recipe as recipe_mod
from spkdiar import change_detection as pcd, clustering as pcl


class FakeFeat(object):
    def __init__(self, n):
        self.n = n

    def close(self):
        pass


def python_replay(lines, rate, nframes, win, first, merges_for):
    """The two stages as ``corpus._BatchJob`` runs them through the Python classes."""
    feat = FakeFeat(nframes)
    parsed = recipe_mod.parse(lines)
    det = pcd.Detector(rate, ctx=object(), **corpus.D2_CHANGE)
    groups = det.gw_chains(parsed, lambda l: feat)
    chains = [c for _, ch in groups for c in ch]
    if groups:
        det.prefetch(feat, chains, (win, first))
    seg = io.StringIO()
    det.writer.record = []
    det.detect_changes(parsed, seg, loader=lambda l: feat)
    seg_lines = seg.getvalue().splitlines(True)
    seg_parsed = recipe_mod.lines_from_records(det.writer.record, seg_lines)
    cl = pcl.Clusterer(rate, variant=1, ctx=object(), **corpus.D2_CLUSTER)
    problem = cl.initial_segments(seg_parsed, nframes)
    merges = merges_for(len(problem))
    clu = io.StringIO()
    if problem:
        cl.prefetch(feat, problem, (merges, np.zeros(4)))
        cl.process_recipe(seg_parsed, clu, loader=lambda l: feat)
    return chains, seg.getvalue(), problem, clu.getvalue(), len(cl.speakers), det.windows_visited


def random_merges(rng, n, nm=None):
    nm = rng.randint(0, max(n - 1, 0)) if nm is None else nm
    out = np.zeros(nm, dtype=_abi.MERGE_DTYPE)
    alive = n
    for m in range(nm):
        a = rng.randrange(0, alive - 1)
        b = rng.randrange(a + 1, alive)
        out[m] = (a, b, -rng.random() * 1000)
        alive -= 1
    return out


def random_windows(rng, chains, rate):
    """Window records shaped like the search's: a few windows per chain, some positive, positions on the
    half-frame grid of a 125 fps search or the integer grid of a 100 fps one."""
    recs, first = [], [0]
    for a, b in chains:
        n = b - a
        start = 0.0
        for w in range(rng.randint(0, 12)):
            r = np.zeros((), dtype=_abi.GW_WINDOW_DTYPE)
            r['start'] = start
            r['end'] = start + 2 * rate
            r['ncand'] = 5
            r['maxd'] = rng.random() * 100 - 50
            if rng.random() < 0.5 and start + rate < max(n, 1):
                step = rng.randint(int(rate / 2), int(rate * 3)) + (0.5 if rate == 125 and rng.random() < 0.5 else 0.0)
                r['positive'] = 1
                r['maxi'] = step
                r['maxi_fine'] = step
                r['maxd_fine'] = float(r['maxd']) + 1
                start += step
            recs.append(r)
        first.append(len(recs))
    win = np.array(recs, dtype=_abi.GW_WINDOW_DTYPE) if recs else np.zeros(0, dtype=_abi.GW_WINDOW_DTYPE)
    return win, np.array(first, dtype=np.int64)


def native_replay(lines, rate, nframes, win_for, merges_for):
    rp = _abi.Replay(rate, lines)
    try:
        sa, sb = rp.chains(nframes)
        chains = list(zip(sa.tolist(), sb.tolist()))
        win, first = win_for(chains)
        nturns = rp.segment(win, first)
        seg = rp.text(0)
        ta, tb = rp.turns(nframes, nturns)
        problem = list(zip(ta.tolist(), tb.tolist()))
        if nturns:
            nspk = rp.cluster(merges_for(nturns))
            clu = rp.text(1)
        else:
            nspk, clu = 0, ''
        return chains, seg, problem, clu, nspk, int(rp.info()[5]), (win, first)
    finally:
        rp.close()


LNAS = ['a_1', 'a_2', 'b_7', 'spk_x_3', 'x', 'noscore', '_', 'a', 'ab', 'a_', 'zz_1', 'a_1']


def random_recipe(rng, audio='/data/rec.wav'):
    lines = []
    t = rng.random() * 3
    for k in range(rng.randint(1, 9)):
        kind = rng.random()
        dur = rng.choice([0.4, 3.0, 12.5, 61.25, 200.0]) * (0.5 + rng.random())
        s, e = t, t + dur
        t = e + rng.choice([0.0, 0.25, 1.5])
        lna = rng.choice(LNAS)
        fmt = rng.choice(['%.3f', '%.2f', '%.1f', '%r', '%.6f'])
        fs, fe = fmt % s, fmt % e
        if kind < 0.08:
            lines.append('# comment line without fields\n')
            continue
        if kind < 0.14:
            lines.append('audio=%s lna=%s start-time=%s\n' % (audio, lna, fs))         # no end-time: skipped
            continue
        if kind < 0.2:
            fs = '%d' % int(s * 100 + 10)                                             # "123": digits only still match
        extra = rng.choice(['', ' speaker=spk_turn', ' alignment=/x/y.seg', '\t'])
        order = rng.random()
        if order < 0.8:
            lines.append('audio=%s lna=%s start-time=%s end-time=%s%s\n' % (audio, lna, fs, fe, extra))
        else:
            lines.append('end-time=%s start-time=%s lna=%s audio=%s%s\n' % (fe, fs, lna, audio, extra))
    return lines


@pytest.mark.parametrize('rate', [100, 125])
def test_native_replay_equals_python_replay_on_random_recordings(rate):
    rng = random.Random(1234 + rate)
    done = 0
    for case in range(300):
        lines = random_recipe(rng)
        nframes = rng.choice([3000, 60000, 75000, 10 ** 6])
        seed = rng.random()
        state = {}

        def win_for(chains):
            r2 = random.Random(seed)
            state['w'] = random_windows(r2, chains, float(rate))
            return state['w']

        def merges_for(n):
            return random_merges(random.Random(seed + 1), n)
        got = native_replay(lines, rate, nframes, win_for, merges_for)
        win, first = got[6]
        want = python_replay(lines, rate, nframes, win, first, merges_for)
        assert got[0] == [tuple(c) for c in want[0]], (case, lines)
        assert got[1] == want[1], (case, lines)
        assert got[2] == want[2], (case, lines)
        assert got[3] == want[3], (case, lines)
        assert got[4] == want[4] and got[5] == want[5], (case, lines)
        done += 1
    assert done == 300


def test_native_parser_follows_the_reference_patterns():
    """The four independent searches of spk-change-detection.py:11-28, including the unescaped dot of
    ``\\d+.\\d+`` and keys inside other fields."""
    cases = [
        'audio=/a/b.wav lna=x_1 start-time=1.5 end-time=20.25\n',
        'audio=/a/b.wav lna=x_1 start-time=15 end-time=2000\n',                  # digits only: \d+ backs off
        'audio=/a/b.wav lna=x_1 start-time=5 end-time=7.0\n',                    # "5" alone does not match: skipped
        'audio=/a/lna=inner.wav lna=x_1 start-time=1.0 end-time=2.0\n',          # lna= inside the audio path wins
        'audio= lna=x_1 audio=/late.wav start-time=1.0 end-time=2.0\n',          # "audio= " has no \S: later key
        'lna=q_9 audio=/a.wav end-time=9.5 start-time=0.125 junk\n',
        'audio=/a.wav lna=x_1 start-time=1.5.7 end-time=3.25.1\n',               # first digits.digits only
        'audio=/a.wav lna=x_1 start-time=12x3 end-time=4.0\n',                   # float("12x3") raises: unsupported
        'audio=/a.wav lna=x_1 start-time=x end-time=4.0 start-time=2.5\n',       # leftmost MATCHING occurrence
        'nothing here\n',
        'audio=/a.wav lna=x_1 start-time=0.0 end-time=0.0\n',
    ]
    for text in cases:
        try:
            want = recipe_mod.parse([text])
            raised = None
        except ValueError as e:
            want, raised = None, e
        try:
            rp = _abi.Replay(100.0, [text])
        except _abi.ReplayUnsupported:
            assert raised is not None, text
            continue
        assert raised is None, text
        try:
            assert rp.nlines == len(want), text
            if want:
                sa, sb = rp.chains(10 ** 9)
                assert (sa[0], sb[0]) == (int(want[0].start * 100.0), max(int(want[0].start * 100.0), int(want[0].end * 100.0)))
                if 'lna=' in want[0].audio:
                    # the segmentation recipe would not parse back field by field: left to the Python replay
                    with pytest.raises(_abi.ReplayUnsupported):
                        rp.segment(np.zeros(0, dtype=_abi.GW_WINDOW_DTYPE), np.zeros(2, dtype=np.int64))
                    continue
                rp.segment(np.zeros(0, dtype=_abi.GW_WINDOW_DTYPE), np.zeros(2, dtype=np.int64))
                det_line = 'audio=%s lna=' % want[0].audio
                assert rp.text(0).startswith(det_line), (text, rp.text(0))
        finally:
            rp.close()


def test_native_replay_declines_what_it_does_not_reproduce():
    with pytest.raises(_abi.ReplayUnsupported):
        _abi.Replay(100.0, ['audio=/a/ä.wav lna=x_1 start-time=1.0 end-time=2.0\n'])
    with pytest.raises(_abi.ReplayUnsupported):
        _abi.Replay(100.0, ['audio=/a.wav lna=x_1 start-time=1.0 end-time=2.0\nlna=y_1\n', 'audio=/b.wav\n'])
    rp = _abi.Replay(100.0, ['audio=/a.wav lna=x_1 start-time=1.0 end-time=2.0\n',
                             'audio=/b.wav lna=x_2 start-time=1.0 end-time=2.0\n'])
    try:
        assert not rp.single_wav
        with pytest.raises(_abi.ReplayUnsupported):
            rp.segment(np.zeros(0, dtype=_abi.GW_WINDOW_DTYPE), np.zeros(3, dtype=np.int64))
    finally:
        rp.close()
    # a time that str(float) prints with an exponent is not digits.digits: the Python replay takes over
    rp = _abi.Replay(100.0, ['audio=/a.wav lna=x_1 start-time=0.0 end-time=99999999999999999.0\n'])
    try:
        with pytest.raises(_abi.ReplayUnsupported):
            rp.segment(np.zeros(0, dtype=_abi.GW_WINDOW_DTYPE), np.zeros(2, dtype=np.int64))
    finally:
        rp.close()
