"""Synthetic test cases shared by the golden generator and the tests.

A case is regenerated from its seed (``spkdiar.synth``); the golden files keep
a SHA-256 of the frames so that a drift of the generator is noticed."""

import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import spkdiar                      # noqa: E402,F401
from spkdiar import synth           # noqa: E402

D2_GW = ['-m', 'gw', '-d', 'BIC', '-w', '1.0', '-st', '3.0', '-dws', '0.1', '-l', '1.0']
GW = ['-m', 'gw', '-w', '1.0', '-st', '3.0', '-dws', '0.1']

# name -> (script, variant, wavs, flags)
#   wavs: list of (wav name, synth kwargs, recipe kind)
CASES = {
    'gw_bic_f100': ('cd', 0, [('g1', dict(seed=101, n_frames=6000, n_speakers=3), 'one')],
                    ['-f', '100'] + D2_GW),
    'gw_bic_f125': ('cd', 0, [('g2', dict(seed=102, n_frames=7500, n_speakers=3, rate=125), 'one')],
                    ['-f', '125'] + D2_GW),
    'gw_glr': ('cd', 0, [('g3', dict(seed=103, n_frames=5000, n_speakers=3), 'one')],
               ['-f', '100'] + GW + ['-d', 'GLR', '-t', '1000']),
    'gw_kl2': ('cd', 0, [('g4', dict(seed=104, n_frames=4000, n_speakers=3), 'one')],
               ['-f', '100'] + GW + ['-d', 'KL2', '-t', '400']),
    'sw_glr': ('cd', 0, [('g5', dict(seed=105, n_frames=12000, n_speakers=4), 'one')],
               ['-f', '100', '-m', 'sw', '-d', 'GLR', '-t', '2000']),
    'sw_kl2_tt': ('cd', 0, [('g6', dict(seed=106, n_frames=6000, n_speakers=3), 'one')],
                  ['-f', '100', '-m', 'sw', '-d', 'KL2', '-t', '25', '-w', '3.0', '-tt']),
    'sw_bic_strict': ('cd', 0, [('g7', dict(seed=107, n_frames=3000, n_speakers=2), 'one')],
                      ['-f', '100', '-m', 'sw', '-d', 'BIC']),
    'merge_bic': ('cd', 0, [('g8', dict(seed=108, n_frames=8000, n_speakers=3, turn_lo=4, turn_hi=9), 'halves')],
                  ['-f', '100', '-m', 'm', '-d', 'BIC', '-t', '0', '-l', '1.0']),
    'merge_glr_dlr': ('cd', 0, [('g9', dict(seed=109, n_frames=8000, n_speakers=3, turn_lo=4, turn_hi=9), 'halves'),
                                ('g10', dict(seed=119, n_frames=5000, n_speakers=2, turn_lo=4, turn_hi=9), 'halves')],
                      ['-f', '100', '-m', 'm', '-d', 'GLR', '-t', '1500', '-dlr']),
    'gw_bic_multi': ('cd', 0, [('h1', dict(seed=110, n_frames=4000, n_speakers=2), 'vad'),
                               ('h2', dict(seed=111, n_frames=3500, n_speakers=3), 'vad')],
                     ['-f', '100'] + D2_GW + ['-tt']),
    'cl1_hi_bic': ('cl', 1, [('c1', dict(seed=201, n_frames=12000, n_speakers=4, turn_lo=2, turn_hi=6), 'turns')],
                   ['-f', '100', '-m', 'hi', '-l', '1.3']),
    'cl2_hi_bic': ('cl', 2, [('c1', dict(seed=201, n_frames=12000, n_speakers=4, turn_lo=2, turn_hi=6), 'turns')],
                   ['-f', '100', '-m', 'hi', '-l', '1.3']),
    # on this input variant 2's stale matrix entries cause extra merges
    'cl1_hi_q5': ('cl', 1, [('c2', dict(seed=301, n_frames=9000, n_speakers=5, turn_lo=1, turn_hi=4), 'turns')],
                  ['-f', '100', '-m', 'hi', '-l', '2.0']),
    'cl2_hi_q5': ('cl', 2, [('c2', dict(seed=301, n_frames=9000, n_speakers=5, turn_lo=1, turn_hi=4), 'turns')],
                  ['-f', '100', '-m', 'hi', '-l', '2.0']),
    'cl1_hi_glr_ms': ('cl', 1, [('c3', dict(seed=203, n_frames=9000, n_speakers=3, turn_lo=2, turn_hi=6), 'turns')],
                      ['-f', '100', '-m', 'hi', '-d', 'GLR', '-t', '500', '-ms', '2']),
    'cl2_hi_kl2': ('cl', 2, [('c4', dict(seed=204, n_frames=7000, n_speakers=3, turn_lo=2, turn_hi=6), 'turns')],
                   ['-f', '100', '-m', 'hi', '-d', 'KL2', '-t', '20']),
    'cl1_in_bic': ('cl', 1, [('c5', dict(seed=205, n_frames=9000, n_speakers=3, turn_lo=2, turn_hi=6), 'turns')],
                   ['-f', '100', '-m', 'in', '-l', '1.3']),
    'cl2_in_glr_tt': ('cl', 2, [('c6', dict(seed=206, n_frames=6000, n_speakers=3, turn_lo=2, turn_hi=6), 'turns')],
                      ['-f', '100', '-m', 'in', '-d', 'GLR', '-t', '600', '-tt']),
}


def _vad_recipe(audio, rec, letter):
    """A VAD-like recipe: a few speech turns with gaps, plus one junk line."""
    total = rec.frames.shape[0] / float(rec.rate)
    cuts = [0.0, round(total * 0.31, 2), round(total * 0.36, 2), round(total * 0.74, 2),
            round(total * 0.78, 2), total]
    lines = []
    for k in range(3):
        lines.append('audio=%s lna=%s_%d start-time=%r end-time=%r\n'
                     % (audio, letter, k + 1, cuts[2 * k], cuts[2 * k + 1]))
    lines.insert(1, 'this line has no fields\n')
    return lines


def _halves_recipe(audio, rec, letter):
    """Every true turn cut in two lines: consecutive same-speaker segments for
    the merge mode to join."""
    lines = []
    k = 0
    for a, b, _ in rec.turns:
        mid = (a + b) // 2
        for lo, hi in ((a, mid), (mid, b)):
            k += 1
            lines.append('audio=%s lna=%s_%d start-time=%r end-time=%r speaker=spk_turn\n'
                         % (audio, letter, k, lo / float(rec.rate), hi / float(rec.rate)))
    return lines


def materialise(name, workdir):
    """Write the case's feature files and recipe under ``workdir``;
    -> (recipe_path, feapath, sha256 of all frames, list of Recording)."""
    script, variant, wavs, flags = CASES[name]
    lines = []
    sha = hashlib.sha256()
    recs = []
    feadir = os.path.join(workdir, 'fea')
    os.makedirs(feadir, exist_ok=True)
    for k, (wav, kw, kind) in enumerate(wavs):
        rec = synth.make_recording(**kw)
        recs.append(rec)
        sha.update(rec.frames.tobytes())
        from spkdiar.feacat import write_features
        write_features(os.path.join(feadir, wav + '.fea'), rec.frames)
        audio = '/syn/%s.wav' % wav
        letter = chr(ord('a') + k)
        if kind == 'one':
            lines += synth.one_line_recipe(audio, rec, letter + '_1')
        elif kind == 'turns':
            lines += synth.turn_recipe(audio, rec, letter)
        elif kind == 'halves':
            lines += _halves_recipe(audio, rec, letter)
        else:
            lines += _vad_recipe(audio, rec, letter)
    rpath = os.path.join(workdir, name + '.recipe')
    with open(rpath, 'w') as f:
        f.writelines(lines)
    return rpath, feadir, sha.hexdigest(), recs
