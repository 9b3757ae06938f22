"""Seeded synthetic recordings for the BASELINE.json configurations.

There is no network, hence no corpus: every test and benchmark runs on
synthetic feacat-format MFCC-like features (d = 39) generated here, following
SURVEY.md section 8(d): K Gaussian "speakers" with ``A_k = I + 0.15*G_k``,
``mu_k = 0.5*g_k`` (G, g i.i.d. N(0, 1)); a recording is a sequence of speaker
turns; frames are drawn in fp64 and cast to fp32.  ``numpy.random.default_rng``
(PCG64) makes the files reproducible from the seed alone.
"""

import os
import os.path as op

import numpy as np

from .feacat import write_features

DIM = 39            # fconfig.cfg:78-83 (12 MFCC + power, delta, delta-delta)


class Recording(object):
    """frames (N, DIM) float32, the true turns [(first_frame, last_frame+1,
    speaker)], the frame rate."""

    def __init__(self, frames, turns, rate):
        self.frames = frames
        self.turns = turns
        self.rate = rate

    @property
    def seconds(self):
        return self.frames.shape[0] / float(self.rate)


def make_recording(seed, n_frames, n_speakers, rate=100, turn_lo=3, turn_hi=19,
                   dim=DIM, unit=None):
    """A recording of ``n_frames`` frames.  Turn lengths are whole multiples of
    ``unit`` frames (default: one second), uniform in [turn_lo, turn_hi]
    units; consecutive turns never share a speaker; the last turn is cut to
    fit."""
    rng = np.random.default_rng(seed)
    unit = int(rate) if unit is None else int(unit)
    mix = np.eye(dim)[None] + 0.15 * rng.standard_normal((n_speakers, dim, dim))
    mean = 0.5 * rng.standard_normal((n_speakers, dim))
    frames = np.empty((n_frames, dim), dtype=np.float32)
    turns = []
    pos = 0
    prev = -1
    while pos < n_frames:
        spk = int(rng.integers(n_speakers))
        if n_speakers > 1 and spk == prev:
            spk = (spk + 1 + int(rng.integers(n_speakers - 1))) % n_speakers
        length = int(rng.integers(turn_lo, turn_hi + 1)) * unit
        length = min(length, n_frames - pos)
        z = rng.standard_normal((length, dim))
        frames[pos:pos + length] = (mean[spk] + z @ mix[spk].T).astype(np.float32)
        turns.append((pos, pos + length, spk))
        pos += length
        prev = spk
    return Recording(frames, turns, rate)


def one_line_recipe(audio, rec, lna='a_1'):
    """A single speech turn covering the whole recording (configs 1, 2, 4)."""
    return ['audio=%s lna=%s start-time=0.0 end-time=%s\n'
            % (audio, lna, repr(float(rec.seconds)))]


def turn_recipe(audio, rec, letter='a', tag='spk_turn'):
    """One line per true turn - the shape the change detector hands to the
    clustering stage (configs 3, 5)."""
    lines = []
    for k, (a, b, _) in enumerate(rec.turns):
        lines.append('audio=%s lna=%s_%d start-time=%s end-time=%s speaker=%s\n'
                     % (audio, letter, k + 1, repr(a / float(rec.rate)),
                        repr(b / float(rec.rate)), tag))
    return lines


def truth_recipe(audio, rec, letter='a'):
    """The ground-truth recipe (speaker labels = true speakers), for the
    scoring tools."""
    lines = []
    for k, (a, b, s) in enumerate(rec.turns):
        lines.append('audio=%s lna=%s_%d start-time=%s end-time=%s speaker=speaker_%d\n'
                     % (audio, letter, k + 1, repr(a / float(rec.rate)),
                        repr(b / float(rec.rate)), s + 1))
    return lines


def write_case(dirname, name, rec, recipe_lines, feaext='.fea'):
    """Write ``<dirname>/fea/<name>.fea`` and ``<dirname>/<name>.recipe``;
    returns (recipe_path, feapath)."""
    feadir = op.join(dirname, 'fea')
    os.makedirs(feadir, exist_ok=True)
    write_features(op.join(feadir, name + feaext), rec.frames)
    rpath = op.join(dirname, name + '.recipe')
    with open(rpath, 'w') as f:
        f.writelines(recipe_lines)
    return rpath, feadir


# ---- the BASELINE.json configurations -----------------

def config1(n_frames=60000, seed=1001):
    """10 min, K = 6, turns 3..19 s, one-line recipe (sliding window)."""
    return make_recording(seed, n_frames, 6)


def config2(n_frames=360000, seed=1002):
    """1 h, K = 8, turns 3..19 s, one-line recipe (growing window)."""
    return make_recording(seed, n_frames, 8)


def config3(n_frames=1080000, seed=1003):
    """3 h, K = 10, ~2,000 segments of 2..9 s, one recipe line per segment."""
    return make_recording(seed, n_frames, 10, turn_lo=2, turn_hi=9)


def config4_file(index, n_frames=60000):
    """File ``index`` of the 1,000-recording corpus: seeds 2000.., K in 2..8."""
    seed = 2000 + index
    k = 2 + (np.random.default_rng(seed ^ 0x5eed).integers(7))
    return make_recording(seed, n_frames, int(k))


def config4_frames(index):
    """Frames of ``config4_file(index)`` alone (a picklable job for a process pool: the benchmark
    generates the 1,000 recordings on all host cores)."""
    return config4_file(index).frames


def config5(n_frames=8640000, seed=1005):
    """24 h, K = 20, ~50,000 segments of 100..246 frames."""
    return make_recording(seed, n_frames, 20, turn_lo=100, turn_hi=246, unit=1)
