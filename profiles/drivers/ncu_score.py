"""Small driver for ncu: frame statistics + one batched BIC scoring call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spkdiar
from spkdiar import synth, _abi
rec = synth.make_recording(1002, 120000, 8)
ctx = _abi.Context(0)
f = ctx.upload(rec.frames)
rng = np.random.default_rng(0)
n = 6000
a = rng.integers(0, 110000, n); m = a + rng.integers(50, 2000, n); b = np.minimum(m + rng.integers(50, 2000, n), 120000)
for _ in range(3):
    d = f.score_windows(a, m, b, _abi.BIC, 1.0)
print('ok', float(np.nanmean(d)))
