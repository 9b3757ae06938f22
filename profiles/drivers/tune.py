"""Threshold tuning for the GLR / KL2 growing-window passes of bench config 2
(run once on the GPU; the chosen values are frozen in bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spkdiar
from spkdiar import synth, _abi
rec = synth.make_recording(1002, 360000, 8)
truth = np.array([t[0] for t in rec.turns[1:]], dtype=float)
ctx = _abi.Context(0)
t0 = time.time(); f = ctx.upload(rec.frames); print('upload+stats %.1f ms' % ((time.time() - t0) * 1e3), 'true turns', len(rec.turns))
def score(win):
    cp = np.array([w['start'] + w['maxi_fine'] for w in win if w['positive']])
    hit = sum(np.min(np.abs(truth - c)) <= 25 for c in cp) if len(cp) else 0
    found = sum(np.min(np.abs(cp - t)) <= 25 for t in truth) if len(cp) else 0
    return len(cp), hit, found
for name, met, ths in (('BIC', _abi.BIC, [0.0]), ('GLR', _abi.GLR, [1500]),
                       ('KL2', _abi.KL2, [700, 4000])):
    for th in ths:
        t0 = time.time()
        win, _ = f.gw_run([0], [360000], 100.0, 100.0, 300.0, 10.0, float(th), 1.0, met)
        dt = time.time() - t0
        n, hit, found = score(win)
        print('%s t=%-6g windows %5d changes %4d precise %4d recall %4d/%d  %.1f ms' % (name, th, len(win), n, hit, found, len(truth), dt * 1e3), flush=True)
