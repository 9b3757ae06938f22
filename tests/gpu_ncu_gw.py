"""Small driver for ncu: statistics + growing-window BIC on the 1-hour recording."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spkdiar
from spkdiar import synth, _abi
rec = synth.make_recording(1002, 360000, 8)
ctx = _abi.Context(0)
f = ctx.upload(rec.frames)
for _ in range(2):
    win, _first = f.gw_run([0], [360000], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
print('ok', len(win), int(win['positive'].sum()))
