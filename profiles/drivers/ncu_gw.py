"""Small driver for ncu: statistics + one growing-window pass (BIC by default; argv[1] =
BIC | GLR | KL2) on the 1-hour recording of bench config 2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi
name = (sys.argv[1] if len(sys.argv) > 1 else 'BIC').upper()
met, thr = {'BIC': (_abi.BIC, 0.0), 'GLR': (_abi.GLR, 1500.0), 'KL2': (_abi.KL2, 4000.0)}[name]
rec = synth.make_recording(1002, 360000, 8)
ctx = _abi.Context(0)
f = ctx.upload(rec.frames)
for _ in range(2):
    win, _first = f.gw_run([0], [360000], 100.0, 100.0, 300.0, 10.0, thr, 1.0, met)
print('ok', name, len(win), int(win['positive'].sum()))
