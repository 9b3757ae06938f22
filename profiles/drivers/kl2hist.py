import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar
from spkdiar import synth, _abi
rec = synth.make_recording(1002, 360000, 8)
ctx = _abi.Context(0)
f = ctx.upload(rec.frames)
for name, met, th in (('KL2', _abi.KL2, 4000.0), ('GLR', _abi.GLR, 1500.0), ('BIC', _abi.BIC, 0.0)):
    win, _ = f.gw_run([0], [360000], 100.0, 100.0, 300.0, 10.0, th, 1.0, met)
    h = collections.Counter(); run = 0
    for r in win:
        if r['positive']:
            h[run] += 1; run = 0
        else:
            run += 1
    tot = sum(h.values())
    print(name, 'positives', tot, 'by number of negative windows before them:', sorted(h.items())[:12],
          'within first 3 windows: %.0f%%' % (100.0 * sum(v for k, v in h.items() if k <= 2) / tot))
