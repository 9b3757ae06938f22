import sys, os, io, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter('ignore')
import numpy as np
import spkdiar
from spkdiar import synth, _abi
from oracle import change_detection as OCD
ctx = _abi.Context(0)
rec = synth.make_recording(seed=101, n_frames=6000, n_speakers=3)
f = ctx.upload(rec.frames)
win, first = f.gw_run([0], [6000], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
tr = []
cd = OCD.ChangeDetection(100, 'gw', 'BIC', 1.0, 3.0, 0.1, 0.0, 1.0, trace=tr)
cd.dist_gw(rec.frames, ('/x.wav', 'a_1', 0.0, 60.0), io.StringIO())
print(len(win), len(tr))
for r, w in zip(tr, win):
    flag = '' if abs(r['maxd'] - w['maxd']) <= 1e-9 * abs(r['maxd']) else '  <<<<'
    print(r['start'], r['end'], r['maxi'], w['maxi'], r['maxd'], w['maxd'], w['ncand'], w['positive'], flag)
print('oracle total', cd.total_dist, 'gpu total', sum(float(w['maxd']) for w in win))
