"""Timing driver (not a test): the three growing-window passes and the clustering of
bench config 2 on the 1-hour recording, with the kernel's phase counters
(SPKDIAR_GW_DEBUG=1).  Usage: python profiles/drivers/gw_time.py [BIC GLR KL2 CL]"""
import hashlib
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi

which = [a.upper() for a in sys.argv[1:]] or ['BIC', 'GLR', 'KL2', 'CL']
rec = synth.make_recording(1002, 360000, 8)
ctx = _abi.Context(0)
f = ctx.upload(rec.frames)
bic = None
for name, met, th in (('BIC', _abi.BIC, 0.0), ('GLR', _abi.GLR, 1500.0), ('KL2', _abi.KL2, 4000.0)):
    if name not in which and not (name == 'BIC' and 'CL' in which):
        continue
    for rep in range(3):
        t0 = time.perf_counter()
        win, _ = f.gw_run([0], [360000], 100.0, 100.0, 300.0, 10.0, th, 1.0, met)
        dt = time.perf_counter() - t0
    if name == 'BIC':
        bic = win
    h = hashlib.sha256(win.tobytes()).hexdigest()[:12]
    print('%s windows %d changes %d  %.2f ms  sha %s' % (name, len(win), int(win['positive'].sum()), dt * 1e3, h), flush=True)
if 'CL' in which:
    cuts = [0.0] + [float(r['start']) + float(r['maxi_fine']) for r in bic if r['positive']] + [360000.0]
    a = [int(c) for c in cuts[:-1]]
    b = [int(c) for c in cuts[1:]]
    for rep in range(3):
        t0 = time.perf_counter()
        with f.cluster(a, b, _abi.BIC, 1.3) as cl:
            merges, _ = cl.run(0.0, 0, 1)
        dt = time.perf_counter() - t0
    print('CL segments %d merges %d  %.2f ms  sha %s' % (len(a), len(merges), dt * 1e3,
                                                        hashlib.sha256(merges.tobytes()).hexdigest()[:12]), flush=True)
