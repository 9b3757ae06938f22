"""Timing driver (not a test): agglomerative clustering of BASELINE config 3 (3-hour
recording, ~2,000 segments) or a prefix of config 5 (argv: nsegments).  SPKDIAR_CL_DEBUG=1
prints the merge loop's phase counters.  Usage: python profiles/drivers/cluster_time.py [c3 | <nseg>] [variant]"""
import hashlib
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi

what = sys.argv[1] if len(sys.argv) > 1 else 'c3'
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.perf_counter()
if what == 'c3':
    rec = synth.config3()
else:
    nseg = int(what)
    rec = synth.config5(n_frames=nseg * 173)
print('generated %d frames, %d segments in %.1f s' % (rec.frames.shape[0], len(rec.turns), time.perf_counter() - t0), flush=True)
ctx = _abi.Context(0)
f = ctx.upload_frames(rec.frames) if os.environ.get('SPKDIAR_FRAMES_ONLY', '1') != '0' else ctx.upload(rec.frames)
a = [t[0] for t in rec.turns]
b = [t[1] for t in rec.turns]
for rep in range(2):
    ctx.profile(True)
    t0 = time.perf_counter()
    with f.cluster(a, b, _abi.BIC, 1.3) as cl:
        merges, stats = cl.run(0.0, 0, variant)
    dt = time.perf_counter() - t0
    prof = ctx.profile_read()
    ctx.profile(False)
print('segments %d merges %d final speakers %d  wall %.1f ms  fill %.2f ms  merge loop %.2f ms  sha %s'
      % (len(a), len(merges), len(a) - len(merges), dt * 1e3, prof['score'][0], prof['merge'][0],
         hashlib.sha256(merges.tobytes()).hexdigest()[:12]), flush=True)
hours = rec.frames.shape[0] / 100.0 / 3600.0
print('audio %.2f h -> %.2f audio-hours/s (clustering only)' % (hours, hours / dt))
