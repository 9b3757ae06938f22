"""Timeline of the two-lane corpus path (run on the GPU box): which stage of which batch runs when.
python profiles/drivers/lanes_trace.py [files] [batch] [lanes]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar                                   # noqa: F401,E402
from spkdiar import _abi, synth, corpus          # noqa: E402

files = int(sys.argv[1]) if len(sys.argv) > 1 else 296
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 74
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
items = []
for k in range(files):
    r = synth.config4_file(k % 37)
    items.append((synth.one_line_recipe('/syn/c4_%d.wav' % k, r), torch.from_numpy(r.frames).pin_memory()))
parts = [[(lines, (t.data_ptr(), t.shape[0])) for lines, t in items[b0:b0 + batch]] for b0 in range(0, files, batch)]
log = []
t00 = [0.0]
for name in ('stage_a', 'stage_b', 'stage_c', 'stage_d', 'close'):
    def wrap(fn, name=name):
        def inner(self, *a, **k):
            t0 = time.perf_counter()
            out = fn(self, *a, **k)
            log.append((threading.current_thread().name[-3:], name, id(self) % 1000, (t0 - t00[0]) * 1e3, (time.perf_counter() - t0) * 1e3))
            return out
        return inner
    setattr(corpus._BatchJob, name, wrap(getattr(corpus._BatchJob, name)))
# finer: the ABI calls inside stage_a
for name in ('upload_batch',):
    fn = getattr(_abi.Context, name)
    def inner(self, *a, _fn=fn, _name=name, **k):
        t0 = time.perf_counter()
        out = _fn(self, *a, **k)
        log.append((threading.current_thread().name[-3:], '  ' + _name, 0, (t0 - t00[0]) * 1e3, (time.perf_counter() - t0) * 1e3))
        return out
    setattr(_abi.Context, name, inner)
ctx = _abi.Context(0)
for rep in range(3):
    del log[:]
    t00[0] = time.perf_counter()
    for got in corpus.diarize_batches(ctx, parts, 100, lanes=lanes):
        pass
    wall = (time.perf_counter() - t00[0]) * 1e3
    print('rep %d: wall %.1f ms = %.3f ms per recording' % (rep, wall, wall / files))
for rec in sorted(log, key=lambda r: r[3]):
    print('%s %-14s job %3d  start %7.1f  dur %6.1f' % rec)
ctx.close()
