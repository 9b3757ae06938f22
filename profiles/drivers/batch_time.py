"""Timing breakdown of the batched corpus path (run on the GPU box):
python profiles/drivers/batch_time.py [files] [batch]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar                                   # noqa: F401,E402
from spkdiar import _abi, synth, corpus          # noqa: E402

files = int(sys.argv[1]) if len(sys.argv) > 1 else 148
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 148
items = []
for k in range(files):
    r = synth.config4_file(k)
    items.append((synth.one_line_recipe('/syn/c4_%d.wav' % k, r), torch.from_numpy(r.frames).pin_memory()))
ctx = _abi.Context(0)
part = [(lines, (t.data_ptr(), t.shape[0])) for lines, t in items]
for rep in range(3):
    ctx.profile(True)
    t0 = time.perf_counter()
    for b0 in range(0, files, batch):
        out = corpus.diarize_batch(ctx, part[b0:b0 + batch], 100)
    t1 = time.perf_counter()
    prof = ctx.profile_read()
    ctx.profile(False)
    dev = sum(v[0] for v in prof.values())
    print('rep %d: wall %.1f ms  device %.1f ms  host %.1f ms  %s' % (
        rep, (t1 - t0) * 1e3, dev, (t1 - t0) * 1e3 - dev,
        ' '.join('%s %.2f ms/%d' % (k, v[0], v[1]) for k, v in prof.items())))
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for b0 in range(0, files, batch):
    corpus.diarize_batch(ctx, part[b0:b0 + batch], 100)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
ctx.close()
