"""Small driver for ncu: one device batch of ten-minute recordings (BASELINE config 4) through
corpus.diarize_batch - packed statistics, one growing-window launch (one CTA per recording),
one clustering launch (one CTA per recording).  argv[1] = recordings (default 148)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi, corpus
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
items = []
for k in range(n):
    r = synth.config4_file(k)
    items.append((synth.one_line_recipe('/syn/c4_%d.wav' % k, r), r.frames))
ctx = _abi.Context(0)
for _ in range(2):
    out = corpus.diarize_batch(ctx, items, 100)
print('ok', n, out[-1][2])
