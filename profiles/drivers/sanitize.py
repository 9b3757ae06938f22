"""Small driver for compute-sanitizer: the batched corpus path and a split growing-window
search on small inputs (every new kernel of round 1d launches at least once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import spkdiar                                   # noqa: F401
from spkdiar import synth, _abi, corpus
ctx = _abi.Context(0)
items = []
for k in range(5):
    r = synth.make_recording(60 + k, 3000 + 300 * k, 3, turn_lo=3, turn_hi=6)
    items.append((synth.one_line_recipe('/syn/s%d.wav' % k, r), r.frames))
out = corpus.diarize_batch(ctx, items, 100)
print('batch ok', [o[2]['turns'] for o in out])
os.environ['SPKDIAR_GW_MINLEN'] = '4'
os.environ['SPKDIAR_GW_SPLIT_KL2'] = '1'
rec = synth.make_recording(77, 6000, 3, turn_lo=3, turn_hi=6)
with ctx.upload(rec.frames) as f:
    for met, thr in ((_abi.BIC, 0.0), (_abi.KL2, 3000.0)):
        win, _ = f.gw_run([0], [6000], 100.0, 100.0, 300.0, 10.0, thr, 1.0, met)
        print('split ok', met, len(win), int(win['positive'].sum()))
ctx.close()
