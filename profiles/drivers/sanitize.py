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
# round 2: frames-only handles (K5), KL2 in the clustering engine, the owner-warp merge engine, the two
# device-resident sequential modes, fewer than 39 dimensions
ctx = _abi.Context(0)
rec = synth.make_recording(78, 6000, 3, turn_lo=2, turn_hi=5)
sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
with ctx.upload_frames(rec.frames) as f:
    for met, thr, ms in ((_abi.BIC, 0.0, 0), (_abi.KL2, -1.0, 2)):
        with f.cluster(sa, sb, met, 1.3) as cl:
            m, _ = cl.run(thr, ms, 1)
        print('frames-only cluster ok', met, len(m))
    d, first, best = f.cluster_inorder([], sa, sb, _abi.BIC, 1.3, 0.0)
    print('in-order ok', len(d), int((best >= 0).sum()))
    t, dist, merged, memo = f.merge_chain(sa, sb, _abi.BIC, 1.0, 0.0, True, None)
    print('merge chain ok', len(dist), int(merged.sum()))
os.environ['SPKDIAR_CL_GENERAL'] = '1'
with ctx.upload(rec.frames) as f, f.cluster(sa, sb, _abi.GLR, 1.3) as cl:
    print('general engine ok', len(cl.run(1e9, 2, 2)[0]))
del os.environ['SPKDIAR_CL_GENERAL']
r13 = synth.make_recording(79, 3000, 2, dim=13)
with ctx.upload(r13.frames) as f:
    win, _ = f.gw_run([0], [3000], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
    print('dim 13 ok', len(win))
ctx.close()
