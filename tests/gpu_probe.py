"""Ad-hoc GPU probe used during bring-up (not a pytest file)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spkdiar
from spkdiar import synth, _abi
from oracle import distances as OD
from oracle import change_detection as OCD, clustering as OCL
import io, warnings
warnings.simplefilter('ignore')

def rel(a, b, scale=None):
    a = np.asarray(a, float); b = np.asarray(b, float)
    s = np.maximum(np.abs(b), 1e-300) if scale is None else scale
    return np.max(np.abs(a - b) / s)

ctx = _abi.Context(0)
print('SMs', ctx.sm_count)
rec = synth.make_recording(5, 20000, 4)
x = rec.frames
t = time.time(); f = ctx.upload(x); print('upload+stats %.1f ms' % ((time.time() - t) * 1e3))
# K1
for (a, b) in [(0, 20000), (100, 160), (19990, 20000), (12345, 17000), (7, 7)]:
    s, q, sh = f.stats_window(a, b)
    xc = x[a:b].astype(np.float64) - sh
    s_ref = xc.sum(0); Q = xc.T @ xc
    q_ref = Q[np.tril_indices(39)]
    print('K1', a, b, 'sum err %.2e' % np.max(np.abs(s - s_ref)), 'Q err %.2e' % np.max(np.abs(q - q_ref)), 'Qmax %.1f' % (np.max(np.abs(q_ref)) if b > a else 0))
# K2
rng = np.random.default_rng(1)
n = 300
a = rng.integers(0, 15000, n); l1 = rng.integers(40, 2000, n); l2 = rng.integers(40, 2000, n)
m = a + l1; b = np.minimum(m + l2, 20000)
for name, met in (('BIC', _abi.BIC), ('GLR', _abi.GLR), ('KL2', _abi.KL2)):
    t = time.time(); d, tr = f.score_windows(a, m, b, met, 1.3, terms=True); dt = time.time() - t
    ref = []
    for k in range(n):
        a1, a2 = x[a[k]:m[k]], x[m[k]:b[k]]
        if name == 'BIC': ref.append(OD.bic_cd(a1, a2, x[a[k]:b[k]], 1.3))
        elif name == 'GLR': ref.append(OD.glr(a1, a2))
        else: ref.append(OD.kl2(a1, a2))
    ref = np.array(ref)
    print('K2', name, 'max rel err %.3e' % rel(d, ref), 'median |d| %.3g' % np.median(np.abs(ref)), '%.1f ms' % (dt * 1e3))
    bad = np.argsort(-np.abs(d - ref) / np.abs(ref))[:3]
    print('   worst', [(int(l1[k]), int(b[k] - m[k]), float(d[k]), float(ref[k])) for k in bad])
# gw
for name, met, thr, lam in (('BIC', _abi.BIC, 0.0, 1.0), ('GLR', _abi.GLR, 350.0, 1.0), ('KL2', _abi.KL2, 60.0, 1.0)):
    for mg in (0, 1):
        segs = [(0, 20000)] if mg == 0 else [(0, 6000), (6000, 13000), (13000, 20000)]
        t = time.time()
        win, first = f.gw_run([s[0] for s in segs], [s[1] for s in segs], 100.0, 100.0, 300.0, 10.0, thr, lam, met)
        dt = time.time() - t
        # oracle
        tr = []
        cd = OCD.ChangeDetection(100, 'gw', name, 1.0, 3.0, 0.1, thr, lam, trace=tr)
        out = io.StringIO()
        for k, (sa, sb) in enumerate(segs):
            cd.dist_gw(x[sa:sb], ('/x.wav', 'a_%d' % (k + 1), sa / 100.0, sb / 100.0), out)
        ok = len(tr) == len(win)
        if ok:
            for r, w in zip(tr, win):
                if not (r['start'] == w['start'] and r['end'] == w['end'] and bool(r['positive']) == bool(w['positive'])):
                    ok = False; break
                if r['positive'] and r['maxi_fine'] != w['maxi_fine']: ok = False; break
                if (not r['positive']) and r['maxi'] is not None and r['maxi'] != w['maxi']: ok = False; break
        errs = [abs(r['maxd'] - w['maxd']) / abs(r['maxd']) for r, w in zip(tr, win)]
        print('GW', name, 'chains', len(segs), 'windows', len(win), 'oracle', len(tr), 'identical decisions:', ok,
              'max rel maxd err %.2e' % (max(errs) if errs else 0), 'changes', int(win['positive'].sum()), '%.1f ms' % (dt * 1e3))
# clustering
rec2 = synth.make_recording(11, 16000, 4, turn_lo=2, turn_hi=6)
f2 = ctx.upload(rec2.frames)
sa = [t_[0] for t_ in rec2.turns]; sb = [t_[1] for t_ in rec2.turns]
for variant in (1, 2):
    for name, met, thr in (('BIC', _abi.BIC, 0.0), ('GLR', _abi.GLR, 500.0)):
        cl = f2.cluster(sa, sb, met, 1.3)
        t = time.time(); merges, stats = cl.run(thr, 0, variant); dt = time.time() - t
        tr = []
        oc = OCL.Clustering(100, variant, 'hi', name, thr, 0, 1.3, trace=tr)
        oc.speakers = [[(float(a_), float(b_), k)] for k, (a_, b_) in enumerate(zip(sa, sb))]
        recipe = [('/x.wav', 'a_%d' % (k + 1), a_ / 100.0, b_ / 100.0) for k, (a_, b_) in enumerate(zip(sa, sb))]
        oc.spk_cluster_hi(rec2.frames, recipe, io.StringIO())
        same = len(tr) == len(merges) and all(t_[0] == m_['a'] and t_[1] == m_['b'] for t_, m_ in zip(tr, merges))
        err = max([abs(t_[2] - m_['d']) / abs(t_[2]) for t_, m_ in zip(tr, merges)] or [0])
        print('CL v%d' % variant, name, 'merges', len(merges), 'oracle', len(tr), 'same sequence:', same, 'max rel d err %.2e' % err,
              'stats', stats, 'oracle stats', oc.max_dist, oc.min_dist, '%.1f ms' % (dt * 1e3))
        cl.close()
print('launches', ctx.launches)
