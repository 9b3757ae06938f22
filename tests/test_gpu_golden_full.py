"""GPU parity against the FULL-SIZE golden fixtures (tests/golden/full/, made by
tests/golden/make_golden_full.py from the reference's own scripts and the oracle's trace):

* BASELINE config 2, the WHOLE hour, growing window with BIC, GLR (-t 1500) and KL2 (-t 4000):
  every window record (start, end, best offset, decision, fine-tuned offset) bit-identical,
  distances to 1e-9, the recipe byte-identical, the log to 1e-9;
* config 3 cut to 200 / 400 segments, spk-clustering.py and spk-clustering2.py: the merge
  sequence identical, distances to 1e-9, recipe byte-identical;
* three ten-minute files of config 4 through both stages (one at a time and as a device batch);
* config 3 at FULL size (1,978 segments): the agglomeration replayed on the host from the rows the
  device wrote - every merge must be ``ndarray.argmin`` of the live matrix at that moment.

Distances are compared as everywhere else: |d_gpu - d_ref| <= 1e-9 * max(|d_ref|, largest
0.5 N ln|S| term the distance is a difference of).  Every test also writes a MARGIN AUDIT
(gpurun_out/audit_*.json; copied to profiles/): the observed errors and how far the reference's own
decisions (maxd vs threshold, winner vs runner-up) are from flipping, in units of that tolerance."""

import io
import json
import os

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from conftest import GOLDEN_DIR, logs_match, run_product
from spkdiar import _abi, synth

pytestmark = pytest.mark.gpu

FULL = os.path.join(GOLDEN_DIR, 'full')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AUDIT_DIR = os.environ.get('SPKDIAR_AUDIT_DIR', os.path.join(ROOT, 'gpurun_out'))
REL = 1e-9
EPS = 2.220446049250313e-16


def load_full(name):
    p = os.path.join(FULL, name + '.json')
    if not os.path.isfile(p):
        pytest.skip('fixture %s not generated' % name)
    with open(p) as f:
        return json.load(f)


def write_audit(name, doc):
    try:
        os.makedirs(AUDIT_DIR, exist_ok=True)
        with open(os.path.join(AUDIT_DIR, 'audit_%s.json' % name), 'w') as f:
            json.dump(doc, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope='module')
def ctx():
    c = _abi.Context(0)
    yield c
    c.close()


def _half_n_logdet(x):
    """0.5 N ln|S| of a frame slice, the size of the terms a BIC / GLR distance is a difference of."""
    if x.shape[0] < 2:
        return 0.0
    sign, ld = np.linalg.slogdet(np.cov(x, rowvar=0))
    return 0.5 * x.shape[0] * abs(ld) if np.isfinite(ld) else 0.0


def _sha(frames):
    import hashlib
    return hashlib.sha256(frames.tobytes()).hexdigest()


# ---------------------------------------------------------------- config 2: the whole hour ----------

@pytest.fixture(scope='module')
def hour(ctx):
    rec = synth.config2()
    feat = ctx.upload(rec.frames)
    yield rec, feat
    feat.close()


def _logdet_x(x):
    """ln|cov(x)| in 80-bit arithmetic (two-pass covariance, LDL^T): the arbiter between two fp64 results."""
    xs = x.astype(np.longdouble)
    xs = xs - xs.mean(0)
    return _logdet_of(xs.T @ xs / np.longdouble(x.shape[0] - 1))


def _logdet_of(A):
    A = A.copy()
    out = np.longdouble(0)
    for c in range(A.shape[0]):
        out += np.log(A[c, c])
        l = A[c + 1:, c] / A[c, c]
        A[c + 1:, c + 1:] -= np.outer(l, A[c + 1:, c])
    return out


def _inv_diag_x(A):
    """diag(A^-1) of a symmetric positive definite matrix in 80-bit arithmetic: A = L D L^T,
    A^-1 = L^-T D^-1 L^-1, so entry i is the sum over k of (L^-1)[k, i]^2 / D[k]."""
    A = A.copy()
    d = A.shape[0]
    L = np.eye(d, dtype=np.longdouble)
    D = np.zeros(d, dtype=np.longdouble)
    for c in range(d):
        D[c] = A[c, c]
        L[c + 1:, c] = A[c + 1:, c] / A[c, c]
        A[c + 1:, c + 1:] -= np.outer(L[c + 1:, c], A[c + 1:, c])
    Li = np.eye(d, dtype=np.longdouble)
    for c in range(d):                       # forward substitution, column by column of the identity
        for r in range(c + 1, d):
            Li[r, c] = -np.dot(L[r, c:r], Li[c:r, c])
    return np.sum(Li * Li / D[:, None], axis=0)


def _cov_x(x):
    xs = x.astype(np.longdouble)
    xs = xs - xs.mean(0)
    return xs.T @ xs / np.longdouble(x.shape[0] - 1)


def exact_distance(metric, x1, x2, lam):
    """The reference's formula (CD:72-121) evaluated in 80-bit arithmetic."""
    n1, n2 = np.longdouble(x1.shape[0]), np.longdouble(x2.shape[0])
    n = n1 + n2
    if metric == _abi.BIC:
        pen = np.longdouble(lam) * np.longdouble(0.5) * (39 + np.longdouble(0.5) * 39 * 40) * np.log(n)
        return float(0.5 * n * _logdet_x(np.concatenate((x1, x2))) - 0.5 * n1 * _logdet_x(x1)
                     - 0.5 * n2 * _logdet_x(x2) - pen)
    if metric == _abi.KL2:
        # CD:124-133: only the diagonals of S and of its inverse reach the trace (Q3); the means are the
        # reference's own float32 sequential sums (Q4) - they are part of the definition, not of the rounding
        s1, s2 = _cov_x(x1), _cov_x(x2)
        p1, p2 = _inv_diag_x(s1), _inv_diag_x(s2)
        delta = (np.mean(x1, 0) - np.mean(x2, 0)).astype(np.longdouble)
        return float(0.5 * np.sum((np.diag(s1) - np.diag(s2)) * (p2 - p1)) + 0.5 * np.sum((p1 + p2) * delta * delta))
    mix = (n1 / n) * _cov_x(x1) + (n2 / n) * _cov_x(x2)
    return float(-(n / 2) * ((n1 / n) * _logdet_x(x1) + (n2 / n) * _logdet_x(x2) - _logdet_of(mix)))


def check_windows(x, win, g, metric, thr, lam, name):
    """Records of one chain against the fixture's: positions and decisions bit-identical, distances to 1e-9 of
    the largest term - and where two fp64 evaluations differ by more than that (covariances with a condition
    number of 1e9: the reference's own LU determinant carries cond * eps), the 80-bit value arbitrates: the
    device must be at least as close to it as the reference is.  Writes the margin audit; returns the largest
    relative deviation from the reference (what a log line may differ by)."""
    assert len(win) == len(g['start']), (len(win), len(g['start']))
    errs, units, refs, thr_margin, gap_margin = [], [], [], [], []
    arbitrated = []
    for k, r in enumerate(win):
        want_pos = bool(g['positive'][k])
        assert r['start'] == g['start'][k] and r['end'] == g['end'][k], (k, r, g['start'][k], g['end'][k])
        assert bool(r['positive']) == want_pos, (k, r)
        assert g['maxi'][k] is not None and r['maxi'] == g['maxi'][k], (k, r, g['maxi'][k])
        s, e = int(r['start']), int(r['end'])
        m = int(r['start'] + r['maxi'])
        pairs = [(r['maxd'], g['maxd'][k], g['gap'][k], m)]
        if want_pos:
            assert r['maxi_fine'] == g['maxi_fine'][k], (k, r, g['maxi_fine'][k])
            pairs.append((r['maxd_fine'], g['maxd_fine'][k], g['gap_fine'][k], int(r['start'] + r['maxi_fine'])))
        for got, ref, gap, mm in pairs:
            if metric == _abi.KL2:
                # the reference's KL2 inverts the covariance of each side: conditioning-limited below 2 d frames
                unit = (REL if min(mm - s, e - mm) >= 78 else 1e-6) * abs(ref)
            else:
                scale = max(_half_n_logdet(x[s:mm]), _half_n_logdet(x[mm:e]), _half_n_logdet(x[s:e]))
                unit = REL * max(abs(ref), scale)
            err = abs(got - ref)
            if err > unit:
                # Two fp64 evaluations part by more than 1e-9: the 80-bit value arbitrates.  The device must be
                # as close to it as the reference is (x2) - or within what ANY fp64 evaluation can promise for
                # these covariances: ln|S| carries cond(S) * eps (a 40-frame side in 39 dimensions has
                # cond(S) ~ 1e12 and the reference's own LU determinant is 15 tolerances off there).
                truth = exact_distance(metric, x[s:mm], x[mm:e], lam)
                e_gpu, e_ref = abs(got - truth), abs(ref - truth)
                sides = [x[s:mm], x[mm:e]] if metric == _abi.KL2 else [x[s:mm], x[mm:e], x[s:e]]
                conds = [float(np.linalg.cond(np.cov(v, rowvar=0))) for v in sides]
                if metric == _abi.KL2:
                    # diag(pinv(S)) carries cond(S) * eps relative to itself, and so does the distance
                    limit = 2.0 * EPS * max(conds) * abs(truth)
                else:
                    limit = 2.0 * EPS * max(0.5 * v.shape[0] * c for v, c in zip(sides, conds))
                assert e_gpu <= max(2.0 * e_ref, unit, limit), (k, got, ref, truth, e_gpu / unit, e_ref / unit, conds)
                arbitrated.append(dict(window=k, frames_left=mm - s, frames_right=e - mm, cond_max=max(conds),
                                       reference=ref, device=float(got), exact80=truth,
                                       device_error_in_tolerances=e_gpu / unit,
                                       reference_error_in_tolerances=e_ref / unit,
                                       fp64_conditioning_limit_in_tolerances=limit / unit,
                                       margin_to_threshold_in_tolerances=abs(ref - thr) / unit,
                                       runner_up_gap_in_tolerances=gap / unit))
            errs.append(err)
            units.append(unit)
            refs.append(abs(ref))
            thr_margin.append(abs(ref - thr))
            gap_margin.append(gap)
    errs, units, refs = np.array(errs), np.array(units), np.array(refs)
    thr_margin, gap_margin = np.array(thr_margin), np.array(gap_margin)
    rel = errs / np.maximum(refs, 1e-300)
    write_audit(name, dict(
        windows=len(win), changes=int(np.sum(win['positive'])), distances_compared=len(errs),
        positions_and_decisions='bit-identical',
        tolerance='1e-9 * max(|d|, largest 0.5 N ln|S| term)' if metric != _abi.KL2
        else '1e-9 * |d| (1e-6 * |d| when a side has fewer than 78 frames)',
        err_over_tolerance_max=float(np.max(errs / units)), err_over_tolerance_median=float(np.median(errs / units)),
        err_abs_max=float(errs.max()), err_rel_to_d_max=float(rel.max()), err_rel_to_d_median=float(np.median(rel)),
        beyond_tolerance=len(arbitrated),
        beyond_tolerance_shortest_side_max_frames=max([min(a['frames_left'], a['frames_right']) for a in arbitrated] or [0]),
        beyond_tolerance_cond_min=min([a['cond_max'] for a in arbitrated] or [0.0]),
        beyond_tolerance_reference_itself_off_by_tolerances_max=max([a['reference_error_in_tolerances'] for a in arbitrated] or [0.0]),
        device_closer_to_exact_than_reference=int(sum(1 for a in arbitrated
                                                      if a['device_error_in_tolerances'] <= a['reference_error_in_tolerances'])),
        beyond_tolerance_arbitrated_by_80bit=sorted(arbitrated, key=lambda a: -a['device_error_in_tolerances'])[:25],
        threshold_margin_min_abs=float(thr_margin.min()),
        threshold_margin_min_in_tolerances=float(np.min(thr_margin / units)),
        runner_up_gap_min_abs=float(gap_margin.min()),
        runner_up_gap_min_in_tolerances=float(np.min(gap_margin / units)),
        decisions_inside_tolerance=int(np.sum(thr_margin <= units) + np.sum(gap_margin <= units)),
        decisions_inside_observed_error=int(np.sum(thr_margin <= errs) + np.sum(gap_margin <= errs))))
    # positions are bit-identical (asserted above); and no decision of the reference lies inside the deviation
    assert np.all(thr_margin > errs)
    return float(rel.max())


@pytest.mark.parametrize('name,metric,thr', [('c2_bic', _abi.BIC, 0.0), ('c2_glr', _abi.GLR, 1500.0),
                                             ('c2_kl2', _abi.KL2, 4000.0)])
def test_config2_whole_hour_records_equal_reference(hour, name, metric, thr):
    gold = load_full(name)
    rec, feat = hour
    n = rec.frames.shape[0]
    assert n == gold['frames'] == 360000 and _sha(rec.frames) == gold['frames_sha256']
    win, _ = feat.gw_run([0], [n], 100.0, 100.0, 300.0, 10.0, thr, 1.0, metric)
    _WINDOW_REL[name] = check_windows(rec.frames, win, gold['windows'], metric, thr, 1.0, name)


_WINDOW_REL = {}        # largest relative deviation of a window distance (arbitrated above), per fixture


@pytest.mark.parametrize('name', ['c2_bic', 'c2_glr', 'c2_kl2'])
def test_config2_whole_hour_cli_equals_reference(name, tmp_path, ctx):
    gold = load_full(name)
    rec = synth.config2()
    rpath, feadir = synth.write_case(str(tmp_path), 'c2', rec, synth.one_line_recipe('/syn/c2.wav', rec))
    out = str(tmp_path / 'out.recipe')
    stdout, _ = run_product('cd', 0, [rpath, feadir, '-o', out] + gold['flags'], ctx)
    assert open(out).read() == gold['recipe']
    # a log line may differ from the reference's by what the window behind it differs by (checked, and
    # arbitrated in 80-bit arithmetic where beyond the tolerance, by the test above)
    bad = logs_match(stdout.replace(str(tmp_path), '<TMP>'), gold['stdout'],
                     max(1e-6, 2 * _WINDOW_REL.get(name, 0.0)) if name == 'c2_kl2' else REL)
    assert bad is None, bad


def check_merges(x, sa, sb, merges, want, variant, name):
    """A merge sequence against the fixture's: pairs identical, distances to 1e-9 of the largest term - or, for
    clusters whose covariance is nearly singular (a 40-frame segment in 39 dimensions), within what fp64 can
    promise (cond * eps), arbitrated by the 80-bit value as in check_windows.  Returns the largest relative
    deviation (what a `Merging:` log line may differ by)."""
    assert len(merges) == len(want), (len(merges), len(want))
    nseg = len(sa)
    members = [[k] for k in range(nseg)]
    ratios, gaps, rels, beyond = [], [], [0.0], []
    for k, (m, w) in enumerate(zip(merges, want)):
        a, b = int(m['a']), int(m['b'])
        assert (a, b) == (w[0], w[1]), (k, m, w)
        fa = np.concatenate([x[sa[i]:sb[i]] for i in members[a]])
        fb = np.concatenate([x[sa[i]:sb[i]] for i in members[b]])
        scale = max(_half_n_logdet(fa), _half_n_logdet(fb), _half_n_logdet(np.concatenate((fa, fb))))
        unit = REL * max(abs(w[2]), scale)
        err = abs(m['d'] - w[2])
        if variant == 1 or np.isfinite(w[2]):
            if err > unit:
                truth = exact_distance(_abi.BIC, fa, fb, 1.3)
                e_gpu, e_ref = abs(m['d'] - truth), abs(w[2] - truth)
                sides = [fa, fb, np.concatenate((fa, fb))]
                conds = [float(np.linalg.cond(np.cov(v, rowvar=0))) for v in sides]
                limit = 2.0 * EPS * max(0.5 * v.shape[0] * c for v, c in zip(sides, conds))
                assert e_gpu <= max(2.0 * e_ref, unit, limit), (k, m, w, truth, e_gpu / unit, e_ref / unit, conds)
                beyond.append(dict(merge=k, frames=[int(v.shape[0]) for v in sides[:2]], cond_max=max(conds),
                                   device_error_in_tolerances=e_gpu / unit, reference_error_in_tolerances=e_ref / unit,
                                   runner_up_gap_in_tolerances=w[3] / unit))
            ratios.append(err / unit)
            gaps.append(w[3] / unit)
            rels.append(err / max(abs(w[2]), 1e-300))
        members[a].extend(members[b])
        members.pop(b)
    write_audit(name + '_merges' if name.startswith('c4') else name, dict(
        segments=nseg, merges=len(merges), merge_pairs='identical', err_over_tolerance_max=float(max(ratios or [0.0])),
        err_over_tolerance_median=float(np.median(ratios or [0.0])), beyond_tolerance=len(beyond),
        beyond_tolerance_arbitrated_by_80bit=beyond[:25],
        runner_up_gap_min_in_tolerances=float(min(gaps or [0.0])),
        merges_with_gap_inside_tolerance=int(sum(1 for v in gaps if v <= 1.0))))
    return float(max(rels))


# ---------------------------------------------------------------- config 3 cut to 200 / 400 ----------

def _c3_cut(nseg):
    full = synth.config3()
    cut = full.turns[nseg - 1][1]
    return synth.Recording(full.frames[:cut].copy(), full.turns[:nseg], full.rate)


@pytest.mark.parametrize('name', ['c3_cl1_200', 'c3_cl2_200', 'c3_cl1_400', 'c3_cl2_400', 'c3_cl1_800'])
def test_config3_cut_merge_sequence_equals_reference(name, tmp_path, ctx):
    gold = load_full(name)
    variant = gold['variant']
    nseg = int(name.rsplit('_', 1)[1])
    rec = _c3_cut(nseg)
    assert _sha(rec.frames) == gold['frames_sha256']
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    with ctx.upload(rec.frames) as feat, feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
        merges, stats = cl.run(0.0, 0, variant)
    check_merges(rec.frames, sa, sb, merges, gold['merges'], variant, name)
    # and through the command line: recipe byte for byte, log numbers to 1e-9
    rpath, feadir = synth.write_case(str(tmp_path), 'c3', rec, synth.turn_recipe('/syn/c3.wav', rec))
    assert open(rpath).read() == gold['recipe_in']
    out = str(tmp_path / 'out.recipe')
    stdout, _ = run_product('cl', variant, [rpath, feadir + '/', '-o', out] + gold['flags'], ctx)
    assert open(out).read() == gold['recipe']
    bad = logs_match(stdout.replace(str(tmp_path), '<TMP>'), gold['stdout'], REL)
    assert bad is None, bad


# ---------------------------------------------------------------- config 4: both stages ----------

@pytest.mark.parametrize('index', [0, 1, 2])
def test_config4_file_both_stages_equal_reference(index, tmp_path, ctx):
    gold = load_full('c4_f%d' % index)
    rec = synth.config4_file(index)
    assert _sha(rec.frames) == gold['frames_sha256']
    n = rec.frames.shape[0]
    with ctx.upload(rec.frames) as feat:
        win, _ = feat.gw_run([0], [n], 100.0, 100.0, 300.0, 10.0, 0.0, 1.0, _abi.BIC)
    # window by window (with the 80-bit arbiter where fp64 evaluations part); the log's sums and averages may
    # then differ from the reference's by what the windows differ by
    rel = check_windows(rec.frames, win, gold['windows'], _abi.BIC, 0.0, 1.0, 'c4_f%d' % index)
    wav = 'c4_%d' % index
    rpath, feadir = synth.write_case(str(tmp_path), wav, rec, synth.one_line_recipe('/syn/%s.wav' % wav, rec))
    mid, out = str(tmp_path / 'turns.recipe'), str(tmp_path / 'out.recipe')
    s1, _ = run_product('cd', 0, [rpath, feadir, '-o', mid] + gold['flags_cd'], ctx)
    assert open(mid).read() == gold['recipe_cd']
    bad = logs_match(s1.replace(str(tmp_path), '<TMP>'), gold['stdout_cd'], max(REL, 2 * rel))
    assert bad is None, bad
    s2, _ = run_product('cl', 1, [mid, feadir + '/', '-o', out] + gold['flags_cl'], ctx)
    assert open(out).read() == gold['recipe']
    # the clustering stage merge by merge (segments = the turns of the recipe just written, CL1:281, 46-52)
    from spkdiar import recipe as recipe_mod
    turns = recipe_mod.parse(gold['recipe_cd'].splitlines(True))
    sa, sb = [int(t.start * 100.0) for t in turns], [int(t.end * 100.0) for t in turns]
    with ctx.upload(rec.frames) as feat, feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
        merges, _ = cl.run(0.0, 0, 1)
    rel2 = check_merges(rec.frames, sa, sb, merges, gold['merges'], 1, 'c4_f%d' % index)
    bad = logs_match(s2.replace(str(tmp_path), '<TMP>'), gold['stdout_cl'], max(REL, 2 * rel2))
    assert bad is None, bad


def test_config4_files_as_device_batch_equal_reference(ctx):
    from spkdiar import corpus
    golds = [load_full('c4_f%d' % k) for k in range(3)]
    batch = []
    for k in range(3):
        rec = synth.config4_file(k)
        batch.append((synth.one_line_recipe('/syn/c4_%d.wav' % k, rec), rec.frames))
    res = corpus.diarize_batch(ctx, batch, frame_rate=100)
    for (turns, clustered, _), g in zip(res, golds):
        assert turns == g['recipe_cd']
        assert clustered == g['recipe']


# ---------------------------------------------------------------- config 3 full size: argmin replay ----------

@pytest.mark.parametrize('variant', [1, 2])
def test_config3_full_size_every_merge_is_argmin_of_live_matrix(variant, ctx):
    """spk-clustering.py:203-205 / spk-clustering2.py:187-191 at N = 1,978: replay the agglomeration on the
    host.  The device supplies the initial matrix and, per merge, the row it rewrote; the host keeps the
    live matrix (dead rows / columns +inf: deleting them preserves the flat order) and requires every merge
    of the device to be numpy's argmin of it, with the very same double."""
    rec = synth.config3()
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    n = len(sa)
    with ctx.upload(rec.frames) as feat, feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
        none, _ = cl.run(-np.inf, 0, variant)              # no merge: the initial fill
        assert len(none) == 0
        M, alive = cl.matrix()
        assert alive.all()
        rows = cl.rowlog(n)
        merges, stats = cl.run(0.0, 0, variant)
        Mend, alive_end = cl.matrix()
    assert len(merges) > 1900
    live = list(range(n))                                   # compacted index -> original index
    dead = np.zeros(n, dtype=bool)
    gaps = []
    for k, m in enumerate(merges):
        idx = int(np.argmin(M))                             # NaN first, first flat index wins
        r, c = divmod(idx, n)
        a, b = (r, c) if r < c else (c, r)
        assert (live.index(a), live.index(b)) == (int(m['a']), int(m['b'])), (k, m, a, b)
        assert M[r, c] == m['d'] or (np.isnan(M[r, c]) and np.isnan(m['d'])), (k, M[r, c], m['d'])
        if k % 16 == 0:
            flat = M.ravel().copy()
            flat[idx] = np.inf
            if variant == 1:
                flat[c * n + r] = np.inf
            gaps.append(float(np.min(flat) - m['d']))
        # the merge: b dies, row a (and column a in variant 1) as the device rewrote it
        dead[b] = True
        live.remove(b)
        M[b, :] = np.inf
        M[:, b] = np.inf
        new = rows[k].copy()
        keep = M[a, a]
        new[dead] = np.inf
        M[a, :] = new
        M[a, a] = keep
        if variant == 1:
            M[:, a] = new
            M[a, a] = keep
    # the loop stopped rightly: nothing at or below the threshold is left
    assert not (np.min(M) <= 0.0)
    # and the host's live matrix is the device's final matrix
    al = ~dead
    assert np.array_equal(al, alive_end)
    assert np.array_equal(M[np.ix_(al, al)], Mend[np.ix_(al, al)])
    write_audit('c3_full_argmin_replay_v%d' % variant,
                dict(segments=n, merges=len(merges), every_merge_is_numpy_argmin=True,
                     runner_up_gap_min_abs_sampled=float(min(gaps)), merge_distance_min=float(merges['d'].min()),
                     merge_distance_max=float(merges['d'].max())))
