"""Speaker-turn change detection: the host side of ``spk-change-detection.py``.

Same command line, same recipe in / recipe out, same stdout text as the
reference script; the numeric work (covariances, log-determinants, the
growing-window search itself) runs on the GPU through ``_abi``.

Division of labour
  device  frame statistics, every distance, the whole sequential growing-window
          loop (``dist_gw``, spk-change-detection.py:180-288) as one persistent
          kernel that returns one record per visited window;
  host    recipe parsing, chain set-up (one chain per distinct consecutive lna,
          spk-change-detection.py:370-374), replay of the window records into
          recipe lines and the run statistics, the sliding-window peak picking
          (spk-change-detection.py:324-354) over the device-scored distances,
          the merge-mode chain (136-177, 375-394), Python-2 text formatting.

Two reference defects are ruled on here exactly as SURVEY.md (Q1, Q2) asks:
``--sw-bic strict`` reproduces the ``ValueError`` of ``-m sw -d BIC``; the
default ``intent`` pools the double window.  ``--bic-cache reference`` (default)
replays the process-wide mutable-default memo of ``bic`` for the callers that
rely on it (sw, merge); ``correct`` scores every window on its own.
"""

import argparse
import os.path as op
import sys

import numpy as np

from . import _abi
from .feacat import feature_file_name, read_features
from .py2fmt import MAXINT, p2line
from .recipe import Writer, parse

NEG_INIT = -MAXINT - 1          # maxd before any candidate, CD:203 (a Python int there too)


def _is_inf(d):
    return d == np.inf or d == -np.inf


class RunStats(object):
    """The module-level counters of the script, CD:542-549."""

    def __init__(self):
        self.total_dist = 0
        self.max_dist = 0
        self.min_dist = MAXINT
        self.total_windows = 0
        self.total_det_dist = 0
        self.max_det_dist = 0
        self.min_det_dist = MAXINT
        self.total_segments = 0

    def window(self, d):                       # CD:155-161, 223-229, 317-323
        if not _is_inf(d):
            self.total_dist += d
            self.total_windows += 1
            if d > self.max_dist:
                self.max_dist = d
            if d < self.min_dist:
                self.min_dist = d

    def detected(self, d):                     # CD:167-172, 257-262, 328-333
        self.total_det_dist += d
        self.total_segments += 1
        if d > self.max_det_dist:
            self.max_det_dist = d
        if d < self.min_det_dist:
            self.min_det_dist = d


def bic_from_terms(n1, n2, ld_left, ld_right, ld_pooled, lambdac, c1=None, p=_abi.DIM):
    """Delta-BIC from log-determinants in the reference's operation order
    (CD:89, 95-99); ``c1`` overrides the left term (the memo of CD:84-90)."""
    n1 = np.float64(n1)
    n2 = np.float64(n2)
    n = n1 + n2
    if c1 is None:
        c1 = 0.5 * n1 * ld_left
    d = 0.5 * n * ld_pooled - c1 - 0.5 * n2 * ld_right
    d -= lambdac * 0.5 * (p + 0.5 * p * (p + 1)) * np.log(n)
    return d, c1


class Detector(object):
    """One run of the change detector over a parsed recipe."""

    def __init__(self, frame_rate=125, method='sw', distance='GLR', winsize=5.0, winstep=0.5,
                 deltaws=0.05, threshold=0.0, lambdac=1.3, tt=False, dlr=False, segpath=None,
                 feapath='.', feaext='.fea', sw_bic='intent', bic_cache='reference',
                 device=0, ctx=None, log=None, gw_on_device=True):
        self.rate = float(frame_rate)                              # CD:499
        self.method = method
        self.distance = distance
        self.metric = _abi.METRIC[distance]
        self.deltaws = np.floor(self.rate * deltaws)               # CD:508
        self.winsize = np.floor(winsize * self.rate)               # CD:526
        self.winstep = np.floor(winstep * self.rate)               # CD:527
        if method in ('sw', 'gw') and not self.winstep >= 1:
            # `start += winstep` (CD:341) / `end += ws` with ws clamped to winstep (CD:273-280) stop advancing:
            # the reference spins until it is interrupted; a device-resident search cannot be, so this is refused
            raise ValueError('window step of %r s is below one frame at %r fps: the search would never end'
                             % (winstep, self.rate))
        self.threshold = threshold
        self.lambdac = lambdac
        self.tt = tt
        self.segpath = segpath
        self.feapath = feapath
        self.feaext = feaext
        self.sw_bic = sw_bic
        self.bic_cache = bic_cache
        self.log = log if log is not None else (lambda *a: None)
        self.stats = RunStats()
        self.writer = Writer(self.rate, rename=not dlr, segprefix=segpath)
        self.memo_c1 = None            # the shared `saved[0]` of CD:72
        self.prev = None               # merge_rec.prev
        self.gw_on_device = gw_on_device
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else _abi.Context(device)
        self.windows_visited = 0
        self.dim = _abi.DIM            # of the feature file in use (set when it is loaded)
        self._merge_plan = {}          # id(next line) -> step of the device-resident merge chain
        self.merge_on_device = gw_on_device
        self._prefetch = {}

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    # ---- features ------------------------------------------------------------
    def load(self, line):
        dim, frames = read_features(feature_file_name(line.audio, self.feapath, self.feaext))
        return self.ctx.upload(frames)

    def _bounds(self, line, nframes):
        """``feas[int(s*rate):int(e*rate)]`` (CD:373) as a clamped frame range."""
        a = min(max(int(line.start * self.rate), 0), nframes)
        b = min(max(int(line.end * self.rate), 0), nframes)
        return a, max(a, b)

    # ---- growing window ------------------------------------------------------
    def gw_chains(self, recipe, loader):
        """The chains ``detect_changes`` will hand to the device, without running them:
        [(feat, [(a, b), ...])] per run of lines of one wav (the walk of CD:360-374).  The
        corpus driver uses it to put the chains of a whole batch of recordings into one launch
        and ``prefetch`` the records."""
        groups = []
        this_wav = this_lna = ''
        feat = None
        chains = []
        for line in recipe:
            if line.audio != this_wav:
                if chains:
                    groups.append((feat, chains))
                    chains = []
                this_wav = line.audio
                feat = loader(line)
            if line.lna != this_lna:
                this_lna = line.lna
                chains.append(self._bounds(line, feat.n))
        if chains:
            groups.append((feat, chains))
        return groups

    def prefetch(self, feat, chains, result):
        """``result`` = what ``feat.gw_run`` would return for these chains."""
        self._prefetch[(id(feat), tuple(c[0] for c in chains), tuple(c[1] for c in chains))] = result

    def _gw_device(self, feat, chains, outf, segf):
        """All chains of one wav in one persistent-kernel launch, then replay."""
        if not chains:
            return
        seg_a = [c[1] for c in chains]
        seg_b = [c[2] for c in chains]
        got = self._prefetch.pop((id(feat), tuple(seg_a), tuple(seg_b)), None)     # detect_changes_multi
        if got is None:
            got = feat.gw_run(seg_a, seg_b, self.rate, self.winsize, self.winstep, self.deltaws,
                              self.threshold, self.lambdac, self.metric)
        win, first = got
        for k, (line, a, b) in enumerate(chains):
            self._gw_replay(feat, line, a, b, win[first[k]:first[k + 1]], outf, segf)

    def _gw_replay(self, feat, line, a, b, recs, outf, segf):
        st = self.stats
        start = 0
        # plain Python lists: item access on a structured array costs microseconds per field
        ninf, ncand, pos = recs['ninf'].tolist(), recs['ncand'].tolist(), recs['positive'].tolist()
        maxds, maxd_f, maxi_f, starts = (recs['maxd'].tolist(), recs['maxd_fine'].tolist(),
                                         recs['maxi_fine'].tolist(), recs['start'].tolist())
        self.windows_visited += len(ninf)
        inf = np.inf
        # RunStats.window (CD:223-229) inlined: this loop runs once per window of the corpus
        tdist, twin, dmax, dmin = st.total_dist, st.total_windows, st.max_dist, st.min_dist
        for i in range(len(ninf)):
            if ninf[i] > 0:
                self._gw_inf_lines(feat, a, recs[i])
            d = maxds[i] if ncand[i] >= 0 else NEG_INIT
            if d != inf and d != -inf:
                tdist += d
                twin += 1
                if d > dmax:
                    dmax = d
                if d < dmin:
                    dmin = d
            if pos[i]:
                maxi = maxi_f[i]
                start = starts[i]
                self.writer.write(line, start, start + maxi, line.start, 'spk_turn', outf, segf)
                st.detected(maxd_f[i])
                start += maxi
        st.total_dist, st.total_windows, st.max_dist, st.min_dist = tdist, twin, dmax, dmin
        end = (line.end - line.start) * self.rate                 # CD:287
        self.writer.write(line, start, end, line.start, 'spk_turn', outf, segf)

    def _gw_offsets(self, start, end):
        minfeas = self.rate / 2
        istep = self.rate / 10
        out = []
        i = minfeas
        while i < end - start - minfeas:
            out.append(i)
            i += istep
        return out

    def _score_offsets(self, feat, a, start, end, offs):
        s0 = a + int(start)
        e0 = a + int(end)
        m = [a + int(start + i) for i in offs]
        return feat.score_windows([s0] * len(m), m, [e0] * len(m), self.metric, self.lambdac), m, s0, e0

    def _gw_inf_lines(self, feat, a, r):
        """The ``Inf:`` lines of CD:219-220 / 249-250 for a window whose record
        reports infinite distances (rare; re-scored so that the text carries the
        shapes and values)."""
        start, end = float(r['start']), float(r['end'])
        offs = self._gw_offsets(start, end)
        passes = [offs]
        if r['positive']:
            istep = self.rate / 10
            i = float(r['maxi']) - istep
            fine = []
            while i < float(r['maxi']) + istep:
                fine.append(i)
                i += 1
            passes.append(fine)
        for off in passes:
            if not off:
                continue
            d, m, s0, e0 = self._score_offsets(feat, a, start, end, off)
            for dk, mk in zip(d, m):
                if _is_inf(dk):
                    self.log(p2line('Inf:', (mk - s0, feat.dim), (e0 - mk, feat.dim), dk))

    def _gw_host_driven(self, feat, line, a, b, outf, segf):
        """The same search with the LOOP on the host and every distance on the
        device (one batched scoring call per window).  Used for ``-tt``, which
        needs every candidate distance in order, and as a cross-check of the
        persistent kernel in the tests.  Control flow: CD:180-288."""
        st = self.stats
        rate = self.rate
        n = b - a
        minfeas = rate / 2
        istep = rate / 10
        start = 0
        end = start + self.winsize * 2
        ws = minfeas
        dws = self.deltaws
        while end <= n:
            self.windows_visited += 1
            offs = self._gw_offsets(start, end)
            maxd = NEG_INIT
            maxi = None
            if offs:
                d, m, s0, e0 = self._score_offsets(feat, a, start, end, offs)
                for i, dk, mk in zip(offs, d, m):
                    if self.tt:
                        self.log(p2line('Time:', start / rate + i / rate + line.start,
                                        '- Distance:', dk))
                    if dk > maxd and dk != np.inf:
                        maxd = dk
                        maxi = i
                    elif _is_inf(dk):
                        self.log(p2line('Inf:', (mk - s0, feat.dim), (e0 - mk, feat.dim), dk))
            st.window(maxd)
            if maxd > self.threshold and not _is_inf(maxd):
                fine = []
                i = maxi - istep
                while i < maxi + istep:
                    fine.append(i)
                    i += 1
                d, m, s0, e0 = self._score_offsets(feat, a, start, end, fine)
                for i, dk, mk in zip(fine, d, m):
                    if dk > maxd and dk != np.inf:
                        maxd = dk
                        maxi = i
                    elif _is_inf(dk):
                        self.log(p2line('Inf:', (mk - s0, feat.dim), (e0 - mk, feat.dim), dk))
                self.writer.write(line, start, start + maxi, line.start, 'spk_turn', outf, segf)
                st.detected(maxd)
                start += maxi
                if start + self.winsize * 2 <= n:
                    end = start + self.winsize * 2
                    ws = minfeas
                    dws = self.deltaws
                else:
                    break
            else:
                if end + ws <= n:
                    end += ws
                    if ws < self.winstep:
                        ws += dws
                        dws *= 2
                    if ws > self.winstep:
                        ws = self.winstep
                elif end != n:
                    end = n
                else:
                    break
        self.writer.write(line, start, (line.end - line.start) * rate, line.start, 'spk_turn',
                          outf, segf)

    # ---- sliding window ------------------------------------------------------
    def _sw(self, feat, line, a, b, outf, segf):
        """CD:291-357: every window is scored in one device call, the
        positive-run peak picking is the reference's scan over those distances."""
        st = self.stats
        rate = self.rate
        W = self.winsize
        step = self.winstep
        n = b - a
        starts = []
        start = 0
        while start + 2 * W <= n:
            starts.append(start)
            start += step
        dist = []
        if starts:
            if self.metric == _abi.BIC and self.sw_bic == 'strict':
                # CD:309: arr = features[start:end] with end == 0 -> empty covariance ->
                # scipy.linalg.det raises; the reference dies here
                raise ValueError('array must not contain infs or NaNs')
            wa = [a + int(s) for s in starts]
            wm = [a + int(s + W) for s in starts]
            wb = [a + int(s + 2 * W) for s in starts]
            if self.metric == _abi.BIC:
                _, terms = feat.score_windows(wa, wm, wb, self.metric, self.lambdac, terms=True)
                for k in range(len(starts)):
                    dist.append(self._bic_memo(wm[k] - wa[k], wb[k] - wm[k], terms[k]))
            else:
                dist = list(feat.score_windows(wa, wm, wb, self.metric, self.lambdac))
        end = 0
        bestd = -1
        best_position = -1
        last_positive = -1
        for start, d in zip(starts, dist):
            self.windows_visited += 1
            if self.tt:
                self.log(p2line('Time:', (start + W) / rate + line.start, '- Distance:', d))
            st.window(d)
            if d < self.threshold or _is_inf(d):
                if start - step == last_positive:
                    self.writer.write(line, end, best_position, line.start, 'spk_turn', outf, segf)
                    st.detected(bestd)
                    bestd = 0
                    end = best_position
            else:
                if d > bestd:
                    bestd = d
                    best_position = start + W
                last_positive = start
        start = starts[-1] + step if starts else 0
        if start - step == last_positive:
            self.writer.write(line, end, best_position, line.start, 'spk_turn', outf, segf)
            st.detected(bestd)
            bestd = 0
            end = best_position
        self.writer.write(line, end, (line.end - line.start) * rate, line.start, 'spk_turn',
                          outf, segf)

    def _bic_memo(self, n1, n2, terms):
        """BIC through the shared memo of CD:72 (Q2) or, with ``--bic-cache
        correct``, on the window's own left term."""
        if self.bic_cache == 'reference':
            d, c1 = bic_from_terms(n1, n2, terms[0], terms[1], terms[2], self.lambdac, self.memo_c1, p=self.dim)
            if self.memo_c1 is None:
                self.memo_c1 = c1
            return d
        return bic_from_terms(n1, n2, terms[0], terms[1], terms[2], self.lambdac, p=self.dim)[0]

    # ---- merge mode ----------------------------------------------------------
    def _plan_merge(self, feat, lines):
        """The whole chain over the lines of one wav in ONE launch (``spkdiar_merge_chain``): per step the three
        ln|S| terms, the distance and the decision.  ``_merge_step`` then replays the script's bookkeeping from
        them and checks at every step that it stands where the device chain stood (same two frame ranges, same
        decision); from a step where it does not - a tie between numpy's and the device's logarithm in the BIC
        penalty - it goes on with one scoring call per line.  BIC and GLR; KL2 keeps the call per line."""
        self._merge_plan.clear()
        if self.metric == _abi.KL2 or not self.merge_on_device or len(lines) < 2 or not hasattr(feat, 'merge_chain'):
            return
        rate, n = self.rate, feat.n
        a = [min(max(int(l.start * rate), 0), n) for l in lines]
        b = [min(max(int(l.end * rate), 0), n) for l in lines]
        terms, dist, merged, _ = feat.merge_chain(a, b, self.metric, self.lambdac, self.threshold,
                                                  self.bic_cache == 'reference', self.memo_c1)
        pa, pb = a[0], b[0]
        for k in range(len(lines) - 1):
            r1 = (pa, max(pa, pb))
            r2 = (a[k + 1], max(a[k + 1], b[k + 1]))
            self._merge_plan[id(lines[k + 1])] = ((r1, r2), terms[k], dist[k], bool(merged[k]))
            if merged[k]:
                pb = b[k + 1]
            else:
                pa, pb = a[k + 1], b[k + 1]

    def _merge_step(self, feat, nxt, outf, segf):
        """CD:136-177: previous (possibly already merged) segment against the
        next one; whole-file frame indices, float bounds truncated."""
        st = self.stats
        rate = self.rate
        prev = self.prev
        n = feat.n

        def clamp(v):
            return min(max(int(v), 0), n)
        r1 = (clamp(prev.start * rate), clamp(prev.end * rate))
        r2 = (clamp(nxt.start * rate), clamp(nxt.end * rate))
        r1 = (r1[0], max(r1))
        r2 = (r2[0], max(r2))
        plan = self._merge_plan.pop(id(nxt), None)
        if plan is not None and plan[0] != (r1, r2):
            plan = None                                  # the device chain stands elsewhere: its steps no longer apply
            self._merge_plan.clear()
        if plan is not None:
            # scored (and decided) by the device-resident chain, spkdiar_merge_chain
            if self.metric == _abi.BIC:
                d = self._bic_memo(r1[1] - r1[0], r2[1] - r2[0], plan[1])
            else:
                d = plan[2]
            if bool(d < self.threshold and not _is_inf(d)) != plan[3]:
                self._merge_plan.clear()                 # (a rounding tie between the host's and the device's log)
        elif self.metric == _abi.BIC:
            _, terms = feat.score_sets([[r1]], [[r2]], self.metric, self.lambdac, terms=True)
            d = self._bic_memo(r1[1] - r1[0], r2[1] - r2[0], terms[0])
        else:
            d = feat.score_sets([[r1]], [[r2]], self.metric, self.lambdac)[0]
        self.windows_visited += 1
        if self.tt:
            self.log(p2line('Time:', prev.end * rate, '- Distance:', d))
        st.window(d)
        if d < self.threshold and not _is_inf(d):
            self.prev = prev._replace(end=nxt.end)
            st.detected(d)
        else:
            self.writer.write(prev, prev.start * rate, prev.end * rate, 0, 'spk_turn', outf, segf)
            self.prev = nxt

    # ---- dispatcher ----------------------------------------------------------
    def detect_changes(self, recipe, outf, segf=None, loader=None):
        """CD:360-395.  ``loader(line)`` -> Features overrides the file reader."""
        load = loader if loader is not None else self.load
        this_wav = ''
        this_lna = ''
        feat = None
        owned = []
        chains = []                 # pending gw chains of the current wav

        def flush():
            if chains:
                self._gw_device(feat, chains, outf, segf)
                del chains[:]
        try:
            l = 0
            wav_start = True
            while l < len(recipe):
                line = recipe[l]
                if line.audio != this_wav:
                    flush()
                    # the reference holds one wav at a time (CD:367-369); a handle keeps 6.7 KB of statistics
                    # per frame on the device, so the previous wav is released before the next is read
                    while owned:
                        owned.pop().close()
                    this_wav = line.audio
                    feat = load(line)
                    self.dim = getattr(feat, 'dim', _abi.DIM)      # p of the BIC penalty (CD:97)
                    if loader is None:
                        owned.append(feat)
                if self.method != 'm':
                    if line.lna != this_lna:
                        this_lna = line.lna
                        a, b = self._bounds(line, feat.n)
                        if self.method == 'gw':
                            if self.tt or not self.gw_on_device:
                                self._gw_host_driven(feat, line, a, b, outf, segf)
                            else:
                                chains.append((line, a, b))
                        else:
                            self._sw(feat, line, a, b, outf, segf)
                else:
                    if l + 1 < len(recipe):
                        if recipe[l + 1].audio != this_wav:
                            l += 1
                            wav_start = True
                            continue
                        if wav_start:
                            wav_start = False
                            self.prev = line
                            m = l
                            while m < len(recipe) and recipe[m].audio == this_wav:
                                m += 1
                            self._plan_merge(feat, recipe[l:m])
                        self._merge_step(feat, recipe[l + 1], outf, segf)
                    else:
                        p = self.prev
                        self.writer.write(p, p.start * self.rate, p.end * self.rate, 0,
                                          'spk_turn', outf, segf)
                l += 1
            flush()
        finally:
            for f in owned:
                f.close()

    def summary(self, nrecipe):
        """CD:563-579."""
        st = self.stats
        log = self.log
        log('Useful metrics for determining the right threshold:')
        log('---------------------------------------------------')
        if st.total_windows > 0:
            log(p2line('Average between windows distance:', float(st.total_dist) / st.total_windows))
        log(p2line('Maximum between windows distance:', st.max_dist))
        if st.min_dist < MAXINT:
            log(p2line('Minimum between windows distance:', st.min_dist))
        log(p2line('Total windows:', st.total_windows))
        log(p2line('Total segments:', st.total_segments + nrecipe))
        if st.total_segments > 0:
            log(p2line('Average between detected segments distance:',
                       float(st.total_det_dist) / st.total_segments))
        log(p2line('Maximum between detected segments distance:', st.max_det_dist))
        if st.min_det_dist < MAXINT:
            log(p2line('Minimum between detected segments distance:', st.min_det_dist))
        log(p2line('Total detected speaker changes:', st.total_segments))


def detect_changes_multi(detectors, recipe, outfs, loader, after=None):
    """Several detectors (e.g. -d BIC, -d GLR and -d KL2 with otherwise the script's flags) over
    the SAME recipe and resident features: the growing-window searches of all of them run side
    by side on the GPU (``spkdiar_gw_multi_*``), and every detector replays its own records
    exactly as ``detect_changes`` does, in the order given.  Outputs are identical to calling
    ``detect_changes`` on each detector in turn.

    ``after``: ``{k: fn}`` - ``fn(detector, where)`` is called as soon as detector k has written
    its output, while the searches of the later detectors are still running on their own SMs;
    ``where`` = (stream, sms) the search of detector k ran on, or None (so that e.g. the
    clustering of its turns can be queued there: ``Context.exec_on``)."""
    after = after or {}
    dets = [d for d in detectors if d.method == 'gw' and not d.tt and d.gw_on_device]
    handles = []
    try:
        last = None
        if len(dets) >= 2 and len({d.rate for d in dets}) == 1:
            groups = dets[0].gw_chains(recipe, loader)
            runs = [dict(rate=d.rate, winsize=d.winsize, winstep=d.winstep, deltaws=d.deltaws,
                         threshold=d.threshold, lambdac=d.lambdac, metric=d.metric) for d in dets]
            for g, (feat, ch) in enumerate(groups):
                h = feat.gw_multi_begin([c[0] for c in ch], [c[1] for c in ch], runs)
                handles.append((feat, ch, h))
                if g + 1 < len(groups):              # one asynchronous object per context: finish this wav first
                    for k, d in enumerate(dets):
                        d.prefetch(feat, ch, h.wait(k))
                    h.close()
                else:
                    last = (feat, ch, h)
        for k, (d, outf) in enumerate(zip(detectors, outfs)):
            where = None
            if last is not None and d in dets:
                feat, ch, h = last
                d.prefetch(feat, ch, h.wait(dets.index(d)))
                where = h.where(dets.index(d))
            d.detect_changes(recipe, outf, loader=loader)
            if k in after:
                after[k](d, where)
    finally:
        for _, _, h in handles:
            h.close()


def build_parser():
    """The command line of spk-change-detection.py:399-465 (same flags, defaults
    and help intent) plus additive options."""
    p = argparse.ArgumentParser(description='Perform speaker turn segmentation, using a '
                                'distance measure.')
    p.add_argument('recfile', type=str, help='Specifies the input recipe file')
    p.add_argument('feapath', type=str, help='Specifies the features files path')
    p.add_argument('-seg', dest='segpath', type=str, default=None,
                   help='Specifies the alignment segmentation files path and generates '
                   '"alignment=" information, default empty (not generate)')
    p.add_argument('-o', dest='outfile', type=str, default='stdout',
                   help='Specifies an output file, default stdout. With "-seg" a second output '
                   'file is created with "-seg" appended to the name before the extension')
    p.add_argument('-fe', dest='feaext', type=str, default='.fea',
                   help='Specifies feature file extension, default ".fea"')
    p.add_argument('-se', dest='segext', type=str, default='.seg',
                   help='Specifies segmentation files extension, default ".seg"')
    p.add_argument('-f', dest='frame_rate', type=int, default=125,
                   help='Specifies the frame rate, default 125')
    p.add_argument('-m', dest='method', type=str, choices=['sw', 'gw', 'm'], default='sw',
                   help='Sliding window (sw, default), growing window (gw) or merge of '
                   'consecutive same-speaker turns (m)')
    p.add_argument('-d', dest='distance', type=str, choices=['GLR', 'BIC', 'KL2'], default='GLR',
                   help='Distance measure, default GLR')
    p.add_argument('-w', dest='winsize', type=float, default=5.0,
                   help='Window size in seconds (sliding) / minimum window (growing), default 5.0')
    p.add_argument('-st', dest='winstep', type=float, default=0.5,
                   help='Window step / maximum growth in seconds, default 0.5')
    p.add_argument('-dws', dest='deltaws', type=float, default=0.05,
                   help='Minimum growth for growing windows, default 0.05 seconds')
    p.add_argument('-t', dest='threshold', type=float, default=0.0,
                   help='Threshold distance for detection, default 0.0')
    p.add_argument('-l', dest='lambdac', type=float, default=1.3,
                   help='Lambda penalty weight for BIC, default 1.3')
    p.add_argument('-tt', action='store_true',
                   help='Output every decision distance, to choose a threshold')
    p.add_argument('-dlr', action='store_true', help='Disable lna renaming')
    # additive
    p.add_argument('--device', type=int, default=0, help='CUDA device ordinal (default 0)')
    p.add_argument('--sw-bic', dest='sw_bic', choices=['intent', 'strict'], default='intent',
                   help='-m sw -d BIC: "strict" reproduces the reference ValueError, "intent" '
                   '(default) pools the double window')
    p.add_argument('--bic-cache', dest='bic_cache', choices=['reference', 'correct'],
                   default='reference',
                   help='sw / merge BIC: replay the reference mutable-default memo (default) or not')
    return p


def main(argv=None, stdout=None, ctx=None):
    """spk-change-detection.py:398-579.  Returns the Detector after the run."""
    out = stdout if stdout is not None else sys.stdout
    args = build_parser().parse_args(argv)

    def log(*items):
        out.write(p2line(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as recfile:
        recipe = parse(recfile, log)
    log('Reading feature files from:', args.feapath)
    if args.segpath:
        log('Setting alignment segmentation files path to:', args.segpath)
        log('Segmentation files extension:', args.segext)
    log('Feature files extension:', args.feaext)
    segfile = False
    if args.outfile != 'stdout':
        log('Writing output to:', args.outfile)
        if args.segpath:
            segfile = op.splitext(op.basename(args.outfile))[0]
            segfile += '-seg' + op.splitext(args.outfile)[1]
            segfile = op.join(args.segpath, segfile)
            log('Writing seg output to:', segfile)
    else:
        log('Writing output to: stdout')
    det = Detector(args.frame_rate, args.method, args.distance, args.winsize, args.winstep,
                   args.deltaws, args.threshold, args.lambdac, args.tt, args.dlr, args.segpath,
                   args.feapath, args.feaext, args.sw_bic, args.bic_cache, args.device, ctx, log)
    try:
        log('Conversion rate set to frame rate:', det.rate)
        if args.method == 'sw':
            log('Using a fixed-size sliding window')
        elif args.method == 'gw':
            log('Using a growing window')
            log('Deltaws set to:', det.deltaws / det.rate, 'seconds')
        else:
            log('Performing similar-segment merge')
        if args.distance == 'GLR':
            log('Using GLR as distance measure')
        elif args.distance == 'BIC':
            log('Using BIC as distance measure, lambda =', args.lambdac)
        else:
            log('Using KL2 as distance measure')
        if args.method != 'm':
            log('Window size set to:', det.winsize / det.rate, 'seconds')
            log('Window step set to:', det.winstep / det.rate, 'seconds')
        log('Threshold distance:', args.threshold)
        if args.dlr:
            log('Disabling LNA renaming')
        if args.outfile != 'stdout':
            with open(args.outfile, 'w') as outf:
                if segfile:
                    with open(segfile, 'w') as segf:
                        det.detect_changes(recipe, outf, segf)
                else:
                    det.detect_changes(recipe, outf)
        else:
            det.detect_changes(recipe, out)
        det.summary(len(recipe))
    finally:
        det.close()
    return det


if __name__ == '__main__':
    main()
