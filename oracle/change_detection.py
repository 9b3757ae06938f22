"""Oracle for ``spk-change-detection.py`` (CD) - speaker-turn search.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, for Python 3, the recipe/feature I/O (CD:11-69), the three search
drivers ``dist_gw`` (CD:180-288), ``dist_sw`` (CD:291-357), ``merge_rec`` +
merge branch (CD:136-177, 375-394), the dispatcher ``detect_changes``
(CD:360-395) and the command line with its stdout text (CD:398-579).  The
module-level globals of the script become attributes of ``ChangeDetection``.

Two reference defects need a ruling; both behaviours are
implemented and selected by constructor arguments:

* ``sw_bic='strict'`` keeps CD:309 literally (``arr = features[start:end]``
  with ``end == 0``), which makes ``-m sw -d BIC`` raise ``ValueError`` from
  ``scipy.linalg.det`` exactly as the reference does; ``'intent'`` (default)
  uses the pooled double window ``features[start:start+2*winsize]``.
* ``bic_cache='reference'`` (default) keeps the shared mutable-default memo of
  CD:72 for the callers that do not pass their own (sw, merge);
  ``'correct'`` recomputes the left term on every call.
"""

import argparse
import os.path as op
import re
import sys

import numpy as np

from . import distances as D
from .py2compat import MAXINT, py2_print_str, py2_str


def parse_recipe(rfile, log):
    """CD:11-28.  ``log`` receives the lines the reference prints."""
    r = []
    audio_file = re.compile(r'audio=(\S+)')
    lna_name = re.compile(r'lna=(\S+)')
    start_time = re.compile(r'start-time=(\d+.\d+)')
    end_time = re.compile(r'end-time=(\d+.\d+)')
    for line in rfile:
        try:
            audio = audio_file.search(line).groups()[0]
            lna = lna_name.search(line).groups()[0]
            start = float(start_time.search(line).groups()[0])
            end = float(end_time.search(line).groups()[0])
            r.append((audio, lna, start, end))
        except AttributeError:
            log('Recipe line without recognizable data:')
            log(line)
    return r


def load_features(recline, fpath, ext):
    """CD:31-43: int32 ``dim`` header, float32 frames, frame-major."""
    name = op.join(fpath, op.splitext(op.basename(recline[0]))[0] + ext)
    with open(name, 'rb') as f:
        dim = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        feats = np.fromfile(f, dtype=np.float32)
    return dim, feats.reshape(feats.size // dim, dim)


def _fslice(a, lo, hi):
    """``a[lo:hi]`` with FLOAT bounds as numpy < 1.12 evaluated it: truncation
    toward zero (CD:144-145, 305-306, 309)."""
    return a[int(lo):int(hi)]


class ChangeDetection(object):
    """State + drivers of one CD process."""

    def __init__(self, rate, method='sw', distance='GLR', winsize=5.0,
                 winstep=0.5, deltaws=0.05, threshold=0.0, lambdac=1.3,
                 tt=False, dlr=False, segpath=None, feapath='.', feaext='.fea',
                 sw_bic='intent', bic_cache='reference', log=None, trace=None):
        self.rate = float(rate)                                  # CD:499
        self.method = method
        self.distance = distance
        self.deltaws = np.floor(self.rate * deltaws)             # CD:508
        self.winsize = np.floor(winsize * self.rate)             # CD:526
        self.winstep = np.floor(winstep * self.rate)             # CD:527
        self.threshold = threshold
        self.lambdac = lambdac
        self.tt = tt
        self.dlr = dlr
        self.segpath = segpath
        self.feapath = feapath
        self.feaext = feaext
        self.sw_bic = sw_bic
        self.bic_cache = bic_cache
        self.log = log if log is not None else (lambda *a: None)
        self.trace = trace          # optional list receiving per-window records
        self.shared_memo = D.BicMemo()                           # CD:72 default
        self.lna_letter = 'a'                                    # CD:534
        self.lna_count = 0                                       # CD:535
        # CD:542-549
        self.total_dist = 0
        self.max_dist = 0
        self.min_dist = MAXINT
        self.total_windows = 0
        self.total_det_dist = 0
        self.max_det_dist = 0
        self.min_det_dist = MAXINT
        self.total_segments = 0
        self.prev = None                                         # merge_rec.prev

    # ---- output ---------------------------------------------------------
    def write_recipe_line(self, recline, start, end, lna_start, outf, segf=None):
        """CD:46-69 (LNA renaming: SURVEY.md Q11)."""
        lna = recline[1]
        if not self.dlr:
            cut = lna.find('_')
            if lna[:cut] == self.lna_letter:
                self.lna_count += 1
            else:
                self.lna_count = 1
                self.lna_letter = lna[:cut]
            lna = lna[:cut + 1] + str(self.lna_count)
        t0 = py2_str(start / self.rate + lna_start)
        t1 = py2_str(end / self.rate + lna_start)
        outf.write('audio=' + recline[0] + ' lna=' + lna + ' start-time=' + t0 +
                   ' end-time=' + t1 + ' speaker=spk_turn\n')
        if self.segpath and segf is not None:
            segf.write('audio=' + recline[0] + ' alignment=' + self.segpath +
                       lna + '.seg' + ' lna=' + lna + ' start-time=' + t0 +
                       ' end-time=' + t1 + ' speaker=spk_turn\n')

    # ---- distances ------------------------------------------------------
    def _dist3(self, arr1, arr2, arr, i, memo):
        """gw call shape CD:207-211: BIC gets (arr, i, memo)."""
        if self.distance == 'BIC':
            return D.bic_cd(arr1, arr2, arr, self.lambdac, i, memo)
        if self.distance == 'GLR':
            return D.glr(arr1, arr2)
        return D.kl2(arr1, arr2)

    def _dist_shared(self, arr1, arr2, arr_fn):
        """sw / merge call shape (CD:147-151, 308-312): BIC without i/memo."""
        if self.distance == 'BIC':
            memo = self.shared_memo if self.bic_cache == 'reference' else {}
            return D.bic_cd(arr1, arr2, arr_fn(), self.lambdac, 0, memo)
        if self.distance == 'GLR':
            return D.glr(arr1, arr2)
        return D.kl2(arr1, arr2)

    def _window_stats(self, d):
        """CD:155-161 == CD:223-229 == CD:317-323."""
        if d != np.inf and d != -np.inf:
            self.total_dist += d
            self.total_windows += 1
            if d > self.max_dist:
                self.max_dist = d
            if d < self.min_dist:
                self.min_dist = d

    def _det_stats(self, d):
        """CD:167-172 == CD:257-262 == CD:328-333."""
        self.total_det_dist += d
        self.total_segments += 1
        if d > self.max_det_dist:
            self.max_det_dist = d
        if d < self.min_det_dist:
            self.min_det_dist = d

    # ---- growing window -------------------------------------------------
    def dist_gw(self, features, recline, outf, segf=None):
        """CD:180-288 (control flow transcribed in SURVEY.md appendix A.1)."""
        rate = self.rate
        n = features.shape[0]
        lna_start = recline[2]
        lna_end = recline[3]
        start = 0
        end = start + self.winsize * 2
        minfeas = rate / 2
        istep = rate / 10
        ws = minfeas
        dws = self.deltaws
        memo = D.BicMemo()
        while end <= n:
            i = minfeas
            maxd = -MAXINT - 1
            second = -MAXINT - 1        # runner-up, for the margin audit only
            maxi = None
            while i < end - start - minfeas:
                arr1 = features[int(start):int(start + i)]
                arr2 = features[int(start + i):int(end)]
                arr = features[int(start):int(end)]
                d = self._dist3(arr1, arr2, arr, i, memo)
                if self.tt:
                    self.log(py2_print_str('Time:', start / rate + i / rate + lna_start,
                                           '- Distance:', d))
                if d > maxd and d != np.inf:
                    second = maxd
                    maxd = d
                    maxi = i
                elif d == np.inf or d == -np.inf:
                    self.log(py2_print_str('Inf:', arr1.shape, arr2.shape, d))
                elif d > second:
                    second = d
                i += istep
            self._window_stats(maxd)
            rec = None
            if self.trace is not None:
                # gap: winner minus runner-up of the coarse scan (margin audit of the tests;
                # not part of the reference's flow)
                rec = dict(line=recline, start=start, end=end, maxi=maxi,
                           maxd=float(maxd), positive=False,
                           gap=float(maxd) - float(second))
                self.trace.append(rec)
            if maxd > self.threshold and maxd != np.inf and maxd != -np.inf:
                i = maxi - istep
                endtune = maxi + istep
                while i < endtune:
                    arr1 = features[int(start):int(start + i)]
                    arr2 = features[int(start + i):int(end)]
                    arr = features[int(start):int(end)]
                    d = self._dist3(arr1, arr2, arr, i, memo)
                    if d > maxd and d != np.inf:
                        second = maxd
                        maxd = d
                        maxi = i
                    elif d == np.inf or d == -np.inf:
                        self.log(py2_print_str('Inf:', arr1.shape, arr2.shape, d))
                    elif d > second and i != maxi:
                        second = d
                    i += 1
                if rec is not None:
                    rec.update(positive=True, maxi_fine=maxi, maxd_fine=float(maxd),
                               gap_fine=float(maxd) - float(second))
                self.write_recipe_line(recline, start, start + maxi, lna_start,
                                       outf, segf)
                memo = D.BicMemo()
                self._det_stats(maxd)
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
        end = (lna_end - lna_start) * rate
        self.write_recipe_line(recline, start, end, lna_start, outf, segf)

    # ---- sliding window -------------------------------------------------
    def dist_sw(self, features, recline, outf, segf=None):
        """CD:291-357 (appendix A.2; Q1, Q2, Q13)."""
        rate = self.rate
        W = self.winsize
        step = self.winstep
        n = features.shape[0]
        lna_start = recline[2]
        lna_end = recline[3]
        start = 0
        end = 0
        bestd = -1
        best_position = -1
        last_positive = -1
        while start + 2 * W <= n:
            arr1 = _fslice(features, start, start + W)
            arr2 = _fslice(features, start + W, start + 2 * W)
            if self.sw_bic == 'strict':
                arr_fn = (lambda s=start, e=end: _fslice(features, s, e))   # CD:309
            else:
                arr_fn = (lambda s=start: _fslice(features, s, s + 2 * W))
            d = self._dist_shared(arr1, arr2, arr_fn)
            if self.tt:
                self.log(py2_print_str('Time:', (start + W) / rate + lna_start,
                                       '- Distance:', d))
            if self.trace is not None:
                self.trace.append(dict(line=recline, start=float(start), d=float(d)))
            self._window_stats(d)
            if d < self.threshold or d == np.inf or d == -np.inf:
                if start - step == last_positive:
                    self.write_recipe_line(recline, end, best_position, lna_start,
                                           outf, segf)
                    self._det_stats(bestd)
                    bestd = 0
                    end = best_position
            else:
                if d > bestd:
                    bestd = d
                    best_position = start + W
                last_positive = start
            start += step
        if start - step == last_positive:
            self.write_recipe_line(recline, end, best_position, lna_start, outf, segf)
            self._det_stats(bestd)
            bestd = 0
            end = best_position
        this_end = (lna_end - lna_start) * rate
        self.write_recipe_line(recline, end, this_end, lna_start, outf, segf)

    # ---- merge mode -----------------------------------------------------
    def merge_rec(self, features, recline2, outf, segf=None):
        """CD:136-177.  (The reference's ``recline1`` argument is unused.)"""
        rate = self.rate
        start1 = self.prev[2] * rate
        start2 = recline2[2] * rate
        end1 = self.prev[3] * rate
        end2 = recline2[3] * rate
        arr1 = _fslice(features, start1, end1)
        arr2 = _fslice(features, start2, end2)
        d = self._dist_shared(arr1, arr2, lambda: np.concatenate((arr1, arr2)))
        if self.tt:
            self.log(py2_print_str('Time:', end1, '- Distance:', d))
        if self.trace is not None:
            self.trace.append(dict(prev=self.prev, next=recline2, d=float(d)))
        self._window_stats(d)
        if d < self.threshold and d != np.inf and d != -np.inf:
            self.prev = (self.prev[0], self.prev[1], self.prev[2], recline2[3])
            self._det_stats(d)
        else:
            self.write_recipe_line(self.prev, self.prev[2] * rate,
                                   self.prev[3] * rate, 0, outf, segf)
            self.prev = (recline2[0], recline2[1], recline2[2], recline2[3])

    # ---- dispatcher -----------------------------------------------------
    def detect_changes(self, recipe, outf, segf=None, loader=None):
        """CD:360-395.  ``loader(recline)`` -> (dim, features); defaults to the
        feacat file reader."""
        if loader is None:
            loader = lambda rl: load_features(rl, self.feapath, self.feaext)  # noqa: E731
        rate = self.rate
        this_wav = ''
        this_lna = ''
        l = 0
        wav_start = True
        feas = None
        while l < len(recipe):
            if recipe[l][0] != this_wav:
                this_wav = recipe[l][0]
                feas = loader(recipe[l])
            if self.method != 'm':
                if recipe[l][1] != this_lna:
                    this_lna = recipe[l][1]
                    seg = feas[1][int(recipe[l][2] * rate):int(recipe[l][3] * rate)]
                    if self.method == 'gw':
                        self.dist_gw(seg, recipe[l], outf, segf)
                    else:
                        self.dist_sw(seg, recipe[l], outf, segf)
            else:
                if l + 1 < len(recipe):
                    if recipe[l + 1][0] != this_wav:
                        l += 1
                        wav_start = True
                        continue
                    if wav_start:
                        wav_start = False
                        self.prev = recipe[l]
                    self.merge_rec(feas[1], recipe[l + 1], outf, segf)
                else:
                    self.write_recipe_line(self.prev, self.prev[2] * rate,
                                           self.prev[3] * rate, 0, outf, segf)
            l += 1

    # ---- end-of-run text ------------------------------------------------
    def summary(self, nrecipe):
        """CD:563-579."""
        log = self.log
        log('Useful metrics for determining the right threshold:')
        log('---------------------------------------------------')
        if self.total_windows > 0:
            log(py2_print_str('Average between windows distance:',
                              float(self.total_dist) / self.total_windows))
        log(py2_print_str('Maximum between windows distance:', self.max_dist))
        if self.min_dist < MAXINT:
            log(py2_print_str('Minimum between windows distance:', self.min_dist))
        log(py2_print_str('Total windows:', self.total_windows))
        log(py2_print_str('Total segments:', self.total_segments + nrecipe))
        if self.total_segments > 0:
            log(py2_print_str('Average between detected segments distance:',
                              float(self.total_det_dist) / self.total_segments))
        log(py2_print_str('Maximum between detected segments distance:',
                          self.max_det_dist))
        if self.min_det_dist < MAXINT:
            log(py2_print_str('Minimum between detected segments distance:',
                              self.min_det_dist))
        log(py2_print_str('Total detected speaker changes:', self.total_segments))


def build_parser():
    """The flags of CD:399-465 plus the two oracle-only rulings."""
    p = argparse.ArgumentParser(description='Perform speaker turn segmentation, '
                                'using a distance measure (CPU oracle).')
    p.add_argument('recfile', type=str)
    p.add_argument('feapath', type=str)
    p.add_argument('-seg', dest='segpath', type=str, default=None)
    p.add_argument('-o', dest='outfile', type=str, default='stdout')
    p.add_argument('-fe', dest='feaext', type=str, default='.fea')
    p.add_argument('-se', dest='segext', type=str, default='.seg')
    p.add_argument('-f', dest='frame_rate', type=int, default=125)
    p.add_argument('-m', dest='method', type=str, choices=['sw', 'gw', 'm'],
                   default='sw')
    p.add_argument('-d', dest='distance', type=str,
                   choices=['GLR', 'BIC', 'KL2'], default='GLR')
    p.add_argument('-w', dest='winsize', type=float, default=5.0)
    p.add_argument('-st', dest='winstep', type=float, default=0.5)
    p.add_argument('-dws', dest='deltaws', type=float, default=0.05)
    p.add_argument('-t', dest='threshold', type=float, default=0.0)
    p.add_argument('-l', dest='lambdac', type=float, default=1.3)
    p.add_argument('-tt', action='store_true')
    p.add_argument('-dlr', action='store_true')
    p.add_argument('--sw-bic', dest='sw_bic', choices=['intent', 'strict'],
                   default='intent')
    p.add_argument('--bic-cache', dest='bic_cache',
                   choices=['reference', 'correct'], default='reference')
    return p


def main(argv=None, stdout=None, trace=None):
    """CD:398-579 - returns the ``ChangeDetection`` object after the run."""
    out = stdout if stdout is not None else sys.stdout
    args = build_parser().parse_args(argv)

    def log(*items):
        out.write(py2_print_str(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as recfile:
        recipe = parse_recipe(recfile, log)
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
    cd = ChangeDetection(args.frame_rate, args.method, args.distance,
                         args.winsize, args.winstep, args.deltaws,
                         args.threshold, args.lambdac, args.tt, args.dlr,
                         args.segpath, args.feapath, args.feaext,
                         args.sw_bic, args.bic_cache, log, trace)
    log('Conversion rate set to frame rate:', cd.rate)
    if args.method == 'sw':
        log('Using a fixed-size sliding window')
    elif args.method == 'gw':
        log('Using a growing window')
        log('Deltaws set to:', cd.deltaws / cd.rate, 'seconds')
    else:
        log('Performing similar-segment merge')
    if args.distance == 'GLR':
        log('Using GLR as distance measure')
    elif args.distance == 'BIC':
        log('Using BIC as distance measure, lambda =', args.lambdac)
    else:
        log('Using KL2 as distance measure')
    if args.method != 'm':
        log('Window size set to:', cd.winsize / cd.rate, 'seconds')
        log('Window step set to:', cd.winstep / cd.rate, 'seconds')
    log('Threshold distance:', args.threshold)
    if args.dlr:
        log('Disabling LNA renaming')

    if args.outfile != 'stdout':
        with open(args.outfile, 'w') as outf:
            if segfile:
                with open(segfile, 'w') as segf:
                    cd.detect_changes(recipe, outf, segf)
            else:
                cd.detect_changes(recipe, outf)
    else:
        cd.detect_changes(recipe, out)
    cd.summary(len(recipe))
    return cd


if __name__ == '__main__':
    main()
