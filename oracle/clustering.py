"""Oracle for ``spk-clustering.py`` (CL1, variant 1) and ``spk-clustering2.py``
(CL2, variant 2) - agglomerative / in-order speaker clustering.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates ``get_spk_features`` (CL1:46-52 / CL2:47-51), ``spk_cluster_in``
(CL1:136-175 / CL2:135-170), ``spk_cluster_hi`` (CL1:178-260 / CL2:173-229),
``process_recipe`` (CL1:263-292 / CL2:232-261), ``write_recipe_line``
(CL1:55-78) and the command lines (CL1:295-442 / CL2:264-406).  The two
scripts differ in result-affecting ways; ``variant``
selects which one is followed:

* variant 1: symmetric matrix, diagonal ``sys.maxint``, row AND column are
  rescored after a merge, ``-o`` default is the ``sys.stdout`` object so
  ``-o stdout`` writes a FILE named ``stdout``, ``feapath + '/'``.
* variant 2: ``inf``-filled matrix, only the upper triangle is filled and only
  row ``s1`` is rescored, so stale entries keep competing in ``argmin``.
"""

import argparse
import os.path as op
import sys

import numpy as np

from . import distances as D
from .change_detection import parse_recipe
from .py2compat import MAXINT, py2_print_str, py2_str


def load_features(recline, feapath, ext, variant):
    """CL1:31-43 (string concatenation) / CL2:32-44 (``op.join``)."""
    base = op.splitext(op.basename(recline[0]))[0] + ext
    name = feapath + base if variant == 1 else op.join(feapath, base)
    with open(name, 'rb') as f:
        dim = int(np.fromfile(f, dtype=np.int32, count=1)[0])
        feats = np.fromfile(f, dtype=np.float32)
    return dim, feats.reshape(feats.size // dim, dim)


def get_spk_features(spk, features):
    """CL1:46-52 / CL2:47-51: concatenated frame slices of one cluster; float
    bounds truncate (``int()`` in CL1, numpy<1.12 float indices in CL2)."""
    arr = features[int(spk[0][0]):int(spk[0][1])]
    for s in spk[1:]:
        arr = np.concatenate((arr, features[int(s[0]):int(s[1])]))
    return arr


class Clustering(object):
    """State + drivers of one CL1 / CL2 process."""

    def __init__(self, rate, variant=1, method='hi', distance='BIC',
                 threshold=0.0, max_spk=0, lambdac=1.3, tt=False, dlr=False,
                 segpath='', feapath='./', feaext='.fea', log=None, trace=None):
        self.rate = float(rate)
        self.variant = variant
        self.method = method
        self.distance = distance
        self.threshold = threshold
        self.max_spk = max_spk
        self.lambdac = lambdac
        self.tt = tt
        self.dlr = dlr
        self.segpath = segpath
        self.feapath = feapath
        self.feaext = feaext
        self.log = log if log is not None else (lambda *a: None)
        self.trace = trace          # optional list receiving (i, j, d) merges
        self.lna_letter = 'a'
        self.lna_count = 0
        self.max_dist = 0
        self.min_dist = MAXINT
        self.max_det_dist = 0
        self.min_det_dist = MAXINT
        self.speakers = []

    def dist(self, arr1, arr2):
        if self.distance == 'BIC':
            return D.bic_cl(arr1, arr2, self.lambdac)
        if self.distance == 'GLR':
            return D.glr(arr1, arr2)
        return D.kl2(arr1, arr2)

    def write_recipe_line(self, recline, start, end, lna_start, speaker, outf,
                          segf=None):
        """CL1:55-78 / CL2:54-77."""
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
        tail = ' speaker=speaker_' + str(speaker) + '\n'
        outf.write('audio=' + recline[0] + ' lna=' + lna + ' start-time=' + t0 +
                   ' end-time=' + t1 + tail)
        seg_on = (self.segpath != '') if self.variant == 1 else bool(self.segpath)
        if seg_on and segf is not None:
            segf.write('audio=' + recline[0] + ' alignment=' + self.segpath +
                       lna + '.seg' + ' lna=' + lna + ' start-time=' + t0 +
                       ' end-time=' + t1 + tail)

    def _range_stats(self, d):
        """CL1:196-200 == CL1:233-237 == CL1:152-156."""
        if d != np.inf and d != -np.inf:
            if d > self.max_dist:
                self.max_dist = d
            if d < self.min_dist:
                self.min_dist = d

    # ---- in-order -------------------------------------------------------
    def spk_cluster_in(self, features, recline, outf, segf=None):
        """CL1:136-175 / CL2:135-170."""
        speakers = self.speakers
        if self.variant == 1:
            start = int(recline[2] * self.rate)
            end = int(recline[3] * self.rate)
        else:
            start = recline[2] * self.rate
            end = recline[3] * self.rate
        arr2 = features[int(start):int(end)]
        mind = MAXINT
        best_candidate = None
        d = None
        spk = 0
        while spk < len(speakers):
            arr1 = get_spk_features(speakers[spk], features)
            d = self.dist(arr1, arr2)
            if self.tt:
                self.log(py2_print_str('Time:', end, '- Distance:', d,
                                       '- Speaker:', spk + 1))
            if d != np.inf and d != -np.inf:
                if d > self.max_dist:
                    self.max_dist = d
                if d < self.min_dist:
                    self.min_dist = d
                if d < mind:
                    mind = d
                    best_candidate = spk
            spk += 1
        if self.trace is not None:
            self.trace.append((best_candidate, float(mind)))
        if mind <= self.threshold:
            if self.variant == 1:                      # CL1:164-167 (uses d, not mind)
                if d > self.max_det_dist:
                    self.max_det_dist = d
                if d < self.min_det_dist:
                    self.min_det_dist = d
            speakers[best_candidate].append((start, end))
            self.write_recipe_line(recline, start, end, 0, best_candidate + 1,
                                   outf, segf)
        else:
            speakers.append([(start, end)])
            self.write_recipe_line(recline, start, end, 0, len(speakers), outf, segf)

    # ---- hierarchical ---------------------------------------------------
    def spk_cluster_hi(self, features, recipe, outf, segf=None):
        """CL1:178-260 (variant 1) / CL2:173-229 (variant 2); appendix A.4."""
        speakers = self.speakers
        sp = len(speakers)
        v1 = self.variant == 1
        if v1:
            distances = np.empty((sp, sp))
            np.fill_diagonal(distances, MAXINT)
        else:
            distances = np.full((sp, sp), np.inf)
        for s1 in range(sp):
            arr1 = get_spk_features(speakers[s1], features)
            for s2 in range(s1 + 1, sp):
                arr2 = get_spk_features(speakers[s2], features)
                d = self.dist(arr1, arr2)
                distances[s1][s2] = d
                if v1:
                    distances[s2][s1] = d
                    self._range_stats(d)
        while True:
            mind = distances.min()
            if mind <= self.threshold or (self.max_spk > 0
                                          and len(speakers) > self.max_spk):
                index = distances.argmin()
                a = index // len(speakers)
                b = index % len(speakers)
                if a > b:
                    a, b = b, a
                if v1:
                    if mind > self.max_det_dist:
                        self.max_det_dist = mind
                    if mind < self.min_det_dist:
                        self.min_det_dist = mind
                self.log(py2_print_str('Merging:', a + 1, 'and', b + 1,
                                       'distance:', mind))
                if self.trace is not None:
                    # 4th entry: distance of the runner-up pair minus the minimum (margin audit of
                    # the tests; not part of the reference's flow)
                    flat = distances.ravel().copy()
                    flat[index] = np.inf
                    if v1:
                        flat[b * len(speakers) + a] = np.inf   # the mirror entry of a symmetric matrix
                    self.trace.append((int(a), int(b), float(mind), float(np.nanmin(flat)) - float(mind)))
                speakers[a].extend(speakers[b])
                speakers.pop(b)
                distances = np.delete(distances, b, 0)
                distances = np.delete(distances, b, 1)
                arr1 = get_spk_features(speakers[a], features)
                for s2 in range(len(speakers)):
                    if s2 == a:
                        continue
                    arr2 = get_spk_features(speakers[s2], features)
                    d = self.dist(arr1, arr2)
                    distances[a][s2] = d
                    if v1:
                        distances[s2][a] = d
                        self._range_stats(d)
            else:
                if not v1:                                     # CL2:220-221
                    self.max_dist = distances.max()
                    self.min_dist = distances.min()
                break
        self.log(py2_print_str('Final speakers:', len(speakers)))
        if v1:
            # CL1:243-260: repeatedly emit the globally smallest turn tuple
            turns = [(turn, s) for s, ts in enumerate(speakers) for turn in ts]
            # ties between identical tuples resolve to the lowest speaker index
            turns.sort(key=lambda ts: (ts[0], ts[1]))
            for turn, s in turns:
                self.write_recipe_line(recipe[turn[2]], turn[0], turn[1], 0,
                                       s + 1, outf, segf)
            for ts in speakers:
                del ts[:]                                      # CL1:257 empties them
        else:
            turns = [(s, turn) for s, ts in enumerate(speakers) for turn in ts]
            turns.sort(key=lambda st: st[1])                   # CL2:226 (stable)
            for s, turn in turns:
                self.write_recipe_line(recipe[turn[2]], turn[0], turn[1], 0,
                                       s + 1, outf, segf)

    # ---- dispatcher -----------------------------------------------------
    def process_recipe(self, recipe, outf, segf=None, loader=None):
        """CL1:263-292 / CL2:232-261."""
        if loader is None:
            loader = lambda rl: load_features(rl, self.feapath, self.feaext,  # noqa: E731
                                              self.variant)
        rate = self.rate
        speakers = self.speakers
        this_wav = ''
        feas = None
        for l in range(len(recipe)):
            if recipe[l][0] != this_wav:
                this_wav = recipe[l][0]
                feas = loader(recipe[l])
            if speakers == [] and self.method == 'in':
                speakers.append([(recipe[l][2] * rate, recipe[l][3] * rate)])
                self.write_recipe_line(recipe[l], recipe[l][2] * rate,
                                       recipe[l][3] * rate, 0, len(speakers),
                                       outf, segf)
            elif self.method == 'hi':
                speakers.append([(recipe[l][2] * rate, recipe[l][3] * rate, l)])
            else:
                self.spk_cluster_in(feas[1], recipe[l], outf, segf)
        if self.method == 'hi':
            self.log(py2_print_str('Initial cluster with:', len(speakers), 'speakers'))
            self.spk_cluster_hi(feas[1], recipe, outf, segf)

    def summary(self, nrecipe):
        """CL1:436-442 / CL2:400-406."""
        log = self.log
        log('Useful metrics for determining the right threshold:')
        log('---------------------------------------------------')
        log(py2_print_str('Maximum between segments distance:', self.max_dist))
        if self.min_dist < MAXINT:
            log(py2_print_str('Minimum between segments distance:', self.min_dist))
        log(py2_print_str('Total segments:', nrecipe))
        log(py2_print_str('Total detected speakers:', len(self.speakers)))


def build_parser(variant):
    """Flags of CL1:296-347 / CL2:265-316.  The ``-o`` / ``-seg`` defaults
    differ between the two scripts."""
    p = argparse.ArgumentParser(description='Perform speaker clustering, using '
                                'a distance measure (CPU oracle).')
    p.add_argument('recfile', type=str)
    p.add_argument('feapath', type=str)
    p.add_argument('-seg', dest='segpath', type=str,
                   default='' if variant == 1 else None)
    p.add_argument('-o', dest='outfile', type=str,
                   default=None if variant == 1 else 'stdout')
    p.add_argument('-fe', dest='feaext', type=str, default='.fea')
    p.add_argument('-se', dest='segext', type=str, default='.seg')
    p.add_argument('-f', dest='frame_rate', type=int, default=125)
    p.add_argument('-m', dest='method', type=str, choices=['in', 'hi'], default='hi')
    p.add_argument('-d', dest='distance', type=str,
                   choices=['GLR', 'BIC', 'KL2'], default='BIC')
    p.add_argument('-t', dest='threshold', type=float, default=0.0)
    p.add_argument('-ms', dest='max_spk', type=int, default=0)
    p.add_argument('-l', dest='lambdac', type=float, default=1.3)
    p.add_argument('-tt', action='store_true')
    p.add_argument('-dlr', action='store_true')
    return p


def main(argv=None, stdout=None, variant=1, trace=None):
    """CL1:295-442 / CL2:264-406 - returns the ``Clustering`` object."""
    out = stdout if stdout is not None else sys.stdout
    args = build_parser(variant).parse_args(argv)

    def log(*items):
        out.write(py2_print_str(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as recfile:
        recipe = parse_recipe(recfile, log)
    log('Reading feature files from:', args.feapath)
    feapath = args.feapath
    segpath = args.segpath
    if variant == 1:
        if feapath[-1] != '/':
            feapath += '/'
        if segpath != '':
            log('Setting alignment segmentation files path to:', segpath)
            if segpath[-1] != '/':
                segpath += '/'
            log('Segmentation files extension:', args.segext)
        seg_on = segpath != ''
        to_file = args.outfile is not None       # default is the sys.stdout object
    else:
        if segpath:
            log('Setting alignment segmentation files path to:', segpath)
            log('Segmentation files extension:', args.segext)
        seg_on = bool(segpath)
        to_file = args.outfile != 'stdout'
    log('Feature files extension:', args.feaext)
    segfile = False
    if to_file:
        log('Writing output to:', args.outfile)
        if seg_on:
            if variant == 1:
                segfile = op.splitext(args.outfile)[0]
                segfile += '-seg' + op.splitext(args.outfile)[1]
            else:
                segfile = op.splitext(op.basename(args.outfile))[0]
                segfile += '-seg' + op.splitext(args.outfile)[1]
                segfile = op.join(segpath, segfile)
            log('Writing seg output to:', segfile)
    else:
        log('Writing output to: stdout')
    cl = Clustering(args.frame_rate, variant, args.method, args.distance,
                    args.threshold, args.max_spk, args.lambdac, args.tt,
                    args.dlr, segpath, feapath, args.feaext, log, trace)
    log('Conversion rate set to frame rate:', cl.rate)
    if args.method == 'hi':
        log('Using hierarchical clustering')
    else:
        log('Using in-order consecutive clustering')
    if args.distance == 'GLR':
        log('Using GLR as distance measure')
    elif args.distance == 'BIC':
        log('Using BIC as distance measure, lambda =', args.lambdac)
    else:
        log('Using KL2 as distance measure')
    log('Threshold distance:', args.threshold)
    log('Maximum speakers:', args.max_spk)
    if args.dlr:
        log('Disabling LNA renaming')

    if to_file:
        with open(args.outfile, 'w') as outf:
            if segfile:
                with open(segfile, 'w') as segf:
                    cl.process_recipe(recipe, outf, segf)
            else:
                cl.process_recipe(recipe, outf)
    else:
        cl.process_recipe(recipe, out)
    cl.summary(len(recipe))
    return cl


if __name__ == '__main__':
    main()
