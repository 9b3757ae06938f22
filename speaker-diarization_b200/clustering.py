"""Speaker clustering: the host side of ``spk-clustering.py`` (variant 1) and
``spk-clustering2.py`` (variant 2).

Same command lines, recipe formats and stdout text as the reference scripts;
cluster statistics, every pair distance and the agglomerative merge loop run on
the GPU (``_abi.Clusters``).  The host keeps what is bookkeeping in the
reference too: the ``speakers`` list of turn tuples (labels are positions in
that list after every ``pop``), the ``Merging:`` log lines, the
output order, Python-2 text.

The two scripts differ in result-affecting ways and both are reproduced
(SURVEY.md Q5, Q14): variant 1 rescores row and column of a symmetric matrix
(spk-clustering.py:222-231) and its ``-o stdout`` writes a file called
``stdout``; variant 2 rescores the row only and keeps stale entries
(spk-clustering2.py:206-215).
"""

import argparse
import os.path as op
import sys

import numpy as np

from . import _abi
from .feacat import feature_file_name, read_features
from .py2fmt import MAXINT, p2line
from .recipe import Writer, parse


def _is_inf(d):
    return d == np.inf or d == -np.inf


class Clusterer(object):
    """One run of the clustering stage over a parsed recipe."""

    def __init__(self, frame_rate=125, variant=1, method='hi', distance='BIC', threshold=0.0,
                 max_spk=0, lambdac=1.3, tt=False, dlr=False, segpath='', feapath='./',
                 feaext='.fea', device=0, ctx=None, log=None, engine='device'):
        self.rate = float(frame_rate)
        self.variant = variant
        self.method = method
        self.distance = distance
        self.metric = _abi.METRIC[distance]
        self.threshold = threshold
        self.max_spk = max_spk
        self.lambdac = lambdac
        self.tt = tt
        self.feapath = feapath
        self.feaext = feaext
        self.log = log if log is not None else (lambda *a: None)
        self.quiet = log is None       # nobody reads the log: skip formatting it (corpus driver)
        self.writer = Writer(self.rate, rename=not dlr, segprefix=segpath or None)
        self.max_dist = 0
        self.min_dist = MAXINT
        self.max_det_dist = 0
        self.min_det_dist = MAXINT
        self.speakers = []
        self.merges = []               # (a, b, d) in compacted indices, as logged
        # 'device': resident engine (BIC, GLR, KL2); 'host': merge loop on the host with
        # every distance from the device (the cross-check in the tests)
        self.engine = engine
        self._own_ctx = ctx is None
        self.ctx = ctx if ctx is not None else _abi.Context(device)
        self._prefetch = {}
        self._in_plan = {}             # line index -> (distances, decision) of the device-resident in-order loop

    def initial_segments(self, recipe, n):
        """The frame ranges of the initial clusters of a hierarchical run over ``recipe``
        (CL1:282-283 + CL1:47), without running it."""
        return [self._range((line.start * self.rate, line.end * self.rate), n) for line in recipe]

    def prefetch(self, feat, segments, result):
        """``result`` = what ``feat.cluster(segments).run(...)`` would return (the corpus driver
        clusters a whole batch of recordings in one launch)."""
        self._prefetch[(id(feat), tuple(segments))] = result

    def close(self):
        if self._own_ctx and self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    def load(self, line):
        name = feature_file_name(line.audio, self.feapath, self.feaext, concat=self.variant == 1)
        dim, frames = read_features(name)
        # clustering on its own never scores a window: frames only, cluster records straight from them
        # (spk-clustering.py:46-52); `-m in` and KL2 score sets and build the window statistics on first use
        return self.ctx.upload_frames(frames)

    @staticmethod
    def _range(turn, n):
        """``features[int(s):int(e)]`` (CL1:47, CL2:48) as a clamped range."""
        a = int(turn[0])
        b = int(turn[1])
        a = 0 if a < 0 else (n if a > n else a)
        b = 0 if b < 0 else (n if b > n else b)
        return (a, b if b > a else a)

    def _ranges(self, spk, n):
        return [self._range(t, n) for t in spk]

    def _track(self, d):
        if not _is_inf(d):
            if d > self.max_dist:
                self.max_dist = d
            if d < self.min_dist:
                self.min_dist = d

    # ---- in-order clustering --------------------------------------------------
    def _plan_in(self, feat, recipe, l):
        """The device-resident loop over the lines l .. (last line of this wav): every distance and decision of
        ``_cluster_in`` for them in ONE launch (``spkdiar_cluster_inorder``); the per-line code below then replays
        the script's bookkeeping from the distances.  BIC and GLR; KL2 keeps one scoring call per line."""
        rate = self.rate
        m = l
        while m < len(recipe) and recipe[m].audio == recipe[l].audio:
            m += 1
        segs = []
        for line in recipe[l:m]:
            if self.variant == 1:
                segs.append(self._range((int(line.start * rate), int(line.end * rate)), feat.n))
            else:
                segs.append(self._range((line.start * rate, line.end * rate), feat.n))
        dist, first, best = feat.cluster_inorder([self._ranges(s, feat.n) for s in self.speakers],
                                                 [a for a, _ in segs], [b for _, b in segs],
                                                 self.metric, self.lambdac, self.threshold)
        for k in range(m - l):
            self._in_plan[l + k] = (dist[int(first[k]):int(first[k + 1])], int(best[k]))

    def _cluster_in(self, feat, line, outf, segf, index=None):
        """CL1:136-175 / CL2:135-170: the new segment against every speaker,
        one batched device call."""
        speakers = self.speakers
        if self.variant == 1:
            start = int(line.start * self.rate)
            end = int(line.end * self.rate)
        else:
            start = line.start * self.rate
            end = line.end * self.rate
        seg = [self._range((start, end), feat.n)]
        plan = self._in_plan.pop(index, None) if index is not None else None
        if plan is not None:
            dist, device_best = plan                 # scored (and decided) by the device-resident loop
            assert len(dist) == len(speakers), (len(dist), len(speakers))
        else:
            device_best = None
            dist = feat.score_sets([self._ranges(s, feat.n) for s in speakers],
                                   [seg] * len(speakers), self.metric, self.lambdac)
        mind = MAXINT
        best = None
        d = None
        for spk, d in enumerate(dist):
            if self.tt:
                self.log(p2line('Time:', end, '- Distance:', d, '- Speaker:', spk + 1))
            if not _is_inf(d):
                if d > self.max_dist:
                    self.max_dist = d
                if d < self.min_dist:
                    self.min_dist = d
                if d < mind:
                    mind = d
                    best = spk
        if device_best is not None and (best if mind <= self.threshold else -1) != device_best:
            raise RuntimeError('in-order clustering: host replay and device loop part at a line (%r against %r)'
                               % (best if mind <= self.threshold else -1, device_best))
        if mind <= self.threshold:
            if self.variant == 1:                                  # CL1:164-167
                if d > self.max_det_dist:
                    self.max_det_dist = d
                if d < self.min_det_dist:
                    self.min_det_dist = d
            speakers[best].append((start, end))
            self.writer.write(line, start, end, 0, 'speaker_%d' % (best + 1), outf, segf)
        else:
            speakers.append([(start, end)])
            self.writer.write(line, start, end, 0, 'speaker_%d' % len(speakers), outf, segf)

    # ---- hierarchical clustering ------------------------------------------------
    def _merge_sequence_device(self, feat):
        n = feat.n
        seg = [self._range(s[0], n) for s in self.speakers]
        got = self._prefetch.pop((id(feat), tuple(seg)), None)
        if got is not None:
            merges, stats = got
        else:
            with feat.cluster([r[0] for r in seg], [r[1] for r in seg], self.metric, self.lambdac) as cl:
                merges, stats = cl.run(self.threshold, self.max_spk, self.variant)
        if self.variant == 1:
            self.max_dist, self.min_dist = self._stat(stats[0], 0), self._stat(stats[1], MAXINT)
            self.max_det_dist = self._stat(stats[2], 0)
            self.min_det_dist = self._stat(stats[3], MAXINT)
        else:                                               # CL2:220-221 always assigns floats
            self.max_dist, self.min_dist = np.float64(stats[0]), np.float64(stats[1])
        return [(int(m['a']), int(m['b']), np.float64(m['d'])) for m in merges]

    @staticmethod
    def _stat(v, initial):
        """The scripts start max at the int 0 and min at the int sys.maxint and
        print them untouched when nothing replaced them."""
        return initial if float(v) == float(initial) else np.float64(v)

    def _merge_sequence_host(self, feat):
        """Same loop with the matrix on the host (numpy), every distance scored on
        the device in batches.  CL1:184-240 / CL2:178-222."""
        sp = len(self.speakers)
        v1 = self.variant == 1
        clusters = [self._ranges(s, feat.n) for s in self.speakers]
        if v1:
            M = np.empty((sp, sp))
            M[:] = 0.0
            np.fill_diagonal(M, MAXINT)
        else:
            M = np.full((sp, sp), np.inf)
        iu = [(i, j) for i in range(sp) for j in range(i + 1, sp)]
        if iu:
            d = feat.score_sets([clusters[i] for i, _ in iu], [clusters[j] for _, j in iu],
                                self.metric, self.lambdac)
            for (i, j), dk in zip(iu, d):
                M[i, j] = dk
                if v1:
                    M[j, i] = dk
                    self._track(dk)
        merges = []
        while True:
            mind = M.min()
            if mind <= self.threshold or (self.max_spk > 0 and len(clusters) > self.max_spk):
                index = M.argmin()
                a, b = index // len(clusters), index % len(clusters)
                if a > b:
                    a, b = b, a
                if a == b:
                    break
                if v1:
                    if mind > self.max_det_dist:
                        self.max_det_dist = mind
                    if mind < self.min_det_dist:
                        self.min_det_dist = mind
                merges.append((int(a), int(b), mind))
                clusters[a] = clusters[a] + clusters[b]
                clusters.pop(b)
                M = np.delete(np.delete(M, b, 0), b, 1)
                others = [k for k in range(len(clusters)) if k != a]
                if others:
                    d = feat.score_sets([clusters[a]] * len(others), [clusters[k] for k in others],
                                        self.metric, self.lambdac)
                    for k, dk in zip(others, d):
                        M[a, k] = dk
                        if v1:
                            M[k, a] = dk
                            self._track(dk)
            else:
                if not v1:
                    self.max_dist = M.max()
                    self.min_dist = M.min()
                break
        return merges

    def _cluster_hi(self, feat, recipe, outf, segf):
        """CL1:178-260 / CL2:173-229."""
        speakers = self.speakers
        use_device = self.engine == 'device'
        merges = self._merge_sequence_device(feat) if use_device else self._merge_sequence_host(feat)
        for a, b, d in merges:
            if not self.quiet:
                self.log(p2line('Merging:', a + 1, 'and', b + 1, 'distance:', d))
            speakers[a].extend(speakers[b])
            speakers.pop(b)
        self.merges = merges
        self.log(p2line('Final speakers:', len(speakers)))
        turns = [(turn, s) for s, ts in enumerate(speakers) for turn in ts]
        if self.variant == 1:
            turns.sort(key=lambda ts: (ts[0], ts[1]))       # CL1:243-260: globally smallest turn first
        else:
            turns.sort(key=lambda ts: ts[0])                # CL2:225-226: stable sort on the turn tuple
        for turn, s in turns:
            self.writer.write(recipe[turn[2]], turn[0], turn[1], 0, 'speaker_%d' % (s + 1),
                              outf, segf)
        if self.variant == 1:
            for ts in speakers:
                del ts[:]                                   # CL1:257 removes every written turn

    # ---- dispatcher -------------------------------------------------------------
    def process_recipe(self, recipe, outf, segf=None, loader=None):
        """CL1:263-292 / CL2:232-261 (the last-loaded wav serves every segment
        of a hierarchical run)."""
        load = loader if loader is not None else self.load
        rate = self.rate
        this_wav = ''
        feat = None
        owned = []
        try:
            for l, line in enumerate(recipe):
                if line.audio != this_wav:
                    this_wav = line.audio
                    # one wav resident at a time, as in the reference (CL1:268-271): every distance of either
                    # method reads the frames of the wav loaded LAST (Q10), so the previous handle is dead
                    if feat is not None and loader is None:
                        feat.close()
                        owned.remove(feat)
                    feat = load(line)
                    if loader is None:
                        owned.append(feat)
                if self.speakers == [] and self.method == 'in':
                    self.speakers.append([(line.start * rate, line.end * rate)])
                    self.writer.write(line, line.start * rate, line.end * rate, 0,
                                      'speaker_%d' % len(self.speakers), outf, segf)
                elif self.method == 'hi':
                    self.speakers.append([(line.start * rate, line.end * rate, l)])
                else:
                    if self.engine == 'device' and self.metric != _abi.KL2 and l not in self._in_plan \
                            and hasattr(feat, 'cluster_inorder'):
                        self._plan_in(feat, recipe, l)
                    self._cluster_in(feat, line, outf, segf, l)
            if self.method == 'hi':
                self.log(p2line('Initial cluster with:', len(self.speakers), 'speakers'))
                self._cluster_hi(feat, recipe, outf, segf)
        finally:
            for f in owned:
                f.close()

    def summary(self, nrecipe):
        """CL1:436-442 / CL2:400-406."""
        log = self.log
        log('Useful metrics for determining the right threshold:')
        log('---------------------------------------------------')
        log(p2line('Maximum between segments distance:', self.max_dist))
        if self.min_dist < MAXINT:
            log(p2line('Minimum between segments distance:', self.min_dist))
        log(p2line('Total segments:', nrecipe))
        log(p2line('Total detected speakers:', len(self.speakers)))


def build_parser(variant):
    """spk-clustering.py:296-347 / spk-clustering2.py:265-316 plus ``--device``."""
    p = argparse.ArgumentParser(description='Perform speaker clustering, using a distance measure.')
    p.add_argument('recfile', type=str, help='Specifies the input recipe file')
    p.add_argument('feapath', type=str, help='Specifies the features files path')
    p.add_argument('-seg', dest='segpath', type=str, default='' if variant == 1 else None,
                   help='Alignment segmentation files path; generates "alignment=" information')
    # variant 1's default is the sys.stdout OBJECT, so "-o stdout" names a file
    p.add_argument('-o', dest='outfile', type=str, default=None if variant == 1 else 'stdout',
                   help='Specifies an output file, default stdout')
    p.add_argument('-fe', dest='feaext', type=str, default='.fea',
                   help='Specifies feature file extension, default ".fea"')
    p.add_argument('-se', dest='segext', type=str, default='.seg',
                   help='Specifies segmentation files extension, default ".seg"')
    p.add_argument('-f', dest='frame_rate', type=int, default=125,
                   help='Specifies the frame rate, default 125')
    p.add_argument('-m', dest='method', type=str, choices=['in', 'hi'], default='hi',
                   help='Hierarchical agglomerative (hi, default) or in-order (in) clustering')
    p.add_argument('-d', dest='distance', type=str, choices=['GLR', 'BIC', 'KL2'], default='BIC',
                   help='Distance measure, default BIC')
    p.add_argument('-t', dest='threshold', type=float, default=0.0,
                   help='Threshold distance for detection, default 0.0')
    p.add_argument('-ms', dest='max_spk', type=int, default=0,
                   help='Maximum speakers stopping criterion for hierarchical clustering, default 0')
    p.add_argument('-l', dest='lambdac', type=float, default=1.3,
                   help='Lambda penalty weight for BIC, default 1.3')
    p.add_argument('-tt', action='store_true', help='Output every decision distance')
    p.add_argument('-dlr', action='store_true', help='Disable lna renaming')
    p.add_argument('--device', type=int, default=0, help='CUDA device ordinal (default 0)')
    return p


def main(argv=None, stdout=None, variant=1, ctx=None, engine='device'):
    """spk-clustering.py:295-442 / spk-clustering2.py:264-406."""
    out = stdout if stdout is not None else sys.stdout
    args = build_parser(variant).parse_args(argv)

    def log(*items):
        out.write(p2line(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as recfile:
        recipe = parse(recfile, log)
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
        to_file = args.outfile is not None
    else:
        if segpath:
            log('Setting alignment segmentation files path to:', segpath)
            log('Segmentation files extension:', args.segext)
        to_file = args.outfile != 'stdout'
    log('Feature files extension:', args.feaext)
    segfile = False
    if to_file:
        log('Writing output to:', args.outfile)
        if segpath:
            if variant == 1:
                segfile = op.splitext(args.outfile)[0] + '-seg' + op.splitext(args.outfile)[1]
            else:
                segfile = op.splitext(op.basename(args.outfile))[0] + '-seg' + \
                    op.splitext(args.outfile)[1]
                segfile = op.join(segpath, segfile)
            log('Writing seg output to:', segfile)
    else:
        log('Writing output to: stdout')
    cl = Clusterer(args.frame_rate, variant, args.method, args.distance, args.threshold,
                   args.max_spk, args.lambdac, args.tt, args.dlr, segpath or '', feapath,
                   args.feaext, args.device, ctx, log, engine)
    try:
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
    finally:
        cl.close()
    return cl


if __name__ == '__main__':
    main()
