"""Corpus driver: ``spk-diarization2.py``'s hot half over many recordings, one
resident process per GPU.

The reference diarizes one media file per invocation of ``spk-diarization2.py``
(lines 122-128: ``spk-change-detection.py -m gw -d BIC -w 1.0 -st 3.0 -dws 0.1
-l 1.0`` then ``spk-clustering.py -m hi -l 1.3``), and every stage is its own
process.  CUDA context creation alone would dominate that on a GPU, so a corpus
is processed by ONE process per GPU that keeps its context and streams the
recordings through it.  Recordings are independent (the reference itself cannot
mix wavs in one hierarchical clustering run), so they are
sharded by file index over the ranks with no data-path communication; the only
collective is the final gather of the per-file summaries.

Per recording the two stages run exactly as the drop-in scripts run them (same
``Detector`` / ``Clusterer`` objects, fresh LNA-renaming state per stage, as a
fresh process would have), so the recipes are byte-identical to running the
scripts one file at a time.
"""

import io
import os
import os.path as op

from . import recipe as recipe_mod

D2_CHANGE = dict(method='gw', distance='BIC', winsize=1.0, winstep=3.0, deltaws=0.1, lambdac=1.0)
D2_CLUSTER = dict(method='hi', distance='BIC', lambdac=1.3)
BATCH_MAX_SEGMENTS = 512        # larger clustering problems leave the one-CTA-per-recording engine


def shard(n_items, rank, world):
    """Indices of the items rank ``rank`` of ``world`` processes owns: round robin
    ("file index mod nGPU")."""
    return list(range(rank, n_items, world))


def diarize_recording(ctx, recipe_lines, frames_loader, frame_rate=125, threshold=0.0):
    """One recording through change detection + clustering on an existing
    context.  ``frames_loader(line)`` -> ``Features``.  Returns (segmentation
    recipe text, clustered recipe text, summary dict)."""
    from . import change_detection as pcd, clustering as pcl
    parsed = recipe_mod.parse(recipe_lines)
    det = pcd.Detector(frame_rate, threshold=threshold, ctx=ctx, **D2_CHANGE)
    seg = io.StringIO()
    det.detect_changes(parsed, seg, loader=frames_loader)
    seg_lines = seg.getvalue().splitlines(True)
    cl = pcl.Clusterer(frame_rate, variant=1, threshold=threshold, ctx=ctx, **D2_CLUSTER)
    out = io.StringIO()
    cl.process_recipe(recipe_mod.parse(seg_lines), out, loader=frames_loader)
    summary = dict(turns=len(seg_lines), speakers=len(cl.speakers), windows=det.windows_visited,
                   merges=len(cl.merges))
    return seg.getvalue(), out.getvalue(), summary


class _PythonRec(object):
    """One recording of a device batch replayed by the general Python classes (``Detector`` /
    ``Clusterer``): every recipe the scripts accept, at about 0.7 ms of host time per ten minutes."""

    def __init__(self, ctx, lines, frame_rate, threshold):
        from . import change_detection as pcd
        self.ctx, self.rate, self.threshold = ctx, frame_rate, threshold
        self.parsed = recipe_mod.parse(lines)
        self.det = pcd.Detector(frame_rate, threshold=threshold, ctx=ctx, **D2_CHANGE)

    def chains(self, view):
        """-> (seg_a, seg_b): the chains of this recording in packed frame rows."""
        import numpy as np
        self.view = view
        self.groups = self.det.gw_chains(self.parsed, lambda l: view)
        flat = [c for _, ch in self.groups for c in ch]
        return (np.array([view.off + int(a) for a, _ in flat], dtype=np.int64),
                np.array([view.off + int(b) for _, b in flat], dtype=np.int64))

    def segment(self, win, first, chain0=0):
        """``win``: the records of the whole launch, ``first``: this recording's offsets into it,
        ``chain0``: the launch's number of this recording's first chain.
        -> (seg_a, seg_b) of the initial clusters of the clustering stage, packed frame rows."""
        import numpy as np
        from . import clustering as pcl
        det, v = self.det, self.view
        c0 = 0
        for feat, ch in self.groups:                         # usually one group: one wav per recipe
            lo, hi = int(first[c0]), int(first[c0 + len(ch)])
            sub = win[lo:hi].copy()
            sub['chain'] -= chain0 + c0
            det.prefetch(feat, ch, (sub, first[c0:c0 + len(ch) + 1] - first[c0]))
            c0 += len(ch)
        seg = io.StringIO()
        det.writer.record = []
        det.detect_changes(self.parsed, seg, loader=lambda l: v)
        self.seg_text = seg.getvalue()
        seg_lines = self.seg_text.splitlines(True)
        # the clustering stage reads the segmentation RECIPE (text, times rounded to 12 digits): the same
        # values without the regular-expression searches
        self.seg_parsed = recipe_mod.lines_from_records(det.writer.record, seg_lines)
        self.cl = pcl.Clusterer(self.rate, variant=1, threshold=self.threshold, ctx=self.ctx, **D2_CLUSTER)
        self.problem = self.cl.initial_segments(self.seg_parsed, v.n)
        self.nturns = len(seg_lines)
        return (np.array([v.off + a for a, _ in self.problem], dtype=np.int64),
                np.array([v.off + b for _, b in self.problem], dtype=np.int64))

    def finish(self, merged):
        """``merged``: (merges, stats) of this recording's clustering problem, or None without turns."""
        v = self.view
        if merged is not None:
            self.cl.prefetch(v, self.problem, merged)
        clu = io.StringIO()
        self.cl.process_recipe(self.seg_parsed, clu, loader=lambda l: v)
        summary = dict(turns=self.nturns, speakers=len(self.cl.speakers), windows=self.det.windows_visited,
                       merges=len(self.cl.merges))
        return self.seg_text, clu.getvalue(), summary

    def close(self):
        pass


class _NativeRec(object):
    """The same three steps through ``spkdiar_replay_*`` (csrc/spkdiar_replay.cu): native code,
    tens of microseconds per recording.  Raises ``_abi.ReplayUnsupported`` for what it does not
    reproduce to the byte; the job then replays that recording with ``_PythonRec``."""

    def __init__(self, lines, frame_rate):
        from . import _abi
        self.rp = _abi.Replay(frame_rate, lines)
        if not self.rp.single_wav or self.rp.nchains == 0:
            self.rp.close()
            raise _abi.ReplayUnsupported('several wavs or no line in one recipe')

    def chains(self, view):
        self.view = view
        return self.rp.chains(view.n, view.off)

    def segment(self, win, first, chain0=0):
        self.nturns = self.rp.segment(win, first)
        return self.rp.turns(self.view.n, self.nturns, self.view.off)

    def finish(self, merged):
        info0 = self.rp.info()
        nspk = self.rp.cluster(merged[0]) if merged is not None else 0
        summary = dict(turns=self.nturns, speakers=nspk, windows=int(info0[5]),
                       merges=len(merged[0]) if merged is not None else 0)
        return self.rp.text(0), self.rp.text(1), summary

    def close(self):
        self.rp.close()


class _BatchJob(object):
    """One device batch on its way through the four stages of ``diarize_batch``:
    A (device) packed upload + statistics + ONE growing-window launch over every chain,
    B (host)   replay of the window records into segmentation recipes,
    C (device) ONE clustering launch, one CTA per recording,
    D (host)   replay of the merge sequences into clustered recipes.
    The host stages run in native code (``_NativeRec``) unless ``native=False`` or a recording
    needs the general replay (``_PythonRec``); results are identical."""

    def __init__(self, ctx, batch, frame_rate, threshold, native=True):
        from . import _abi
        self.ctx, self.batch, self.rate, self.threshold = ctx, batch, frame_rate, threshold
        self.recs = []
        for lines, _ in batch:
            rec = None
            if native:
                try:
                    rec = _NativeRec(lines, frame_rate)
                except _abi.ReplayUnsupported:
                    rec = None
            self.recs.append(rec if rec is not None else _PythonRec(ctx, lines, frame_rate, threshold))
        self.pack = None

    def _python_again(self, r):
        """Recording r left the native replay: the general one takes over from the start."""
        self.recs[r].close()
        rec = _PythonRec(self.ctx, self.batch[r][0], self.rate, self.threshold)
        sa, sb = rec.chains(self.views[r])
        c0, c1 = self.owner[r], self.owner[r + 1]
        if sa.tolist() != self.seg_a[c0:c1].tolist() or sb.tolist() != self.seg_b[c0:c1].tolist():
            raise RuntimeError('native and Python replay disagree on the chains of recording %d' % r)
        self.recs[r] = rec
        return rec

    def stage_a(self):
        import numpy as np
        from . import change_detection as pcd
        self.pack = self.ctx.upload_batch([frames() if callable(frames) else frames for _, frames in self.batch])
        self.views = [self.pack.view(r) for r in range(len(self.batch))]
        parts = [rec.chains(v) for rec, v in zip(self.recs, self.views)]
        self.owner = np.cumsum([0] + [len(sa) for sa, _ in parts])
        self.seg_a = np.concatenate([sa for sa, _ in parts]) if parts else np.zeros(0, dtype=np.int64)
        self.seg_b = np.concatenate([sb for _, sb in parts]) if parts else np.zeros(0, dtype=np.int64)
        d0 = pcd.Detector(self.rate, threshold=self.threshold, ctx=self.ctx, **D2_CHANGE)
        self.win, self.first = self.pack.gw_run(self.seg_a, self.seg_b, d0.rate, d0.winsize, d0.winstep, d0.deltaws,
                                                d0.threshold, d0.lambdac, d0.metric)
        self.metric_cd = d0.metric
        return self

    def stage_b(self):
        import numpy as np
        from . import _abi
        parts = []
        for r, rec in enumerate(self.recs):
            first = self.first[self.owner[r]:self.owner[r + 1] + 1]
            try:
                parts.append(rec.segment(self.win, first, int(self.owner[r])))
            except _abi.ReplayUnsupported:
                parts.append(self._python_again(r).segment(self.win, first, int(self.owner[r])))
        self.pfirst = np.cumsum([0] + [len(sa) for sa, _ in parts]).astype(np.int64)
        self.turn_a = np.concatenate([sa for sa, _ in parts]) if parts else np.zeros(0, dtype=np.int64)
        self.turn_b = np.concatenate([sb for _, sb in parts]) if parts else np.zeros(0, dtype=np.int64)
        return self

    def stage_c(self):
        from . import _abi
        metric, lam = _abi.METRIC[D2_CLUSTER['distance']], D2_CLUSTER['lambdac']
        size = self.pfirst[1:] - self.pfirst[:-1]
        # The batched engine runs ONE CTA per recording and sizes its workspaces from the largest problem of
        # the batch (148 x nmax^2 doubles): recordings with many turns go through the resident engine (whole
        # GPU per recording) so that one long file neither starves nor overflows the batch.
        small = [r for r in range(len(self.recs)) if 0 < size[r] <= BATCH_MAX_SEGMENTS]
        self.merged = [None] * len(self.recs)
        if small:
            got = self.pack.cluster_batch_rows([(self.turn_a[self.pfirst[r]:self.pfirst[r + 1]],
                                                 self.turn_b[self.pfirst[r]:self.pfirst[r + 1]]) for r in small],
                                               metric, lam, self.threshold, 0, 1)
            for r, g in zip(small, got):
                self.merged[r] = g
        for r in range(len(self.recs)):
            if size[r] > BATCH_MAX_SEGMENTS:
                with self.pack.cluster(self.turn_a[self.pfirst[r]:self.pfirst[r + 1]],
                                       self.turn_b[self.pfirst[r]:self.pfirst[r + 1]], metric, lam) as cl:
                    self.merged[r] = cl.run(self.threshold, 0, 1)
        return self

    def stage_d(self):
        return [rec.finish(self.merged[r]) for r, rec in enumerate(self.recs)]

    def close(self):
        for rec in self.recs:
            rec.close()
        if self.pack is not None:
            self.pack.close()
            self.pack = None


def diarize_batch(ctx, batch, frame_rate=125, threshold=0.0, native=True):
    """A batch of recordings through change detection + clustering with the device work of the
    WHOLE batch in a handful of launches: packed upload + statistics, one growing-window launch over
    the chains of all recordings (one CTA per chain), one clustering launch (one CTA per recording).
    ``batch``: list of (recipe_lines, frames).  Returns [(segmentation recipe text, clustered
    recipe text, summary)] - byte-identical to ``diarize_recording`` on each item (the packed
    statistics restart per recording; the host replay is the same code)."""
    job = _BatchJob(ctx, batch, frame_rate, threshold, native)
    try:
        return job.stage_a().stage_b().stage_c().stage_d()
    finally:
        job.close()


def diarize_batches(ctx, batches, frame_rate=125, threshold=0.0, native=True, lanes=2):
    """``diarize_batch`` over a sequence of batches in ``lanes`` LANES that overlap each other:
    every lane is a thread with its own context (its own stream) on the same device and takes
    batches k, k + lanes, ... through the four stages one after the other (every ABI call is
    synchronous and releases the GIL).  While the growing-window search of one batch occupies the
    SMs, the next batch's frames cross PCIe on the other lane's stream (26 of the 88 ms a batch of
    148 ten-minute recordings needs on the device are that copy), and the short host stages
    (native replay) of one lane hide behind device work of the other.  Lane 0 uses ``ctx``; the
    others use contexts that ``ctx`` keeps for this purpose (``Context.lane_contexts``).  One packed batch per lane is resident
    (6,560 B per frame).  Yields the result list of every batch, in order; results are those of
    ``diarize_batch``."""
    from concurrent.futures import ThreadPoolExecutor
    batches = list(batches)
    if not batches:
        return
    lanes = max(1, min(int(lanes), len(batches)))
    pools = []
    try:
        ctxs = [ctx] + ctx.lane_contexts(lanes - 1)
        pools = [ThreadPoolExecutor(max_workers=1) for _ in range(lanes)]

        def run(k):
            job = _BatchJob(ctxs[k % lanes], batches[k], frame_rate, threshold, native)
            try:
                return job.stage_a().stage_b().stage_c().stage_d()
            finally:
                job.close()
        # a lane holds at most one finished batch ahead of the consumer
        futs = {}
        nxt = 0
        for k in range(len(batches)):
            while nxt < len(batches) and nxt < k + 2 * lanes:
                futs[nxt] = pools[nxt % lanes].submit(run, nxt)
                nxt += 1
            yield futs.pop(k).result()
    finally:
        for pool in pools:
            pool.shutdown(wait=True, cancel_futures=True)


def run_corpus(items, rank=0, world=1, device=None, outdir=None, frame_rate=125, runner=None,
               gather=None, batch=0, overlap=False):
    """Diarize the items this rank owns.

    ``items``  list of (name, recipe_lines, frames) with ``frames`` a float32
               (n, 39) array or a callable returning one (lazy file reads);
    ``runner`` overrides the per-recording function (the CPU tests inject one;
               the default uploads the frames and calls ``diarize_recording``);
    ``gather`` ``gather(obj) -> list of every rank's obj`` (e.g. built on
               ``torch.distributed.all_gather_object``); None = single process.
    ``batch``  recordings per device batch (``diarize_batch``: the device work of the whole
               batch in four launches); 0 = one recording at a time.
    ``overlap`` with ``batch``: device stages of the next batch run while the host replays the
               current one (``diarize_batches``).

    Returns {name: summary} for the WHOLE corpus on every rank (after the
    gather) - results do not depend on the number of ranks.
    """
    mine = shard(len(items), rank, world)
    ctx = None
    local = {}
    try:
        if runner is None:
            from . import _abi
            ctx = _abi.Context(rank if device is None else device)

            def runner(name, lines, frames):
                data = frames() if callable(frames) else frames
                feat = ctx.upload(data)
                try:
                    return diarize_recording(ctx, lines, lambda l: feat, frame_rate)
                finally:
                    feat.close()
        results = []
        if batch > 0 and ctx is not None:
            parts = [[items[k] for k in mine[b0:b0 + batch]] for b0 in range(0, len(mine), batch)]
            plain = [[(lines, frames) for _, lines, frames in part] for part in parts]
            gen = diarize_batches(ctx, plain, frame_rate) if overlap else \
                (diarize_batch(ctx, b, frame_rate) for b in plain)
            for part, got in zip(parts, gen):
                results.extend((name,) + g for (name, _, _), g in zip(part, got))
        else:
            for k in mine:
                name, lines, frames = items[k]
                results.append((name,) + tuple(runner(name, lines, frames)))
        for name, seg, clu, summary in results:
            if outdir is not None:
                os.makedirs(outdir, exist_ok=True)
                with open(op.join(outdir, name + '.spkc.recipe'), 'w') as f:
                    f.write(seg)
                with open(op.join(outdir, name + '.recipe'), 'w') as f:
                    f.write(clu)
            local[name] = summary
    finally:
        if ctx is not None:
            ctx.close()
    if gather is None:
        return local
    merged = {}
    for part in gather(local):
        merged.update(part)
    return merged


def items_from_recipes(recipe_paths, feapath, feaext='.fea'):
    """Corpus items for ``run_corpus`` from recipe files as ``voice-detection2.py`` writes them (one
    recipe per recording, every line of it naming the same wav): the item is called after the recipe
    file, the frames are read lazily from ``feapath`` / <wav basename> + ``feaext`` when the
    recording's batch is uploaded."""
    from .feacat import feature_file_name, read_features
    items = []
    seen = {}
    for path in recipe_paths:
        with open(path, 'r') as f:
            lines = f.readlines()
        parsed = recipe_mod.parse(lines)
        if not parsed:
            raise ValueError('%s: no recipe line with audio / lna / start-time / end-time' % path)
        wavs = {l.audio for l in parsed}
        if len(wavs) != 1:
            raise ValueError('%s names %d wavs; the corpus driver takes one recording per recipe' % (path, len(wavs)))
        fea = feature_file_name(parsed[0].audio, feapath, feaext)

        def frames(fea=fea):
            dim, x = read_features(fea)
            return x
        name = op.splitext(op.basename(path))[0]
        if name in seen:
            raise ValueError('recipes %s and %s would both write %s.recipe: item names must be unique'
                             % (seen[name], path, name))
        seen[name] = path
        items.append((name, lines, frames))
    return items


def main(argv=None):
    """``spk-diarization-corpus.py recipes... feapath -o outdir``: the two hot-path calls of
    ``spk-diarization2.py`` (lines 122-128) over many recordings.  Started once, or as one process
    per GPU under ``torchrun`` (recordings are dealt to the ranks by index, no communication but
    the final gather of the summaries)."""
    import argparse
    ap = argparse.ArgumentParser(description='Speaker-turn segmentation + clustering of a corpus of recordings '
                                 '(spk-diarization2.py flags), recordings sharded over the GPUs.')
    ap.add_argument('recipes', nargs='+', help='one speech-turn recipe per recording (voice-detection2.py output)')
    ap.add_argument('feapath', help='directory of the feacat feature files')
    ap.add_argument('-o', dest='outdir', required=True, help='directory for <name>.spkc.recipe and <name>.recipe')
    ap.add_argument('-fe', dest='feaext', default='.fea', help='feature file extension, default ".fea"')
    ap.add_argument('-f', dest='frame_rate', type=int, default=125, help='frame rate, default 125')
    ap.add_argument('--batch', type=int, default=64, help='recordings per device batch (0: one at a time)')
    ap.add_argument('--no-overlap', dest='overlap', action='store_false',
                    help='do not overlap the device stages of the next batch with the host replay')
    args = ap.parse_args(argv)
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    gather = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('gloo')           # summaries only: a few hundred bytes per recording

        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    items = items_from_recipes(args.recipes, args.feapath, args.feaext)
    done = run_corpus(items, rank=rank, world=world, device=local, outdir=args.outdir, frame_rate=args.frame_rate,
                      gather=gather, batch=args.batch, overlap=args.overlap)
    if rank == 0:
        for name in sorted(done):
            s = done[name]
            print('%s: %d turns, %d speakers (%d windows, %d merges)'
                  % (name, s['turns'], s['speakers'], s['windows'], s['merges']))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return done

