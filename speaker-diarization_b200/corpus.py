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


class _BatchJob(object):
    """One device batch on its way through the four stages of ``diarize_batch``:
    A (device) packed upload + statistics + ONE growing-window launch over every chain,
    B (host)   replay of the window records into segmentation recipes,
    C (device) ONE clustering launch, one CTA per recording,
    D (host)   replay of the merge sequences into clustered recipes."""

    def __init__(self, ctx, batch, frame_rate, threshold):
        from . import change_detection as pcd
        self.ctx, self.batch, self.rate, self.threshold = ctx, batch, frame_rate, threshold
        self.parsed = [recipe_mod.parse(lines) for lines, _ in batch]
        self.dets = [pcd.Detector(frame_rate, threshold=threshold, ctx=ctx, **D2_CHANGE) for _ in batch]
        self.pack = None

    def stage_a(self):
        self.pack = self.ctx.upload_batch([frames() if callable(frames) else frames for _, frames in self.batch])
        self.views = [self.pack.view(r) for r in range(len(self.batch))]
        self.groups = [det.gw_chains(rec, lambda l, v=v: v)
                       for det, rec, v in zip(self.dets, self.parsed, self.views)]
        chains = [[c for _, ch in grp for c in ch] for grp in self.groups]
        d0 = self.dets[0]
        self.gw = self.pack.gw_run_batch(chains, d0.rate, d0.winsize, d0.winstep, d0.deltaws, d0.threshold,
                                         d0.lambdac, d0.metric)
        return self

    def stage_b(self):
        from . import clustering as pcl
        self.seg_lines = []
        self.seg_records = []
        for r, (det, rec, v) in enumerate(zip(self.dets, self.parsed, self.views)):
            win, first = self.gw[r]
            c0 = 0
            for feat, ch in self.groups[r]:                  # usually one group: one wav per recipe
                lo, hi = int(first[c0]), int(first[c0 + len(ch)])
                sub = win[lo:hi].copy()
                sub['chain'] -= c0
                det.prefetch(feat, ch, (sub, first[c0:c0 + len(ch) + 1] - first[c0]))
                c0 += len(ch)
            seg = io.StringIO()
            det.writer.record = []
            det.detect_changes(rec, seg, loader=lambda l, v=v: v)
            self.seg_lines.append(seg.getvalue().splitlines(True))
            self.seg_records.append(det.writer.record)
        self.cls = [pcl.Clusterer(self.rate, variant=1, threshold=self.threshold, ctx=self.ctx, **D2_CLUSTER)
                    for _ in self.batch]
        # the clustering stage reads the segmentation RECIPE (text, times rounded to 12 digits): the same
        # values without the regular-expression searches
        self.seg_parsed = [recipe_mod.lines_from_records(recs, lines)
                           for recs, lines in zip(self.seg_records, self.seg_lines)]
        self.problems = [cl.initial_segments(rec, v.n)
                         for cl, rec, v in zip(self.cls, self.seg_parsed, self.views)]
        self.live = [r for r, p in enumerate(self.problems) if p]
        return self

    def stage_c(self):
        c0 = self.cls[0]
        # The batched engine runs ONE CTA per recording and sizes its workspaces from the largest problem of
        # the batch (148 x nmax^2 doubles): recordings with many turns go through the resident engine (whole
        # GPU per recording) so that one long file neither starves nor overflows the batch.
        small = [r for r in self.live if len(self.problems[r]) <= BATCH_MAX_SEGMENTS]
        got = self.pack.cluster_batch([self.problems[r] for r in small], c0.metric, c0.lambdac,
                                      self.threshold, 0, 1) if small else []
        by_rec = dict(zip(small, got))
        for r in self.live:
            if r not in by_rec:
                sa = [a for a, _ in self.problems[r]]
                sb = [b for _, b in self.problems[r]]
                with self.views[r].cluster(sa, sb, c0.metric, c0.lambdac) as cl:
                    by_rec[r] = cl.run(self.threshold, 0, 1)
        self.merged = [by_rec[r] for r in self.live]
        return self

    def stage_d(self):
        out = []
        where = {r: k for k, r in enumerate(self.live)}
        for r, (cl, rec, v) in enumerate(zip(self.cls, self.seg_parsed, self.views)):
            if r in where:
                cl.prefetch(v, self.problems[r], self.merged[where[r]])
            clu = io.StringIO()
            cl.process_recipe(rec, clu, loader=lambda l, v=v: v)
            summary = dict(turns=len(self.seg_lines[r]), speakers=len(cl.speakers),
                           windows=self.dets[r].windows_visited, merges=len(cl.merges))
            out.append((''.join(self.seg_lines[r]), clu.getvalue(), summary))
        return out

    def close(self):
        if self.pack is not None:
            self.pack.close()
            self.pack = None


def diarize_batch(ctx, batch, frame_rate=125, threshold=0.0):
    """A batch of recordings through change detection + clustering with the device work of the
    WHOLE batch in a handful of launches: packed upload + statistics, one growing-window launch over
    the chains of all recordings (one CTA per chain), one clustering launch (one CTA per recording).
    ``batch``: list of (recipe_lines, frames).  Returns [(segmentation recipe text, clustered
    recipe text, summary)] - byte-identical to ``diarize_recording`` on each item (the packed
    statistics restart per recording; the host replay is the same code)."""
    job = _BatchJob(ctx, batch, frame_rate, threshold)
    try:
        return job.stage_a().stage_b().stage_c().stage_d()
    finally:
        job.close()


def diarize_batches(ctx, batches, frame_rate=125, threshold=0.0):
    """``diarize_batch`` over a sequence of batches with the stages OVERLAPPED: a worker thread
    queues the device stages (the context's calls are serialised; ctypes releases the GIL inside
    them) while this thread replays records.  While the host replays batch k, the device already
    uploads batch k + 1 and searches it; the clustering launch of batch k queues behind.  Up to
    three packed batches are resident at a time (6,560 B per frame each).  Yields the result list
    of every batch, in order; results are those of ``diarize_batch``."""
    from concurrent.futures import ThreadPoolExecutor
    batches = list(batches)
    if not batches:
        return
    with ThreadPoolExecutor(max_workers=1) as dev:
        jobs = [_BatchJob(ctx, b, frame_rate, threshold) for b in batches]
        try:
            fa = dev.submit(jobs[0].stage_a)
            prev = None                                         # (job, future of its stage C)
            for k, job in enumerate(jobs):
                fa.result()
                if k + 1 < len(jobs):
                    fa = dev.submit(jobs[k + 1].stage_a)        # device: upload + search of the next batch ...
                job.stage_b()                                   # ... while the host replays this one
                fc = dev.submit(job.stage_c)
                if prev is not None:
                    prev[1].result()
                    out = prev[0].stage_d()
                    dev.submit(prev[0].close)
                    yield out
                prev = (job, fc)
            prev[1].result()
            out = prev[0].stage_d()
            dev.submit(prev[0].close).result()
            yield out
        finally:
            for job in jobs:
                try:
                    dev.submit(job.close).result()
                except Exception:           # pragma: no cover - the first error is the one to report
                    pass


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

