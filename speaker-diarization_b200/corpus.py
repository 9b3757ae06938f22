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


def diarize_batch(ctx, batch, frame_rate=125, threshold=0.0):
    """A batch of recordings through change detection + clustering with the device work of the
    WHOLE batch in four launches: packed upload + statistics, one growing-window launch over the
    chains of all recordings (one CTA per chain), one clustering launch (one CTA per recording).
    ``batch``: list of (recipe_lines, frames).  Returns [(segmentation recipe text, clustered
    recipe text, summary)] - byte-identical to ``diarize_recording`` on each item (the packed
    statistics restart per recording; the host replay is the same code)."""
    from . import change_detection as pcd, clustering as pcl
    pack = ctx.upload_batch([frames() if callable(frames) else frames for _, frames in batch])
    try:
        views = [pack.view(r) for r in range(len(batch))]
        parsed = [recipe_mod.parse(lines) for lines, _ in batch]
        dets = [pcd.Detector(frame_rate, threshold=threshold, ctx=ctx, **D2_CHANGE) for _ in batch]
        # ---- change detection: every chain of every recording in one launch ----
        groups = [det.gw_chains(rec, lambda l, v=v: v) for det, rec, v in zip(dets, parsed, views)]
        flat = [(r, k) for r, grp in enumerate(groups) for k in range(len(grp))]
        chains = [[] for _ in batch]
        for r, k in flat:
            chains[r].extend(groups[r][k][1])
        d0 = dets[0]
        results = pack.gw_run_batch(chains, d0.rate, d0.winsize, d0.winstep, d0.deltaws, d0.threshold,
                                    d0.lambdac, d0.metric)
        seg_lines = []
        for r, (det, rec, v) in enumerate(zip(dets, parsed, views)):
            win, first = results[r]
            c0 = 0
            for feat, ch in groups[r]:                       # usually one group: one wav per recipe
                lo, hi = int(first[c0]), int(first[c0 + len(ch)])
                sub = win[lo:hi].copy()
                sub['chain'] -= c0
                det.prefetch(feat, ch, (sub, first[c0:c0 + len(ch) + 1] - first[c0]))
                c0 += len(ch)
            seg = io.StringIO()
            det.detect_changes(rec, seg, loader=lambda l, v=v: v)
            seg_lines.append(seg.getvalue().splitlines(True))
        # ---- clustering: one problem per recording, one launch ----
        cls = [pcl.Clusterer(frame_rate, variant=1, threshold=threshold, ctx=ctx, **D2_CLUSTER) for _ in batch]
        seg_parsed = [recipe_mod.parse(lines) for lines in seg_lines]
        problems = [cl.initial_segments(rec, v.n) for cl, rec, v in zip(cls, seg_parsed, views)]
        live = [r for r, p in enumerate(problems) if p]
        merged = pack.cluster_batch([problems[r] for r in live], cls[0].metric, cls[0].lambdac,
                                    threshold, 0, 1) if live else []
        out = []
        for r, (cl, rec, v) in enumerate(zip(cls, seg_parsed, views)):
            if r in live:
                cl.prefetch(v, problems[r], merged[live.index(r)])
            clu = io.StringIO()
            cl.process_recipe(rec, clu, loader=lambda l, v=v: v)
            summary = dict(turns=len(seg_lines[r]), speakers=len(cl.speakers), windows=dets[r].windows_visited,
                           merges=len(cl.merges))
            out.append((''.join(seg_lines[r]), clu.getvalue(), summary))
        return out
    finally:
        pack.close()


def run_corpus(items, rank=0, world=1, device=None, outdir=None, frame_rate=125, runner=None,
               gather=None, batch=0):
    """Diarize the items this rank owns.

    ``items``  list of (name, recipe_lines, frames) with ``frames`` a float32
               (n, 39) array or a callable returning one (lazy file reads);
    ``runner`` overrides the per-recording function (the CPU tests inject one;
               the default uploads the frames and calls ``diarize_recording``);
    ``gather`` ``gather(obj) -> list of every rank's obj`` (e.g. built on
               ``torch.distributed.all_gather_object``); None = single process.
    ``batch``  recordings per device batch (``diarize_batch``: the device work of the whole
               batch in four launches); 0 = one recording at a time.

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
            for b0 in range(0, len(mine), batch):
                part = [items[k] for k in mine[b0:b0 + batch]]
                got = diarize_batch(ctx, [(lines, frames) for _, lines, frames in part], frame_rate)
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
