#!/usr/bin/env python3
"""bench.py - the headline benchmark: audio-hours/sec diarized.

Workload (BASELINE.json configs[1], SURVEY.md section 8d "config 2"): ONE
synthetic 1-hour recording (360,000 feacat-format MFCC frames, d = 39, 100 fps,
K = 8 speakers, turns 3..19 s, seed 1002).  One STEP is the whole hot path over
that recording:

  1. frame statistics (fp64 prefix sums)                         K1
  2. growing-window change detection with BIC  (spk-diarization2's flags:
     -m gw -d BIC -w 1.0 -st 3.0 -dws 0.1 -l 1.0)                K3
  3. the same search with GLR (-t GLR_T) and with KL2 (-t KL2_T)  K3
  4. agglomerative BIC clustering of the BIC turns
     (spk-clustering.py -m hi -l 1.3)                            K6 / K7

and counts as ONE audio-hour diarized (the GLR and KL2 passes are extra work
the config names, not extra audio).  `value` runs it with the frames already in
HBM through the low-level C-ABI; `e2e` runs it through the public drop-in API
(`Detector` / `Clusterer`, recipe text in, recipe text out) with the frames in
pinned HOST memory, the host->device copy and the host-side replay inside the
timed region.

N > 1 (torchrun): recordings are independent, so every rank diarizes its own
copy of the 1-hour recording (seed 1002) with no data-path collective - weak scaling.

`--impl reference` times the reference's own CPU implementation (the oracle: the
reference scripts are Python 2 and cannot run here; the oracle is their
Python-3 restatement, pinned against them - see oracle/__init__.py) on a bounded
sample of the same workload.
"""

import argparse
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RATE = 100
FRAMES = 360000
GW_FLAGS = dict(winsize=1.0, winstep=3.0, deltaws=0.1)
# thresholds tuned once on seed 1002 with the script's own end-of-run metrics so that the
# number of detected changes is about the number of true turns, then frozen
GLR_T = 1500.0
KL2_T = 4000.0
CPU_SAMPLE_SECONDS = 150            # audio seconds of the CPU-baseline sample
SEQUENTIAL = bool(os.environ.get('SPKDIAR_BENCH_SEQUENTIAL'))      # one search after the other (for comparison)

# algorithmic work per unit (restated in DESIGN.md)
FLOP_LOGDET = 22893.0               # form 39x39 from prefix differences + factorise
FLOP_KL2 = 85332.0
BYTES_STATS_PER_FRAME = 156 + 6560          # frames read once, one prefix record written per frame


def make_recording(rank):
    """Every rank gets the SAME synthetic recording (seed 1002): weak scaling with identical work per
    GPU, so that the per-N values measure the machine and not the luck of a seed."""
    import spkdiar                                   # noqa: F401
    from spkdiar import synth
    return synth.make_recording(1002, FRAMES, 8, rate=RATE)


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[k] for r in self.rows if len(r) >= 9 for k in range(4)
                          if r[5 + k].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def segments_from_windows(win, nframes):
    """Turn boundaries (frames) from the growing-window records of one chain."""
    cuts = [0.0]
    for r in win:
        if r['positive']:
            cuts.append(float(r['start']) + float(r['maxi_fine']))
    cuts.append(float(nframes))
    a = [int(c) for c in cuts[:-1]]
    b = [int(c) for c in cuts[1:]]
    return a, b


def device_step(ctx, dev_ptr, nframes):
    """One step with the frames resident in HBM (low-level C-ABI)."""
    from spkdiar import _abi
    W, ST, DW = GW_FLAGS['winsize'] * RATE, GW_FLAGS['winstep'] * RATE, float(int(RATE * GW_FLAGS['deltaws']))
    feat = ctx.adopt(dev_ptr, nframes)
    out = {}
    try:
        # the three searches are independent dependent chains: side by side on disjoint SM subsets
        runs = [dict(rate=float(RATE), winsize=W, winstep=ST, deltaws=DW, threshold=thr, lambdac=1.0, metric=met)
                for met, thr in ((_abi.BIC, 0.0), (_abi.GLR, GLR_T), (_abi.KL2, KL2_T))]
        def cluster_bic():
            sa, sb = segments_from_windows(out['BIC'], nframes)
            with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
                out['merges'] = cl.run(0.0, 0, 1)[0]
            out['nseg'] = len(sa)
        if SEQUENTIAL:
            for name, r in zip(('BIC', 'GLR', 'KL2'), runs):
                out[name] = feat.gw_run([0], [nframes], **r)[0]
            cluster_bic()
        else:
            # all three searches are launched together; the BIC one (split into sub-chains, done first) is
            # collected and its turns are clustered on the SMs it ran on while the KL2 chain is still running
            with feat.gw_multi_begin([0], [nframes], runs) as h:
                out['BIC'] = h.wait(0)[0]
                ctx.exec_on(*h.where(0))
                try:
                    cluster_bic()
                finally:
                    ctx.exec_on()
                out['GLR'] = h.wait(1)[0]
                out['KL2'] = h.wait(2)[0]
    finally:
        feat.close()
    return out


def d2_step(ctx, dev_ptr, nframes):
    """The part of a step spk-diarization2.py itself runs (lines 122-128): growing-window BIC
    change detection + CL1 clustering, frames resident."""
    from spkdiar import _abi
    W, ST, DW = GW_FLAGS['winsize'] * RATE, GW_FLAGS['winstep'] * RATE, float(int(RATE * GW_FLAGS['deltaws']))
    feat = ctx.adopt(dev_ptr, nframes)
    try:
        win, _ = feat.gw_run([0], [nframes], float(RATE), W, ST, DW, 0.0, 1.0, _abi.BIC)
        sa, sb = segments_from_windows(win, nframes)
        with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
            merges, _ = cl.run(0.0, 0, 1)
    finally:
        feat.close()
    return len(win), len(merges)


def measured_fp64_peak():
    """fp64 FMA peak of this GPU from tests/micro/bench_dfma (8 independent DFMA chains per
    thread, 2 flop each; built by __graft_entry__.build()); the nominal figure when the
    binary is missing."""
    exe = os.path.join(ROOT, 'tests', 'micro', 'bench_dfma')
    try:
        out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
        return float(json.loads(out)['fp64_fma_tflops']), 'measured in this run (tests/micro/bench_dfma: independent DFMA chains)'
    except (OSError, ValueError, KeyError, subprocess.SubprocessError):
        return 37.2, 'nominal 148 SM x 64 DFMA/clk x 1.965 GHz'


def e2e_step(ctx, host_frames, recipe_lines):
    """One step through the public drop-in API with HOST buffers: recipe text in,
    recipe text out; H2D copy, device work, D2H of the records and the host-side
    replay are all inside."""
    from spkdiar import change_detection as pcd, clustering as pcl, recipe
    parsed = recipe.parse(recipe_lines)
    feat = ctx.upload_ptr(host_frames.data_ptr(), host_frames.shape[0])
    texts = {}
    try:
        names = ('BIC', 'GLR', 'KL2')
        dets = [pcd.Detector(RATE, 'gw', name, GW_FLAGS['winsize'], GW_FLAGS['winstep'], GW_FLAGS['deltaws'],
                             thr, 1.0, ctx=ctx) for name, thr in zip(names, (0.0, GLR_T, KL2_T))]
        bufs = [io.StringIO() for _ in names]
        def cluster_bic(det, where):
            # spk-clustering.py on the recipe the BIC detector has just written - queued on the stream and
            # SMs its search ran on, while the GLR and KL2 searches are still running
            if where is not None:
                ctx.exec_on(*where)
            try:
                cl = pcl.Clusterer(RATE, 1, 'hi', 'BIC', 0.0, 0, 1.3, ctx=ctx)
                buf = io.StringIO()
                cl.process_recipe(recipe.parse(bufs[0].getvalue().splitlines(True)), buf, loader=lambda l: feat)
                texts['clusters'] = buf.getvalue()
            finally:
                ctx.exec_on()
        if SEQUENTIAL:
            for det, buf in zip(dets, bufs):
                det.detect_changes(parsed, buf, loader=lambda l: feat)
            cluster_bic(dets[0], None)
        else:
            pcd.detect_changes_multi(dets, parsed, bufs, loader=lambda l: feat, after={0: cluster_bic})
        for name, buf in zip(names, bufs):
            texts[name] = buf.getvalue()
    finally:
        feat.close()
    return texts


def cpu_oracle_sample(seconds, rank=0, parallel=False, replicas=1):
    """The reference's CPU path (oracle) on the first `seconds` of the recording:
    the same four passes.  ``parallel``: the three searches of a recording as three processes;
    ``replicas``: that many recordings side by side (the N-GPU line diarizes N recordings at a time, so the
    reference arm at N runs 3 N processes).  -> (audio_seconds of all replicas, cpu_wall_seconds, processes)."""
    import warnings
    warnings.simplefilter('ignore')
    rec = make_recording(rank)
    n = int(seconds * RATE)
    x = rec.frames[:n]
    jobs = [('BIC', 0.0, True), ('GLR', GLR_T, False), ('KL2', KL2_T, False)]
    t0 = time.perf_counter()
    if parallel:
        import multiprocessing as mp
        nproc = len(jobs) * replicas
        with mp.get_context('fork').Pool(nproc) as pool:
            pool.starmap(_cpu_job, [(x, j) for _ in range(replicas) for j in jobs])
        cores = nproc
    else:
        for j in jobs:
            _cpu_job(x, j)
        cores = 1
    return replicas * n / float(RATE), time.perf_counter() - t0, cores


def _cpu_job(x, job):
    from oracle import change_detection as ocd, clustering as ocl
    name, thr, then_cluster = job
    n = x.shape[0]
    cd = ocd.ChangeDetection(RATE, 'gw', name, GW_FLAGS['winsize'], GW_FLAGS['winstep'], GW_FLAGS['deltaws'],
                             thr, 1.0)
    out = io.StringIO()
    line = ('/syn/c2.wav', 'a_1', 0.0, n / float(RATE))
    cd.dist_gw(x, line, out)
    if then_cluster:
        recipe = ocd.parse_recipe(out.getvalue().splitlines(True), lambda *a: None)
        cl = ocl.Clustering(RATE, 1, 'hi', 'BIC', 0.0, 0, 1.3)
        cl.process_recipe(recipe, io.StringIO(), loader=lambda rl: (39, x))


def run_reference(args):
    """The reference arm.  Like for like with our line at N GPUs (N recordings diarized at a time): N replicas of
    the bounded sample, each as three processes (the reference is single-process Python; its three searches are
    independent scripts), i.e. 3 N host processes - capped by the host's cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    ncpu = os.cpu_count() or 1
    replicas = max(1, min(args.gpus, ncpu // 3))
    vals = []
    for k in range(args.warmup + args.steps):
        audio_s, wall, cores = cpu_oracle_sample(args.cpu_seconds, 0, parallel=True, replicas=replicas)
        if k >= args.warmup:
            vals.append((audio_s / 3600.0) / wall)
    v = sum(vals) / len(vals)
    sample = ('first %d s of the 1-hour recording: oracle gw BIC -> CL1 clustering, gw GLR, gw KL2 as 3 parallel '
              'processes per recording x %d recordings side by side = %d processes on %d host cores (the reference is '
              'single-process Python; 1 BLAS thread each).  same_config holds for the LABEL: the sample is the first '
              '%d s, which favours the reference (its clustering grows like N^2.7 with the number of turns)'
              % (args.cpu_seconds, replicas, 3 * replicas, ncpu, args.cpu_seconds))
    print(json.dumps({
        'impl': 'reference', 'metric': 'audio_hours_per_sec_diarized', 'value': v, 'unit': 'audio-hours/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * (replicas * args.cpu_seconds / 3600.0) / v, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(),
        'cpu_baseline': {'value': v, 'unit': 'audio-hours/s', 'cores': 3 * replicas, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'audio-hours/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config():
    return {'workload': 'config2: 1-hour synthetic recording, 360000x39 fp32 frames @100fps, growing-window '
                        'BIC+GLR+KL2 change detection + CL1 BIC clustering of the BIC turns',
            'frames': FRAMES, 'dim': 39, 'frame_rate': RATE, 'glr_threshold': GLR_T, 'kl2_threshold': KL2_T,
            'l2': 'inputs larger than L2: 2.36 GB of prefix records rewritten and re-read every step (L2 = 126 MB)',
            'parallelism': 'one recording per GPU (every rank the same synthetic recording), no data-path collective'}


def other_cpu_sample(workload):
    """A bounded sample of the reference's CPU path (the oracle, one process) for the other workloads.
    config4: ONE ten-minute recording through spk-diarization2.py's two calls.  config3: the reference's
    clustering grows like N^2.7, the full 1,978 segments would take hours - the
    first 120 segments of the same recording are clustered and the audio they cover is what counts."""
    import warnings
    warnings.simplefilter('ignore')
    import spkdiar                                   # noqa: F401
    from spkdiar import synth
    from oracle import change_detection as ocd, clustering as ocl
    t0 = time.perf_counter()
    if workload == 'config4':
        r = synth.config4_file(0)
        n = r.frames.shape[0]
        cd = ocd.ChangeDetection(RATE, 'gw', 'BIC', 1.0, 3.0, 0.1, 0.0, 1.0)
        out = io.StringIO()
        cd.dist_gw(r.frames, ('/syn/c4_0.wav', 'a_1', 0.0, n / float(RATE)), out)
        recipe = ocd.parse_recipe(out.getvalue().splitlines(True), lambda *a: None)
        cl = ocl.Clustering(RATE, 1, 'hi', 'BIC', 0.0, 0, 1.3)
        cl.process_recipe(recipe, io.StringIO(), loader=lambda rl: (39, r.frames))
        audio_s, sample = n / float(RATE), 'one ten-minute recording of the corpus: oracle gw BIC (D2 flags) + CL1 clustering'
    else:
        r = synth.config3()
        turns = r.turns[:120]
        recipe = [('/syn/c3.wav', 'a_%d' % (k + 1), t[0] / float(RATE), t[1] / float(RATE)) for k, t in enumerate(turns)]
        cl = ocl.Clustering(RATE, 1, 'hi', 'BIC', 0.0, 0, 1.3)
        cl.process_recipe(recipe, io.StringIO(), loader=lambda rl: (39, r.frames))
        audio_s = (turns[-1][1] - turns[0][0]) / float(RATE)
        sample = ('the first 120 of the 1,978 segments: oracle CL1 clustering (the reference grows like N^2.7: the '
                  'full recording is far slower per audio-hour than this sample)')
    wall = time.perf_counter() - t0
    return {'value': (audio_s / 3600.0) / wall, 'unit': 'audio-hours/s', 'cores': 1, 'kind': 'port', 'sample': sample,
            'seconds': wall}


def run_other_workload(args):
    """The other BASELINE.json configurations (not the headline): one JSON line each, same
    keys.  config3: agglomerative clustering of a 3-hour recording per GPU (weak scaling).
    config4: a corpus of 10-minute recordings, file-sharded, change detection + clustering
    through the drop-in API with host buffers (weak scaling).  config5: ONE long recording,
    its pair matrix sharded over the GPUs, one 16-byte all-gather per merge (strong scaling)."""
    import numpy as np
    import torch
    import spkdiar                                   # noqa: F401
    from spkdiar import _abi, synth, corpus, sharded
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ctx = _abi.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    W = max(args.warmup, 1 if args.workload == 'config5' else 3)     # a full-size config5 step takes seconds

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == 'config3':
        rec = synth.make_recording(1003 + rank, 1080000, 10, turn_lo=2, turn_hi=9)
        sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
        feat = ctx.upload_frames(rec.frames)         # clustering only: no window statistics (K5 records from the frames)
        hours_per_step, scaling = world * 3.0, 'weak'
        desc = 'config3: CL1 agglomerative BIC clustering of a 3-hour recording (%d segments) per GPU' % len(sa)

        def step():
            with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
                return cl.run(0.0, 0, 1)[0]
    elif args.workload == 'config5':
        rec = synth.config5(n_frames=args.segments * 173)
        sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
        feat = ctx.upload_frames(rec.frames)         # clustering only: no window statistics (K5 records from the frames)
        mbx = sharded.Mailboxes(ctx) if world > 1 else None
        hours_per_step, scaling = rec.frames.shape[0] / RATE / 3600.0, 'strong'
        desc = ('config5 (scaled): one %.1f-hour recording, %d segments, pair matrix dealt over %d GPU(s), one '
                'persistent kernel per GPU, 16-byte candidates exchanged through peer-memory mailboxes per merge'
                % (hours_per_step, len(sa), world))

        def step():
            with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
                if world > 1:
                    return cl.run_sharded_p2p(0.0, 0, rank, world, mbx.ptrs, mbx.next_base(len(sa)))[0]
                return cl.run(0.0, 0, 1)[0]
    else:
        items = []
        for k in range(args.files):
            r = synth.config4_file(rank * args.files + k)
            items.append(('c4_%d' % k, synth.one_line_recipe('/syn/c4_%d.wav' % k, r), r.frames))
        hours_per_step, scaling = world * args.files * (60000 / RATE / 3600.0), 'weak'
        desc = ('config4: %d ten-minute recordings per GPU, spk-diarization2 flags (gw BIC change detection + CL1), '
                'host frames through the drop-in API, %s' % (args.files, 'device batches of %d recordings (packed '
                'statistics, one growing-window launch with one CTA per recording, one clustering launch)%s' % (args.batch, ', device stages of the next batch overlapped with the host replay' if args.overlap else '')
                if args.batch > 0 else 'one recording at a time'))
        pinned = [torch.from_numpy(frames).pin_memory() for _, _, frames in items]

        def step():
            out = None
            if args.batch > 0:
                parts = [[(items[k][1], (pinned[k].data_ptr(), pinned[k].shape[0]))
                          for k in range(b0, min(b0 + args.batch, len(items)))]
                         for b0 in range(0, len(items), args.batch)]
                if args.overlap:
                    for got in corpus.diarize_batches(ctx, parts, RATE):
                        out = got[-1]
                else:
                    for part in parts:
                        out = corpus.diarize_batch(ctx, part, RATE)[-1]
                return out
            for name, lines, frames in items:
                f = ctx.upload(frames)
                try:
                    out = corpus.diarize_recording(ctx, lines, lambda l: f, RATE)
                finally:
                    f.close()
            return out
    for _ in range(W):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    barrier()
    value = args.steps * hours_per_step / (float(ms.item()) / 1e3)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline and args.workload in ('config3', 'config4'):
        cpu = other_cpu_sample(args.workload)
    if rank == 0:
        print(json.dumps({'metric': 'audio_hours_per_sec_diarized', 'value': value, 'unit': 'audio-hours/s',
                          'n_gpus': world, 'steps': args.steps, 'warmup': W,
                          'ms_per_step': float(ms.item()) / args.steps, 'higher_is_better': True,
                          'scaling': scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                          'config': {'workload': desc}, 'cpu_baseline': cpu,
                          'result': {'last': (len(res) if hasattr(res, '__len__') and not isinstance(res, tuple)
                                              else res[2] if isinstance(res, tuple) else None)}}))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


C4_FILES = 1000                 # the FIXED corpus of the config4 sub-record: BASELINE config 4, recordings 0..999 of synth.config4_file
C4_BATCH = 148                  # recordings per device batch (one CTA per recording in the growing-window launch)
C5_SEGMENTS = 23940             # the FIXED long recording of the config5 sub-record (BASELINE config 5 scaled by 1/2)


def sub_config4(ctx, rank, world, timed, steps=2):
    """BASELINE config 4 as it shards: a FIXED corpus of 1,000 ten-minute recordings, dealt to the ranks by file
    index (recording k -> rank k % N, spk-diarization2.py:122-128 once per file, no communication), through the
    drop-in corpus driver with HOST buffers: pinned frames in, recipe text out, device batches in two lanes that
    overlap each other (copies of one batch under the kernels of the other), host replay in native code.
    Strong scaling: the corpus does not grow with N."""
    import multiprocessing as mp
    import torch
    from spkdiar import corpus, synth
    mine = corpus.shard(C4_FILES, rank, world)
    # the synthetic recordings are generated on the host cores this rank may use (0.2 s each on one core)
    workers = max(1, min(12, (os.cpu_count() or 1) // max(world, 1)))
    if workers > 1 and len(mine) > 8:
        with mp.get_context('spawn').Pool(workers) as pool:
            frames = pool.map(synth.config4_frames, mine, chunksize=4)
    else:
        frames = [synth.config4_frames(k) for k in mine]
    items = []
    for k, x in zip(mine, frames):
        lines = ['audio=/syn/c4_%d.wav lna=a_1 start-time=0.0 end-time=%s\n' % (k, repr(x.shape[0] / float(RATE)))]
        items.append((lines, torch.from_numpy(x).pin_memory()))
    del frames
    # two lanes: at least two batches per rank whenever there is more than one recording
    batch = max(1, min(C4_BATCH, (len(items) + 1) // 2))
    parts = [[(lines, (t.data_ptr(), t.shape[0])) for lines, t in items[b0:b0 + batch]]
             for b0 in range(0, len(items), batch)]

    def step():
        out = None
        for got in corpus.diarize_batches(ctx, parts, RATE):
            out = got
        return out
    step()
    ms, last = timed(step, steps)
    hours = C4_FILES * (60000 / RATE / 3600.0)
    return {'what': 'FIXED corpus of %d ten-minute recordings (BASELINE config 4), recording k on rank k %% N, gw BIC change '
                    'detection + CL1 clustering (spk-diarization2.py flags) through corpus.diarize_batches: pinned host '
                    'frames in, recipe text out, device batches of <= %d recordings in two overlapping lanes, host replay '
                    'in native code (spkdiar_replay_*)' % (C4_FILES, batch),
            'scaling': 'strong', 'recordings': C4_FILES, 'recordings_this_rank': len(mine), 'steps': steps,
            'ms_per_step': ms / steps, 'value': steps * hours / (ms / 1e3), 'unit': 'audio-hours/s',
            'h2d_bytes_per_step': C4_FILES * 60000 * 39 * 4,
            'last_recording': last[-1][2] if last else None}


def sub_config5(ctx, rank, world, timed, dist, steps=2):
    """BASELINE config 5 as it shards: ONE fixed long recording (%d segments), pair (r, c) of the matrix on rank
    (r + c) %% N, one persistent kernel per GPU, 16-byte candidates through peer-memory mailboxes per merge
    (spk-clustering.py:201-237).  Strong scaling.  The merge sequence is hashed so that the lines of different N
    can be compared.""" % C5_SEGMENTS
    import hashlib
    from spkdiar import _abi, sharded, synth
    rec = synth.config5(n_frames=C5_SEGMENTS * 173)
    sa, sb = [t[0] for t in rec.turns], [t[1] for t in rec.turns]
    feat = ctx.upload_frames(rec.frames)         # clustering only: no window statistics (K5 records from the frames)
    mbx = sharded.Mailboxes(ctx) if world > 1 else None
    box = {}

    def step():
        with feat.cluster(sa, sb, _abi.BIC, 1.3) as cl:
            if world > 1:
                m = cl.run_sharded_p2p(0.0, 0, rank, world, mbx.ptrs, mbx.next_base(len(sa)))[0]
            else:
                m = cl.run(0.0, 0, 1)[0]
            box['counters'] = cl.counters()
        return m
    try:
        step()
        ms, merges = timed(step, steps)
    finally:
        if mbx is not None:
            mbx.close()
        feat.close()
    hours = rec.frames.shape[0] / RATE / 3600.0
    return {'what': 'ONE fixed %.1f-hour recording, %d segments, CL1 -m hi -l 1.3; pair matrix dealt over N GPUs, one '
                    'persistent kernel per GPU, candidates exchanged through peer-memory mailboxes' % (hours, len(sa)),
            'scaling': 'strong', 'segments': len(sa), 'steps': steps, 'ms_per_run': ms / steps,
            'value': steps * hours / (ms / 1e3), 'unit': 'audio-hours/s', 'merges': int(len(merges)),
            'merge_sequence_sha256': hashlib.sha256(merges.tobytes()).hexdigest()[:16],
            'cycles_per_merge_cta0': box.get('counters')}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-seconds', type=int, default=CPU_SAMPLE_SECONDS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='config2', choices=['config2', 'config3', 'config4', 'config5'],
                    help='config2 is the headline (BASELINE.json configs[1]); the others print their own line')
    ap.add_argument('--segments', type=int, default=8000, help='config5: segments of the long recording')
    ap.add_argument('--files', type=int, default=148, help='config4: recordings per GPU')
    ap.add_argument('--batch', type=int, default=74, help='config4: recordings per device batch (0: one at a time)')
    ap.add_argument('--no-overlap', dest='overlap', action='store_false',
                    help='config4 with --batch: do not overlap device stages and host replay')
    ap.add_argument('--no-sharded', action='store_true',
                    help='skip the config4 / config5 sub-records (the two configurations that shard over the GPUs)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    if args.workload != 'config2':
        return run_other_workload(args)

    import numpy as np
    import torch
    import spkdiar                                   # noqa: F401
    from spkdiar import _abi, synth

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    rec = make_recording(rank)
    host = torch.from_numpy(rec.frames).pin_memory()
    dev = host.cuda()
    recipe_lines = synth.one_line_recipe('/syn/c2_%d.wav' % rank, rec)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = _abi.Context(local, stream=stream)
    W = max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; device time by CUDA events on
        the stream the library launches on; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(steps):
            res = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item()), res

    # ---- device-resident arm ----
    for _ in range(W):
        device_step(ctx, dev.data_ptr(), FRAMES)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    l0 = ctx.launches
    ms_total, res = timed(lambda: device_step(ctx, dev.data_ptr(), FRAMES), args.steps)
    launches = ctx.launches - l0
    prof = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    value = world * args.steps * (FRAMES / RATE / 3600.0) / (ms_total / 1e3)

    # ---- end-to-end arm (host buffers, public API) ----
    for _ in range(2):
        e2e_step(ctx, host, recipe_lines)
    ms_e2e, texts = timed(lambda: e2e_step(ctx, host, recipe_lines), args.steps)
    e2e_value = world * args.steps * (FRAMES / RATE / 3600.0) / (ms_e2e / 1e3)
    h2d = FRAMES * 39 * 4
    d2h = sum(len(res[k]) for k in ('BIC', 'GLR', 'KL2')) * 72 + len(res['merges']) * 16

    # ---- the pipeline's own flags alone (reported beside the headline, not part of it) ----
    for _ in range(2):
        d2_step(ctx, dev.data_ptr(), FRAMES)
    ms_d2, _ = timed(lambda: d2_step(ctx, dev.data_ptr(), FRAMES), args.steps)
    fp64_peak, fp64_src = measured_fp64_peak() if rank == 0 else (37.2, 'nominal')

    lt = torch.tensor([launches], device='cuda', dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(lt)

    # ---- roofline of the kernel classes (per-class CUDA-event time inside the library) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    hbm_peak, hbm_src = (peaks['hbm_gbs'], 'measured') if 'hbm_gbs' in peaks else (6650.0, 'fallback')
    steps = float(args.steps)

    def cands(win):
        coarse = int(np.sum(np.abs(win['ncand'])))
        fine = int(np.sum(win['positive'])) * int(2 * RATE / 10)
        return coarse + fine
    flops_gw = (cands(res['BIC']) * 2 * FLOP_LOGDET + cands(res['GLR']) * 3 * FLOP_LOGDET
                + cands(res['KL2']) * FLOP_KL2)
    nseg = res['nseg']
    flops_merge = (len(res['merges']) * nseg / 2.0) * (FLOP_LOGDET + 820)
    flops_score = (nseg * (nseg - 1) / 2.0 + nseg) * (FLOP_LOGDET + 820)
    kern = {
        'stats': {'bound': 'hbm', 'unit': 'GB/s', 'peak': hbm_peak, 'peak_source': hbm_src,
                  'work': 3 * FRAMES * BYTES_STATS_PER_FRAME / 3.0},      # one build per step
        'gw': {'bound': 'fp64', 'unit': 'TFLOP/s', 'peak': fp64_peak, 'peak_source': fp64_src, 'work': flops_gw},
        'score': {'bound': 'fp64', 'unit': 'TFLOP/s', 'peak': fp64_peak, 'peak_source': fp64_src, 'work': flops_score},
        'merge': {'bound': 'fp64', 'unit': 'TFLOP/s', 'peak': fp64_peak, 'peak_source': fp64_src, 'work': flops_merge},
    }
    rl_all = {}
    for k, spec in kern.items():
        ms_k, n_k = prof[k]
        if ms_k <= 0:
            continue
        per_step_s = ms_k / steps / 1e3
        scale = 1e9 if spec['unit'] == 'GB/s' else 1e12
        ach = spec['work'] / per_step_s / scale
        rl_all[k] = {'bound': spec['bound'], 'achieved': ach, 'peak': spec['peak'], 'unit': spec['unit'],
                     'frac': ach / spec['peak'], 'peak_source': spec['peak_source'],
                     'ms_per_step': ms_k / steps, 'launches_per_step': n_k / steps, 'traffic': None}
        if spec['bound'] == 'fp64':         # beside the measured DFMA peak: the nominal one (148 SM x 64 DFMA/clk x 1.965 GHz)
            rl_all[k]['peak_nominal'] = 37.2
            rl_all[k]['frac_of_nominal'] = ach / 37.2
    try:                                    # DRAM traffic per launch from this round's ncu captures
        tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
        if not os.path.isfile(tpath):
            tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        traffic = json.load(open(tpath))
        for k in rl_all:
            if k in traffic:
                rl_all[k]['traffic'] = traffic[k]['bytes_per_launch']
    except (OSError, ValueError, KeyError):
        pass
    dominant = max(rl_all, key=lambda k: rl_all[k]['ms_per_step'])
    roofline = dict(rl_all[dominant], kernel=dominant)

    # ---- the two configurations that SHARD over the GPUs, measured in the same run (sub-records) ----
    sub4 = sub5 = None
    if not args.no_sharded:
        del dev
        torch.cuda.empty_cache()
        sub4 = sub_config4(ctx, rank, world, timed)
        sub5 = sub_config5(ctx, rank, world, timed, dist)

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        audio_s, wall, cores = cpu_oracle_sample(args.cpu_seconds)
        cpu = {'value': (audio_s / 3600.0) / wall, 'unit': 'audio-hours/s', 'cores': cores, 'kind': 'port',
               'sample': 'first %d s of the same recording, same four passes, oracle (py3 restatement of the '
                         'reference scripts), one process' % args.cpu_seconds,
               'seconds': wall}

    if rank == 0:
        line = {
            'metric': 'audio_hours_per_sec_diarized', 'value': value, 'unit': 'audio-hours/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': W, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(),
            'frames_per_sec': value * 3600.0 * RATE,
            'e2e': {'value': e2e_value, 'unit': 'audio-hours/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(lt.item()),
            'clocks': clocks,
            'roofline': roofline,
            'roofline_all': rl_all,
            'cpu_baseline': cpu,
            'd2_flags_only': {'what': 'growing-window BIC + CL1 clustering alone (the two calls of spk-diarization2.py:122-128), '
                                      'frames resident; the BIC search runs split into sub-chains',
                              'ms_per_step': ms_d2 / args.steps,
                              'value': world * args.steps * (FRAMES / RATE / 3600.0) / (ms_d2 / 1e3), 'unit': 'audio-hours/s'},
            'config4': sub4,
            'config5': sub5,
            'result': {'bic_changes': int(np.sum(res['BIC']['positive'])), 'glr_changes': int(np.sum(res['GLR']['positive'])),
                       'kl2_changes': int(np.sum(res['KL2']['positive'])), 'true_turns': len(rec.turns),
                       'windows': {k: int(len(res[k])) for k in ('BIC', 'GLR', 'KL2')},
                       'segments_clustered': nseg, 'merges': int(len(res['merges'])),
                       'final_speakers': nseg - int(len(res['merges']))},
        }
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
