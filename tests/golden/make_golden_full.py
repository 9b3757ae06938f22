#!/usr/bin/env python3
"""Generate tests/golden/full/*.json: the BASELINE.json configurations at (or near) FULL size,
run through the REFERENCE's own scripts (oracle/ref_exec.py) and through the oracle.

Run in the build container only (needs /root/reference; about 35 minutes on 8 cores):

    python tests/golden/make_golden_full.py all          # every job, 7 at a time, then `merge`
    python tests/golden/make_golden_full.py job NAME ARM # one job (ARM = ref | ora)
    python tests/golden/make_golden_full.py merge        # /tmp partial results -> tests/golden/full/

Jobs:

  c2_bic, c2_glr, c2_kl2     config 2, the WHOLE hour (360,000 frames, seed 1002): growing-window
                             search with spk-diarization2.py's flags; GLR -t 1500, KL2 -t 4000
  c3_cl1_200 ... c3_cl2_400  config 3 cut to its first 200 / 400 segments: spk-clustering.py and
                             spk-clustering2.py -m hi -l 1.3; c3_cl1_800: 800 segments, spk-clustering.py
                             (about 80 minutes of CPU per arm: the reference grows like N^2.7)
  c4_f0, c4_f1, c4_f2        three ten-minute files of config 4 through BOTH stages
                             (D2:122-128: gw BIC change detection, then CL1 -m hi -l 1.3)

Arm `ref` executes the reference script (recipe + stdout); arm `ora` runs the oracle with its trace
switched on: every growing-window record (start, end, maxi, maxd, positive, maxi_fine, maxd_fine) with
the winner-minus-runner-up gaps, every merge (a, b, d) with the gap to the runner-up pair.  `merge`
REQUIRES the oracle's recipe and stdout to equal the reference's byte for byte before it writes a
fixture - the full-size pin of the oracle - and stores both the reference text and the oracle records.
The GPU tests (tests/test_gpu_golden_full.py) then require positions / merges bit-identical and
distances to 1e-9, and print the margins in units of that tolerance."""

import io
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, 'full')
TMP = os.environ.get('GOLDEN_FULL_TMP', '/tmp/golden_full')

GW = ['-f', '100', '-m', 'gw', '-w', '1.0', '-st', '3.0', '-dws', '0.1']
JOBS = {
    'c2_bic': ('cd', GW + ['-d', 'BIC', '-l', '1.0']),
    'c2_glr': ('cd', GW + ['-d', 'GLR', '-t', '1500']),
    'c2_kl2': ('cd', GW + ['-d', 'KL2', '-t', '4000']),
    'c3_cl1_200': ('cl', 1, 200), 'c3_cl2_200': ('cl', 2, 200),
    'c3_cl1_400': ('cl', 1, 400), 'c3_cl2_400': ('cl', 2, 400),
    'c3_cl1_800': ('cl', 1, 800),
    'c4_f0': ('d2', 0), 'c4_f1': ('d2', 1), 'c4_f2': ('d2', 2),
}
CL_FLAGS = ['-f', '100', '-m', 'hi', '-l', '1.3']
SCRIPT = {0: 'spk-change-detection.py', 1: 'spk-clustering.py', 2: 'spk-clustering2.py'}


def _case(name, tmp):
    """-> (wav name, recording, recipe lines) of a job, feature file written under tmp/fea."""
    import hashlib
    import spkdiar                                   # noqa: F401
    from spkdiar import synth
    job = JOBS[name]
    if job[0] == 'cd':
        rec = synth.config2()
        lines = synth.one_line_recipe('/syn/c2.wav', rec)
        wav = 'c2'
    elif job[0] == 'cl':
        full = synth.config3()
        nseg = job[2]
        cut = full.turns[nseg - 1][1]
        rec = synth.Recording(full.frames[:cut].copy(), full.turns[:nseg], full.rate)
        lines = synth.turn_recipe('/syn/c3.wav', rec)
        wav = 'c3'
    else:
        rec = synth.config4_file(job[1])
        wav = 'c4_%d' % job[1]
        lines = synth.one_line_recipe('/syn/%s.wav' % wav, rec)
    rpath, feadir = synth.write_case(tmp, wav, rec, lines)
    return wav, rec, rpath, feadir, hashlib.sha256(rec.frames.tobytes()).hexdigest()


def _run(arm, kind, variant, argv, trace=None):
    """One script run through the chosen arm -> stdout text."""
    if arm == 'ref':
        from oracle import ref_exec
        return ref_exec.run(SCRIPT[variant if kind == 'cl' else 0], argv)[0]
    from oracle import change_detection as ocd, clustering as ocl
    out = io.StringIO()
    if kind == 'cd':
        ocd.main(argv, stdout=out, trace=trace)
    else:
        ocl.main(argv, stdout=out, variant=variant, trace=trace)
    return out.getvalue()


def _records(trace):
    cols = {k: [] for k in ('start', 'end', 'maxi', 'maxd', 'positive', 'maxi_fine', 'maxd_fine', 'gap',
                            'gap_fine')}
    for r in trace:
        for k in cols:
            v = r.get(k)
            cols[k].append(None if v is None else (bool(v) if k == 'positive' else float(v)))
    return cols


def job(name, arm):
    import numpy
    import scipy
    spec = JOBS[name]
    t0 = time.time()
    res = dict(name=name, arm=arm, numpy=numpy.__version__, scipy=scipy.__version__)
    with tempfile.TemporaryDirectory() as tmp:
        wav, rec, rpath, feadir, sha = _case(name, tmp)
        res['frames_sha256'] = sha
        res['frames'] = int(rec.frames.shape[0])
        outp = os.path.join(tmp, 'out.recipe')
        if spec[0] == 'cd':
            trace = [] if arm == 'ora' else None
            res['flags'] = spec[1]
            res['stdout'] = _run(arm, 'cd', 0, [rpath, feadir, '-o', outp] + spec[1], trace).replace(tmp, '<TMP>')
            res['recipe'] = open(outp).read()
            if trace is not None:
                res['windows'] = _records(trace)
        elif spec[0] == 'cl':
            trace = [] if arm == 'ora' else None
            res['flags'] = CL_FLAGS
            res['variant'] = spec[1]
            res['recipe_in'] = open(rpath).read()
            res['stdout'] = _run(arm, 'cl', spec[1], [rpath, feadir + '/', '-o', outp] + CL_FLAGS,
                                 trace).replace(tmp, '<TMP>')
            res['recipe'] = open(outp).read()
            if trace is not None:
                res['merges'] = [list(m) for m in trace]
        else:
            cdflags = GW + ['-d', 'BIC', '-l', '1.0']
            mid = os.path.join(tmp, 'turns.recipe')
            tr1 = [] if arm == 'ora' else None
            tr2 = [] if arm == 'ora' else None
            res['flags_cd'], res['flags_cl'] = cdflags, CL_FLAGS
            res['recipe_in'] = open(rpath).read()
            res['stdout_cd'] = _run(arm, 'cd', 0, [rpath, feadir, '-o', mid] + cdflags, tr1).replace(tmp, '<TMP>')
            res['recipe_cd'] = open(mid).read()
            res['stdout_cl'] = _run(arm, 'cl', 1, [mid, feadir + '/', '-o', outp] + CL_FLAGS, tr2).replace(tmp, '<TMP>')
            res['recipe'] = open(outp).read()
            if arm == 'ora':
                res['windows'] = _records(tr1)
                res['merges'] = [list(m) for m in tr2]
    res['seconds'] = round(time.time() - t0, 1)
    os.makedirs(TMP, exist_ok=True)
    with open(os.path.join(TMP, '%s.%s.json' % (name, arm)), 'w') as f:
        json.dump(res, f)
    print(name, arm, 'done in %.0f s' % res['seconds'], flush=True)


TEXT_KEYS = ('stdout', 'recipe', 'stdout_cd', 'recipe_cd', 'stdout_cl')


def merge():
    os.makedirs(OUT, exist_ok=True)
    for name in sorted(JOBS):
        parts = {}
        for arm in ('ref', 'ora'):
            p = os.path.join(TMP, '%s.%s.json' % (name, arm))
            if os.path.isfile(p):
                parts[arm] = json.load(open(p))
        if len(parts) < 2:
            print(name, 'incomplete:', sorted(parts))
            continue
        ref, ora = parts['ref'], parts['ora']
        assert ref['frames_sha256'] == ora['frames_sha256']
        for k in TEXT_KEYS:
            if k in ref:
                assert ref[k] == ora[k], '%s: the oracle does not reproduce the reference (%s)' % (name, k)
        fix = dict(ora)
        del fix['arm']
        fix['seconds_reference'] = ref['seconds']
        fix['seconds_oracle'] = ora['seconds']
        del fix['seconds']
        fix['pinned'] = 'reference script executed by oracle/ref_exec.py; oracle text identical byte for byte'
        with open(os.path.join(OUT, name + '.json'), 'w') as f:
            json.dump(fix, f, indent=None, separators=(',', ':'), sort_keys=True)
        print(name, 'ok: reference %.0f s, oracle %.0f s' % (ref['seconds'], ora['seconds']))


def run_all(par=7):
    env = dict(os.environ, OPENBLAS_NUM_THREADS='1', OMP_NUM_THREADS='1', MKL_NUM_THREADS='1')
    # longest first
    order = ['c3_cl1_800', 'c3_cl1_400', 'c3_cl2_400', 'c2_kl2', 'c2_bic', 'c2_glr', 'c3_cl1_200', 'c3_cl2_200',
             'c4_f0', 'c4_f1', 'c4_f2']
    todo = [(n, a) for n in order for a in ('ref', 'ora')
            if not os.path.isfile(os.path.join(TMP, '%s.%s.json' % (n, a)))]
    running = []
    while todo or running:
        running = [p for p in running if p.poll() is None]
        while todo and len(running) < par:
            n, a = todo.pop(0)
            running.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), 'job', n, a], env=env))
        time.sleep(2)
    merge()


if __name__ == '__main__':
    cmd = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if cmd == 'job':
        job(sys.argv[2], sys.argv[3])
    elif cmd == 'merge':
        merge()
    else:
        run_all()
