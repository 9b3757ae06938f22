#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the REFERENCE's own scripts.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Each fixture holds the output recipe and the stdout text the reference script
(executed through oracle/ref_exec.py, i.e. its own source with the documented
py2->py3 token rewrites) produced for one synthetic case of tests/cases.py,
plus the SHA-256 of the input frames and the numpy / scipy versions.  The tests
then require (a) the oracle to reproduce these byte for byte (CPU) and (b) the
CUDA path to reproduce the recipes byte for byte and the numbers in the log to
1e-9 (GPU)."""

import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy                         # noqa: E402
import scipy                         # noqa: E402
import cases                         # noqa: E402
from oracle import ref_exec          # noqa: E402

SCRIPT = {('cd', 0): 'spk-change-detection.py', ('cl', 1): 'spk-clustering.py',
          ('cl', 2): 'spk-clustering2.py'}


def main():
    if not ref_exec.available():
        raise SystemExit('the reference tree is not available here')
    for name in sorted(cases.CASES):
        kind, variant, wavs, flags = cases.CASES[name]
        with tempfile.TemporaryDirectory() as tmp:
            rpath, feadir, sha, _ = cases.materialise(name, tmp)
            outp = os.path.join(tmp, 'out.recipe')
            fix = dict(name=name, script=SCRIPT[(kind, variant)], flags=flags, frames_sha256=sha,
                       numpy=numpy.__version__, scipy=scipy.__version__)
            try:
                stdout, _ = ref_exec.run(SCRIPT[(kind, variant)], [rpath, feadir, '-o', outp] + flags)
                fix['stdout'] = stdout.replace(tmp, '<TMP>')
                with open(outp) as f:
                    fix['recipe'] = f.read()
            except ValueError as e:
                fix['raises'] = 'ValueError: %s' % e
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(fix, f, indent=1, sort_keys=True)
        print(name, 'raises' in fix and fix['raises'] or '%d recipe lines' % fix['recipe'].count('\n'))


def scoring_fixtures():
    """The two scoring tools of the reference on (truth, detected) recipe pairs."""
    from spkdiar import synth
    for name, src, script, flags in (
            ('score_change', 'gw_bic_f100', 'spk-change-performance.py', ['-t', '0.25', '-sc', '-si', '-sd']),
            ('score_clus', 'cl1_hi_bic', 'clus-performance.py', [])):
        kind, variant, wavs, _ = cases.CASES[src]
        wav, kw, _ = wavs[0]
        rec = synth.make_recording(**kw)
        truth = ''.join(synth.truth_recipe('/syn/%s.wav' % wav, rec))
        proposed = json.load(open(os.path.join(HERE, src + '.json')))['recipe']
        with tempfile.TemporaryDirectory() as tmp:
            bp, pp = os.path.join(tmp, 'truth.recipe'), os.path.join(tmp, 'prop.recipe')
            open(bp, 'w').write(truth)
            open(pp, 'w').write(proposed)
            stdout, _ = ref_exec.run(script, [bp, pp] + flags)
            fix = dict(name=name, script=script, flags=flags, baseline=truth, proposed=proposed,
                       stdout=stdout.replace(tmp, '<TMP>'))
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(fix, f, indent=1, sort_keys=True)
        print(name, len(stdout.splitlines()), 'report lines')


def synthetic_exp(seed, nruns):
    """A speech-activity token stream: alternating speech / silence runs (some shorter than the
    hysteresis thresholds), repeated tokens inside long runs, ending in speech or silence."""
    rng = numpy.random.default_rng(seed)
    frame, toks = 0, []
    speech = False
    for _ in range(nruns):
        n = int(rng.choice([8, 20, 30, 45, 120, 400, 900]))
        tok = 'p' if speech else '<w>'
        toks.append('%d %s' % (frame, tok))
        if n >= 400:                                   # the decoder repeats the token inside long runs
            toks.append('%d %s' % (frame + n // 2, tok))
        frame += n
        speech = not speech
    return ' '.join(toks) + '\n', frame + int(rng.integers(1, 200))


def glue_fixtures():
    """voice-detection2.py and aku2ann.py of the reference on synthetic inputs."""
    with tempfile.TemporaryDirectory() as tmp:
        wavs = ['/syn/v%d.wav' % k for k in range(3)]
        exps = {}
        for k, w in enumerate(wavs):
            text, last = synthetic_exp(3000 + k, 40 + 7 * k + (k % 2))
            exps['v%d' % k] = dict(exp=text, last_frame='%d\n' % last)
            open(os.path.join(tmp, 'v%d.exp' % k), 'w').write(text)
            open(os.path.join(tmp, 'v%d.last_frame' % k), 'w').write('%d\n' % last)
        recipe = ''.join('audio=%s\n' % w for w in wavs[:2]) + 'no audio here\n' + 'audio=%s\n' % wavs[2]
        rp, outp = os.path.join(tmp, 'in.recipe'), os.path.join(tmp, 'out.recipe')
        open(rp, 'w').write(recipe)
        flags = ['-r', '125', '-ms', '0.2', '-mns', '0.3', '-sbe', '0.1', '-see', '0.05']
        stdout, _ = ref_exec.run('voice-detection2.py', [rp, tmp, '-o', outp] + flags)
        fix = dict(name='vad_recipe', script='voice-detection2.py', flags=flags, recipe_in=recipe, exps=exps,
                   stdout=stdout.replace(tmp, '<TMP>'), recipe=open(outp).read())
    json.dump(fix, open(os.path.join(HERE, 'vad_recipe.json'), 'w'), indent=1, sort_keys=True)
    print('vad_recipe', fix['recipe'].count('\n'), 'turns')
    with tempfile.TemporaryDirectory() as tmp:
        src = json.load(open(os.path.join(HERE, 'cl1_hi_glr_ms.json')))['recipe']
        lines = src.splitlines(True)
        lines.insert(2, 'a line without fields\n')
        lines.insert(4, 'audio=/syn/other.wav lna=b_1 start-time=1.5 end-time=2.25\n')
        text = ''.join(lines)
        rp, outp = os.path.join(tmp, 'in.recipe'), os.path.join(tmp, 'out.ann')
        open(rp, 'w').write(text)
        stdout, _ = ref_exec.run('aku2ann.py', [rp, '-o', outp])
        fix = dict(name='aku2ann', script='aku2ann.py', recipe_in=text, stdout=stdout.replace(tmp, '<TMP>'),
                   ann=open(outp).read())
    json.dump(fix, open(os.path.join(HERE, 'aku2ann.json'), 'w'), indent=1, sort_keys=True)
    print('aku2ann', fix['ann'].count('\n'), 'lines')
    # aku2elan.py: the reference's own tree building, serialised by oracle/lxml_shim.py (lxml is absent)
    import re
    with tempfile.TemporaryDirectory() as tmp:
        lines = src.splitlines(True)
        lines.insert(2, 'a line without fields\n')
        lines.insert(4, 'audio=/syn/other.wav lna=b_1 start-time=1.1 end-time=2.3\n')
        lines.insert(5, 'audio=/syn/o<&>"ther.wav lna=b_2 start-time=0.57 end-time=4.35 speaker=sp<&"k\n')
        text = ''.join(lines)
        rp, outp = os.path.join(tmp, 'in.recipe'), os.path.join(tmp, 'out.eaf')
        open(rp, 'w').write(text)
        stdout, _ = ref_exec.run('aku2elan.py', [rp, '-o', outp])
        to_stdout, _ = ref_exec.run('aku2elan.py', [rp])
        mask = lambda t: re.sub(r'DATE="[^"]*"', 'DATE="<NOW>"', t).replace(tmp, '<TMP>')      # noqa: E731
        fix = dict(name='aku2elan', script='aku2elan.py', recipe_in=text, stdout=mask(stdout),
                   eaf=mask(open(outp).read()), to_stdout=mask(to_stdout),
                   note='lxml is not installed here: serialisation by oracle/lxml_shim.py (un-pinned)')
    json.dump(fix, open(os.path.join(HERE, 'aku2elan.json'), 'w'), indent=1, sort_keys=True)
    print('aku2elan', fix['eaf'].count('\n'), 'lines')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'glue':
        glue_fixtures()
    else:
        if len(sys.argv) < 2 or sys.argv[1] != 'scoring':
            main()
        scoring_fixtures()
        glue_fixtures()
