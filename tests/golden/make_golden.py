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


if __name__ == '__main__':
    if len(sys.argv) < 2 or sys.argv[1] != 'scoring':
        main()
    scoring_fixtures()
