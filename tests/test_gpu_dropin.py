"""The drop-in boundary exercised the way the reference's driver uses it (spk-diarization2.py:108-132):
the scripts are looked up RELATIVE TO THE CURRENT DIRECTORY and run as child processes that exchange
recipe files -

    ./voice-detection2.py init.recipe exppath -o vad.recipe -ms 0.5 -mns 1.5
    ./spk-change-detection.py vad.recipe feapath -o spkc.recipe -m gw -d BIC -w 1.0 -st 3.0 -dws 0.1 -l 1.0
    ./spk-clustering.py spkc.recipe feapath -o out.recipe -m hi -l 1.3
    ./aku2ann.py out.recipe -o out.ann ;  ./aku2elan.py out.recipe -o out.eaf

Our scripts are installed into a scratch "checkout" once as symbolic links and once as plain copies
(+ SPKDIAR_HOME), executed from that directory, and every file they write must equal what the CPU oracle
writes for the same inputs.  Includes the reference's `-o stdout` behaviour of spk-clustering.py
(a FILE named `stdout`)."""

import io
import os
import shutil
import subprocess
import sys

import pytest

import spkdiar                              # noqa: F401
from conftest import run_oracle
from spkdiar import feacat, glue, synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ['voice-detection2.py', 'spk-change-detection.py', 'spk-clustering.py', 'aku2ann.py', 'aku2elan.py']


def _exp_for(rec, rate):
    """A speech-activity token stream with two silences inside the recording."""
    n = rec.frames.shape[0]
    cuts = [(0, 'p'), (int(n * 0.30), '<w>'), (int(n * 0.30) + 3 * rate, 'p'), (int(n * 0.72), '<w>'),
            (int(n * 0.72) + 2 * rate, 'p')]
    return ' '.join('%d %s' % c for c in cuts) + '\n', n


@pytest.mark.parametrize('install', ['symlink', 'copy'])
def test_scripts_run_from_cwd_like_spk_diarization2(install, tmp_path):
    work = tmp_path / 'checkout'
    work.mkdir()
    env = dict(os.environ)
    for name in NAMES:
        src = os.path.join(ROOT, 'scripts', name)
        if install == 'symlink':
            os.symlink(src, str(work / name))
        else:
            shutil.copy(src, str(work / name))
            os.chmod(str(work / name), 0o755)
    if install == 'copy':
        env['SPKDIAR_HOME'] = ROOT
    else:
        env.pop('SPKDIAR_HOME', None)
    rate = 125                                                 # D2 passes no -f: the scripts' default
    rec = synth.make_recording(4242, 90 * rate, 3, rate=rate, turn_lo=3, turn_hi=9)
    exppath, feapath, tmpp = work / 'exp', work / 'fea', work / 'tmp'
    for d in (exppath, feapath, tmpp):
        d.mkdir()
    infile = '/media/meeting.wav'
    feacat.write_features(str(feapath / 'meeting.fea'), rec.frames)
    exp, last = _exp_for(rec, rate)
    (exppath / 'meeting.exp').write_text(exp)
    (exppath / 'meeting.last_frame').write_text('%d\n' % last)
    init = str(tmpp / 'init.recipe')
    open(init, 'w').write('audio=' + infile + '\n')
    vad, spkc, outfile = str(tmpp / 'vad.recipe'), str(tmpp / 'spkc.recipe'), str(work / 'out.recipe')

    def call(argv):
        r = subprocess.run(argv, cwd=str(work), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           timeout=600)
        assert r.returncode == 0, (argv, r.stdout[-2000:], r.stderr[-2000:])
        return r.stdout

    # spk-diarization2.py:111-112, 122-128, 131-137 - argv lists exactly as written there; feapath with the
    # trailing separator the clustering script's string concatenation needs (spk-clustering.py:33)
    fp = str(feapath) + os.sep
    call(['./voice-detection2.py', init, str(exppath), '-o', vad, '-ms', '0.5', '-mns', '1.5'])
    s_cd = call(['./spk-change-detection.py', vad, fp, '-o', spkc, '-m', 'gw', '-d', 'BIC', '-w', '1.0',
                 '-st', '3.0', '-dws', '0.1', '-l', '1.0'])
    s_cl = call(['./spk-clustering.py', spkc, fp, '-o', outfile, '-m', 'hi', '-l', '1.3'])
    call(['./aku2ann.py', outfile, '-o', str(work / 'out.ann')])
    call(['./aku2elan.py', outfile, '-o', str(work / 'out.eaf')])

    # the same chain through the CPU oracle (and the host glue, golden-tested against the reference)
    o_vad = str(tmpp / 'o_vad.recipe')
    glue.vad_main([init, str(exppath), '-o', o_vad, '-ms', '0.5', '-mns', '1.5'], stdout=io.StringIO())
    assert open(vad).read() == open(o_vad).read() and open(vad).read().count('\n') >= 2
    o_spkc, o_out = str(tmpp / 'o_spkc.recipe'), str(tmpp / 'o_out.recipe')
    run_oracle('cd', 0, [o_vad, fp, '-o', o_spkc, '-m', 'gw', '-d', 'BIC', '-w', '1.0', '-st', '3.0',
                         '-dws', '0.1', '-l', '1.0'])
    run_oracle('cl', 1, [o_spkc, fp, '-o', o_out, '-m', 'hi', '-l', '1.3'])
    assert open(spkc).read() == open(o_spkc).read()
    assert open(outfile).read() == open(o_out).read()
    assert open(outfile).read().count('\n') >= 6 and 'speaker=speaker_2' in open(outfile).read()
    assert 'Using a growing window' in s_cd and 'Final speakers:' in s_cl
    ann = open(str(work / 'out.ann')).read()
    assert ann.startswith('# ' + infile + '\n') and ann.count('\n') == open(outfile).read().count('\n') + 1
    assert '<ANNOTATION_DOCUMENT' in open(str(work / 'out.eaf')).read()

    # `-o stdout` (what D2 passes when it has no output file): spk-clustering.py writes a FILE named stdout (Q14)
    call(['./spk-clustering.py', spkc, fp, '-o', 'stdout', '-m', 'hi', '-l', '1.3'])
    assert open(str(work / 'stdout')).read() == open(o_out).read()
