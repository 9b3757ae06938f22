"""Text glue around the hot path: ``voice-detection2.py`` and
``aku2ann.py`` restated in ``spkdiar.glue`` must reproduce, byte for byte, what the
reference's own scripts wrote for the committed synthetic inputs (fixtures made by
``tests/golden/make_golden.py glue`` through ``oracle/ref_exec.py``)."""

import io
import os

import pytest

import spkdiar                              # noqa: F401
from conftest import load_golden
from spkdiar import glue
from oracle import ref_exec


def _vad_inputs(fix, tmp):
    for name, d in fix['exps'].items():
        open(os.path.join(tmp, name + '.exp'), 'w').write(d['exp'])
        open(os.path.join(tmp, name + '.last_frame'), 'w').write(d['last_frame'])
    rp = os.path.join(tmp, 'in.recipe')
    open(rp, 'w').write(fix['recipe_in'])
    return rp


def test_vad_recipe_reproduces_reference(tmp_path):
    fix = load_golden('vad_recipe')
    tmp = str(tmp_path)
    rp = _vad_inputs(fix, tmp)
    outp = os.path.join(tmp, 'out.recipe')
    out = io.StringIO()
    glue.vad_main([rp, tmp, '-o', outp] + fix['flags'], stdout=out)
    assert open(outp).read() == fix['recipe']
    assert out.getvalue().replace(tmp, '<TMP>') == fix['stdout']
    # to stdout instead of a file: log and recipe interleave as in the reference
    out2 = io.StringIO()
    glue.vad_main([rp, tmp] + fix['flags'], stdout=out2)
    assert out2.getvalue().endswith(fix['recipe'])


@pytest.mark.skipif(not ref_exec.available(), reason='needs /root/reference')
@pytest.mark.parametrize('flags', [[], ['-r', '100', '-ms', '0.5', '-mns', '0.1'], ['-mns', '1.0', '-see', '0.2']])
def test_vad_recipe_matches_live_reference(tmp_path, flags):
    fix = load_golden('vad_recipe')
    tmp = str(tmp_path)
    rp = _vad_inputs(fix, tmp)
    want_out, got_out = os.path.join(tmp, 'w.recipe'), os.path.join(tmp, 'g.recipe')
    want, _ = ref_exec.run('voice-detection2.py', [rp, tmp, '-o', want_out] + flags)
    out = io.StringIO()
    glue.vad_main([rp, tmp, '-o', got_out] + flags, stdout=out)
    assert open(got_out).read() == open(want_out).read()
    assert out.getvalue().replace('g.recipe', 'w.recipe') == want


def test_lna_names():
    seq, lna = [], 'a'
    for _ in range(28):
        seq.append(lna)
        lna = glue.next_lna(lna)
    assert seq[:3] == ['a', 'b', 'c'] and seq[25:28] == ['z', 'aa', 'ab']
    assert glue.next_lna('az') == 'ba' and glue.next_lna('zz') == 'aaa'


def test_missing_exp_file_exits(tmp_path):
    rp = str(tmp_path / 'in.recipe')
    open(rp, 'w').write('audio=/syn/none.wav\n')
    out = io.StringIO()
    with pytest.raises(SystemExit):
        glue.vad_main([rp, str(tmp_path)], stdout=out)
    assert 'does not exist' in out.getvalue()


def test_aku2ann_reproduces_reference(tmp_path):
    fix = load_golden('aku2ann')
    tmp = str(tmp_path)
    rp, outp = os.path.join(tmp, 'in.recipe'), os.path.join(tmp, 'out.ann')
    open(rp, 'w').write(fix['recipe_in'])
    out = io.StringIO()
    glue.ann_main([rp, '-o', outp], stdout=out)
    assert open(outp).read() == fix['ann']
    assert out.getvalue().replace(tmp, '<TMP>') == fix['stdout']
    out2 = io.StringIO()
    glue.ann_main([rp], stdout=out2)
    assert out2.getvalue().endswith('Writing output to: stdout\n' + fix['ann'])
