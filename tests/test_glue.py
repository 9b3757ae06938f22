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


def _mask(text, tmp):
    import re
    return re.sub(r'DATE="[^"]*"', 'DATE="<NOW>"', text).replace(tmp, '<TMP>')


def test_aku2elan_reproduces_reference(tmp_path):
    """aku2elan.py:45-99.  The fixture is the reference's own tree building serialised by
    oracle/lxml_shim.py (lxml is absent here); the product writes the text directly."""
    fix = load_golden('aku2elan')
    tmp = str(tmp_path)
    rp, outp = os.path.join(tmp, 'in.recipe'), os.path.join(tmp, 'out.eaf')
    open(rp, 'w').write(fix['recipe_in'])
    out = io.StringIO()
    glue.elan_main([rp, '-o', outp], stdout=out)
    assert _mask(open(outp).read(), tmp) == fix['eaf']
    assert _mask(out.getvalue(), tmp) == fix['stdout']
    out2 = io.StringIO()
    glue.elan_main([rp], stdout=out2)
    assert _mask(out2.getvalue(), tmp) == fix['to_stdout']
    # the document is well-formed XML with one annotation and two time slots per recipe line
    import xml.etree.ElementTree as ET
    doc = ET.parse(outp).getroot()
    n = len(doc.findall('./TIER/ANNOTATION'))
    assert n == 24 and len(doc.findall('./TIME_ORDER/TIME_SLOT')) == 2 * n
    assert doc.find('./HEADER/PROPERTY').text == str(n)
    assert doc.find('./HEADER/MEDIA_DESCRIPTOR').get('MIME_TYPE') == 'audio/x-wav'
    vals = [a.findtext('./ALIGNABLE_ANNOTATION/ANNOTATION_VALUE') for a in doc.findall('./TIER/ANNOTATION')]
    assert 'sp<&"k' in vals and None in vals              # escaped tag round-trips; a line without a speaker tag


def test_aku2elan_date_and_errors(tmp_path):
    import re
    assert re.match(r'^\d{4}-\d\d-\d\dT\d\d:\d\d:\d\d(\.\d+)?[+-]\d{1,2}:\d\d$', glue.iso_now())   # '%+02d' (aku2elan.py:17)
    rp = str(tmp_path / 'in.recipe')
    open(rp, 'w').write('nothing here\n')
    out = io.StringIO()
    with pytest.raises(IndexError):                       # recipe[0][0] of an empty recipe (aku2elan.py:60)
        glue.elan_main([rp], stdout=out)
    assert out.getvalue().startswith('Reading recipe from:') and 'Recipe line without recognizable data:' in out.getvalue()
    open(rp, 'w').write('audio=/syn/x.unknownext lna=a_1 start-time=0.0 end-time=1.0\n')
    with pytest.raises(TypeError):                        # MIME_TYPE None is refused by lxml
        glue.elan_main([rp], stdout=io.StringIO())


@pytest.mark.skipif(not ref_exec.available(), reason='needs /root/reference')
def test_aku2elan_matches_live_reference(tmp_path):
    fix = load_golden('aku2elan')
    tmp = str(tmp_path)
    rp = os.path.join(tmp, 'in.recipe')
    open(rp, 'w').write(fix['recipe_in'])
    want, _ = ref_exec.run('aku2elan.py', [rp, '-o', os.path.join(tmp, 'w.eaf')])
    assert _mask(open(os.path.join(tmp, 'w.eaf')).read(), tmp) == fix['eaf']
    assert _mask(want, tmp).replace('w.eaf', 'out.eaf') == fix['stdout']
