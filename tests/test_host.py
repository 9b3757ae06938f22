"""CPU tests of the host layer: Python-2 text, recipe parsing / writing, feacat
I/O, command-line contracts, and that the C-ABI library exports what
include/spkdiar.h declares (no compute: there is no GPU here)."""

import ctypes
import io
import os
import re

import numpy as np
import pytest

import spkdiar
from spkdiar import _abi, feacat, glue, py2fmt, recipe, synth
from spkdiar import change_detection as pcd, clustering as pcl
from oracle import change_detection as ocd, clustering as ocl, py2compat, ref_exec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# known Python-2 ``str(float)`` outputs (tier T1)
PY2_STR = [(0.1 + 0.2, '0.3'), (2.0 / 3 * 100, '66.6666666667'), (1e16, '1e+16'), (20.0, '20.0'),
           (22.290000000000003, '22.29'), (1503.5 / 125, '12.028'), (float('inf'), 'inf'),
           (float('-inf'), '-inf'), (1e-5, '1e-05'), (123456789012.0, '123456789012.0'),
           (1234567890123.0, '1.23456789012e+12'), (0.0, '0.0'), (-0.5, '-0.5')]


@pytest.mark.parametrize('x,want', PY2_STR)
def test_py2_float_str(x, want):
    assert py2fmt.fstr(x) == want
    assert py2compat.py2_float_str(x) == want
    assert py2fmt.p2str(np.float64(x)) == want


def test_py2_print_line():
    assert py2fmt.p2line('Merging:', np.int64(3), 'and', 5, 'distance:', np.float64(-1.5)) == \
        'Merging: 3 and 5 distance: -1.5'
    assert py2fmt.p2line('Inf:', (50, 39), (150, 39), float('-inf')) == 'Inf: (50, 39) (150, 39) -inf'
    assert py2fmt.p2str(float('nan')) == 'nan'
    assert py2fmt.p2line('x', 0, py2fmt.MAXINT) == 'x 0 9223372036854775807'


def test_recipe_parse_quirks():
    msgs = []
    lines = ['audio=/a/b.wav lna=a_1 start-time=0.0 end-time=12.5 speaker=x\n',
             'audio=/a/b.wav lna=a_2 start-time=5 end-time=7.0\n',          # "5" alone does not match \d+.\d+
             'end-time=3.25 start-time=1x5 lna=q audio=z\n',               # any char between the digits
             'junk\n']
    with pytest.raises(ValueError):                                          # float('1x5'), as in the reference
        recipe.parse(lines, msgs.append)
    del msgs[:]
    got = recipe.parse([lines[0], lines[1], lines[3]], msgs.append)
    assert got == [recipe.Line('/a/b.wav', 'a_1', 0.0, 12.5)]
    assert msgs.count('Recipe line without recognizable data:') == 2
    # the oracle parses the same way
    omsgs = []
    assert ocd.parse_recipe([lines[0], lines[1], lines[3]], omsgs.append) == [tuple(got[0])]
    assert omsgs == msgs


def test_lna_renaming():
    """counter per prefix, prefix = lna[:lna.find('_')] (no
    underscore drops the last character), -dlr keeps names."""
    w = recipe.Writer(100.0)
    out = io.StringIO()
    for lna in ['a_7', 'a_9', 'b_1', 'b_1', 'xyz', 'xy_3']:
        w.write(recipe.Line('/f.wav', lna, 1.0, 2.0), 150, 300.0, 1.0, 'spk_turn', out)
    names = re.findall(r'lna=(\S+)', out.getvalue())
    assert names == ['a_1', 'a_2', 'b_1', 'b_2', '1', 'xy_2']
    assert 'start-time=2.5 end-time=4.0 speaker=spk_turn' in out.getvalue()
    w = recipe.Writer(100.0, rename=False, segprefix='/seg/')
    out, seg = io.StringIO(), io.StringIO()
    w.write(recipe.Line('/f.wav', 'k_4', 0.0, 1.0), 0, 50, 0, 'speaker_2', out, seg)
    assert out.getvalue() == 'audio=/f.wav lna=k_4 start-time=0.0 end-time=0.5 speaker=speaker_2\n'
    assert seg.getvalue() == ('audio=/f.wav alignment=/seg/k_4.seg lna=k_4 start-time=0.0 '
                              'end-time=0.5 speaker=speaker_2\n')


def test_writer_matches_oracle_writer():
    o = ocd.ChangeDetection(125)
    w = recipe.Writer(125.0)
    a, b = io.StringIO(), io.StringIO()
    for k, lna in enumerate(['a_1', 'a_1', 'b_2', 'c']):
        line = ('/x/y.wav', lna, 0.37 * k, 9.0)
        o.write_recipe_line(line, 1503.5 * k, 1600.25 + k, line[2], a)
        w.write(recipe.Line(*line), 1503.5 * k, 1600.25 + k, line[2], 'spk_turn', b)
    assert a.getvalue() == b.getvalue()


def test_feacat_roundtrip(tmp_path):
    rec = synth.make_recording(3, 257, 2)
    p = str(tmp_path / 'a.fea')
    feacat.write_features(p, rec.frames)
    dim, back = feacat.read_features(p)
    assert dim == 39 and back.dtype == np.float32 and np.array_equal(back, rec.frames)
    assert os.path.getsize(p) == 4 + 257 * 39 * 4
    assert feacat.feature_file_name('/a/b/c.wav', '/fea', '.fea') == '/fea/c.fea'
    assert feacat.feature_file_name('/a/b/c.wav', '/fea/', '.fea', concat=True) == '/fea/c.fea'
    with open(p, 'ab') as f:
        f.write(b'\0\0\0\0')                                  # ragged tail: not a whole frame
    with pytest.raises(ValueError):
        feacat.read_features(p)


def test_synth_is_reproducible():
    a = synth.make_recording(9, 1000, 3)
    b = synth.make_recording(9, 1000, 3)
    assert np.array_equal(a.frames, b.frames) and a.turns == b.turns
    assert all(t[2] != u[2] for t, u in zip(a.turns, a.turns[1:]))
    assert a.turns[-1][1] == 1000 and a.frames.dtype == np.float32


def _flags(parser):
    out = {}
    for a in parser._actions:
        for s in a.option_strings or [a.dest]:
            out[s] = (a.default, tuple(a.choices) if a.choices else None)
    return out


def _flags_full(parser):
    out = {}
    for a in parser._actions:
        for s in a.option_strings or [a.dest]:
            if s in ('-h', '--help'):
                continue
            out[s] = (a.default, tuple(a.choices) if a.choices else None, a.type, a.nargs, type(a).__name__, a.dest)
    return out


def test_cli_contract_change_detection():
    """Same flags and defaults as spk-change-detection.py:399-465 (checked
    against the oracle's parser, which is checked against the reference)."""
    ref = _flags(ocd.build_parser())
    got = _flags(pcd.build_parser())
    for k, v in ref.items():
        assert got[k] == v, k
    assert got['-f'][0] == 125 and got['-m'] == ('sw', ('sw', 'gw', 'm'))
    assert got['-d'] == ('GLR', ('GLR', 'BIC', 'KL2')) and got['-o'][0] == 'stdout'
    assert set(got) - set(ref) <= {'--device'}


@pytest.mark.parametrize('variant', [1, 2])
def test_cli_contract_clustering(variant):
    ref = _flags(ocl.build_parser(variant))
    got = _flags(pcl.build_parser(variant))
    for k, v in ref.items():
        assert got[k] == v, k
    assert got['-o'][0] == (None if variant == 1 else 'stdout')
    assert got['-seg'][0] == ('' if variant == 1 else None)
    assert set(got) - set(ref) <= {'--device'}


ADDED_FLAGS = {'--device', '--sw-bic', '--bic-cache'}


@pytest.mark.skipif(not ref_exec.available(), reason='needs /root/reference')
@pytest.mark.parametrize('script,build', [
    ('spk-change-detection.py', lambda: pcd.build_parser()),
    ('spk-clustering.py', lambda: pcl.build_parser(1)),
    ('spk-clustering2.py', lambda: pcl.build_parser(2)),
    ('voice-detection2.py', lambda: glue.vad_parser()),
])
def test_cli_tables_equal_the_reference_scripts_own_parsers(script, build):
    """Flag for flag against the parser the REFERENCE script builds (executed up to its parse_args by
    oracle/ref_exec.py): option strings, defaults, choices, value types, arity, destinations."""
    import sys as _sys
    ref = _flags_full(ref_exec.parser_of(script))
    got = _flags_full(build())
    for k, v in ref.items():
        assert k in got, (script, k)
        d, choices, typ, nargs, kind, dest = got[k]
        if v[0] is _sys.stdout or (hasattr(v[0], 'write') and not isinstance(v[0], str)):
            # `-o` default is the sys.stdout OBJECT in spk-clustering.py (Q14): ours is None, tested as "is None"
            assert d is None, (script, k)
        else:
            assert d == v[0], (script, k, d, v[0])
        assert (choices, typ, nargs, kind, dest) == v[1:], (script, k, got[k], v)
    assert set(got) - set(ref) <= ADDED_FLAGS, set(got) - set(ref)
    oref = {'spk-change-detection.py': lambda: ocd.build_parser(), 'spk-clustering.py': lambda: ocl.build_parser(1),
            'spk-clustering2.py': lambda: ocl.build_parser(2)}.get(script)
    if oref:                                                  # the oracle's tables too
        o = _flags_full(oref())
        for k, v in ref.items():
            if not hasattr(v[0], 'write'):
                assert o[k][0] == v[0] and o[k][1:] == v[1:], (script, k)


def test_bic_from_terms_matches_reference_expression():
    n1, n2, l1, l2, l = 150, 250, -31.25, -29.5, -27.125
    d, c1 = pcd.bic_from_terms(n1, n2, l1, l2, l, 1.3)
    n = n1 + n2
    want = 0.5 * n * l - 0.5 * n1 * l1 - 0.5 * n2 * l2
    want -= 1.3 * 0.5 * (39 + 0.5 * 39 * 40) * np.log(n)
    assert d == want and c1 == 0.5 * n1 * l1
    d2, _ = pcd.bic_from_terms(n1, n2, l1, l2, l, 1.3, c1=7.0)
    assert d2 == 0.5 * n * l - 7.0 - 0.5 * n2 * l2 - 1.3 * 0.5 * 819.0 * np.log(n)


def test_library_exports_every_declared_symbol():
    """The C-ABI library is built in-tree and exports each function
    include/spkdiar.h declares (and the Python binding knows all of them)."""
    import __graft_entry__ as g
    g.build()
    header = open(os.path.join(ROOT, 'include', 'spkdiar.h')).read()
    declared = set(re.findall(r'\b(spkdiar_[a-z_0-9]+)\s*\(', header))
    declared -= {'spkdiar_exchange_fn'}
    assert declared == set(_abi.SYMBOLS)
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), s
    lib.spkdiar_abi_version.restype = ctypes.c_int
    assert lib.spkdiar_abi_version() == 1
    assert ctypes.sizeof(_abi.GwParams) == 56


def test_no_gpu_fails_loudly():
    """Without a device the product path raises - there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    with pytest.raises(_abi.SpkdiarError):
        _abi.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'speaker-diarization_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.inc')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f
    for f in os.listdir(os.path.join(ROOT, 'scripts')):
        assert 'oracle' not in open(os.path.join(ROOT, 'scripts', f)).read()


@pytest.mark.parametrize('name,fn', [('score_change', 'change_performance_main'),
                                     ('score_clus', 'clus_performance_main')])
def test_scoring_tools_reproduce_reference_report(name, fn, tmp_path):
    """spk-change-performance.py / clus-performance.py: our Python-3 versions
    print the report the reference printed (golden made from the reference)."""
    import json
    from spkdiar import scoring
    gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', name + '.json')))
    bp, pp = str(tmp_path / 'truth.recipe'), str(tmp_path / 'prop.recipe')
    open(bp, 'w').write(gold['baseline'])
    open(pp, 'w').write(gold['proposed'])
    out = io.StringIO()
    getattr(scoring, fn)([bp, pp] + gold['flags'], stdout=out)
    assert out.getvalue().replace(str(tmp_path), '<TMP>') == gold['stdout']


def test_clus_performance_run_length_equals_per_millisecond_lists():
    from spkdiar import scoring
    base = [('f', [(0.0, 1.5, 'A'), (2.0, 3.25, 'B'), (3.25, 4.0, 'A')])]
    prop = [('f', [(0.1, 1.4, 'x'), (1.4, 3.0, 'y'), (3.0, 3.9, 'x')])]

    def expand(segs):
        out, now = [], 0.0
        for s, e, l in segs:
            if s > now:
                out += [None] * int((s - now) / 0.001)
            out += [l] * int((e - s) / 0.001)
            now = e
        return out
    want = {}
    for a, b in zip(expand(base[0][1]), expand(prop[0][1])):
        want[(a, b)] = want.get((a, b), 0) + 1
    got = scoring._co_occurrence(scoring._label_runs(base[0][1], 0.001), scoring._label_runs(prop[0][1], 0.001))
    assert got == want and list(got) == list(want)


def test_lines_from_records_equal_parse():
    """The corpus driver skips the regular expressions when it reads back a recipe it has just
    written; every case the reference's patterns treat specially must fall back to them."""
    import io
    from spkdiar import recipe
    w = recipe.Writer(100.0)
    w.record = []
    out = io.StringIO()
    src = [recipe.Line('/a/b.wav', 'a_1', 0.0, 10.0), recipe.Line('/a/lna=x.wav', 'a_2', 0.0, 1.0),
           recipe.Line('/a/c.wav', 'b_7', 0.0, 1.0)]
    cases = [(0.0, 1234.5), (1e-7, 3.0), (12.3456789012345, 1e17), (-50.0, 20.0), (7.0, 7.0), (1e5 / 3, 2e5 / 3)]
    for k, (s, e) in enumerate(cases):
        w.write(src[k % 3], s, e, 0.0 if k % 2 else 1.5, 'spk_turn', out)
    texts = out.getvalue().splitlines(True)
    assert len(w.record) == len(texts) == len(cases)
    assert recipe.lines_from_records(w.record, texts) == recipe.parse(texts)
    # the fast path of parse itself: same result as four independent searches
    odd = ['audio=/x/lna=y.wav lna=a_1 start-time=1.0 end-time=2.0\n',
           'lna=a_3 audio=/z.wav end-time=9.5 start-time=4.25 speaker=s\n',
           'audio=/z.wav lna=a_4 start-time=5 end-time=6.0\n',
           'audio=/z.wav lna=a_5 start-time=1e-05 end-time=6.0\n']
    got = recipe.parse(odd)
    assert got == [recipe.Line('/x/lna=y.wav', 'y.wav', 1.0, 2.0), recipe.Line('/z.wav', 'a_3', 4.25, 9.5)]
