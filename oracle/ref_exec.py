"""Run the reference's OWN scripts in this container, to pin the oracle.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference is Python 2 and there is no Python 2 here, so its scripts cannot
be imported or run as they are.  This module reads a reference script from
``/root/reference`` AT RUN TIME, applies the token rewrites listed in ``RULES``
to the text IN MEMORY and executes the result with ``sys.argv`` set, capturing
stdout.  Nothing of the reference is copied into the repository, and nothing
here runs on the GPU box (``/root/reference`` does not exist there; callers
skip when ``available()`` is false).

Each rewrite closes one py2 -> py3 semantic gap and changes no algorithm:

R1  ``print a, b``        -> ``_p2print(a, b)``   print statement; items are
                              rendered with Python-2 ``str`` (12 significant
                              digits for floats)
R2  ``xrange``            -> ``range``
R3  ``sys.maxint``        -> ``(2**63-1)``
R4  ``x.size / dim``, ``index / len(speakers)``  -> ``//``  (py2 int division)
R5  ``int(np.fromfile(..., count=1))`` -> ``int(np.fromfile(...)[0])``
R6  ``str(``              -> ``_p2str(``   (recipe times, Q7)
R7  the loaded feature matrix is viewed as ``_FloatIdx``, an ndarray subclass
    whose ``[a:b]`` truncates float bounds the way numpy < 1.12 did
    (CD:144-145, 305-306, 309; CL2:48-50, 142)
R8  (clus-performance / spk-change-performance only) ``izip`` -> ``zip``,
    ``.iteritems()`` -> ``.items()``
R9  (aku2elan.py only) ``from lxml import etree`` resolves to ``oracle/lxml_shim.py``
    when lxml is not installed: the reference builds the element tree, the shim
    serialises it the way lxml does under Python 2
"""

import contextlib
import io
import os
import re
import sys

import numpy as np

from .py2compat import py2_print_str, py2_str

REFERENCE_ROOT = os.environ.get('SPKDIAR_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'spk-change-detection.py'))


class _FloatIdx(np.ndarray):
    """R7: float slice bounds truncate toward zero (numpy < 1.12)."""

    def __getitem__(self, key):
        if isinstance(key, slice):
            def fix(v):
                return int(v) if isinstance(v, (float, np.floating)) else v
            key = slice(fix(key.start), fix(key.stop), fix(key.step))
        return super(_FloatIdx, self).__getitem__(key)


def _rewrite_prints(src):
    """R1, line based: a print statement continues while the line ends in a
    backslash."""
    out = []
    lines = src.split('\n')
    k = 0
    pat = re.compile(r'^(\s*)print (.*)$')
    while k < len(lines):
        m = pat.match(lines[k])
        if not m or lines[k].lstrip().startswith('#'):
            out.append(lines[k])
            k += 1
            continue
        indent, body = m.group(1), m.group(2)
        while body.rstrip().endswith('\\'):
            body = body.rstrip()[:-1] + ' ' + lines[k + 1].strip()
            out.append('')                  # keep line numbers aligned
            k += 1
        out.append(indent + '_p2print(' + body.rstrip().rstrip(',') + ')')
        k += 1
    return '\n'.join(out)


RULES = [
    (re.compile(r'\bxrange\b'), 'range'),                                   # R2
    (re.compile(r'\bsys\.maxint\b'), '(2**63-1)'),                          # R3
    (re.compile(r'features\.size / dim'), 'features.size // dim'),          # R4
    (re.compile(r'index / len\(speakers\)'), 'index // len(speakers)'),     # R4
    (re.compile(r'int\(np\.fromfile\(ffile, dtype=np\.int32, count=1\)\)'),
     'int(np.fromfile(ffile, dtype=np.int32, count=1)[0])'),                # R5
    (re.compile(r'(?<![\w.])str\('), '_p2str('),                            # R6
    (re.compile(r'(features = features\.reshape\(.*\))$', re.M),
     r'\1.view(_FloatIdx)'),                                                # R7
    (re.compile(r'from itertools import izip'), 'izip = zip'),              # R8
    (re.compile(r'\.iteritems\(\)'), '.items()'),                           # R8
]


def translate(script):
    with open(os.path.join(REFERENCE_ROOT, script), 'r') as f:
        src = f.read()
    src = _rewrite_prints(src)
    for pat, rep in RULES:
        src = pat.sub(rep, src)
    return src


def run(script, argv, cwd=None):
    """Execute ``script`` (a file name under the reference root) with
    ``argv``; returns ``(stdout_text, globals_dict)``.  Exceptions raised by
    the script propagate (e.g. the ``ValueError`` of ``-m sw -d BIC``, Q1)."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')         # '\S' in non-raw regex literals
        code = compile(translate(script), os.path.join(REFERENCE_ROOT, script),
                       'exec')
    buf = io.StringIO()

    def _p2print(*items):
        sys.stdout.write(py2_print_str(*items) + '\n')

    g = {'__name__': '__main__', '_p2print': _p2print, '_p2str': py2_str,
         '_FloatIdx': _FloatIdx}
    old_argv = sys.argv
    old_cwd = os.getcwd()
    sys.argv = [script] + [str(a) for a in argv]
    shim = 'lxml' not in sys.modules and _lxml_missing()
    if shim:                                    # R9: see oracle/lxml_shim.py
        import types
        from . import lxml_shim
        pkg = types.ModuleType('lxml')
        pkg.etree = lxml_shim
        sys.modules['lxml'] = pkg
        sys.modules['lxml.etree'] = lxml_shim
    try:
        if cwd:
            os.chdir(cwd)
        with contextlib.redirect_stdout(buf), np.errstate(all='ignore'):
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                exec(code, g)
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
        if shim:
            del sys.modules['lxml'], sys.modules['lxml.etree']
    return buf.getvalue(), g


def _lxml_missing():
    import importlib.util
    return importlib.util.find_spec('lxml') is None


class _ParserCaptured(Exception):
    pass


def parser_of(script):
    """The ``argparse.ArgumentParser`` the reference script builds (its flags, defaults, choices and
    types): the script is executed up to its ``parse_args()`` call, which is intercepted."""
    import argparse
    box = []
    orig = argparse.ArgumentParser.parse_args

    def grab(self, *a, **k):
        box.append(self)
        raise _ParserCaptured()
    argparse.ArgumentParser.parse_args = grab
    try:
        run(script, [])
    except _ParserCaptured:
        pass
    finally:
        argparse.ArgumentParser.parse_args = orig
    return box[0]
