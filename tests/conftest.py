import io
import json
import os
import re
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

warnings.filterwarnings('ignore', category=RuntimeWarning)
warnings.filterwarnings('ignore', category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
_NUM = re.compile(r'^[-+]?(\d+\.?\d*([eE][-+]?\d+)?|inf|nan)$')


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + '.json')) as f:
        return json.load(f)


def logs_match(got, want, rel=1e-9):
    """Two stdout texts are equal up to numeric tokens, which may differ by
    ``rel`` (the GPU's log-determinants agree with LAPACK's to ~1e-12, and the
    scripts print 12 significant digits)."""
    gl, wl = got.splitlines(), want.splitlines()
    if len(gl) != len(wl):
        return 'line count %d != %d' % (len(gl), len(wl))
    for k, (a, b) in enumerate(zip(gl, wl)):
        ta, tb = a.split(' '), b.split(' ')
        if len(ta) != len(tb):
            return 'line %d: %r != %r' % (k, a, b)
        for x, y in zip(ta, tb):
            if x == y:
                continue
            if _NUM.match(x) and _NUM.match(y):
                fx, fy = float(x), float(y)
                if fx == fy or abs(fx - fy) <= rel * max(abs(fx), abs(fy)) + 5e-12 * max(abs(fx), abs(fy)):
                    continue
            return 'line %d: %r != %r' % (k, a, b)
    return None


@pytest.fixture
def gpu_ctx():
    import spkdiar                      # noqa: F401
    from spkdiar import _abi
    ctx = _abi.Context(0)
    yield ctx
    ctx.close()


def run_oracle(kind, variant, argv, trace=None):
    from oracle import change_detection as ocd, clustering as ocl
    out = io.StringIO()
    if kind == 'cd':
        obj = ocd.main(argv, stdout=out, trace=trace)
    else:
        obj = ocl.main(argv, stdout=out, variant=variant, trace=trace)
    return out.getvalue(), obj


def run_product(kind, variant, argv, ctx=None, **kw):
    from spkdiar import change_detection as pcd, clustering as pcl
    out = io.StringIO()
    if kind == 'cd':
        obj = pcd.main(argv, stdout=out, ctx=ctx)
    else:
        obj = pcl.main(argv, stdout=out, variant=variant, ctx=ctx, **kw)
    return out.getvalue(), obj
