"""Python-2 number formatting used by every recipe and log line.

The reference scripts are Python 2: ``str(float)`` there is ``'%.12g'`` plus a
trailing ``.0`` for integer-looking results, and the ``print`` statement joins
its items with single spaces (spk-change-detection.py:59-60, 563-579;
spk-clustering.py:68-70, 214-215).  Recipes are byte-identical only if the
host reproduces this.
"""

import numpy as np

MAXINT = 2 ** 63 - 1   # sys.maxint of the reference's 64-bit Python 2


def fstr(x):
    """Python-2 ``str`` of a float."""
    s = '%.12g' % x
    if not ('.' in s or 'e' in s or 'n' in s):
        s += '.0'
    return s


def p2str(x):
    """Python-2 ``str`` of the kinds of value the scripts print."""
    if isinstance(x, (bool, np.bool_)):
        return str(bool(x))
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    if isinstance(x, (float, np.floating)):
        return fstr(float(x))
    if isinstance(x, tuple):
        body = ', '.join(str(int(e)) for e in x)
        return '(' + body + (',)' if len(x) == 1 else ')')
    return str(x)


def p2line(*items):
    """Text of ``print a, b, c`` (without the newline)."""
    return ' '.join(p2str(i) for i in items)
