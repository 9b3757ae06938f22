"""Python-2 text semantics the reference's output depends on (oracle side).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference writes every recipe time and every logged number through the
Python-2 ``str()`` / ``print`` statement (CD:59-60, CD:563-579, CL1:68-69,
CL1:214-215).  Under Python 2 (and numpy <= 1.13, which the float slice
indices at CD:144-145 / CD:305-306 / CL2:48 require) ``str(float)`` is
``'%.12g'`` with ``'.0'`` appended when the result looks like an integer.
"""

import numpy as np

MAXINT = 2 ** 63 - 1  # sys.maxint on the 64-bit Python 2 the reference ran on


def py2_float_str(x):
    """``str(float)`` as Python 2 prints it."""
    s = '%.12g' % x
    if '.' not in s and 'e' not in s and 'n' not in s:  # 'n' covers inf / nan
        s += '.0'
    return s


def py2_str(x):
    """``str(x)`` for the value kinds the reference prints or writes."""
    if isinstance(x, (bool, np.bool_)):
        return str(bool(x))
    if isinstance(x, (int, np.integer)):
        return str(int(x))
    if isinstance(x, (float, np.floating)):
        return py2_float_str(float(x))
    if isinstance(x, tuple):
        # print statement uses repr() for container items; only shapes
        # (tuples of ints) are ever printed (CD:220, CD:250)
        inner = ', '.join(repr(int(e)) if isinstance(e, (int, np.integer))
                          else repr(e) for e in x)
        if len(x) == 1:
            inner += ','
        return '(' + inner + ')'
    return str(x)


def py2_print_str(*items):
    """The text a Python-2 ``print a, b, c`` statement emits (no newline)."""
    return ' '.join(py2_str(i) for i in items)
