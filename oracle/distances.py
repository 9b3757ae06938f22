"""The three segment distances of the reference, restated for Python 3.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Arithmetic is deliberately the reference's: ``np.cov(rowvar=0)`` (fp64, ddof
1), ``scipy.linalg.det`` (LAPACK getrf), ``np.log``, ``scipy.linalg.pinv`` and
``np.mean`` on the float32 slice (sequential fp32 accumulation).
"""

import numpy as np
from scipy.linalg import det, pinv


def _cov(a):
    return np.cov(a, rowvar=0)


def bic_penalty(lambdac, p, n):
    """``corr`` of CD:98 / CL1:98 (same expression, same operation order)."""
    return lambdac * 0.5 * (p + 0.5 * p * (p + 1)) * np.log(n)


class BicMemo(dict):
    """The ``saved`` dict of ``bic`` (CD:72, 84-90): left term keyed by ``i``."""


def bic_cd(arr1, arr2, arr, lambdac, i=0, saved=None):
    """Delta-BIC of change detection, CD:72-100.

    ``saved`` is the memo of ``c1 = 0.5*N1*ln|S1|`` keyed by ``i``.  The
    reference's signature has a MUTABLE DEFAULT ``saved={}`` (CD:72): callers
    that do not pass one (``dist_sw`` CD:310, ``merge_rec`` CD:149) all share
    one process-wide dict and all use ``i == 0``, so the very first ``c1`` of
    the process is reused for ever.  The caller reproduces that
    by handing the same ``BicMemo`` to every such call.
    """
    if saved is None:
        saved = {}
    with np.errstate(all='ignore'):
        if i in saved:
            c1 = saved[i]
        else:
            S1 = _cov(arr1)
            N1 = arr1.shape[0]
            c1 = 0.5 * N1 * np.log(det(S1))
            saved[i] = c1
        S2 = _cov(arr2)
        N2 = arr2.shape[0]
        N = arr.shape[0]
        S = _cov(arr)
        d = 0.5 * N * np.log(det(S)) - c1 - 0.5 * N2 * np.log(det(S2))
        p = arr.shape[1]
        d -= bic_penalty(lambdac, p, N)
    return d


def bic_cl(arr1, arr2, lambdac):
    """Delta-BIC of clustering, CL1:81-100 (== CL2:80-99): pooled covariance of
    the concatenated frames, no memo."""
    with np.errstate(all='ignore'):
        arr = np.concatenate((arr1, arr2))
        N1 = arr1.shape[0]
        N2 = arr2.shape[0]
        S1 = _cov(arr1)
        S2 = _cov(arr2)
        N = arr.shape[0]
        S = _cov(arr)
        d = 0.5 * N * np.log(det(S)) - 0.5 * N1 * np.log(det(S1)) \
            - 0.5 * N2 * np.log(det(S2))
        p = arr.shape[1]
        d -= bic_penalty(lambdac, p, N)
    return d


def glr(arr1, arr2):
    """Covariance-only generalised likelihood ratio, CD:103-121 (== CL1:103-121,
    CL2:102-120)."""
    with np.errstate(all='ignore'):
        N1 = arr1.shape[0]
        N2 = arr2.shape[0]
        S1 = _cov(arr1)
        S2 = _cov(arr2)
        N = float(N1 + N2)
        d = -(N / 2.0) * ((N1 / N) * np.log(det(S1)) + (N2 / N) * np.log(det(S2))
                          - np.log(det((N1 / N) * S1 + (N2 / N) * S2)))
    return d


def kl2(arr1, arr2):
    """The reference's "KL2", CD:124-133 (== CL1:124-133, CL2:123-132).

    ``*`` on ndarrays is element-wise, so this is NOT the textbook matrix
    formula: only diag(S) and diag(pinv(S)) reach the trace, and
    the means are float32 sequential sums (Q4)."""
    with np.errstate(all='ignore'):
        S1 = _cov(arr1)
        S2 = _cov(arr2)
        m1 = np.mean(arr1, 0)
        m2 = np.mean(arr2, 0)
        delta = m1 - m2
        d = 0.5 * np.trace((S1 - S2) * (pinv(S2) - pinv(S1))) + \
            0.5 * np.trace((pinv(S1) + pinv(S2)) * delta * delta.T)
    return d
