"""CPU oracle for the hot path of the Aalto speaker-diarization scripts.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or the
timed CPU baseline.  The shipped path (``speaker-diarization_b200``) never
imports this package and has no CPU fallback.

What it is: a Python-3 / numpy / scipy restatement of the statistical core of
``spk-change-detection.py`` (CD), ``spk-clustering.py`` (CL1) and
``spk-clustering2.py`` (CL2) of the reference, following the reference's own
arithmetic call by call (``np.cov`` -> ``scipy.linalg.det`` -> ``np.log``,
``scipy.linalg.pinv``, ``np.mean`` on float32 slices) so that floating-point
results are those the reference computes.  Every function cites the reference
``file:line`` range it restates.

Parity pin: the reference ships NO tests, fixtures or golden vectors
(SURVEY.md section 4), it is Python 2, and its arithmetic lives in un-pinned
numpy/scipy (docker/Dockerfile:49).  The oracle is therefore pinned by
``oracle/ref_exec.py``: the reference's own source files are read from
``/root/reference`` at run time, put through a small documented set of
py2->py3 token rewrites IN MEMORY (never written to the repo), executed in this
container, and their recipes / stdout compared with the oracle's
(``tests/test_oracle.py``, skipped where ``/root/reference`` is
absent).  The same runs produced the committed fixtures under
``tests/golden/`` (generator: ``tests/golden/make_golden.py``; numpy / scipy
versions recorded inside each fixture).
"""

from .py2compat import py2_str, py2_print_str  # noqa: F401
