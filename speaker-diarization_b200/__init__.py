"""speaker-diarization_b200 - B200-native statistical core of the Aalto
speaker-diarization scripts (speaker-turn distance search + agglomerative BIC
clustering), behind the reference's own command lines.

The directory name carries a hyphen (it is the name the build contract asks
for), so it is imported through ``importlib`` - the top-level module
``spkdiar`` does that and aliases this package::

    import spkdiar
    from spkdiar import synth, feacat

Numeric work happens only in ``csrc/`` (CUDA, sm_100a) behind the C-ABI of
``include/spkdiar.h``, reached through ``_abi`` (ctypes).  There is no CPU
fallback: calls raise ``SpkdiarError`` when the library or a GPU is missing.
"""

from . import py2fmt, feacat, synth, recipe            # noqa: F401
from . import _abi                                     # noqa: F401
from . import change_detection, clustering, scoring    # noqa: F401
from . import corpus                                   # noqa: F401
from ._abi import SpkdiarError                         # noqa: F401

__version__ = '0.1.0'
