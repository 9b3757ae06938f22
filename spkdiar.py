"""Import alias: ``import spkdiar`` == the package in ``speaker-diarization_b200/``
(a hyphenated directory cannot be named in an ``import`` statement)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module('speaker-diarization_b200')
sys.modules[__name__] = _pkg
