#!/usr/bin/env python3
"""Drop-in for the reference's ./spk-clustering.py (same flags, recipe in / recipe out,
same stdout text); the numeric work runs on the GPU through libspkdiar.so."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spkdiar  # noqa: E402,F401
from spkdiar import clustering
clustering.main(variant=1)
