#!/usr/bin/env python3
"""Drop-in for the reference's ./spk-clustering2.py (same flags, recipe in / recipe out,
same stdout text); the numeric work runs on the GPU through libspkdiar.so."""
import os
import sys

# the package lives next to scripts/ (symlinked installs resolve through realpath); a COPY of this file,
# e.g. inside a checkout of the reference, finds it through SPKDIAR_HOME
sys.path.insert(0, os.environ.get('SPKDIAR_HOME') or os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
import spkdiar  # noqa: E402,F401
from spkdiar import clustering
clustering.main(variant=2)
