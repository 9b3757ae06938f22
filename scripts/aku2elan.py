#!/usr/bin/env python3
"""Drop-in for the reference's ./aku2elan.py (same flags, same files in and out, same stdout
text); host-only text glue around the hot path (speaker-diarization_b200/glue.py)."""
import os
import sys

# the package lives next to scripts/ (symlinked installs resolve through realpath); a COPY of this file,
# e.g. inside a checkout of the reference, finds it through SPKDIAR_HOME
sys.path.insert(0, os.environ.get('SPKDIAR_HOME') or os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
import spkdiar  # noqa: E402,F401
from spkdiar import glue
glue.elan_main()
