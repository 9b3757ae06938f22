#!/usr/bin/env python3
"""The two hot-path calls of ./spk-diarization2.py (lines 122-128) over a corpus of recordings:
spk-diarization-corpus.py <recipe>... <feapath> -o <outdir> [-f rate] [--batch N]; one process
per GPU under torchrun shards the recordings by index (speaker-diarization_b200/corpus.py)."""
import os
import sys

# the package lives next to scripts/ (symlinked installs resolve through realpath); a COPY of this file,
# e.g. inside a checkout of the reference, finds it through SPKDIAR_HOME
sys.path.insert(0, os.environ.get('SPKDIAR_HOME') or os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
import spkdiar  # noqa: E402,F401
from spkdiar import corpus
corpus.main()
