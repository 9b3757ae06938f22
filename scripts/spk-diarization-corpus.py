#!/usr/bin/env python3
"""The two hot-path calls of ./spk-diarization2.py (lines 122-128) over a corpus of recordings:
spk-diarization-corpus.py <recipe>... <feapath> -o <outdir> [-f rate] [--batch N]; one process
per GPU under torchrun shards the recordings by index (speaker-diarization_b200/corpus.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spkdiar  # noqa: E402,F401
from spkdiar import corpus
corpus.main()
