"""AKU recipe text: the wire format between the stages of the pipeline.

One segment per line, space separated ``key=value`` fields found by regular
expression (spk-change-detection.py:11-28).  Writing
reproduces ``write_recipe_line`` (spk-change-detection.py:46-69,
spk-clustering.py:55-78): LNA renaming with a letter/counter state that
persists over the whole run, and Python-2 float text.
"""

import re
from collections import namedtuple

from .py2fmt import fstr

Line = namedtuple('Line', 'audio lna start end')

_AUDIO = re.compile(r'audio=(\S+)')
_LNA = re.compile(r'lna=(\S+)')
# the reference's pattern is '\d+.\d+' with an UNESCAPED dot: digit(s), any one
# character, digit(s) - "5" alone does not match and the line is reported
_START = re.compile(r'start-time=(\d+.\d+)')
_END = re.compile(r'end-time=(\d+.\d+)')


def parse(lines, report=None):
    """-> list of Line.  Lines lacking one of the four fields are handed to
    ``report`` (the two messages the reference prints) and skipped."""
    out = []
    for text in lines:
        ma, ml, ms, me = (_AUDIO.search(text), _LNA.search(text), _START.search(text),
                          _END.search(text))
        # the reference evaluates the fields in this order and gives up at the
        # first miss; float() of a match such as '1x5' raises ValueError there too
        if ma is None or ml is None or ms is None or me is None:
            if report is not None:
                report('Recipe line without recognizable data:')
                report(text)
            continue
        out.append(Line(ma.group(1), ml.group(1), float(ms.group(1)), float(me.group(1))))
    return out


class Writer(object):
    """Recipe line writer with the LNA renaming state of the scripts.

    ``lna[:lna.find('_')]`` is the prefix; without an underscore ``find`` gives
    -1 and the prefix silently drops the last character (kept: SURVEY.md Q11).
    ``segprefix``: value of ``-seg`` when alignment lines are wanted."""

    def __init__(self, rate, rename=True, segprefix=None):
        self.rate = rate
        self.rename = rename
        self.segprefix = segprefix
        self.letter = 'a'
        self.count = 0

    def _lna(self, lna):
        if not self.rename:
            return lna
        cut = lna.find('_')
        if lna[:cut] == self.letter:
            self.count += 1
        else:
            self.count = 1
            self.letter = lna[:cut]
        return lna[:cut + 1] + str(self.count)

    def write(self, line, start, end, lna_start, tag, outf, segf=None):
        """``start`` / ``end`` in frames; time = frames / rate + lna_start."""
        lna = self._lna(line.lna)
        t0 = fstr(start / self.rate + lna_start)
        t1 = fstr(end / self.rate + lna_start)
        outf.write('audio=%s lna=%s start-time=%s end-time=%s speaker=%s\n'
                   % (line.audio, lna, t0, t1, tag))
        if self.segprefix and segf is not None:
            segf.write('audio=%s alignment=%s%s.seg lna=%s start-time=%s end-time=%s speaker=%s\n'
                       % (line.audio, self.segprefix, lna, lna, t0, t1, tag))
