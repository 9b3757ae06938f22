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
_CANON = re.compile(r'audio=(\S+) lna=(\S+) start-time=(\d+.\d+) end-time=(\d+.\d+)')


def parse(lines, report=None):
    """-> list of Line.  Lines lacking one of the four fields are handed to
    ``report`` (the two messages the reference prints) and skipped."""
    out = []
    for text in lines:
        # Fast path for lines in the writers' own field order: one match instead of four searches - taken
        # only when every key's FIRST occurrence is the one the match used, i.e. when the four independent
        # searches of the reference would return the same groups (a path such as /x/lna=y.wav falls through).
        m = _CANON.match(text)
        if m is not None and text.find('lna=') == m.start(2) - 4 and text.find('start-time=') == m.start(3) - 11 \
                and text.find('end-time=') == m.start(4) - 9:
            out.append(Line(m.group(1), m.group(2), float(m.group(3)), float(m.group(4))))
            continue
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


_TIME = re.compile(r'\d+.\d+')


def lines_from_records(records, texts):
    """What ``parse(texts)`` returns, from the fields a ``Writer`` recorded while it wrote those
    lines (``Writer.record``) - without the regular-expression searches.  A time whose text the
    reference's pattern would not take whole (an exponent, a sign: ``\\d+.\\d+`` with its unescaped
    dot) sends that line through ``parse``, so the result is the same in every case."""
    out = []
    for rec, text in zip(records, texts):
        audio, lna, t0, t1 = rec
        if _TIME.fullmatch(t0) and _TIME.fullmatch(t1) and 'lna=' not in audio and 'start-time=' not in audio \
                and 'end-time=' not in audio and 'start-time=' not in lna and 'end-time=' not in lna:
            out.append(Line(audio, lna, float(t0), float(t1)))
        else:
            out.extend(parse([text]))
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
        self.record = None          # a list: every written line is also kept as (audio, lna, t0 text, t1 text)

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
        if self.record is not None:
            self.record.append((line.audio, lna, t0, t1))
        if self.segprefix and segf is not None:
            segf.write('audio=%s alignment=%s%s.seg lna=%s start-time=%s end-time=%s speaker=%s\n'
                       % (line.audio, self.segprefix, lna, lna, t0, t1, tag))
