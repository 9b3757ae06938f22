"""Text glue on both sides of the hot path: what produces the
recipe ``spk-change-detection.py`` reads and what consumes the recipe ``spk-clustering.py``
writes in ``spk-diarization2.py`` (lines 111-112 and 131-132).

* ``voice-detection2.py``: the ``.exp`` token stream of the speech-activity decoder
  (``<frame> p`` = speech, ``<frame> <w>`` = silence) -> a recipe of speech turns, with
  minimum-speech / minimum-nonspeech hysteresis (voice-detection2.py:44-114) and the LNA naming
  the later stages rely on (one letter sequence per wav, ``a_1, a_2 ...``; 33-41, 125-131);
* ``aku2ann.py``: a recipe -> the simple annotation format (``start<TAB>end<TAB>speaker``
  under a ``# audio`` header per file; aku2ann.py:31-38).

* ``aku2elan.py``: a recipe -> an ELAN 2.7 annotation document (aku2elan.py:45-99), written
  directly as text in the form lxml gives ``tree.write(..., pretty_print=True)`` (lxml is not
  needed).

Pure host code: no device work, same command lines, same stdout text, Python-2 float text.
"""

import argparse
import os.path as op
import re
import sys

from .py2fmt import fstr, p2line

_AUDIO = re.compile(r'audio=(\S+)')
_TOKEN = re.compile(r'(\d+) (p|<w>)')


def next_lna(lna):
    """'a' -> 'b' ... 'z' -> 'aa' -> 'ab' ... (voice-detection2.py:33-41): the rightmost
    letter that is not 'z' is advanced and everything after it restarts at 'a'."""
    for c in range(len(lna) - 1, -1, -1):
        if lna[c] != 'z':
            return lna[:c] + chr(ord(lna[c]) + 1) + 'a' * (len(lna) - c - 1)
    return 'a' * (len(lna) + 1)


class TurnTracker(object):
    """The speech / silence state machine of voice-detection2.py:44-104, one token at a time.

    A speech token opens a candidate turn; the first silence token after it confirms the turn
    if the speech lasted ``ms`` seconds (else the candidate is dropped); inside a turn a silence
    token marks a candidate end, and the next speech token closes the turn there if the silence
    lasted ``mns`` seconds (else the end is forgotten).  Durations are the time advanced by the
    CURRENT token, as in the reference."""

    def __init__(self, rate, ms, mns, sbe=0.0, see=0.0):
        self.rate = float(rate)
        self.ms, self.mns, self.sbe, self.see = ms, mns, sbe, see
        self.turns = []                 # (start - sbe, end + see)
        self.now = 0.0                  # total_time
        self.last_frame = 0.0
        self.start = 0.0
        self.end = 0.0
        self.in_speech = False

    def feed(self, frame, token):
        frame = float(frame)
        step = (frame - self.last_frame) / self.rate
        self.last_frame = frame
        self.now += step
        speech = token == 'p'
        if self.in_speech:
            if not speech:
                self.end = self.now
            elif self.end:
                if step < self.mns:
                    self.end = 0.0
                else:
                    self.in_speech = False
                    self.turns.append((self.start - self.sbe, self.end + self.see))
                    self.start = self.now
        elif speech:
            self.start = self.now
        elif self.start:
            if step < self.ms:
                self.start = 0.0
            else:
                self.end = self.now
                self.in_speech = True

    def finish(self, last_frame):
        """voice-detection2.py:105-113: a turn still open at the end of the file is kept
        if it is at least ``ms`` long, ending at the last frame (no end expansion)."""
        if self.start:
            t_end = float(last_frame) / self.rate
            if t_end - self.start >= self.ms:
                self.turns.append((self.start - self.sbe, t_end))
        return self.turns


def exp_turns(expfile, rate, ms, mns, sbe=0.0, see=0.0):
    tr = TurnTracker(rate, ms, mns, sbe, see)
    with open(expfile, 'r') as f:
        for line in f:
            for m in _TOKEN.finditer(line):
                tr.feed(m.group(1), m.group(2))
    if tr.start:
        with open(op.splitext(expfile)[0] + '.last_frame', 'r') as f:
            return tr.finish(float(f.read()))
    return tr.turns


def vad_parser():
    """The command line of voice-detection2.py:133-170."""
    p = argparse.ArgumentParser(description='Creates a recipe from the Speech Activity Detection '
                                'generate_exp.py output (.exp files), that is, speech/non-speech '
                                'turn detection')
    p.add_argument('recfile', type=str, help='Specifies the input recipe file')
    p.add_argument('exppath', type=str, help='Specifies the input .exp files path')
    p.add_argument('-o', dest='outfile', type=str, default='stdout', help='Output file, default stdout.')
    p.add_argument('-r', dest='rate', type=int, default=125, help='Sample rate, default 125.')
    p.add_argument('-ms', dest='minspeech', type=float, default=0.2,
                   help='Minimum speech turn duration, default 0.2 seconds.')
    p.add_argument('-mns', dest='minnonspeech', type=float, default=0.3,
                   help='Minimum nonspeech between-turns duration, default 0.3 seconds.')
    p.add_argument('-sbe', dest='seg_before_exp', type=float, default=0.0,
                   help='Time removed before each detected segment, default 0.0.')
    p.add_argument('-see', dest='seg_end_exp', type=float, default=0.0,
                   help='Time added after each detected segment, default 0.0.')
    return p


def vad_main(argv=None, stdout=None):
    """voice-detection2.py:133-198."""
    out = stdout if stdout is not None else sys.stdout
    args = vad_parser().parse_args(argv)

    def log(*items):
        out.write(p2line(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    wavs = []
    with open(args.recfile, 'r') as f:
        for line in f:
            m = _AUDIO.search(line)
            if m is None:
                log('Recipe line without recognizable audio files:')
                log(line)
            else:
                wavs.append(m.group(1))
    log('Reading .exp files from:', args.exppath)
    if not op.isdir(args.exppath):
        log('Error,', args.exppath, 'is not a valid directory')
        raise SystemExit
    log('Writing output to:', args.outfile if args.outfile != 'stdout' else 'stdout')
    log('Sample rate set to:', args.rate)
    log('Minimum speech turn duration:', args.minspeech, 'seconds')
    log('Minimum nonspeech between-turns duration:', args.minnonspeech, 'seconds')
    log('Segment before expansion set to:', args.seg_before_exp, 'seconds')
    log('Segment end expansion set to:', args.seg_end_exp, 'seconds')

    def work(outf):
        lna = 'a'
        for wav in wavs:
            expfile = op.join(args.exppath, op.splitext(op.basename(wav))[0] + '.exp')
            if not op.isfile(expfile):
                log('Error,', expfile, 'does not exist')
                raise SystemExit
            turns = exp_turns(expfile, args.rate, args.minspeech, args.minnonspeech,
                              args.seg_before_exp, args.seg_end_exp)
            for k, (s, e) in enumerate(turns):
                outf.write('audio=%s lna=%s_%d start-time=%s end-time=%s\n' % (wav, lna, k + 1, fstr(s), fstr(e)))
            lna = next_lna(lna)
    if args.outfile != 'stdout':
        with open(args.outfile, 'w') as outf:
            work(outf)
    else:
        work(out)


_ANN = [re.compile(r'audio=(\S+)'), re.compile(r'lna=(\S+)'), re.compile(r'start-time=(\d+.\d+)'),
        re.compile(r'end-time=(\d+.\d+)')]
_SPK = re.compile(r'speaker=(\S+)')


def recipe_to_ann(lines, outf, report=None):
    """aku2ann.py:7-38: lines lacking audio / lna / start / end are reported and skipped, a
    missing speaker tag becomes the empty string."""
    audio = ''
    for text in lines:
        m = [r.search(text) for r in _ANN]
        if any(x is None for x in m):
            if report is not None:
                report('Recipe line without recognizable data:', text)
            continue
        spk = _SPK.search(text)
        if audio != m[0].group(1):
            audio = m[0].group(1)
            outf.write('# ' + audio + '\n')
        outf.write('%s\t%s\t%s\n' % (fstr(float(m[2].group(1))), fstr(float(m[3].group(1))),
                                     spk.group(1) if spk else ''))


def ann_main(argv=None, stdout=None):
    """aku2ann.py:41-69."""
    out = stdout if stdout is not None else sys.stdout
    p = argparse.ArgumentParser(description='Converts an AKU recipe to simple annotation format.')
    p.add_argument('recfile', type=str, help='Specifies the input recipe file')
    p.add_argument('-o', dest='outfile', type=str, default=None, help='Output file, default stdout.')
    args = p.parse_args(argv)

    def log(*items):
        out.write(p2line(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as f:
        lines = f.readlines()
    # the reference parses (and reports) before it announces the output
    import io
    body = io.StringIO()
    recipe_to_ann(lines, body, log)
    if args.outfile is not None:
        log('Writing output to:', args.outfile)
        with open(args.outfile, 'w') as outf:
            outf.write(body.getvalue())
    else:
        log('Writing output to: stdout')
        out.write(body.getvalue())


# ---- aku2elan.py ---------------------------------------------------------------------------------

def iso_now():
    """aku2elan.py:10-17: local time in ISO format with the UTC offset ELAN expects."""
    from datetime import datetime, timezone
    now = datetime.now()
    delta = now - datetime.now(timezone.utc).replace(tzinfo=None)
    hh, mm = divmod((delta.days * 24 * 60 * 60 + delta.seconds + 30) // 60, 60)
    return '%s%+02d:%02d' % (now.isoformat(), hh, mm)


def _xml(text, attr=False):
    """Character escaping of libxml2's ASCII serialisation."""
    out = []
    for ch in text:
        o = ord(ch)
        if ch == '&':
            out.append('&amp;')
        elif ch == '<':
            out.append('&lt;')
        elif ch == '>':
            out.append('&gt;')
        elif attr and ch == '"':
            out.append('&quot;')
        elif (attr and ch in '\n\r\t') or o > 126:
            out.append('&#%d;' % o)
        else:
            out.append(ch)
    return ''.join(out)


def recipe_to_elan(lines, outf, report=None, date=None):
    """aku2elan.py:20-99.  Lines lacking audio / lna / start / end are reported and skipped; two time
    slots and one alignable annotation per line (``int(seconds * 1000)`` milliseconds, fp64 product
    truncated as the reference does); the speaker tag, when present, is the annotation value.  The
    media descriptor is the first line's audio (an empty recipe raises IndexError, an unknown media
    type TypeError - as the reference)."""
    from mimetypes import guess_type
    rows = []
    for text in lines:
        m = [r.search(text) for r in _ANN]
        if any(x is None for x in m):
            if report is not None:                      # two print statements (aku2elan.py:41-42)
                report('Recipe line without recognizable data:')
                report(text)
            continue
        spk = _SPK.search(text)
        rows.append((m[0].group(1), float(m[2].group(1)), float(m[3].group(1)), spk.group(1) if spk else ''))
    media = rows[0][0]
    mime = guess_type(media)[0]
    if mime is None:
        raise TypeError("Argument must be bytes or unicode, got 'NoneType'")
    w = outf.write
    w('<ANNOTATION_DOCUMENT xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" AUTHOR="" DATE="%s" '
      'FORMAT="2.7" VERSION="2.7" xsi:noNamespaceSchemaLocation="http://www.mpi.nl/tools/elan/EAFv2.7.xsd">\n'
      % _xml(date if date is not None else iso_now(), True))
    w('  <HEADER MEDIA_FILE="" TIME_UNITS="milliseconds">\n')
    w('    <MEDIA_DESCRIPTOR MEDIA_URL="%s" MIME_TYPE="%s" RELATIVE_MEDIA_URL=""/>\n'
      % (_xml('file://' + media, True), _xml(mime, True)))
    w('    <PROPERTY NAME="lastUsedAnnotationId">%d</PROPERTY>\n' % len(rows))
    w('  </HEADER>\n')
    w('  <TIME_ORDER>\n')
    for k, (_, start, end, _) in enumerate(rows):
        w('    <TIME_SLOT TIME_SLOT_ID="ts%d" TIME_VALUE="%d"/>\n' % (2 * k + 1, int(start * 1000)))
        w('    <TIME_SLOT TIME_SLOT_ID="ts%d" TIME_VALUE="%d"/>\n' % (2 * k + 2, int(end * 1000)))
    w('  </TIME_ORDER>\n')
    w('  <TIER DEFAULT_LOCALE="en" LINGUISTIC_TYPE_REF="default-lt" TIER_ID="Speakers">\n')
    for k, (_, _, _, spk) in enumerate(rows):
        w('    <ANNOTATION>\n')
        head = '      <ALIGNABLE_ANNOTATION ANNOTATION_ID="a%d" TIME_SLOT_REF1="ts%d" TIME_SLOT_REF2="ts%d"' \
            % (k + 1, 2 * k + 1, 2 * k + 2)
        if spk:
            w(head + '>\n')
            w('        <ANNOTATION_VALUE>%s</ANNOTATION_VALUE>\n' % _xml(spk))
            w('      </ALIGNABLE_ANNOTATION>\n')
        else:
            w(head + '/>\n')
        w('    </ANNOTATION>\n')
    w('  </TIER>\n')
    w('  <LINGUISTIC_TYPE GRAPHIC_REFERENCES="false" LINGUISTIC_TYPE_ID="default-lt" TIME_ALIGNABLE="true"/>\n')
    w('  <LOCALE COUNTRY_CODE="US" LANGUAGE_CODE="en"/>\n')
    w('</ANNOTATION_DOCUMENT>\n')


def elan_main(argv=None, stdout=None, date=None):
    """aku2elan.py:102-128."""
    out = stdout if stdout is not None else sys.stdout
    p = argparse.ArgumentParser(description='Converts an AKU recipe to Elan format.')
    p.add_argument('recfile', type=str, help='Specifies the input recipe file')
    p.add_argument('-o', dest='outfile', type=str, default=None, help='Specifies an output file, default stdout.')
    args = p.parse_args(argv)

    def log(*items):
        out.write(p2line(*items) + '\n')

    log('Reading recipe from:', args.recfile)
    with open(args.recfile, 'r') as f:
        lines = f.readlines()
    import io
    body = io.StringIO()
    # the reference parses (and reports bad lines) before it announces the output
    bad = []
    try:
        recipe_to_elan(lines, body, lambda *items: bad.append(items), date)
    finally:
        for items in bad:
            log(*items)
        if args.outfile is not None:
            log('Writing output to:', args.outfile)
        else:
            log('Writing output to: stdout')
    if args.outfile is not None:
        with open(args.outfile, 'w') as outf:
            outf.write(body.getvalue())
    else:
        out.write(body.getvalue())
