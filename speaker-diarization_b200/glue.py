"""Text glue on both sides of the hot path: what produces the
recipe ``spk-change-detection.py`` reads and what consumes the recipe ``spk-clustering.py``
writes in ``spk-diarization2.py`` (lines 111-112 and 131-132).

* ``voice-detection2.py``: the ``.exp`` token stream of the speech-activity decoder
  (``<frame> p`` = speech, ``<frame> <w>`` = silence) -> a recipe of speech turns, with
  minimum-speech / minimum-nonspeech hysteresis (voice-detection2.py:44-114) and the LNA naming
  the later stages rely on (one letter sequence per wav, ``a_1, a_2 ...``; 33-41, 125-131);
* ``aku2ann.py``: a recipe -> the simple annotation format (``start<TAB>end<TAB>speaker``
  under a ``# audio`` header per file; aku2ann.py:31-38).

Pure host code: no device work, same command lines, same stdout text, Python-2 float text.
``aku2elan.py`` (ELAN XML through lxml) is not built.
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


def vad_main(argv=None, stdout=None):
    """voice-detection2.py:133-198."""
    out = stdout if stdout is not None else sys.stdout
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
    args = p.parse_args(argv)

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
