"""Scoring tools: Python-3 versions of ``spk-change-performance.py`` (segment
boundary precision / recall / F1) and ``clus-performance.py`` (1-ms frame-level
diarization error with a greedy label mapping).

CPU tools in the reference and here (they read two recipes and print a
report); they are part of the package because the acceptance criterion is that
both report identical scores for our recipes and the reference's.  Same
command lines and report text.  ``clus-performance`` is evaluated on runs of
identical labels instead of materialising one list entry per millisecond (a
24-hour recording would need 86 M entries per recipe) - the counts are the
same integers.
"""

import argparse
import re
import sys

from .py2fmt import p2line

_AUDIO = re.compile(r'audio=(\S+)')
_START = re.compile(r'start-time=(\d+.\d+)')
_END = re.compile(r'end-time=(\d+.\d+)')
_SPK = re.compile(r'speaker=(\S+)')


def parse_scored_recipe(lines, report):
    """Both scripts' ``parse_recipe`` (spk-change-performance.py:6-37,
    clus-performance.py:9-40): segments grouped per consecutive audio file.
    Kept quirk: when the file changes, the finished group is stored under the
    NEW file's name."""
    groups = []
    cur = []
    this_file = ''
    total = 0
    for text in lines:
        ma, ms, me = _AUDIO.search(text), _START.search(text), _END.search(text)
        if ma is None or ms is None or me is None:
            report('Recipe line without recognizable data:')
            report(text)
            continue
        audio = ma.group(1)
        start, end = float(ms.group(1)), float(me.group(1))
        mk = _SPK.search(text)
        speaker = mk.group(1) if mk else ''
        if audio != this_file:
            if this_file != '':
                groups.append((audio, cur))
                cur = []
            this_file = audio
        cur.append((start, end, speaker))
        total += 1
    groups.append((this_file, cur))
    return groups, total


# ---- spk-change-performance.py ---------------------------------------------------

class _Matcher(object):
    """The greedy alignment spk-change-performance.py:83-107 runs three times
    (all boundaries, starts, ends): walk the proposed values in order, advance
    through the baseline values; within +-threshold = correct, baseline behind =
    deletion, baseline ahead = insertion.  The baseline cursor is NOT reset
    between files (the script never resets ``countb``)."""

    def __init__(self, threshold, say, show, what):
        self.t = threshold
        self.say = say
        self.sc, self.si, self.sd = show
        self.what = what
        self.cursor = 0
        self.correct = self.ins = self.dele = 0

    def feed(self, value, baseline):
        while self.cursor < len(baseline):
            b = baseline[self.cursor]
            if value - self.t <= b <= value + self.t:
                if self.sc:
                    self.say(p2line('Correct %s:' % self.what, b, 'proposed:', value))
                self.correct += 1
                self.cursor += 1
                return
            elif b < value:
                if self.sd:
                    self.say(p2line('%s deletion:' % self.what.capitalize(), b, 'proposed:', value))
                self.dele += 1
            elif b > value:
                if self.si:
                    self.say(p2line('%s insertion:' % self.what.capitalize(), b, 'proposed:', value))
                self.ins += 1
                return
            self.cursor += 1
        if self.si:
            self.say(p2line('%s insertion:' % self.what.capitalize(), 0, 'proposed:', value))
        self.ins += 1


def _durations(groups):
    total = mx = mn = 0.0
    for _, segs in groups:
        for s in segs:
            d = s[1] - s[0]
            if d > mx:
                mx = d
            if d < mn or mn == 0.0:
                mn = d
            total += d
    return total, mx, mn


def change_performance(baseline, totalb, proposed, totalp, threshold, say, show=(False, False, False)):
    """spk-change-performance.py:60-246."""
    times_b = [[v for s in segs for v in (s[0], s[1])] for _, segs in baseline]
    times_p = [[v for s in segs for v in (s[0], s[1])] for _, segs in proposed]
    total_timesb = sum(len(t) for t in times_b)
    m_all = _Matcher(threshold, say, show, 'time')
    pos = 0                                   # the script's countr is not reset between files either
    for f in range(len(times_b)):
        while pos < len(times_p[f]):
            v = times_p[f][pos]
            pos += 1
            m_all.feed(v, times_b[f])
    m_s = _Matcher(threshold, say, show, 'start')
    m_e = _Matcher(threshold, say, show, 'end')
    pos = 0
    for f in range(len(baseline)):
        bs = [s[0] for s in baseline[f][1]]
        be = [s[1] for s in baseline[f][1]]
        while pos < len(proposed[f][1]):
            seg = proposed[f][1][pos]
            pos += 1
            m_s.feed(seg[0], bs)
            m_e.feed(seg[1], be)
    sb, sp = _durations(baseline), _durations(proposed)
    c = (float(m_all.correct), float(m_s.correct), float(m_e.correct))
    ins = (m_all.ins, m_s.ins, m_e.ins)
    dele = (m_all.dele, m_s.dele, m_e.dele)
    prec = [c[k] / (c[k] + ins[k]) for k in range(3)]
    rec = [c[k] / (c[k] + dele[k]) for k in range(3)]
    f1 = [2 * (prec[k] * rec[k] / (prec[k] + rec[k])) for k in range(3)]
    say('{0:>58}'.format('Benchmark results:          '))
    say('{0:>58}'.format('------------------          '))
    say('{0:<36} {1:^5}'.format('Correct:', int(c[0])))
    say('{0:<36} {1:^5}'.format('Inserted:', int(ins[0])))
    say('{0:<36} {1:^5}'.format('Deleted:', int(dele[0])))
    say('{0:<36} {1:^5}'.format('Substituted:', 0))
    say('{0:<36} {1:^5.3f}'.format('Accuracy:', c[0] / total_timesb))
    say('{0:<36} {1:^5.3f}'.format('Precision:', prec[0]))
    say('{0:<36} {1:^5.3f}'.format('Recall:', rec[0]))
    say('{0:<36} {1:^5.3f}'.format('F1:', f1[0]))
    say('')
    say('{0:>58}'.format('Detailed results:           '))
    say('{0:>58}'.format('-----------------           '))
    say('{0:>58}'.format('Start points                 End points'))
    say('{0:>58}'.format('---------------------------------------'))
    row = '{0:<23} {1:^5}          -           {2:^5}'
    rowf = '{0:<23} {1:^5.3f}          -           {2:^5.3f}'
    say(row.format('Correct:', int(c[1]), int(c[2])))
    say(row.format('Inserted:', int(ins[1]), int(ins[2])))
    say(row.format('Deleted:', int(dele[1]), int(dele[2])))
    say(row.format('Substituted:', 0, 0))
    say(rowf.format('Accuracy:', c[1] / totalb, c[2] / totalb))
    say(rowf.format('Precision:', prec[1], prec[2]))
    say(rowf.format('Recall:', rec[1], rec[2]))
    say(rowf.format('F1:', f1[1], f1[2]))
    say('')
    say('{0:>56}'.format('Segment stats (seconds):    '))
    say('{0:>56}'.format('------------------------    '))
    say('{0:>56}'.format('Baseline                 Proposed'))
    say('{0:>56}'.format('---------------------------------'))
    rowd = '{0:<23} {1:>5.2f}          -           {2:>5.2f}'
    say(rowd.format('Average duration:', sb[0] / totalb, sp[0] / totalp))
    say(rowd.format('Maximum duration:', sb[1], sp[1]))
    say(rowd.format('Minimum duration:', sb[2], sp[2]))
    return dict(correct=c, inserted=ins, deleted=dele, precision=prec, recall=rec, f1=f1)


def _score_parser():
    p = argparse.ArgumentParser(description='Rate a recipe against another, typically to benchmark '
                                'diarization performance.')
    p.add_argument('baseline', type=str, help='Specifies the baseline recipe file.')
    p.add_argument('proposed', type=str, help='Specifies the proposed recipe file, to benchmark.')
    p.add_argument('-o', dest='outfile', type=str, default=None,
                   help='Specifies an output file, default stdout.')
    p.add_argument('-t', dest='threshold', type=float, default=0.25,
                   help='Threshold to determine when a time is incorrect, default 0.25 seconds.')
    p.add_argument('-sc', action='store_true', help='If set, shows the time of each correct')
    p.add_argument('-si', action='store_true', help='If set, shows the time of each insertion')
    p.add_argument('-sd', action='store_true', help='If set, shows the time of each deletion')
    p.add_argument('-ss', action='store_true', help='If set, shows the time of each substitution')
    return p


def _score_prologue(args, say):
    say(p2line('Reading baseline recipe from:', args.baseline))
    with open(args.baseline, 'r') as f:
        base, totalb = parse_scored_recipe(f, say)
    say(p2line('Reading proposed recipe from:', args.proposed))
    with open(args.proposed, 'r') as f:
        prop, totalp = parse_scored_recipe(f, say)
    if args.outfile is not None:
        say(p2line('Writing output to:', args.outfile))
        open(args.outfile, 'w').close()        # the scripts open (and never write) the file
    else:
        say('Writing output to: stdout')
    say(p2line('Threshold:', args.threshold))
    return base, totalb, prop, totalp


def change_performance_main(argv=None, stdout=None):
    """spk-change-performance.py:248-303."""
    out = stdout if stdout is not None else sys.stdout
    args = _score_parser().parse_args(argv)

    def say(text):
        out.write(text + '\n')
    base, totalb, prop, totalp = _score_prologue(args, say)
    return change_performance(base, totalb, prop, totalp, args.threshold, say,
                              (args.sc, args.si, args.sd))


# ---- clus-performance.py ----------------------------------------------------------

def _label_runs(segments, resolution):
    """Run-length form of ``_generate_labeled_list`` (clus-performance.py:54-74):
    [(label or None, count)], count = int((end - start) / resolution)."""
    runs = []
    now = 0.0
    for start, end, label in segments:
        if start > now:
            runs.append((None, int((start - now) / resolution)))
        runs.append((label, int((end - start) / resolution)))
        now = end
    return [r for r in runs if r[1] > 0]


def _co_occurrence(runs_a, runs_b):
    """Counts of (label_a, label_b) over ``zip(list_a, list_b)`` - in order of
    first appearance, as the script's dict of dicts fills."""
    counts = {}
    ia = ib = 0
    left_a = runs_a[0][1] if runs_a else 0
    left_b = runs_b[0][1] if runs_b else 0
    while ia < len(runs_a) and ib < len(runs_b):
        n = min(left_a, left_b)
        key = (runs_a[ia][0], runs_b[ib][0])
        counts[key] = counts.get(key, 0) + n
        left_a -= n
        left_b -= n
        if left_a == 0:
            ia += 1
            left_a = runs_a[ia][1] if ia < len(runs_a) else 0
        if left_b == 0:
            ib += 1
            left_b = runs_b[ib][1] if ib < len(runs_b) else 0
    return counts


def clus_performance(baseline, proposed, resolution, say):
    """clus-performance.py:77-108: every baseline label maps to the proposed
    label it co-occurs with most; time on the mapped label is correct."""
    counts = _co_occurrence(_label_runs(baseline[0][1], resolution),
                            _label_runs(proposed[0][1], resolution))
    # sorted(..., key=count, reverse=True) is stable: ties keep dict order, which is
    # baseline-label-major in order of first appearance
    order = {}
    for (a, b) in counts:
        order.setdefault(a, len(order))
    ranked = sorted(counts.items(), key=lambda kv: order[kv[0][0]])
    ranked.sort(key=lambda kv: kv[1], reverse=True)
    best = {}
    for (a, b), _ in ranked:
        if a not in best:
            best[a] = b
    correct = float(sum(n for (a, b), n in counts.items() if best[a] == b))
    incorrect = float(sum(n for (a, b), n in counts.items() if best[a] != b))
    say(p2line('Correct time:', correct * resolution))
    say(p2line('Incorrect time:', incorrect * resolution))
    say(p2line('Total time:', (incorrect + correct) * resolution))
    say(p2line('DER:', incorrect / (incorrect + correct)))
    return dict(correct=correct * resolution, incorrect=incorrect * resolution,
                der=incorrect / (incorrect + correct))


def clus_performance_main(argv=None, stdout=None):
    """clus-performance.py:111-155 (resolution fixed at 1 ms, line 153)."""
    out = stdout if stdout is not None else sys.stdout
    args = _score_parser().parse_args(argv)

    def say(text):
        out.write(text + '\n')
    base, _, prop, _ = _score_prologue(args, say)
    return clus_performance(base, prop, 0.001, say)
