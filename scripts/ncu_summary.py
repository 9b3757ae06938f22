#!/usr/bin/env python3
"""Turn ncu output brought back in gpurun_out/ into the text summaries kept under profiles/.

  ncu_summary.py launches <launches.csv>            per-kernel totals and shares of a
                                                    `--metrics gpu__time_duration.sum` launch list
  ncu_summary.py report <file.ncu-rep> [regex]      key metrics + top stall reasons / opcodes / source
                                                    lines of every profiled launch (needs `ncu` on PATH)
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__cycles_active.avg']


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO(''.join(lines))))
    agg = collections.OrderedDict()
    order = []
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = re.sub(r'\(.*', '', r['Kernel Name'])
        ns = float(r['Metric Value'].replace(',', ''))
        order.append((k, ns, r['Grid Size'], r['Block Size']))
        a = agg.setdefault(k, [0, 0.0, r['Grid Size'], r['Block Size']])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print('launch list: %s' % path)
    print('%d launches, %.3f ms of kernel time in total (ncu serialises and replays with cold caches: '
          'compare shares, not absolutes)\n' % (len(order), tot / 1e6))
    print('%-36s %5s %12s %7s  %-14s %s' % ('kernel', 'n', 'total ms', 'share', 'grid', 'block'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-36s %5d %12.3f %6.1f%%  %-14s %s' % (k, a[0], a[1] / 1e6, 100 * a[1] / tot, a[2], a[3]))
    print('\nin launch order:')
    for k, ns, g, b in order:
        print('  %-36s %10.3f ms  grid %s block %s' % (k, ns / 1e6, g, b))


def report(path, pattern=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print('ncu report: %s' % path)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if pattern and not re.search(pattern, d['Kernel Name']):
            continue
        print('\n== %s  grid %s block %s' % (d['Kernel Name'], d.get('Grid Size'), d.get('Block Size')))
        for k in KEYS:
            if k in d and d[k] not in ('', 'n/a'):
                print('   %-70s %s %s' % (k, d[k], units[hdr.index(k)]))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks = re.split(r'(?m)^"Kernel Name",', src)
    for blk in blocks[1:]:
        lines = blk.splitlines()
        name = lines[0].strip('",')
        if pattern and not re.search(pattern, name):
            continue
        rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
        hdr, data = rows[0], [r for r in rows[1:] if len(r) == len(rows[0])]
        ix = {h: i for i, h in enumerate(hdr)}
        if '# Samples' not in ix:
            continue
        tot = sum(int(r[ix['# Samples']] or 0) for r in data) or 1
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
        print('\n-- source page of %s: %d samples over %d SASS instructions' % (name, tot, len(data)))
        print('   stall reasons: ' + ', '.join('%s %.1f%%' % (s[6:], 100.0 * v / tot)
                                               for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
        op = collections.Counter()
        ex = collections.Counter()
        for r in data:
            t = r[ix['Source']].strip().split()
            o = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
            op[o] += int(r[ix['# Samples']] or 0)
            ex[o] += int(r[ix['Instructions Executed']] or 0)
        print('   by opcode (samples share, warp-level executions):')
        for o, v in op.most_common(14):
            print('      %-24s %5.1f%%  %d' % (o, 100.0 * v / tot, ex[o]))
        print('   hottest instructions:')
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:12]:
            print('      %5.1f%%  %s' % (100.0 * int(r[ix['# Samples']]) / tot, r[ix['Source']].strip()[:110]))


if __name__ == '__main__':
    if len(sys.argv) < 3 or sys.argv[1] not in ('launches', 'report'):
        sys.exit(__doc__)
    if sys.argv[1] == 'launches':
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
