#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report (read here, no GPU needed):
per profiled launch the duration, DRAM traffic (read+write = `traffic`), the
throughput figures, occupancy and the dominant warp-stall reasons.

    python tools/ncu_summary.py gpurun_out/r1f_prof.ncu-rep "title" > profiles/r1f_ncu_full.md
"""
import csv
import io
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__throughput.avg.pct_of_peak_sustained_elapsed',
       'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
       'sm__warps_active.avg.pct_of_peak_sustained_active',
       'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
       'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
       'smsp__inst_executed.sum']
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle',
          'mio_throttle', 'lg_throttle', 'barrier', 'not_selected', 'selected',
          'dispatch_stall', 'no_instruction', 'branch_resolving', 'membar',
          'tex_throttle', 'drain', 'imc_miss', 'sleeping']


def main(rep, title='', traffic_json=None):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print('# ncu --set full: ' + title)
    print()
    print('source: `%s` (`--clock-control none`, caches flushed per replay: '
          'durations are cold-cache)' % rep)
    print()
    print('| # | kernel | grid x block | regs | us | DRAM rd MB | DRAM wr MB | '
          'traffic MB | DRAM % | L2 % | L1 % | SM % | L1 hit % | L2 hit % | '
          'occupancy % | fp64 pipe % | top stalls (share of samples) |')
    print('|' + '---|'*17)
    traffic = {}
    for r in rows[2:]:
        g = lambda k: r[col[k]] if k in col else ''
        f = lambda k: float(g(k).replace(',', '')) if g(k) not in ('', 'n/a') else float('nan')
        dur = f('gpu__time_duration.sum')
        if units[col['gpu__time_duration.sum']] == 'ns':
            dur /= 1e3
        elif units[col['gpu__time_duration.sum']] == 'ms':
            dur *= 1e3

        def mb(k):
            v, u = f(k), units[col[k]]
            return v*{'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1., 'Gbyte': 1e3}.get(u, 1.)
        st = {}
        for s in STALLS:
            k = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if k in col and g(k) not in ('', 'n/a'):
                st[s] = f(k)
        tot = sum(st.values()) or 1.
        top = ', '.join('%s %.0f%%' % (k, 100*v/tot) for k, v in
                        sorted(st.items(), key=lambda kv: -kv[1])[:3])
        name = g('Kernel Name').split('(')[0].replace('void ', '')
        rd, wr = mb('dram__bytes_read.sum'), mb('dram__bytes_write.sum')
        t = traffic.setdefault(name, dict(launches=0, traffic_bytes=0., us=0.))
        t['traffic_bytes'] = (t['traffic_bytes']*t['launches'] + 1e6*(rd + wr))/(t['launches'] + 1)
        t['us'] = (t['us']*t['launches'] + dur)/(t['launches'] + 1)
        t['launches'] += 1
        print('| %s | `%s` | %s x %s | %s | %.1f | %.1f | %.1f | %.1f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %s |' % (
            g('ID'), name, g('launch__grid_size'), g('launch__block_size'),
            g('launch__registers_per_thread'), dur, rd, wr, rd + wr,
            f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
            f('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
            f('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
            f('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
            f('l1tex__t_sector_hit_rate.pct'), f('lts__t_sector_hit_rate.pct'),
            f('sm__warps_active.avg.pct_of_peak_sustained_active'),
            f('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), top))


    if traffic_json:
        import json
        with open(traffic_json, 'w') as f:
            json.dump(dict(source=rep, note='mean dram__bytes_read.sum + '
                           'dram__bytes_write.sum per launch (ncu --set full)',
                           **traffic), f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main(*sys.argv[1:4])
