#!/usr/bin/env python3
"""Summarise an .ncu-rep: key raw metrics of each profiled launch + the hottest source lines by stall
samples (barrier stalls excluded so that waiting warps do not drown the working ones)."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum', 'smsp__inst_executed_pipe_lsu.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'launch__occupancy_per_block_size',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print('==', d.get('Kernel Name', '?')[:90])
        for h, u in zip(hdr, units):
            if h in WANT:
                print('   %-70s %s %s' % (h, d[h], u))


def source(rep, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    for si in start[:1]:
        hdr = rows[si]
        ci = {h: i for i, h in enumerate(hdr)}
        data = [r for r in rows[si + 1:] if len(r) == len(hdr)]
        stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = sum(int(r[ci['# Samples']]) for r in data)
        agg = sorted(((sum(int(r[ci[s]]) for r in data), s) for s in stalls), reverse=True)
        print('total samples', tot, [(s, n) for n, s in agg[:8]])
        nb = sum(int(r[ci['# Samples']]) - int(r[ci['stall_barrier']]) for r in data)
        print('non-barrier samples', nb)
        scored = sorted(((int(r[ci['# Samples']]) - int(r[ci['stall_barrier']]), i) for i, r in enumerate(data)), reverse=True)
        for sc, i in scored[:top]:
            r = data[i]
            st = sorted(((int(r[ci[s]]), s[6:]) for s in stalls if s != 'stall_barrier'), reverse=True)[:3]
            print('%6d  #%-5d %-60s exec %-9s %s' % (sc, i, r[ci['Source']].strip()[:60], r[ci['Instructions Executed']],
                                                    ' '.join('%s=%d' % (n, v) for v, n in st if v)))
        # cumulative samples by instruction index range (to see phase shares)
        cum = 0
        marks = []
        for i, r in enumerate(data):
            cum += int(r[ci['# Samples']]) - int(r[ci['stall_barrier']])
            if 'BAR.' in r[ci['Source']]:
                marks.append((i, cum, r[ci['Source']].strip()[:40]))
        prev = 0
        print('non-barrier samples between barriers:')
        for i, c, s in marks:
            print('   up to #%d %-40s +%d' % (i, s, c - prev))
            prev = c
        print('   tail +%d' % (cum - prev))


if __name__ == '__main__':
    rep = sys.argv[1]
    raw(rep)
    source(rep, int(sys.argv[2]) if len(sys.argv) > 2 else 30)
