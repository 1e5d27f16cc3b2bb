#!/usr/bin/env python
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into profiles/<tag>_ncu_summary.csv and
profiles/<tag>_source_hotspots.txt.   usage: python profiles/summarize.py gpurun_out/prof_X.ncu-rep TAG "note" """
import collections, csv, io, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tmem.sum.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
with open(f'profiles/{tag}_ncu_summary.csv', 'w') as f:
    f.write(f'# {note}\n')
    f.write('metric,unit,' + ','.join(f'launch{i}' for i in range(len(rows) - 2)) + '\n')
    for i, h in enumerate(hdr):
        if h in KEEP:
            f.write(','.join([h, rows[1][i]] + ['"%s"' % r[i] if ',' in r[i] else r[i] for r in rows[2:]]) + '\n')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass,cuda'], capture_output=True, text=True).stdout
by = collections.defaultdict(collections.Counter); samp = collections.defaultdict(collections.Counter); text = {}
stall = collections.defaultdict(collections.Counter)
fname = func = None; ix = None
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No':
        ix = {}
        for i, n in enumerate(r): ix.setdefault(n, i)
        continue
    if ix is None or func is None: continue
    try:
        ln = int(r[0]); n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    except Exception:
        continue
    by[func][(fname, ln)] += n; samp[func][(fname, ln)] += s; text[(fname, ln)] = r[1]
    for c, i in ix.items():
        if c.startswith('stall_') and 'Not Issued' not in c:
            try: stall[func][c] += int(r[i])
            except Exception: pass
with open(f'profiles/{tag}_source_hotspots.txt', 'w') as f:
    f.write(f'# {note}\n')
    for func in by:
        tot = sum(by[func].values()); st = max(1, sum(samp[func].values()))
        f.write(f'\n## {func}\ninstructions executed: {tot}   stall samples: {st}\n')
        ss = max(1, sum(stall[func].values()))
        f.write('stall reasons: ' + ', '.join(f'{c[6:]} {100*n/ss:.1f}%' for c, n in stall[func].most_common(8)) + '\n')
        for k, n in by[func].most_common(45):
            f.write(f'{k[0]:18s}:{k[1]:4d} inst {100*n/tot:5.1f}% samp {100*samp[func][k]/st:5.1f}%  {text[k].strip()[:120]}\n')
print(open(f'profiles/{tag}_ncu_summary.csv').read())
