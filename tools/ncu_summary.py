"""Summarise an .ncu-rep into the handful of metrics DESIGN.md / bench.py quote.
usage: ncu_summary.py <report.ncu-rep> [out.txt] [launch index, default 0]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + which]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio']
out = [f'# ncu --set full --clock-control none  ({rep.split("/")[-1]})']
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        out.append(f'{h:88s} {v:>22s} {u}')
# instruction mix from the source page
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
# one section per (kernel, view): a "Kernel Name" row, a header row, then instructions.  Keep the
# first section (SASS view) of the `which`-th distinct launch.
sections, cur = [], None
for row in r:
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1] if len(row) > 1 else '', 'hdr': None, 'rows': []}
        sections.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = row
    elif cur is not None:
        cur['rows'].append(row)
want = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
norm = lambda t: t.replace('(int)', '').replace('(bool)', '').replace(' ', '')
cands = [sec for sec in sections if norm(sec['name']).startswith(norm(want)[:60])]
seen_before = sum(1 for rr in rows[2:2 + which] if rr[hdr.index('Kernel Name')] == want)
sass = [sec for sec in cands if sec['hdr'] and 'Source' in sec['hdr']]
pick = None
if sass:
    per = max(1, len(sass) // max(1, sum(1 for rr in rows[2:] if rr[hdr.index('Kernel Name')] == want)))
    pick = sass[min(len(sass) - 1, seen_before * per)]
if pick:
    import collections, re
    h2 = pick['hdr']; ix = {k: i for i, k in enumerate(h2)}
    mix = collections.Counter(); samples = collections.Counter()
    for row in pick['rows']:
        if len(row) < len(h2): continue
        t = re.sub(r'^@!?U?P\w+\s+', '', row[ix['Source']].strip())
        op = t.split()[0].split('.')[0] if t else '?'
        try:
            mix[op] += int(row[ix['Instructions Executed']] or 0)
            samples[op] += int(row[ix['# Samples']] or 0)
        except ValueError:
            continue
    tot = sum(mix.values())
    out.append(f'# executed warp instructions by opcode (total {tot})')
    for op, n in mix.most_common(16):
        out.append(f'  {op:10s} {n:14d}  {100.0 * n / max(tot, 1):5.1f} %   stall samples {samples[op]}')
txt = '\n'.join(out) + '\n'
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(txt)
print(txt)
