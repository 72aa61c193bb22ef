"""Summarise an ncu report (raw page + hottest SASS lines of the source page). Usage: ncu_summary.py rep [kernel_idx]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__cycles_active.avg', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'lts__t_sectors_srcunit_tex_op_read.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {rows[1][i]}")
    print('---')
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else len(rows) - 3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(kidx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; secs.append(cur); continue
    if cur is not None: cur['rows'].append(r)
for sec in secs[:1]:
    h = sec['rows'][0]; idx = {x: i for i, x in enumerate(h)}
    data = [r for r in sec['rows'][1:] if len(r) > 10]
    si, ei = idx['# Samples'], idx['Instructions Executed']
    print(sec['name'][:80], 'samples', sum(int(r[si]) for r in data), 'inst', sum(int(r[ei]) for r in data), 'nsass', len(data))
    for i in sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:28]:
        r = data[i]
        print(str(i).rjust(5), r[si].rjust(7), r[ei].rjust(9), r[idx['Source']].strip()[:100])
    # block summary
    for i in range(0, len(data), 50):
        ex = sum(int(data[j][ei]) for j in range(i, min(i + 50, len(data))))
        sm = sum(int(data[j][si]) for j in range(i, min(i + 50, len(data))))
        print('blk', i, ex, sm, data[i][idx['Source']].strip()[:50])
