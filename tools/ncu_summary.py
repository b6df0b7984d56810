"""Summarise an .ncu-rep (read on the CPU box): key pipe/throughput metrics per captured launch and the top stall
instructions of one launch. usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_top] [launch index, default 0]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f'{w} [{units[i]}]:', [r[i][:60] for r in data])
for i, h in enumerate(hdr):
    if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        v = float(data[launch][i] or 0)
        if v > 0.2:
            print('  stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = []
        blocks.append(cur)
        continue
    if cur is not None:
        cur.append(r)
b = blocks[launch]
print('launch', launch, data[launch][hdr.index('Kernel Name')][:70])
h, d = b[0], b[1:]
iS, iSrc, iEx = h.index('Warp Stall Sampling (All Samples)'), h.index('Source'), h.index('Instructions Executed')
tot = sum(int(r[iS]) for r in d if len(r) > iS and r[iS].isdigit())
print('total samples', tot, 'instructions', len(d))
top = sorted([(int(r[iS]), i, r[iSrc].strip(), r[iEx]) for i, r in enumerate(d) if len(r) > iS and r[iS].isdigit()], reverse=True)[:ntop]
for s, i, sr, ex in top:
    print(f'{s:7d} {100 * s / tot:5.1f}% idx={i:5d} exec={ex:>9s} {sr[:90]}')
