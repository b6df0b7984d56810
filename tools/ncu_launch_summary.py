"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv > profiles/summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    name = re.sub(r'\(.*$', '', r[ik]).replace('vited::<unnamed>::', 'vited::')
    v = float(r[iv].replace(',', ''))
    scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[iu], 1.0)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print('kernel,launches,total_us,share,avg_us')
for k in sorted(tot, key=lambda k: -tot[k]):
    print(f'"{k}",{cnt[k]},{tot[k]:.1f},{tot[k] / total:.4f},{tot[k] / cnt[k]:.2f}')
