"""profiles/ncu_traffic.json from one `ncu --set full` capture of tools/profile_ops.py (read on the CPU box):
per engine kernel class, dram__bytes_read.sum + dram__bytes_write.sum of its launch and the algorithmic bytes of that
launch. bench.py fills `roofline.traffic` from this file when its own launches have the same algorithmic bytes.
usage: python tools/ncu_traffic.py capture.ncu-rep manifest.json out.json"""
import csv
import io
import json
import subprocess
import sys

rep, manifest_path, out_path = sys.argv[1:4]
manifest = json.load(open(manifest_path))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def to_bytes(v, unit):
    scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[unit]
    return float(v) * scale


launches = manifest['launches']
want = [l for l in launches]
captured = [r for r in data]
# the capture may hold fewer kernels than the manifest (-k filter): pair in order by kernel-name substring
out, ci = {}, 0
for l in want:
    while ci < len(captured) and l['kernel'] not in captured[ci][col['Kernel Name']]:
        ci += 1
    if ci == len(captured):
        break
    r = captured[ci]
    ci += 1
    rd = to_bytes(r[col['dram__bytes_read.sum']], units[col['dram__bytes_read.sum']])
    wr = to_bytes(r[col['dram__bytes_write.sum']], units[col['dram__bytes_write.sum']])
    out[l['cls']] = dict(kernel=r[col['Kernel Name']][:80], dram_bytes=rd + wr, dram_read=rd, dram_write=wr,
                         algorithmic_bytes=l['algorithmic_bytes'], ratio=(rd + wr) / l['algorithmic_bytes'],
                         duration_us=float(r[col['gpu__time_duration.sum']]) *
                         {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(units[col['gpu__time_duration.sum']], 1.0),
                         tensor_pipe_pct=float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]))
json.dump(dict(source=f'ncu --set full --clock-control none of tools/profile_ops.py at {manifest["rows"]} rows per launch '
                      f'({rep.split("/")[-1]}); dram__bytes_read.sum + dram__bytes_write.sum per launch',
               rows=manifest['rows'], kernels=out), open(out_path, 'w'), indent=1)
for k, v in out.items():
    print(f"{k:24s} dram {v['dram_bytes'] / 1e9:7.3f} GB  algorithmic {v['algorithmic_bytes'] / 1e9:7.3f} GB  ratio {v['ratio']:.3f}  "
          f"{v['duration_us']:8.1f} us  tensor pipe {v['tensor_pipe_pct']:.1f} %")
