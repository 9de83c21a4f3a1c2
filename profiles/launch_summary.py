"""Aggregates an ncu launch list (--csv, metrics gpu__time_duration.sum [+ dram__bytes_*]) per kernel:

    python profiles/launch_summary.py launches.csv [--by-id]
"""
import collections
import csv
import io
import json
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(io.StringIO(''.join(lines))))
per_id = collections.OrderedDict()
for r in rows:
  d = per_id.setdefault(r['ID'], {'name': r['Kernel Name'], 'grid': r['Grid Size'], 'block': r['Block Size']})
  d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
if '--by-id' in sys.argv:
  for i, d in per_id.items():
    print(i, f"{d.get('gpu__time_duration.sum', 0) / 1e3:9.2f} us", f"rd {d.get('dram__bytes_read.sum', 0) / 1e6:8.2f} MB",
          f"wr {d.get('dram__bytes_write.sum', 0) / 1e6:8.2f} MB", d['grid'], d['block'], d['name'][:110])
agg = collections.OrderedDict()
for d in per_id.values():
  a = agg.setdefault(d['name'][:120], collections.Counter())
  a['n'] += 1
  a['us'] += d.get('gpu__time_duration.sum', 0) / 1e3
  a['rd_mb'] += d.get('dram__bytes_read.sum', 0) / 1e6
  a['wr_mb'] += d.get('dram__bytes_write.sum', 0) / 1e6
tot = sum(a['us'] for a in agg.values())
out = []
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
  out.append(dict(kernel=k, n=a['n'], us=round(a['us'], 2), share=round(a['us'] / tot, 4),
                  dram_read_mb=round(a['rd_mb'], 3), dram_write_mb=round(a['wr_mb'], 3)))
print(json.dumps(dict(total_us=round(tot, 2), launches=len(per_id), kernels=out), indent=1))
