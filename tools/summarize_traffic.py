"""Summarise an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch into
per-kernel-family DRAM traffic, and write the JSON bench.py reads for `roofline.traffic`:

  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file gpurun_out/traffic.csv python tools/ncu_step.py --batch 1024 --precision f16
  python tools/summarize_traffic.py gpurun_out/traffic.csv profiles/traffic_f16.json
"""
import collections
import csv
import json
import re
import sys

SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}


def main():
  lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
  per = collections.defaultdict(lambda: dict(read=0.0, write=0.0, us=0.0))
  for row in csv.DictReader(lines):
    m = row.get('Metric Name')
    if m not in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'):
      continue
    v = float(row['Metric Value'].replace(',', '')) * SCALE.get(row['Metric Unit'], 1.0)
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('b200::', '').replace('<unnamed>::', '')
    key = (int(row['ID']), name)
    per[key]['read' if m.endswith('read.sum') else 'write' if m.endswith('write.sum') else 'us'] += v
  fam = collections.defaultdict(lambda: dict(launches=0, read=0.0, write=0.0, us=0.0))
  for (_, name), d in per.items():
    f = 'tcgen05 contraction (gemm_tc_kernel / gemm_tc2_kernel)' if name.startswith('gemm_tc') else name
    fam[f]['launches'] += 1
    for k in ('read', 'write', 'us'):
      fam[f][k] += d[k]
  out = {}
  print('| kernel | launches | DRAM read MB | DRAM write MB | bytes / launch (MB) | ms |')
  print('|---|---:|---:|---:|---:|---:|')
  for f, d in sorted(fam.items(), key=lambda kv: -kv[1]['us']):
    per_launch = (d['read'] + d['write']) / d['launches']
    out[f] = dict(launches=d['launches'], dram_read_bytes=d['read'], dram_write_bytes=d['write'],
                  dram_bytes_per_launch=per_launch, ms=d['us'] / 1e3)
    print(f"| `{f[:64]}` | {d['launches']} | {d['read'] / 1e6:.1f} | {d['write'] / 1e6:.1f} | {per_launch / 1e6:.2f} | {d['us'] / 1e3:.3f} |")
  if len(sys.argv) > 2:
    json.dump(dict(source=sys.argv[1], note='one eagerly launched PC step (2 network evaluations) at batch 1024 under '
                   'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum (serialised launches)', kernels=out),
              open(sys.argv[2], 'w'), indent=1)


if __name__ == '__main__':
  main()
