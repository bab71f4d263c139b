"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares."""
import collections
import csv
import re
import sys


def load(path):
  lines = [l for l in open(path) if not l.startswith('==')]
  rows = []
  for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
      continue
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1e-3)
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('b200::', '').replace('<unnamed>::', '')
    rows.append((int(row['ID']), name, v, row['Grid Size']))
  return rows


def main():
  rows = load(sys.argv[1])
  tot, cnt = collections.defaultdict(float), collections.Counter()
  for _, n, us, _ in rows:
    tot[n] += us
    cnt[n] += 1
  T = sum(tot.values())
  print(f'# {sys.argv[1]}: {len(rows)} launches, {T / 1e3:.3f} ms total (cold-cache, serialised: compare shares)')
  print('| kernel | launches | total ms | share |')
  print('|---|---:|---:|---:|')
  for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f'| `{k[:70]}` | {cnt[k]} | {v / 1e3:.3f} | {100 * v / T:.1f}% |')
  if len(sys.argv) > 2 and sys.argv[2] == '--all':
    for i, n, us, g in rows:
      print(f'{i:5d} {n[:40]:40s} {us:10.1f} us {g}')


if __name__ == '__main__':
  main()
