"""Warp-state samples per SASS instruction of one kernel in an `.ncu-rep` (`ncu --set full --import-source on`):
the most-sampled instructions of the whole kernel, and every instruction with samples in the MMA-issue region (from
60 instructions before the first UTCHMMA to 10 after the last UTCBAR), with its stall reasons.

  python tools/ncu_source_samples.py gpurun_out/ncu_h1.ncu-rep [kernel-regex [launch-index]] > profiles/....md

This is how DESIGN.md section 4.13 located the bound of the convolutions: the issuer warp's samples were not on a
barrier wait and not on the MMA instructions, but spread over the ALU / uniform-datapath instructions the compiler
had wrapped around every UTCHMMA."""
import csv
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else 'gemm_tc'
idx = sys.argv[3] if len(sys.argv) > 3 else '1'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f'::regex:{kre}:{idx}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
name = rows[0][1] if rows and len(rows[0]) > 1 else '?'
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
total = sum(int(r[isamp]) for r in data)
print(f'# {rep}: `{name[:100]}` (launch {idx}) - {total} warp-state samples over {len(data)} SASS instructions\n')
print('## most-sampled instructions\n\n| # | samples | executed | instruction |\n|---:|---:|---:|---|')
for i in sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:16]:
  print(f'| {i} | {data[i][isamp]} | {data[i][iex]} | `{data[i][isrc].strip()[:90]}` |')
mma = [i for i, r in enumerate(data) if 'UTCHMMA' in r[isrc] or 'UTCBAR' in r[isrc]]
if mma:
  lo, hi = max(0, min(mma) - 60), min(len(data), max(mma) + 10)
  reg = sum(int(data[i][isamp]) for i in range(lo, hi))
  nins = sum(1 for i in range(lo, hi) if int(data[i][iex]) > 0)
  nm = sum(1 for i in range(lo, hi) if 'UTCHMMA' in data[i][isrc])
  loops = sum(1 for i in range(lo, hi) if 'BRA.U.ANY' in data[i][isrc])
  print(f'\n## MMA-issue region (instructions {lo}..{hi}): {reg} samples; {nins} executed instructions, {nm} UTCHMMA sites, '
        f'{loops} BRA.U.ANY (lane-serialising) loops\n\n| # | samples | executed | instruction | stall reasons |\n|---:|---:|---:|---|---|')
  for i in range(lo, hi):
    s = int(data[i][isamp])
    if s >= 6:
      st = ', '.join(f'{hdr[c][6:]} {data[i][c]}' for c in stall if data[i][c] not in ('0', ''))
      print(f'| {i} | {s} | {data[i][iex]} | `{data[i][isrc].strip()[:70]}` | {st} |')
