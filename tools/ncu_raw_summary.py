"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (`ncu --set full`) as markdown."""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__cycles_active.avg',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
name_i = hdr.index('Kernel Name')
print(f'# {sys.argv[1]} (ncu --set full --clock-control none)')
for r in rows[2:]:
  print(f'\n## `{r[name_i][:110]}`\n')
  print('| metric | value | unit |\n|---|---:|---|')
  for w in WANT:
    if w in hdr:
      i = hdr.index(w)
      print(f'| {w} | {r[i]} | {units[i]} |')
