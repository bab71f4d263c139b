#!/bin/bash
# one bench line per secondary workload (bench.py --workload), 1 GPU
mkdir -p gpurun_out; : > gpurun_out/bench_workloads.jsonl
for w in cifar10_ddpmpp_vp celebahq_256_ve ffhq_1024_ve celebahq_256_ddpmpp_subvp_ode; do
  timeout 400 python bench.py --workload $w --steps 8 --warmup 3 >> gpurun_out/bench_workloads.jsonl 2> gpurun_out/bench_workload_$w.err; echo "$w exit $?"
  tail -c 300 gpurun_out/bench_workload_$w.err
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_workloads.jsonl'):
  l=l.strip()
  if not l.startswith('{'): continue
  d=json.loads(l); print(d['config']['workload'][:70], '|', d['value'], d['unit'], '|', d['ms_per_step'], 'ms/step | batch', d['config']['batch_per_gpu'], '|', d['dtype'], '| finite', d['finite'], '| nfe', d['config'].get('nfe'))
PY
