#!/bin/bash
# GPU check of the package default `halo='pairs'`: the tests that run CTA-pair launches through the engine, and the bench's
# in-run parity at batch 1024.   tools/gpu_pairs.sh TAG
TAG=${1:-p1}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
timeout 110 python -m pytest tests/test_gpu_round2.py -q --tb=short -p no:cacheprovider -s -k "batch256 or plans_stay_valid or (groupnorm_on_load and 40)" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> $S
grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -2; grep -h "^E  \|^FAILED\|batch-256\|halo=" gpurun_out/pytest_${TAG}.log | head -12
timeout 80 python bench.py --steps 6 --warmup 3 --no-cpu --no-strong --no-variants > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
cat $S; tail -c 200 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; frac',r.get('frac'),'stf',r.get('step_tensor_fraction'), d['config'].get('conv3x3_mainloop'))
  print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k!='oracle'})
except Exception as e:
  print('bench parse failed', e)
PY
