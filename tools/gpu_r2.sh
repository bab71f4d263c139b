#!/bin/bash
# One GPU session of round 2: parity tests, smoke, bench (with in-run parity / strong probe / reference CPU arm).
#   tools/gpu_r2.sh TAG [quick]
TAG=${1:-x}; MODE=${2:-full}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $S
if [ "$MODE" = "quick" ]; then
  timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x -k "not full_1000" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest -m gpu (quick) exit $?" >> $S
else
  timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest -m gpu exit $?" >> $S
fi
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
cat $S; grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -5; grep -h "rel-L2" gpurun_out/pytest_${TAG}.log gpurun_out/smoke_$TAG.log | tail -40
tail -c 1500 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; e2e',d.get('e2e',{}).get('value'),'; frac',r.get('frac'),'step_tensor_fraction',r.get('step_tensor_fraction'))
  print('by kind', {k:v['ms'] for k,v in r.get('forward_ms_by_kind',{}).items()})
  print('parity', d.get('parity')); print('clocks', d.get('clocks')); print('variants', d.get('variants'))
  print('cpu', d.get('cpu_baseline')); s=d.get('strong_scaling') or {}; print('strong', {k:v for k,v in s.items() if k!='underfilled_contractions'})
  print('underfilled', (s.get('underfilled_contractions') or {}).get('count'), (s.get('underfilled_contractions') or {}).get('ms'))
except Exception as e:
  print('bench parse failed', e)
PY
