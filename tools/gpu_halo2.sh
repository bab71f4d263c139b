#!/bin/bash
# GPU session: halo modes A/B per op on one box (profile_ops --halo-mode = raw no_halo: 0 swapped kernel only, 1 off, 2 pairs too,
# +4 with the L2 prefetch of the next tile - at the commit of profiles/r02_h2_* the prefetch bit was inverted), f4 tests, short bench.
#   tools/gpu_halo2.sh TAG
TAG=${1:-h2}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $S
timeout 500 python -m pytest tests/test_gpu_f4.py tests/test_gpu_tc.py tests/test_gpu_round2.py -q --tb=short -p no:cacheprovider -s \
  -k "f4 or halo or batch256 or upfirdn2d or fused_leaky or evaluation" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> $S
grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -3; grep -h "^E  \|^FAILED\|halo=\|batch-256" gpurun_out/pytest_${TAG}.log | head -30
for m in 0 4 2 6 1; do
  timeout 150 python tools/profile_ops.py --batch 1024 --halo-mode $m --md gpurun_out/ops_${TAG}_mode$m.md > /dev/null 2> gpurun_out/ops_${TAG}_mode$m.err; echo "profile_ops mode $m exit $?" >> $S
  head -1 gpurun_out/ops_${TAG}_mode$m.md; grep -h "conv3x3" gpurun_out/ops_${TAG}_mode$m.md | head -12
done
timeout 400 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-strong > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
cat $S; tail -c 300 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; e2e',d.get('e2e',{}).get('value'),'; frac',r.get('frac'),'step_tensor_fraction',r.get('step_tensor_fraction'))
  print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k!='oracle'}); print('clocks', d.get('clocks')); print('variants', d.get('variants'))
except Exception as e:
  print('bench parse failed', e)
PY
