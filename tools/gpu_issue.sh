#!/bin/bash
# GPU session: whole-warp (elect.sync) TMA producer / MMA issuer roles - correctness subset, per-op profile, short bench.
#   tools/gpu_issue.sh TAG
TAG=${1:-i1}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $S
timeout 200 python -m pytest tests/test_gpu_tc.py -q -x --tb=short -p no:cacheprovider -s -k "halo_mainloop or matches_cuda_core or attention_core" > gpurun_out/canary_${TAG}.log 2>&1; rc=$?
echo "canary exit $rc" >> $S; grep -h "passed\|failed" gpurun_out/canary_${TAG}.log | tail -2
if [ $rc -ne 0 ]; then cat $S; tail -40 gpurun_out/canary_${TAG}.log; exit 0; fi
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py -q --tb=short -p no:cacheprovider -s -k "not full_1000 and not vp_subvp and not ancestral and not halo_mainloop and not matches_cuda_core" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> $S
grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -3; grep -h "^E  \|^FAILED\|batch-256\|rel-L2" gpurun_out/pytest_${TAG}.log | head -20
timeout 150 python tools/profile_ops.py --batch 1024 --md gpurun_out/ops_${TAG}.md > /dev/null 2> gpurun_out/ops_${TAG}.err; echo "profile_ops exit $?" >> $S
head -40 gpurun_out/ops_${TAG}.md
timeout 400 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-strong --no-variants > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
cat $S; tail -c 300 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; e2e',d.get('e2e',{}).get('value'),'; frac',r.get('frac'),'step_tensor_fraction',r.get('step_tensor_fraction'))
  print('by kind', {k:v['ms'] for k,v in r.get('forward_ms_by_kind',{}).items()})
  print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k!='oracle'}); print('clocks', d.get('clocks'))
except Exception as e:
  print('bench parse failed', e)
PY
