#!/bin/bash
# GPU session for the halo form of the 3x3 mainloop: canary, conv / network A-B tests, per-op profiles with and without it,
# one ncu --set full capture of the new launches, short bench with variants.
#   tools/gpu_halo.sh TAG
TAG=${1:-h1}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $S
timeout 240 python -m pytest tests/test_gpu_tc.py -q -x --tb=short -p no:cacheprovider -s -k "halo_mainloop" > gpurun_out/canary_${TAG}.log 2>&1; rc=$?
echo "canary exit $rc" >> $S
grep -h "halo vs\|passed\|failed\|Error\|error" gpurun_out/canary_${TAG}.log | head -40
if [ $rc -ne 0 ]; then cat $S; tail -40 gpurun_out/canary_${TAG}.log; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_round2.py -q --tb=short -p no:cacheprovider -s -k "not full_1000 and not vp_subvp and not ancestral" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> $S
grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -3; grep -h "halo vs\|batch-256\|rel-L2" gpurun_out/pytest_${TAG}.log | head -30
timeout 200 python tools/profile_ops.py --batch 1024 --md gpurun_out/ops_${TAG}_halo.md > /dev/null 2> gpurun_out/ops_${TAG}_halo.err; echo "profile_ops halo exit $?" >> $S
timeout 200 python tools/profile_ops.py --batch 1024 --no-halo --md gpurun_out/ops_${TAG}_nine.md > /dev/null 2> gpurun_out/ops_${TAG}_nine.err; echo "profile_ops nine exit $?" >> $S
head -30 gpurun_out/ops_${TAG}_halo.md; head -12 gpurun_out/ops_${TAG}_nine.md
timeout 400 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-strong > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 6 -f -o gpurun_out/ncu_${TAG} \
  python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu exit $?" >> $S
cat $S; tail -c 400 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; e2e',d.get('e2e',{}).get('value'),'; frac',r.get('frac'),'step_tensor_fraction',r.get('step_tensor_fraction'))
  print('by kind', {k:v['ms'] for k,v in r.get('forward_ms_by_kind',{}).items()})
  print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k!='oracle'}); print('clocks', d.get('clocks')); print('variants', d.get('variants'))
except Exception as e:
  print('bench parse failed', e)
PY
