#!/bin/bash
# Final GPU session of a round: every -m gpu test (incl. the 1000-step sampler), smoke, the bench line, the ncu launch
# list + DRAM traffic of one PC step, and an `ncu --set full` capture of the large tcgen05 launches.
#   tools/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $S
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest -m gpu exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches_${TAG}.csv python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo "ncu launch list exit $?" >> $S
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 8 -f -o gpurun_out/ncu_gemm_${TAG} \
  python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_gemm_${TAG}.log 2>&1; echo "ncu --set full exit $?" >> $S
timeout 300 python tools/profile_ops.py --batch 1024 --md gpurun_out/ops_${TAG}.md > /dev/null 2> gpurun_out/ops_${TAG}.err; echo "profile_ops exit $?" >> $S
# memcheck over the halo-form convolutions (new shared-memory layout, halo TMA boxes) and the loss / op-gradient kernels
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_tc.py tests/test_gpu_f4.py -q -m gpu -x --tb=line \
  -p no:cacheprovider -k "(halo_mainloop and not B160 and not B20) or upfirdn2d_first or fused_leaky or ddpm_evaluation" > gpurun_out/memcheck_${TAG}.log 2>&1
echo "memcheck exit $?; $(grep -h 'passed\|failed' gpurun_out/memcheck_${TAG}.log | tail -1); $(grep -h 'ERROR SUMMARY' gpurun_out/memcheck_${TAG}.log | tail -1)" >> $S
cat $S; grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -3; grep -h "smoke ok" gpurun_out/smoke_$TAG.log
python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d.get('roofline',{})
  print(d.get('value'),'img/s',d.get('ms_per_step'),'ms/step; e2e',d.get('e2e',{}).get('value'),'; frac',r.get('frac'),'step_tensor_fraction',r.get('step_tensor_fraction'))
  print('parity', {k:v for k,v in (d.get('parity') or {}).items() if k!='oracle'}); print('clocks', d.get('clocks')); print('variants', d.get('variants'))
  print('cpu', d.get('cpu_baseline'))
except Exception as e:
  print('bench parse failed', e)
PY
