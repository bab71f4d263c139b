#!/bin/bash
# GPU session: selected tests, per-op profile, ncu --set full of one kernel family, short bench
#   tools/gpu_r2c.sh TAG "KEXPR" NCU_REGEX [bench-steps]
TAG=${1:-x}; KEXPR=${2:-gn_on_load}; NCUK=${3:-}; STEPS=${4:-10}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
# canary: one small forward through the newest kernels under a short timeout, so a deadlock costs 3 minutes, not the session
if [ -n "$CANARY" ]; then
  timeout 180 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider -s -k "$CANARY" > gpurun_out/canary_${TAG}.log 2>&1; rc=$?
  echo "canary '$CANARY' exit $rc" >> $S
  if [ $rc -ne 0 ]; then cat $S; tail -30 gpurun_out/canary_${TAG}.log; exit 0; fi
fi
timeout ${PYTEST_TO:-700} python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s -k "$KEXPR" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest -k '$KEXPR' exit $?" >> $S
timeout 240 python tools/profile_ops.py --batch 1024 --md gpurun_out/ops_${TAG}.md > /dev/null 2> gpurun_out/ops_${TAG}.err; echo "profile_ops exit $?" >> $S
if [ -n "$NCUK" ]; then
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$NCUK -s 2 -c 3 -f -o gpurun_out/ncu_${TAG} python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_${TAG}.log 2>&1; echo "ncu exit $?" >> $S
fi
if [ "$STEPS" != "0" ]; then
  timeout 500 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
fi
cat $S; grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -5; grep -h "gn-on-load\|batch-256\|rel-L2\|Error\|error\|ode \[" gpurun_out/pytest_${TAG}.log | head -40
head -40 gpurun_out/ops_${TAG}.md
if [ "$STEPS" != "0" ]; then
tail -c 600 gpurun_out/bench_$TAG.err
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
fi
