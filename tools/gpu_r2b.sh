#!/bin/bash
# focused GPU session: selected tests (-k EXPR), optional compute-sanitizer pass, short bench
#   tools/gpu_r2b.sh TAG "KEXPR" [sanitize-kexpr] [bench-steps]
TAG=${1:-x}; KEXPR=${2:-gn_on_load}; SAN=${3:-}; STEPS=${4:-10}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s -k "$KEXPR" > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest -k '$KEXPR' exit $?" >> $S
if [ -n "$SAN" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider -k "$SAN" > gpurun_out/memcheck_${TAG}.log 2>&1; echo "memcheck '$SAN' exit $?" >> $S
fi
if [ "$STEPS" != "0" ]; then
  timeout 900 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
fi
cat $S; grep -h "passed\|failed\|error" gpurun_out/pytest_${TAG}.log | tail -5; grep -h "gn-on-load\|batch-256\|rel-L2\|Error\|error" gpurun_out/pytest_${TAG}.log | head -40
[ -n "$SAN" ] && tail -15 gpurun_out/memcheck_${TAG}.log
if [ "$STEPS" != "0" ]; then
tail -c 800 gpurun_out/bench_$TAG.err
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
