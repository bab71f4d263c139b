#!/bin/bash
# experiments: fp16 operand mode
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "f16" -s >> $L 2>&1; echo "f16 tests exit $?" >> $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -k "not f16" >> $L 2>&1; echo "other tests exit $?" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke exit $?" >> $L
run() { name=$1; shift; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/bench_${TAG}_$name.json 2>> $L; }
run f16 --precision f16
run tf32 --precision tf32
run f16b --precision f16
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
grep -v "^$" $L | tail -60
for f in gpurun_out/bench_${TAG}_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step peak',r['peak'],'frac',r['frac'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
