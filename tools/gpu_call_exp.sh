#!/bin/bash
# fp16 operand mode: gn_apply prefetch, 16-byte fp16 epilogue stores, bench variants
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider >> $L 2>&1; echo "tests exit $?" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke exit $?" >> $L
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2>> $L; echo "bench exit $?" >> $L
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "nculist exit $?" >> $L
grep -v "^$" $L | tail -30
python -c "
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step e2e',d['e2e']['value'],'peak',r['peak'],'frac',r['frac'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks'], d.get('variants'), d.get('cpu_baseline'))
"
