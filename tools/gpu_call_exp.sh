#!/bin/bash
# explicit shared-space epilogue accesses: tests + bench + profile
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -k "not full_1000" 2>&1 | grep -v "^$" | tail -5 >> $L; echo "tests exit $?" >> $L
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants > gpurun_out/bench_${TAG}_$name.json 2>> $L; }
run default A=1
run default2 A=1
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
grep -v "^==" $L | grep -v "^$" | tail -8; grep "nchw-out\|attention\|+res" gpurun_out/ops_${TAG}_f16.md | head -6
for f in gpurun_out/bench_${TAG}_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step peak',r['peak'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
