#!/bin/bash
# tensor-core head conv: tests + bench; ncu capture of the attention kernel (after the residual hoist)
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -12 >> $L; echo "tests exit $?" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke exit $?" >> $L
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants > gpurun_out/bench_${TAG}_$name.json 2>> $L; }
run default A=1
run default2 A=1
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 1 -o gpurun_out/prof_attn_$TAG python tools/ncu_step.py --batch 1024 --precision f16 >> $L 2>&1
grep -v "^==" $L | grep -v "^$" | tail -16; grep "nchw-out\|attention" gpurun_out/ops_${TAG}_f16.md
for f in gpurun_out/bench_${TAG}_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step peak',r['peak'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
