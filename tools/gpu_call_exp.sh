#!/bin/bash
# experiments: fused attention core
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu -x --tb=short -p no:cacheprovider -k "fused_attention" >> $L 2>&1; echo "attn tests exit $?" >> $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -q -m gpu --tb=short -p no:cacheprovider -s >> $L 2>&1; echo "tests exit $?" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "smoke exit $?" >> $L
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_$name.json 2>> $L; }
run default A=1
run noattn B200_FUSED_ATTN=0
run default2 A=1
timeout 300 python tools/profile_ops.py --batch 1024 --md gpurun_out/ops_${TAG}.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
grep -v "^$" $L | tail -40
for f in gpurun_out/bench_${TAG}_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step tf32peak',r['peak'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
