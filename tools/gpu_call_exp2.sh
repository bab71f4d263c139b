#!/bin/bash
# A/B: shared-memory carve-out of the streaming kernels (lane overlap), lanes on/off
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider >> $L 2>&1; echo "tests exit $?" >> $L
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-variants > gpurun_out/bench_${TAG}_$name.json 2>> $L; }
run default A=1
run nocarve B200_CARVEOUT=0
run lanes1 B200_LANES=1
run default2 A=1
grep -v "^$" $L | tail -12
for f in gpurun_out/bench_${TAG}_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step peak',r['peak'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
