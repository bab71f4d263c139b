#!/bin/bash
# Final check of HEAD: default parity tests, smoke, bench, and one ncu --set full capture each of a CTA-pair and a
# swapped-operand convolution inside a real step.
TAG=${1:-x}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/pytest_${TAG}_all.log 2>&1; echo "pytest -m gpu exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 20 -c 2 -o gpurun_out/prof_pair_$TAG python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_pair_$TAG.log 2>&1; echo "ncu pair exit $?" >> $S
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 2 -o gpurun_out/prof_swap_$TAG python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_swap_$TAG.log 2>&1; echo "ncu swap exit $?" >> $S
cat $S; tail -3 gpurun_out/pytest_${TAG}_all.log; cat gpurun_out/smoke_$TAG.log | tail -2
python -c "
import json
d=json.loads(open('gpurun_out/bench_$TAG.json').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step e2e',d['e2e']['value'],'peak',r['peak'],'tc_frac',r['frac'],d['clocks'], d.get('variants'), r.get('traffic'), r.get('alg_hbm_bytes_per_launch'), d.get('cpu_baseline'))
"
