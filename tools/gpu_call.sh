#!/bin/bash
# One GPU session: parity tests, smoke, bench (+ A/B variants), ncu launch list + one full capture.
TAG=${1:-x}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
for f in kernels engine tc; do
  timeout 700 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_${TAG}_$f.log 2>&1; echo "$f exit $?" >> $S
done
B200_TC_2CTA=0 B200_TC_SWAP=0 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "tcgen05 or tf32_cifar10" > gpurun_out/pytest_${TAG}_tc_plain.log 2>&1; echo "tc_plain exit $?" >> $S
B200_TC_EPILOGUE=staged timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "tcgen05 or tf32_cifar10_matches" > gpurun_out/pytest_${TAG}_tc_staged.log 2>&1; echo "tc_staged exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_again.json 2> gpurun_out/bench_${TAG}_again.err; echo "bench_again exit $?" >> $S
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "nculist exit $?" >> $S
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 3 -o gpurun_out/prof_gemm_tc_$TAG python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncufull exit $?" >> $S
cat $S; for f in gpurun_out/bench_${TAG}*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step tf32peak',r['peak'],'tc_frac',r['frac'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks'])
"; done
