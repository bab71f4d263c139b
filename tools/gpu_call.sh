#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + one full capture.  Usage: tools/gpu_call.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
for f in kernels engine tc; do
  timeout 700 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_${TAG}_$f.log 2>&1; echo "$f exit $?" >> $S
done
B200_TC_EPILOGUE=direct timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "tcgen05 or tf32_cifar10_matches" > gpurun_out/pytest_${TAG}_tc_direct.log 2>&1; echo "tc_direct exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "nculist exit $?" >> $S
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o gpurun_out/prof_gemm_tc_$TAG python tools/ncu_step.py --batch 1024 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncufull exit $?" >> $S
cat $S; cat gpurun_out/bench_$TAG.json
