#!/bin/bash
# One GPU session: parity tests (default + alternative kernel forms), smoke, bench, ncu launch list, DRAM-traffic pass,
# one ncu --set full capture of the contraction kernels inside a real step.
TAG=${1:-x}
mkdir -p gpurun_out; S=gpurun_out/summary_$TAG.txt; rm -f $S
for f in kernels engine tc; do
  timeout 700 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/pytest_${TAG}_$f.log 2>&1; echo "$f exit $?" >> $S
done
B200_TC_2CTA=0 B200_TC_SWAP=0 B200_LANES=1 B200_FUSE_SKIP=0 B200_FUSED_ATTN=0 B200_TC_SPLIT_SMALL=0 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "(tcgen05 or cifar10) and not f16" > gpurun_out/pytest_${TAG}_tc_plain.log 2>&1; echo "tc_plain exit $?" >> $S
B200_TC_EPILOGUE=staged timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "(tcgen05 or tf32_cifar10_matches) and not f16" > gpurun_out/pytest_${TAG}_tc_staged.log 2>&1; echo "tc_staged exit $?" >> $S
B200_H1_F16=0 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "cifar10 and f16" -s > gpurun_out/pytest_${TAG}_tc_h1fp32.log 2>&1; echo "tc_h1fp32 exit $?" >> $S
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?" >> $S
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-variants > gpurun_out/bench_${TAG}_again.json 2> gpurun_out/bench_${TAG}_again.err; echo "bench_again exit $?" >> $S
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> gpurun_out/bench_$TAG.err; echo "profile_ops exit $?" >> $S
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "nculist exit $?" >> $S
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/traffic_$TAG.csv python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_traffic_$TAG.log 2>&1; echo "ncutraffic exit $?" >> $S
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 4 -o gpurun_out/prof_gemm_tc_$TAG python tools/ncu_step.py --batch 1024 --precision f16 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncufull exit $?" >> $S
cat $S; grep -h "passed\|failed" gpurun_out/pytest_${TAG}_*.log; grep -h "rel-L2" gpurun_out/pytest_${TAG}_tc_h1fp32.log gpurun_out/smoke_$TAG.log
for f in gpurun_out/bench_${TAG}.json gpurun_out/bench_${TAG}_again.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step e2e',d['e2e']['value'],'peak',r['peak'],'tc_frac',r['frac'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks'], d.get('variants'), r.get('traffic'), r.get('alg_hbm_bytes_per_launch'))
"; done
