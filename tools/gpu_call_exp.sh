#!/bin/bash
# experiments: swap for 1x1 convs, truncating skip conv accuracy, 2-GPU bench
mkdir -p gpurun_out
L=gpurun_out/exp_$1.log; rm -f $L
for args in "--c1 256 --cout 256 --hw 16 --k 1" "--c1 256 --cout 256 --hw 16 --k 1 --residual" "--c1 512 --cout 256 --hw 16 --k 1" "--c1 256 --cout 256 --hw 32 --k 1"; do
  python tools/ncu_conv.py $args >> $L 2>&1
  B200_TC_SWAP=2 python tools/ncu_conv.py $args >> $L 2>&1
done
B200_SKIP_TRUNC=1 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -s --tb=short -p no:cacheprovider -k "tf32_cifar10" >> $L 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu -s --tb=short -p no:cacheprovider -k "tf32_cifar10" >> $L 2>&1
B200_SKIP_TRUNC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$1_skiptrunc.json 2>> $L
B200_TC_SWAP=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$1_swap2.json 2>> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$1_default.json 2>> $L
cat $L | grep -v "^$" | tail -40
for f in gpurun_out/bench_$1_*.json; do echo $f; python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); r=d['roofline']
print(d['value'],'img/s',d['ms_per_step'],'ms/step tf32peak',r['peak'],'tc_ms',r['forward_ms_by_kind']['tcgen05_contraction']['ms'],'gn_ms',r['forward_ms_by_kind']['groupnorm']['ms'], d['clocks']['sm_mhz'])
"; done
