#!/bin/bash
# full 1000-step sampler parity, tensor-core head vs CUDA-core head
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -s -k "full_1000 or cifar10_matches" 2>&1 | grep -v "^$" | tail -8 >> $L; echo "tc-head exit $?" >> $L
B200_TC_HEAD=0 timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -s -k "full_1000 or cifar10_matches" 2>&1 | grep -v "^$" | tail -8 >> $L; echo "cuda-core-head exit $?" >> $L
cat $L
