#!/bin/bash
# re-run the alternative-form parity tests after the head fix
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
B200_TC_2CTA=0 B200_TC_SWAP=0 B200_LANES=1 B200_FUSE_SKIP=0 B200_FUSED_ATTN=0 B200_TC_SPLIT_SMALL=0 timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "(tcgen05 or cifar10) and not f16" 2>&1 | grep -v "^$" | tail -8 >> $L; echo "tc_plain exit $?" >> $L
B200_TC_HEAD=0 B200_GN_STREAM=0 timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -q -m gpu --tb=short -p no:cacheprovider -k "not full_1000" 2>&1 | grep -v "^$" | tail -4 >> $L; echo "alt(head cuda-core, generic gn) exit $?" >> $L
cat $L
