#!/bin/bash
# attention phase stamps; im2col rewrite check
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
B200_ATTN_DBG=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "fused_attention_core_f16 and 150" -s >> $L 2>&1
B200_ATTN_DBG=1 timeout 300 python -m pytest tests/test_gpu_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "fused_attention_core_matches_torch and 200" -s >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | tail -5 >> $L; echo "tests exit $?" >> $L
timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_f16.md > /dev/null 2>> $L; echo "profile_ops exit $?" >> $L
grep -v "^$" $L | tail -16; grep "im2col\|attention" gpurun_out/ops_${TAG}_f16.md
