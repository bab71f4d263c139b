#!/bin/bash
# ncu --set full captures of single convolution launches (fp16 operands): swap form 128->128 @32 (+res), pair form 256->256 @16
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/ncu_$TAG.log; rm -f $L
for args in "--c1 128 --cout 128 --hw 32 --batch 512" "--c1 128 --cout 128 --hw 32 --batch 512 --residual" "--c1 256 --cout 256 --hw 16 --batch 512" "--c1 256 --cout 256 --hw 16 --batch 512 --residual"; do
  python tools/ncu_conv.py --f16 $args >> $L 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_swap_${TAG} python tools/ncu_conv.py --f16 --c1 128 --cout 128 --hw 32 --batch 512 --residual >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/prof_pair_${TAG} python tools/ncu_conv.py --f16 --c1 256 --cout 256 --hw 16 --batch 512 >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 40 -c 1 -o gpurun_out/prof_gn_${TAG} python tools/ncu_step.py --batch 1024 --precision f16 >> $L 2>&1
cat $L | grep -v "^==PROF==" | tail -20
ls -la gpurun_out/*.ncu-rep
