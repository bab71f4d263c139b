#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/conv_isolated_$1.log; rm -f $L
for args in "--c1 256 --cout 256 --hw 16" "--c1 256 --cout 256 --hw 16 --residual" "--c1 512 --cout 256 --hw 16" "--c1 256 --cout 256 --hw 8" "--c1 256 --cout 256 --hw 4" "--c1 256 --cout 256 --hw 32" "--c1 256 --cout 256 --hw 16 --k 1"; do
  python tools/ncu_conv.py $args >> $L 2>&1
  B200_TC_2CTA=0 python tools/ncu_conv.py $args >> $L 2>&1
done
for args in "--c1 128 --cout 128 --hw 32" "--c1 128 --cout 128 --hw 32 --residual" "--c1 256 --cout 128 --hw 32" "--c1 384 --cout 128 --hw 32 --k 1"; do
  python tools/ncu_conv.py $args >> $L 2>&1
  B200_TC_SWAP=0 python tools/ncu_conv.py $args >> $L 2>&1
done
cat $L
