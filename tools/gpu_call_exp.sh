#!/bin/bash
# GroupNorm-apply work per thread under the single-lane default
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
for w in 8 16 32; do
  B200_GN_WORK=$w timeout 200 python tools/profile_ops.py --batch 1024 --precision f16 --reps 3 --md gpurun_out/ops_${TAG}_gnwork$w.md > /dev/null 2>> $L
  python - <<PY >> $L
t=0.0
for l in open('gpurun_out/ops_${TAG}_gnwork$w.md'):
    if l.startswith('| \`gn_apply'):
        t+=float(l.split('|')[3])
print('GN_WORK=$w: gn_apply total %.3f ms' % t)
PY
done
cat $L | grep GN_WORK
