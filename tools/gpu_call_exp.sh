#!/bin/bash
# experiments: GroupNorm-apply work per thread; ncu capture of the fused attention kernel
mkdir -p gpurun_out
TAG=$1; L=gpurun_out/exp_$TAG.log; rm -f $L
for w in 16 32 64 128; do
  B200_GN_WORK=$w timeout 300 python tools/profile_ops.py --batch 1024 --precision f16 --md gpurun_out/ops_${TAG}_gnwork$w.md > /dev/null 2>> $L
  echo "GN_WORK=$w: $(python - <<PY
import re
t=0.0
for l in open('gpurun_out/ops_${TAG}_gnwork$w.md'):
    if l.startswith('| \`gn_apply'):
        t+=float(l.split('|')[3])
print(f'gn_apply total {t:.3f} ms')
PY
)" >> $L
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc -s 2 -c 1 -o gpurun_out/prof_attn_$TAG python tools/ncu_step.py --batch 1024 --precision f16 >> $L 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 10 -c 3 -o gpurun_out/prof_gn_$TAG python tools/ncu_step.py --batch 1024 --precision f16 >> $L 2>&1
grep -v "^==" $L | grep -v "^$" | tail -12
