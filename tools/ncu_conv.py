"""One tcgen05 convolution launch in isolation (timed with CUDA events; wrap in ncu for counters).

  python tools/ncu_conv.py --c1 256 --cout 256 --hw 16 --batch 1024 [--residual]
  B200_TC_2CTA=1 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 2 -o gpurun_out/prof python tools/ncu_conv.py
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import gpu_util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--c1', type=int, default=256)
ap.add_argument('--c2', type=int, default=0)
ap.add_argument('--cout', type=int, default=256)
ap.add_argument('--hw', type=int, default=16)
ap.add_argument('--k', type=int, default=3)
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--residual', action='store_true')
ap.add_argument('--reps', type=int, default=5)
ap.add_argument('--f16', action='store_true', help='fp16 operands (kind::f16)')
a = ap.parse_args()
dev = torch.device('cuda:0')
torch.manual_seed(0)
rt = gpu_util.round_tf32
x1 = rt(torch.randn(a.batch, a.hw, a.hw, a.c1, device=dev))
x2 = rt(torch.randn(a.batch, a.hw, a.hw, a.c2, device=dev)) if a.c2 else None
w = rt(torch.randn(a.cout, a.c1 + a.c2, a.k, a.k, device=dev) / np.sqrt((a.c1 + a.c2) * a.k * a.k))
bias = torch.randn(a.cout, device=dev)
res = torch.randn(a.batch, a.hw, a.hw, a.cout, device=dev) if a.residual else None
wp = gpu_util.pack_conv_weight(w, f16=a.f16)
if a.f16:
  x1 = x1.half()
  x2 = x2.half() if x2 is not None else None
flops = 2.0 * a.batch * a.hw * a.hw * a.cout * (a.c1 + a.c2) * a.k * a.k
for impl_env in ('',):
  ts = []
  for i in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = gpu_util.conv_nhwc(x1, x2, wp, bias, a.cout, a.k, residual=res, scale=0.7071067690849304, impl=2 if a.f16 else 1)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
  best = min(ts[1:])
  print(f'conv{" f16" if a.f16 else ""} {a.c1}+{a.c2}->{a.cout} k{a.k} @{a.hw}x{a.hw} B={a.batch} res={a.residual} '
        f'2cta={os.environ.get("B200_TC_2CTA", "0")} epi={os.environ.get("B200_TC_EPILOGUE", "auto")}: '
        f'{best * 1e3:.1f} us  {flops / best / 1e9:.0f} TFLOP/s  (all: {[round(t * 1e3) for t in ts]})')
