"""One eagerly-launched PC iteration between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/ncu_step.py --batch 1024
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc \
      -s 20 -c 3 -o gpurun_out/prof_gemm_tc python tools/ncu_step.py --batch 1024
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import headline_config                                    # noqa: E402
from score_sde_pytorch_b200 import native, sampling, sde_lib         # noqa: E402
from score_sde_pytorch_b200.models.ncsnpp import NCSNpp              # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=1024)
ap.add_argument('--precision', default='f16')
args = ap.parse_args()
dev = torch.device('cuda:0')
cfg = headline_config()
torch.manual_seed(0)
model = NCSNpp(cfg, precision=args.precision).to(dev)
sde = sde_lib.VESDE(0.01, 50, 1000)
shape = (args.batch, 3, 32, 32)
plan = native.match_pc_plan(sde=sde, model=model, predictor=sampling.ReverseDiffusionPredictor,
                            corrector=sampling.LangevinCorrector, shape=shape, snr=0.16, n_steps=1,
                            probability_flow=False, continuous=True, eps=1e-5, device=dev)
plan.use_graph = False
torch.manual_seed(1); torch.cuda.manual_seed(1)
x0 = sde.prior_sampling(shape).to(dev)
plan.run(x0, first_step=0, num_steps=2, clone=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
plan.run(x0, first_step=2, num_steps=1, clone=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('launches per PC step:', plan.launches_per_step())
