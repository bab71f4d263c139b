#!/usr/bin/env python
"""Benchmark of the predictor–corrector sampling hot path (BASELINE.json metric).

Metric: PC-sampler images/sec, NCSN++ cont. CIFAR-10, VE-SDE, 1000 steps
(``configs/ve/cifar10_ncsnpp_continuous.py`` with ``model.init_scale = 1``, random-init
weights, Gaussian-noise inputs).  A bench *step* is one PC iteration over one batch
(Langevin corrector + reverse-diffusion predictor = 2 network evaluations + both state
updates with in-kernel noise); a full sample is 1000 such steps, so

    images/s = n_gpus * batch / (1000 * seconds_per_step).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1     # CPU arm (oracle port)

One JSON line on stdout (rank 0).  `value` times the CUDA-graph replay loop with the state
resident in HBM; `e2e` goes through the public sampler plan with pinned host buffers
(H2D of the state and D2H of the result inside every timed step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ALG_FLOP_PER_IMG_STEP = 43.58e9      # 2 evals x 21.788 GFLOP (2*MAC), measured on the reference
ALG_BYTES_PER_IMG_STEP = 153.9e6     # fp32 contraction operands only, everything else fused
N_SAMPLER_STEPS = 1000


def load_peaks():
  p = os.path.join(REPO, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    with open(p) as fh:
      d = json.load(fh)
    return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'], bf16_tflops_sustained=d['bf16_tflops_sustained'],
                source='measured (MEASURED_PEAKS.json)')
  return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
  """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

  def __init__(self, index):
    self.index, self.rows, self.proc = index, [], None

  def start(self):
    q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    try:
      self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                    '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      threading.Thread(target=self._pump, daemon=True).start()
    except OSError:
      self.proc = None

  def _pump(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.proc is None:
      return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
    self.proc.terminate()
    sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
    reasons = set()
    for r in self.rows:
      if len(r) >= 7:
        for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
          if v.lower().startswith('active'):
            reasons.add(name)
    return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                reasons=sorted(reasons), samples=len(sm))


def headline_config():
  from score_sde_pytorch_b200 import configs
  cfg = configs.ve_cifar10_ncsnpp_continuous()
  cfg.model.init_scale = 1.0          # init_scale=0 zeroes every block's last conv
  return cfg


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's sampler on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_pc_steps(batch, steps, warmup, threads=None):
  """Time PC iterations of the oracle (plain PyTorch fp32 restatement of the reference path)."""
  from oracle import ncsnpp_oracle as NO
  from oracle import sampling_oracle as SO
  from score_sde_pytorch_b200.models.ncsnpp import NCSNpp
  # torch's own default (one thread per physical core) is the fastest setting for this workload: using every
  # hyper-thread of the 128-way host made the oracle ~70x slower (127 s vs 1.8 s per iteration).  torchrun exports
  # OMP_NUM_THREADS=1 to every rank, which would reduce the CPU arm to one core: undo that.
  if threads:
    torch.set_num_threads(threads)
  elif torch.get_num_threads() <= 1:
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
  cfg = headline_config()
  torch.manual_seed(0)
  sd = NCSNpp(cfg).state_dict()
  model = lambda x, l: NO.ncsnpp_forward(sd, cfg, x, l)
  sde = SO.VE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
  shape = (batch, 3, 32, 32)
  torch.manual_seed(1)
  x = sde.prior_sampling(shape)
  ts = torch.linspace(sde.T, 1e-5, sde.N)
  times = []
  with torch.no_grad():
    for i in range(warmup + steps):
      t0 = time.perf_counter()
      vec_t = torch.ones(batch) * ts[i]
      x, _ = SO.langevin_step(sde, model, x, vec_t, cfg.sampling.snr, 1)
      x, _ = SO.reverse_diffusion_step(sde, model, x, vec_t)
      if i >= warmup:
        times.append(time.perf_counter() - t0)
  t_step = float(np.mean(times))
  return dict(value=batch / (N_SAMPLER_STEPS * t_step), unit='images/s', cores=torch.get_num_threads(), kind='port',
              sample=f'{steps} PC iterations (2 network evals each) of the oracle port at batch {batch} after {warmup} warm-up, '
                     f'{t_step:.3f} s/iteration, extrapolated to {N_SAMPLER_STEPS} iterations',
              host_cpus=os.cpu_count(), ms_per_step=t_step * 1e3)


def run_reference_arm(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  batch = args.cpu_batch
  r = cpu_pc_steps(batch, args.steps, args.warmup)
  line = dict(impl='reference', metric='PC-sampler images/sec, NCSN++ CIFAR-10 1000-step VE', value=r['value'],
              unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=r['ms_per_step'],
              higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32', data='synthetic',
              config=dict(workload='NCSN++ cont. CIFAR-10 32x32 VE-SDE PC sampler (1000 steps); CPU arm: bounded sample at '
                                   f'batch {batch}', batch_per_gpu=batch, sampler_steps=N_SAMPLER_STEPS),
              cpu_baseline=dict(value=r['value'], unit='images/s', cores=r['cores'], kind='port', sample=r['sample']),
              e2e=dict(value=r['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
              gpu_launches=0)
  print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def load_traffic(precision):
  """DRAM bytes per launch of the contraction kernels from the committed ncu metrics pass of one PC step
  (profiles/traffic_<precision>.json, written by tools/summarize_traffic.py); None when no capture is committed."""
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', f'traffic_{precision}.json')
  try:
    k = json.load(open(path))['kernels']
    for name, d in k.items():
      if name.startswith('tcgen05'):
        return round(d['dram_bytes_per_launch'])
  except (OSError, KeyError, ValueError):
    pass
  return None


def measure_tf32_peak(dev, f16=False):
  """cuBLAS GEMM throughput on this device in the operand format the engine computes in: TF32 (kind::tf32
  work) or fp16 with fp32 accumulation (kind::f16 work) - the measured denominator of the tensor roofline."""
  n = 8192
  a = torch.randn(n, n, device=dev)
  b = torch.randn(n, n, device=dev)
  if f16:
    a, b = a.half(), b.half()
  old = torch.backends.cuda.matmul.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = True
  try:
    for _ in range(3):
      a @ b
    best = 0.0
    for _ in range(6):
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
      best = max(best, 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
  finally:
    torch.backends.cuda.matmul.allow_tf32 = old
  del a, b
  return best


def run_gpu_arm(args):
  import torch.distributed as dist
  from score_sde_pytorch_b200 import native, sampling, sde_lib, _lib
  from score_sde_pytorch_b200.models.ncsnpp import NCSNpp
  import ctypes

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if args.gpus != world and world > 1:
    raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}')
  dev = torch.device('cuda', local)
  torch.cuda.set_device(dev)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)

  cfg = headline_config()
  cfg.device = dev
  B = args.batch
  shape = (B, 3, 32, 32)
  torch.manual_seed(0)                       # same random-init weights on every rank ...
  model = NCSNpp(cfg, precision=args.precision).to(dev)
  if world > 1:                              # ... and broadcast once over NCCL/NVLink anyway (the production path)
    from score_sde_pytorch_b200 import distributed as bdist
    bdist.broadcast_parameters(model, src=0)
  sde = sde_lib.VESDE(cfg.model.sigma_min, cfg.model.sigma_max, cfg.model.num_scales)
  plan = native.match_pc_plan(sde=sde, model=model, predictor=sampling.ReverseDiffusionPredictor,
                              corrector=sampling.LangevinCorrector, shape=shape, snr=cfg.sampling.snr, n_steps=1,
                              probability_flow=False, continuous=True, eps=1e-5, device=dev)
  assert plan is not None
  torch.manual_seed(1 + rank)                # independent chains per rank
  torch.cuda.manual_seed(1 + rank)
  x_host = sde.prior_sampling(shape).pin_memory()
  out_host = torch.empty(shape).pin_memory()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # ---- device-resident timing: K graph replays ----
  x0 = x_host.to(dev)
  plan.run(x0, first_step=0, num_steps=args.warmup, clone=False)  # warm-up (captures the graph)
  clocks = ClockSampler(local)
  barrier()
  clocks.start()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  x, x_mean = plan.run(x0, first_step=args.warmup, num_steps=args.steps, clone=False)
  e1.record()
  barrier()
  clk = clocks.stop()
  ms = e0.elapsed_time(e1)
  finite = bool(torch.isfinite(x_mean).all().item())

  # ---- end to end through the public plan API with host buffers, copies inside the timed region ----
  e2e_steps = args.steps
  barrier()
  f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  f0.record()
  for i in range(e2e_steps):
    # public plan API on host buffers: H2D of the step's input state (pinned), one PC iteration, D2H of the result
    xd, xm = plan.run(x_host, first_step=args.warmup + i, num_steps=1, clone=False)
    out_host.copy_(xm, non_blocking=True)
  f1.record()
  barrier()
  ms_e2e = f0.elapsed_time(f1)

  t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  ms, ms_e2e = t.tolist()
  ms_per_step = ms / args.steps
  value = world * B / (N_SAMPLER_STEPS * ms_per_step * 1e-3)
  e2e_value = world * B / (N_SAMPLER_STEPS * (ms_e2e / e2e_steps) * 1e-3)

  if rank == 0:
    peaks = load_peaks()
    # ---- per-kind device time of one eager forward (events around every launch) ----
    eng = model.engine(B, dev)
    ms_k = (ctypes.c_float * 8)(); fl_k = (ctypes.c_double * 8)(); n_k = (ctypes.c_longlong * 8)()
    xin = x0.clone(); lab = torch.full((B,), 1.0, device=dev); out = torch.empty_like(xin)
    for _ in range(2):
      _lib.call('b200_ncsnpp_profile_forward', eng['h'], _lib.ptr(xin), _lib.ptr(lab), 1, _lib.ptr(out),
                _lib.stream_ptr(dev), ms_k, fl_k, n_k)
    # algorithmic HBM bytes of the contraction launches (operands once + outputs once), from the plan
    n_ops = int(_lib.load().b200_ncsnpp_num_ops(eng['h']))
    tc_alg_bytes = 0.0
    kind_c, bytes_c = ctypes.c_int(), ctypes.c_double()
    for i in range(n_ops):
      _lib.call('b200_ncsnpp_op_info', eng['h'], i, None, 0, ctypes.byref(kind_c), None)
      if kind_c.value == 0:
        _lib.call('b200_ncsnpp_op_bytes', eng['h'], i, ctypes.byref(bytes_c))
        tc_alg_bytes += bytes_c.value
    kinds = ['tcgen05_contraction', 'cuda_core_contraction', 'groupnorm', 'fir', 'softmax', 'time_embedding', 'misc']
    by_kind = {k: dict(ms=round(ms_k[i], 4), gflop=round(fl_k[i] / 1e9, 2), launches=int(n_k[i])) for i, k in enumerate(kinds)}
    fwd_ms = sum(ms_k[i] for i in range(7))
    tc_ms, tc_flops, tc_n = ms_k[0], fl_k[0], max(int(n_k[0]), 1)
    f16 = args.precision == 'f16'
    tf32_peak = measure_tf32_peak(dev, f16)
    achieved = (tc_flops / tc_n) / ((tc_ms / tc_n) * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    t_hbm_ms = ALG_BYTES_PER_IMG_STEP * B / (peaks['hbm_gbs'] * 1e9) * 1e3
    roofline = dict(bound='tensor', kernel=f"gemm_tc_kernel / gemm_tc2_kernel (tcgen05 kind::{'f16' if f16 else 'tf32'} implicit GEMM)",
                    achieved=round(achieved, 2), peak=round(tf32_peak, 2), unit='TFLOP/s',
                    frac=round(achieved / tf32_peak, 4) if tf32_peak else None,
                    peak_source=(f"cuBLAS {'fp16' if f16 else 'TF32'} 8192^3 GEMM measured in this run (MEASURED_PEAKS.json holds bf16 only: "
                                 f"{peaks['bf16_tflops_sustained']} TF/s sustained, {peaks['source']})"),
                    frac_of_measured_bf16_sustained=round(achieved / peaks['bf16_tflops_sustained'], 4),
                    alg_flop_per_launch=tc_flops / tc_n, avg_launch_ms=tc_ms / tc_n, launches_per_forward=tc_n,
                    kernel_share_of_forward=round(tc_ms / fwd_ms, 4) if fwd_ms else None,
                    traffic=load_traffic(args.precision), alg_hbm_bytes_per_launch=round(tc_alg_bytes / tc_n),
                    hbm_fraction_of_step=round(t_hbm_ms / ms_per_step, 4),
                    step_tensor_fraction=round((ALG_FLOP_PER_IMG_STEP * B / (tf32_peak * 1e12) * 1e3) / ms_per_step, 4) if tf32_peak else None,
                    forward_ms_by_kind=by_kind)
    # ---- the other tensor-core operand format, same model / batch / steps, device-resident timing only ----
    variants = None
    if world == 1 and not args.no_variants and args.precision in ('f16', 'tf32'):
      other = 'tf32' if args.precision == 'f16' else 'f16'
      torch.manual_seed(0)
      model2 = NCSNpp(cfg, precision=other).to(dev)
      plan2 = native.match_pc_plan(sde=sde, model=model2, predictor=sampling.ReverseDiffusionPredictor,
                                   corrector=sampling.LangevinCorrector, shape=shape, snr=cfg.sampling.snr, n_steps=1,
                                   probability_flow=False, continuous=True, eps=1e-5, device=dev)
      x2 = x_host.to(dev)
      plan2.run(x2, first_step=0, num_steps=args.warmup, clone=False)
      torch.cuda.synchronize()
      g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      g0.record()
      plan2.run(x2, first_step=args.warmup, num_steps=args.steps, clone=False)
      g1.record()
      torch.cuda.synchronize()
      ms2 = g0.elapsed_time(g1) / args.steps
      variants = {other: dict(value=round(B / (N_SAMPLER_STEPS * ms2 * 1e-3), 4), unit='images/s', ms_per_step=round(ms2, 4))}
      del plan2, model2
    cpu = cpu_pc_steps(args.cpu_batch, 2, 1) if world == 1 and not args.no_cpu else None
    line = dict(metric='PC-sampler images/sec, NCSN++ CIFAR-10 1000-step VE', value=round(value, 4), unit='images/s',
                n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=round(ms_per_step, 4),
                higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype={'tf32': 'tf32', 'f16': 'f16 operands (11-bit significand, as tf32), f32 accumulate and activations', 'fp32': 'f32'}[args.precision],
                data='synthetic',
                config=dict(workload='NCSN++ cont. CIFAR-10 32x32 VE-SDE PC sampler (1000 steps), batch 1024 per GPU'
                            if B == 1024 else f'NCSN++ cont. CIFAR-10 32x32 VE-SDE PC sampler (1000 steps), batch {B} per GPU',
                            batch_per_gpu=B, global_batch=B * world, sampler_steps=N_SAMPLER_STEPS,
                            step='one PC iteration = Langevin corrector + reverse-diffusion predictor (2 score evaluations)',
                            parallelism=f'{world} independent chains shards (weights broadcast once, no in-loop collective)',
                            l2='per-step working set (activations) exceeds the 126 MB L2 by >100x; no explicit flush',
                            weights='random init, init_scale=1, torch.manual_seed(0)', precision=args.precision),
                clocks=clk,
                e2e=dict(value=round(e2e_value, 4), unit='images/s', h2d_bytes_per_step=int(np.prod(shape)) * 4,
                         d2h_bytes_per_step=int(np.prod(shape)) * 4, ms_per_step=round(ms_e2e / e2e_steps, 4)),
                gpu_launches=int(plan.launches_per_step()) * args.steps,
                roofline=roofline, finite=finite)
    if variants is not None:
      line['variants'] = variants
    if cpu is not None:
      line['cpu_baseline'] = dict(value=cpu['value'], unit='images/s', cores=cpu['cores'], kind='port', sample=cpu['sample'])
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--batch', type=int, default=1024, help='images per GPU (BASELINE.json configs[1]: 1024)')
  ap.add_argument('--cpu-batch', type=int, default=8, help='batch of the bounded CPU sample')
  ap.add_argument('--precision', default='f16', choices=['tf32', 'f16', 'fp32'],
                  help="tensor-core operand format: 'f16' (default) and 'tf32' both carry 11-bit significands with fp32 "
                       "accumulation and meet the same 1e-3 parity bound (tests/test_gpu_tc.py); 'fp32' = CUDA cores")
  ap.add_argument('--no-variants', action='store_true', help='skip timing the other operand format')
  ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
  args = ap.parse_args()
  if args.warmup < 3 and args.impl == 'ours':
    args.warmup = 3
  if args.impl == 'reference':
    run_reference_arm(args)
  else:
    run_gpu_arm(args)


if __name__ == '__main__':
  main()
