"""Round-2 golden fixtures from the REAL reference (same recipe as tools/make_golden.py, which is left untouched so the
round-1 fixtures stay byte-identical): sampler cases the oracle did not pin yet.

  pc_extra_tiny.npz
    ve_none_langevin   VE, predictor=None (NonePredictor) + LangevinCorrector, denoise=True  -> pins x_mean := x
                       after a 'none' predictor (sampling.py:241-250, :409)
    subvp_em_none      sub-VP (sde_lib.py:167-204), EulerMaruyama + NoneCorrector
    subvp_rd_none      sub-VP, ReverseDiffusion (the base-class Euler discretisation, sde_lib.py:52-69) + NoneCorrector

    python tools/make_golden_r2.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as MG   # noqa: E402


def main():
  torch.set_num_threads(8)
  sde_lib, sampling, ncsnpp, mutils, _ = MG.import_reference()
  cfgs = MG.golden_configs()
  out = {}

  def ref_model_for(name):
    cfg, B = cfgs[name]
    cfg.device = torch.device('cpu')
    sd = MG.our_weights(cfg)
    torch.manual_seed(0)
    m = mutils.get_model('ncsnpp')(cfg).eval()
    m.load_state_dict(sd, strict=True)
    return cfg, B, m

  cfg, B, m = ref_model_for('tiny')
  shape = (B, cfg.data.num_channels, cfg.data.image_size, cfg.data.image_size)
  sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=12)
  fn = sampling.get_pc_sampler(sde, shape, None, sampling.LangevinCorrector, lambda v: v, snr=0.16, n_steps=1,
                               probability_flow=False, continuous=True, denoise=True, eps=1e-5, device='cpu')
  torch.manual_seed(31)
  s, nfe = fn(m)
  out['ve_none_langevin'] = s.numpy(); out['ve_none_langevin_nfe'] = nfe

  cfg, B, m = ref_model_for('tiny_vp')
  shape = (B, cfg.data.num_channels, cfg.data.image_size, cfg.data.image_size)
  sde = sde_lib.subVPSDE(beta_min=0.1, beta_max=20., N=20)
  for tag, pred, seed in (('subvp_em_none', sampling.EulerMaruyamaPredictor, 32),
                          ('subvp_rd_none', sampling.ReverseDiffusionPredictor, 33)):
    fn = sampling.get_pc_sampler(sde, shape, pred, sampling.NoneCorrector, lambda v: v, snr=0.16, n_steps=1,
                                 probability_flow=False, continuous=True, denoise=True, eps=1e-3, device='cpu')
    torch.manual_seed(seed)
    s, nfe = fn(m)
    out[tag] = s.numpy(); out[tag + '_nfe'] = nfe
  np.savez_compressed(os.path.join(MG.OUT, 'pc_extra_tiny.npz'), **out)

  # ---- pc_ancestral_ald_tiny.npz: AncestralSamplingPredictor (sampling.py:204-239) and AnnealedLangevinDynamics (:286-319)
  out2 = {}
  cfg, B, m = ref_model_for('tiny')
  shape = (B, cfg.data.num_channels, cfg.data.image_size, cfg.data.image_size)
  sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=12)
  for tag, pred, corr, seed in (('ve_ancestral_langevin', sampling.AncestralSamplingPredictor, sampling.LangevinCorrector, 34),
                                ('ve_rd_ald', sampling.ReverseDiffusionPredictor, sampling.AnnealedLangevinDynamics, 35)):
    fn = sampling.get_pc_sampler(sde, shape, pred, corr, lambda v: v, snr=0.16, n_steps=1, probability_flow=False,
                                 continuous=True, denoise=True, eps=1e-5, device='cpu')
    torch.manual_seed(seed)
    s, nfe = fn(m)
    out2[tag] = s.numpy()
  cfg, B, m = ref_model_for('tiny_vp')
  shape = (B, cfg.data.num_channels, cfg.data.image_size, cfg.data.image_size)
  # N = 100: with N = 20 the last discrete beta is 20/20 = 1 and the ancestral update divides by sqrt(1 - beta) = 0
  sde = sde_lib.VPSDE(beta_min=0.1, beta_max=20., N=100)
  for tag, pred, corr, snr, seed in (('vp_ancestral_ald', sampling.AncestralSamplingPredictor, sampling.AnnealedLangevinDynamics, 0.05, 36),
                                     ('vp_ancestral_none', sampling.AncestralSamplingPredictor, sampling.NoneCorrector, 0.16, 37)):
    fn = sampling.get_pc_sampler(sde, shape, pred, corr, lambda v: v, snr=snr, n_steps=1, probability_flow=False,
                                 continuous=True, denoise=True, eps=1e-3, device='cpu')
    torch.manual_seed(seed)
    s, nfe = fn(m)
    assert torch.isfinite(s).all(), tag
    out2[tag] = s.numpy()
  np.savez_compressed(os.path.join(MG.OUT, 'pc_ancestral_ald_tiny.npz'), **out2)
  print('written', {k: v.shape for k, v in out2.items()})
  print('written', {k: (v.shape if hasattr(v, 'shape') else v) for k, v in out.items()})


if __name__ == '__main__':
  main()
