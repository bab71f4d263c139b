"""CPU: inpainting / colorization (controllable_generation.py:8-198).  Goldens from the REAL reference
(tools/make_golden_controllable.py); both the oracle restatement and this package's host-side mirror (driven by a
user-style nn.Module carrying the oracle arithmetic) must reproduce them draw for draw."""
import torch

from helpers import golden, golden_config, seeded_model, rel_l2
from oracle import ncsnpp_oracle as NO
from oracle import sampling_oracle as SO


class _TorchModel(torch.nn.Module):
  def __init__(self, cfg):
    super().__init__()
    self.cfg, self.sd = cfg, seeded_model(cfg).state_dict()

  def forward(self, x, labels):
    return NO.ncsnpp_forward(self.sd, self.cfg, x, labels)


def test_oracle_inpaint_and_colorize_match_reference():
  g = golden('controllable_tiny.npz')
  cfg = golden_config('tiny')
  model = _TorchModel(cfg)
  data, mask, gray = (torch.from_numpy(g[k]) for k in ('data', 'mask', 'gray'))
  torch.manual_seed(71)
  out = SO.inpaint_sample(SO.VE(0.01, 50, 12), model, data, mask)
  assert rel_l2(out, torch.from_numpy(g['inpainted'])) < 1e-5
  torch.manual_seed(72)
  out = SO.colorize_sample(SO.VE(0.01, 50, 12), model, gray)
  assert rel_l2(out, torch.from_numpy(g['colorized'])) < 1e-5


def test_package_inpainter_and_colorizer_match_reference():
  from score_sde_pytorch_b200 import controllable_generation as CG, sampling, sde_lib
  g = golden('controllable_tiny.npz')
  cfg = golden_config('tiny')
  model = _TorchModel(cfg)
  data, mask, gray = (torch.from_numpy(g[k]) for k in ('data', 'mask', 'gray'))
  sde = sde_lib.VESDE(0.01, 50, 12)
  inp = CG.get_pc_inpainter(sde, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, lambda v: v, snr=0.16,
                            n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-5)
  torch.manual_seed(71)
  out = inp(model, data, mask)
  assert rel_l2(out, torch.from_numpy(g['inpainted'])) < 1e-5
  known = mask.bool()
  assert torch.allclose(out[known], data[known], atol=1e-4)        # known pixels come back (denoised mean of the data marginal)
  col = CG.get_pc_colorizer(sde, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, lambda v: v, snr=0.16,
                            n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-5)
  torch.manual_seed(72)
  out = col(model, gray)
  assert rel_l2(out, torch.from_numpy(g['colorized'])) < 1e-5
