"""`-m gpu`: the high-resolution NCSN++ family (second variant): progressive='output_skip',
progressive_input='input_skip', Combine 'sum' (models/ncsnpp.py:163-166, 190-203, 289-292, 325-341, 366-367;
layerspp.py:44-59) through the engine - against the oracle, against a golden written by the REAL reference
(tools/make_golden_progressive.py), and at full size for the two reference configurations of the family."""
import pytest
import torch

from helpers import golden, golden_config, seeded_model, rel_l2
from oracle import ncsnpp_oracle as NO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
  import gpu_util
  gpu_util.strict_fp32()
  return torch.device('cuda:0')


@pytest.mark.parametrize('precision', ['fp32', 'tf32'])     # (32/64-channel layers: off the fp16 tiling; full-size f16 below)
def test_progressive_tiny_matches_oracle_and_reference_golden(dev, precision):
  g = golden('ncsnpp_tiny_progressive.npz')
  cfg = golden_config('tiny_progressive')
  model = seeded_model(cfg, precision=precision, keep_activations=True).to(dev)
  sd = {k: v.to(dev) for k, v in model.state_dict().items()}
  x, sigma = torch.from_numpy(g['x']).to(dev), torch.from_numpy(g['sigma']).to(dev)
  taps = {}
  with torch.no_grad():
    ref = NO.ncsnpp_forward(sd, cfg, x, sigma, taps=taps)
    y = model(x, sigma)
  tol_mod, tol_out = (1e-4, 1e-4) if precision == 'fp32' else (5e-3, 2.5e-3)
  rows = []
  for i in sorted(taps):
    if taps[i].dim() != 4:
      continue
    try:
      rows.append((i, rel_l2(model.tap(i), taps[i])))
    except RuntimeError:
      continue
  worst = sorted(rows, key=lambda r: -r[1])[:4]
  e_or, e_gold = rel_l2(y, ref), rel_l2(y, torch.from_numpy(g['y']).to(dev))
  print(f'progressive tiny [{precision}]: rel-L2 vs oracle {e_or:.3e}, vs reference CPU golden {e_gold:.3e}; {len(rows)} module taps, worst {worst}')
  assert len(rows) >= 10
  assert all(r[1] < tol_mod for r in rows), f'worst modules: {worst}'
  assert e_or < tol_out and e_gold < tol_out


def test_progressive_nofir_matches_oracle(dev):
  """fir=False member (avg-pool / nearest pyramids): oracle only - the reference's own Upsample(fir=False) does not run
  on current PyTorch (F.interpolate called with the mode in the scale_factor slot, layerspp.py:116)."""
  from score_sde_pytorch_b200 import configs
  cfg = configs.tiny_progressive(fir=False)
  model = seeded_model(cfg, precision='fp32').to(dev)
  sd = {k: v.to(dev) for k, v in model.state_dict().items()}
  torch.manual_seed(3)
  x = torch.randn(2, 3, 32, 32, device=dev) * 2
  sigma = torch.tensor([9.0, 0.2], device=dev)
  with torch.no_grad():
    assert rel_l2(model(x, sigma), NO.ncsnpp_forward(sd, cfg, x, sigma)) < 1e-4


@pytest.mark.parametrize('name,precision', [('celebahq_256', 'tf32'), ('celebahq_256', 'f16'), ('ffhq_1024', 'tf32')])
def test_high_resolution_reference_configs_full_size(dev, name, precision):
  """configs/ve/celebahq_256_ncsnpp_continuous.py (65.6 M parameters, 256x256, seven levels) and
  configs/ve/ffhq_ncsnpp_continuous.py (105.8 M parameters, 1024x1024, eight levels, nf=16: the 16/32-channel levels run on
  the CUDA-core convolution, the rest on tcgen05), one evaluation at batch 1 against the strict-fp32 oracle."""
  cfg = golden_config(name)
  model = seeded_model(cfg, precision=precision).to(dev)
  sd = {k: v.to(dev) for k, v in model.state_dict().items()}
  R = cfg.data.image_size
  torch.manual_seed(4)
  x = torch.randn(1, 3, R, R, device=dev) * 5
  sigma = torch.tensor([3.0], device=dev)
  with torch.no_grad():
    y = model(x, sigma)
    ref = NO.ncsnpp_forward(sd, cfg, x, sigma)
  e = rel_l2(y, ref)
  print(f'{name} [{precision}] 1x3x{R}x{R}: rel-L2 vs oracle {e:.3e}, {model.launches_per_forward()} launches')
  assert torch.isfinite(y).all()
  assert e < 2.5e-3


@pytest.mark.parametrize('precision', ['tf32', 'f16'])
def test_deep_cifar10_variant_matches_oracle(dev, precision):
  """configs/ve/cifar10_ncsnpp_deep_continuous.py: eight residual blocks per level ("deep")."""
  cfg = golden_config('cifar10_deep')
  model = seeded_model(cfg, precision=precision).to(dev)
  sd = {k: v.to(dev) for k, v in model.state_dict().items()}
  torch.manual_seed(6)
  x = torch.randn(2, 3, 32, 32, device=dev) * 4
  sigma = torch.tensor([12.0, 0.4], device=dev)
  with torch.no_grad():
    e = rel_l2(model(x, sigma), NO.ncsnpp_forward(sd, cfg, x, sigma))
  print(f'cifar10 deep [{precision}]: rel-L2 vs oracle {e:.3e}, {model.launches_per_forward()} launches, '
        f'{sum(p.numel() for p in model.parameters())} parameters')
  assert e < 2.5e-3
