"""`-m gpu`: inpainting / colorization (controllable_generation.py) over the engine-backed network: this package's
host-side loop on the GPU against the oracle loop on the same GPU (same prior draw, same CUDA noise stream), and
against the reference's own CPU result (tests/golden/controllable_tiny.npz), loosely (different RNG device)."""
import pytest
import torch

from helpers import golden, golden_config, seeded_model, rel_l2
from oracle import ncsnpp_oracle as NO
from oracle import sampling_oracle as SO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
  import gpu_util
  gpu_util.strict_fp32()
  return torch.device('cuda:0')


@pytest.mark.parametrize('task', ['inpaint', 'colorize'])
def test_controllable_generation_on_the_engine_matches_oracle(dev, task):
  from score_sde_pytorch_b200 import controllable_generation as CG, sampling, sde_lib
  g = golden('controllable_tiny.npz')
  cfg = golden_config('tiny')
  model = seeded_model(cfg, precision='fp32').to(dev)
  sd = {k: v.to(dev) for k, v in model.state_dict().items()}
  net = lambda a, l: NO.ncsnpp_forward(sd, cfg, a, l)
  sde, osde = sde_lib.VESDE(0.01, 50, 12), SO.VE(0.01, 50, 12)
  data, mask, gray = (torch.from_numpy(g[k]).to(dev) for k in ('data', 'mask', 'gray'))
  kw = dict(snr=0.16, n_steps=1, probability_flow=False, continuous=True, denoise=True, eps=1e-5)
  if task == 'inpaint':
    fn = CG.get_pc_inpainter(sde, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, lambda v: v, **kw)
    torch.manual_seed(71); torch.cuda.manual_seed(71)
    out = fn(model, data, mask)
    torch.manual_seed(71); torch.cuda.manual_seed(71)
    ref = SO.inpaint_sample(osde, net, data, mask)
    known = mask.bool()
    assert torch.allclose(out[known], data[known], atol=1e-4)
  else:
    fn = CG.get_pc_colorizer(sde, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, lambda v: v, **kw)
    torch.manual_seed(72); torch.cuda.manual_seed(72)
    out = fn(model, gray)
    torch.manual_seed(72); torch.cuda.manual_seed(72)
    ref = SO.colorize_sample(osde, net, gray)
  e = rel_l2(out, ref)
  print(f'{task} on the engine: rel-L2 vs oracle loop {e:.2e}')
  assert e < 2e-4
