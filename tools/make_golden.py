"""Generate the golden fixtures under tests/golden/ by running the REAL reference.

Runs only in the build container (needs /root/reference, read-only).  It imports the
reference's own modules (``models.ncsnpp``, ``sampling``, ``sde_lib``,
``models.utils``, ``op.upfirdn2d``) with an ``ml_collections`` shim, loads into the
reference network a ``state_dict`` produced deterministically by this repository's
``NCSNpp(config)`` constructor under ``torch.manual_seed`` (same key names, so
``load_state_dict(strict=True)`` is the check that the layouts agree), and stores
seeded inputs and the reference's outputs as small ``.npz`` files.  The oracle in
``oracle/`` is then pinned against these files by ``tests/test_oracle_golden.py``.

    python tools/make_golden.py            # ~3 min on 8 cores (first run JIT-builds the reference's op/)
"""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('SCORE_SDE_REFERENCE', '/root/reference')
OUT = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, REPO)

from score_sde_pytorch_b200 import configs as our_configs            # noqa: E402
from score_sde_pytorch_b200.models.ncsnpp import NCSNpp as OurNCSNpp  # noqa: E402


def import_reference():
  class ConfigDict(dict):
    def __getattr__(self, k):
      try:
        return self[k]
      except KeyError as e:
        raise AttributeError(k) from e
    __setattr__ = dict.__setitem__
  shim = types.ModuleType('ml_collections')
  shim.ConfigDict = ConfigDict
  sys.modules['ml_collections'] = shim
  sys.path.insert(0, REF)
  import sde_lib, sampling                      # noqa: E401
  from models import ncsnpp, utils as mutils    # first import JIT-builds op/ (cached afterwards)
  from op import upfirdn2d as ref_upfirdn2d
  return sde_lib, sampling, ncsnpp, mutils, ref_upfirdn2d


def golden_configs():
  """name -> (config, batch).  Each is also constructible on the GPU box from its name."""
  tiny = our_configs.tiny_ncsnpp()
  tiny_vp = our_configs.tiny_ncsnpp()
  tiny_vp.model.scale_by_sigma = False
  tiny_vp.data.centered = True
  tiny_noattn = our_configs.tiny_ncsnpp(nf=16, image_size=8, ch_mult=(1, 1, 2), attn_resolutions=())
  tiny_noattn.model.skip_rescale = False
  tiny_noattn.model.progressive_input = 'none'
  cifar = our_configs.ve_cifar10_ncsnpp_continuous()
  cifar.model.init_scale = 1.0
  return {'tiny': (tiny, 2), 'tiny_vp': (tiny_vp, 2), 'tiny_noattn': (tiny_noattn, 3), 'cifar10_ve': (cifar, 2)}


def our_weights(cfg, seed=0):
  torch.manual_seed(seed)
  return OurNCSNpp(cfg).state_dict()


def main():
  os.makedirs(OUT, exist_ok=True)
  torch.set_num_threads(8)
  sde_lib, sampling, ncsnpp, mutils, ref_upfirdn2d = import_reference()

  # ---- 1. upfirdn2d: the three parameterisations NCSN++ uses + a generic one -------------
  g = torch.Generator().manual_seed(123)
  k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32)
  k /= k.sum()
  fir = {}
  for name, (shape, kk, up, down, pad) in {
      'down2': ((2, 8, 12, 12), k, 1, 2, (1, 1)),
      'up2': ((2, 8, 6, 6), k * 4, 2, 1, (2, 1)),
      'pad22': ((2, 3, 9, 9), k, 1, 1, (2, 2)),
      'generic': ((1, 2, 7, 5), np.arange(1, 7, dtype=np.float32).reshape(2, 3) / 21., 3, 2, (1, 2)),
  }.items():
    x = torch.randn(*shape, generator=g)
    y = ref_upfirdn2d(x, torch.tensor(kk), up=up, down=down, pad=pad)
    fir[name + '_x'] = x.numpy(); fir[name + '_k'] = kk; fir[name + '_y'] = y.numpy()
    fir[name + '_p'] = np.array([up, down, pad[0], pad[1]], dtype=np.int32)
  np.savez_compressed(os.path.join(OUT, 'upfirdn2d.npz'), **fir)

  # ---- 2. network forwards ---------------------------------------------------------------
  for name, (cfg, B) in golden_configs().items():
    cfg.device = torch.device('cpu')
    sd = our_weights(cfg)
    torch.manual_seed(0)
    ref_model = mutils.get_model('ncsnpp')(cfg).eval()
    ref_model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(7)
    R, C = cfg.data.image_size, cfg.data.num_channels
    x = torch.randn(B, C, R, R, generator=g) * 3.0
    sigma = torch.exp(torch.linspace(np.log(40.0), np.log(0.02), B))
    if name == 'tiny_vp':
      sigma = torch.linspace(999., 3., B)       # VP label range (999 t)
    with torch.no_grad():
      y = ref_model(x, sigma)
      taps = {}
      hooks = []
      for i, mod in enumerate(ref_model.all_modules):
        hooks.append(mod.register_forward_hook(lambda m, a, o, i=i: taps.__setitem__(i, o.detach().clone())))
      ref_model(x, sigma)
      for h in hooks:
        h.remove()
    rec = dict(x=x.numpy(), sigma=sigma.numpy(), y=y.numpy())
    if name != 'cifar10_ve':     # per-module activations for the small models (a few hundred KB)
      for i, v in taps.items():
        rec[f'tap{i}'] = v.numpy()
    else:                        # checksums only for the 62.8 M-parameter model
      rec['tap_norms'] = np.array([float(taps[i].double().norm()) if i in taps else 0.0
                                   for i in range(len(ref_model.all_modules))])
    np.savez_compressed(os.path.join(OUT, f'ncsnpp_{name}.npz'), **rec)
    print(name, 'forward done', float(y.abs().mean()))

    # ---- 3. PC sampler trajectories on the small models -----------------------------------
    if name == 'tiny':
      sde = sde_lib.VESDE(sigma_min=0.01, sigma_max=50, N=12)
      shape = (B, C, R, R)
      fn = sampling.get_pc_sampler(sde, shape, sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector,
                                   lambda v: v, snr=0.16, n_steps=1, probability_flow=False, continuous=True,
                                   denoise=True, eps=1e-5, device='cpu')
      torch.manual_seed(11)
      s, nfe = fn(ref_model)
      fn2 = sampling.get_pc_sampler(sde, shape, sampling.EulerMaruyamaPredictor, sampling.NoneCorrector,
                                    lambda v: v, snr=0.16, n_steps=1, probability_flow=False, continuous=True,
                                    denoise=False, eps=1e-5, device='cpu')
      torch.manual_seed(12)
      s2, nfe2 = fn2(ref_model)
      np.savez_compressed(os.path.join(OUT, 'pc_ve_tiny.npz'), rd_langevin=s.numpy(), nfe=nfe,
                          em_none=s2.numpy(), nfe2=nfe2)
    if name == 'tiny_vp':
      sde = sde_lib.VPSDE(beta_min=0.1, beta_max=20., N=20)
      shape = (B, C, R, R)
      out = {}
      for tag, pred, corr, seed in (('em_none', sampling.EulerMaruyamaPredictor, sampling.NoneCorrector, 21),
                                    ('rd_langevin', sampling.ReverseDiffusionPredictor, sampling.LangevinCorrector, 22)):
        fn = sampling.get_pc_sampler(sde, shape, pred, corr, lambda v: v, snr=0.16, n_steps=1,
                                     probability_flow=False, continuous=True, denoise=True, eps=1e-3, device='cpu')
        torch.manual_seed(seed)
        s, nfe = fn(ref_model)
        out[tag] = s.numpy(); out[tag + '_nfe'] = nfe
      np.savez_compressed(os.path.join(OUT, 'pc_vp_tiny.npz'), **out)

  # ---- 4. SDE scalar tables ---------------------------------------------------------------
  ve = sde_lib.VESDE(0.01, 50, 1000)
  t = torch.linspace(1, 1e-5, 1000)
  _, G = ve.discretize(torch.zeros(1000, 1, 1, 1), t)
  sig = ve.marginal_prob(torch.zeros(1000, 1, 1, 1), t)[1]
  vp = sde_lib.VPSDE(0.1, 20., 1000)
  t3 = torch.linspace(1, 1e-3, 1000)
  fvp, Gvp = vp.discretize(torch.ones(1000, 1, 1, 1), t3)
  np.savez_compressed(os.path.join(OUT, 'sde_tables.npz'), ve_G=G.numpy(), ve_sigma=sig.numpy(),
                      vp_f=fvp.reshape(-1).numpy(), vp_G=Gvp.numpy(), vp_std=vp.marginal_prob(torch.zeros(1000, 1, 1, 1), t3)[1].numpy())
  print('golden fixtures written to', OUT)


if __name__ == '__main__':
  main()
