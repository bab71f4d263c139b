"""Shared test helpers: golden loading, config lookup by golden name, deterministic weights."""
import os

import numpy as np
import torch

from score_sde_pytorch_b200 import configs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden(name):
  return dict(np.load(os.path.join(GOLDEN, name)))


def golden_config(name):
  """Same constructions as tools/make_golden.py:golden_configs()."""
  if name == 'tiny':
    return configs.tiny_ncsnpp()
  if name == 'tiny_vp':
    c = configs.tiny_ncsnpp()
    c.model.scale_by_sigma = False
    c.data.centered = True
    return c
  if name == 'tiny_noattn':
    c = configs.tiny_ncsnpp(nf=16, image_size=8, ch_mult=(1, 1, 2), attn_resolutions=())
    c.model.skip_rescale = False
    c.model.progressive_input = 'none'
    return c
  if name == 'cifar10_ve':
    c = configs.ve_cifar10_ncsnpp_continuous()
    c.model.init_scale = 1.0
    return c
  if name == 'tiny_ddpmpp':
    return configs.tiny_ddpmpp()
  if name == 'cifar10_ddpmpp':
    c = configs.vp_cifar10_ddpmpp_continuous()
    c.model.init_scale = 1.0
    return c
  if name == 'celebahq_256_ddpmpp_subvp':
    c = configs.subvp_celebahq_256_ddpmpp_continuous()
    c.model.init_scale = 1.0
    return c
  if name == 'cifar10_deep':
    c = configs.ve_cifar10_ncsnpp_deep_continuous()
    c.model.init_scale = 1.0
    return c
  if name == 'tiny_progressive':
    return configs.tiny_progressive()
  if name in ('celebahq_256', 'ffhq_1024'):
    c = configs.ve_celebahq_256_ncsnpp_continuous() if name == 'celebahq_256' else configs.ve_ffhq_1024_ncsnpp_continuous()
    c.model.init_scale = 1.0
    return c
  raise KeyError(name)


def seeded_model(cfg, seed=0, **kw):
  """The engine-backed module with the deterministic weights the goldens were made with."""
  from score_sde_pytorch_b200.models.ncsnpp import NCSNpp
  torch.manual_seed(seed)
  return NCSNpp(cfg, **kw)


def rel_l2(a, b):
  a = a.double().reshape(a.shape[0], -1)
  b = b.double().reshape(b.shape[0], -1)
  return ((a - b).norm(dim=1) / b.norm(dim=1).clamp_min(1e-30)).max().item()
