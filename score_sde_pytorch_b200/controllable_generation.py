"""Inpainting and colorization with PC samplers — host-side mirror of ``controllable_generation.py:8-198``.

Same factories, signatures and results: ``get_pc_inpainter(...) -> pc_inpainter(model, data, mask)`` and
``get_pc_colorizer(...) -> pc_colorizer(model, gray_scale_img)``.  Both are the PC loop of ``sampling.py`` with one extra
step after every corrector / predictor update: the known part of the image (the masked pixels, or the luminance channel
of an orthonormal colour transform) is replaced by a fresh draw from the forward marginal of the data at time ``t``.
The update functions are this package's ``shared_*_update_fn``, so the score network may be the engine-backed NCSN++
(one engine evaluation per score call; the per-step blend is a handful of elementwise torch ops on 12 KB per image) or any
user module.  RNG consumption follows the reference draw for draw (prior sample, then per update: the update's own
noise, then ``randn_like`` for the data marginal).
"""
import functools

import torch

from .sampling import shared_corrector_update_fn, shared_predictor_update_fn

# Orthonormal colour transform of ``controllable_generation.py:109-111``: channel 0 of ``decouple(x)`` is the grey level.
_M = ((5.7735014e-01, -8.1649649e-01, 4.7008697e-08),
      (5.7735026e-01, 4.0824834e-01, 7.0710671e-01),
      (5.7735026e-01, 4.0824822e-01, -7.0710683e-01))


def _update_fns(sde, predictor, corrector, snr, n_steps, probability_flow, continuous):
  pred = functools.partial(shared_predictor_update_fn, sde=sde, predictor=predictor, probability_flow=probability_flow,
                           continuous=continuous)
  corr = functools.partial(shared_corrector_update_fn, sde=sde, corrector=corrector, continuous=continuous, snr=snr,
                           n_steps=n_steps)
  return corr, pred


def _constrained_pc_loop(sde, model, x, known, mask, update_fns, to_latent, from_latent, eps):
  """``sde.N`` iterations of (corrector, predictor), each followed by the data-consistency step
  (``controllable_generation.py:43-52`` / ``:137-146``): in the latent space given by ``to_latent``, the coordinates
  selected by ``mask`` are overwritten with a noisy copy of ``known`` at the current noise level.  As in the reference,
  ``x_mean`` is rebuilt from the *blended* ``x`` (not from the update's own mean)."""
  timesteps = torch.linspace(sde.T, eps, sde.N)
  x_mean = x
  for i in range(sde.N):
    t = timesteps[i]
    for update_fn in update_fns:
      vec_t = torch.ones(x.shape[0], device=x.device) * t
      x, _ = update_fn(x, vec_t, model=model)
      known_mean, std = sde.marginal_prob(known, vec_t)
      known_noisy = known_mean + torch.randn_like(x) * std[:, None, None, None]
      x = from_latent(to_latent(x) * (1. - mask) + known_noisy * mask)
      x_mean = from_latent(to_latent(x) * (1. - mask) + known_mean * mask)
  return x, x_mean


def get_pc_inpainter(sde, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False, continuous=False,
                     denoise=True, eps=1e-5):
  """``controllable_generation.py:8-80``.  ``mask`` is 1 on known pixels, 0 where the image is to be generated."""
  update_fns = _update_fns(sde, predictor, corrector, snr, n_steps, probability_flow, continuous)
  ident = lambda v: v

  def pc_inpainter(model, data, mask):
    with torch.no_grad():
      x = data * mask + sde.prior_sampling(data.shape).to(data.device) * (1. - mask)
      x, x_mean = _constrained_pc_loop(sde, model, x, data, mask, update_fns, ident, ident, eps)
      return inverse_scaler(x_mean if denoise else x)

  return pc_inpainter


def get_pc_colorizer(sde, predictor, corrector, inverse_scaler, snr, n_steps=1, probability_flow=False, continuous=False,
                     denoise=True, eps=1e-5):
  """``controllable_generation.py:83-198``.  ``gray_scale_img`` has identical R, G, B channels."""
  update_fns = _update_fns(sde, predictor, corrector, snr, n_steps, probability_flow, continuous)
  M = torch.tensor(_M)
  invM = torch.inverse(M)
  decouple = lambda v: torch.einsum('bihw,ij->bjhw', v, M.to(v.device))
  couple = lambda v: torch.einsum('bihw,ij->bjhw', v, invM.to(v.device))

  def pc_colorizer(model, gray_scale_img):
    with torch.no_grad():
      g = gray_scale_img
      mask = torch.cat([torch.ones_like(g[:, :1, ...]), torch.zeros_like(g[:, 1:, ...])], dim=1)
      x = couple(decouple(g) * mask + decouple(sde.prior_sampling(g.shape).to(g.device) * (1. - mask)))
      x, x_mean = _constrained_pc_loop(sde, model, x, decouple(g), mask, update_fns, decouple, couple, eps)
      return inverse_scaler(x_mean if denoise else x)

  return pc_colorizer
