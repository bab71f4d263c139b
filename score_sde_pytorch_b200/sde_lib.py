"""Forward SDEs, their reverse-time counterparts and discretisations.

Host-side mirror of the reference's ``sde_lib.py`` surface (``SDE`` ABC
``sde_lib.py:7-109``; ``VPSDE`` ``:112-164``; ``subVPSDE`` ``:167-204``;
``VESDE`` ``:207-254``): same class names, constructor arguments, method names
and return conventions, so user code written against the reference (custom
SDE subclasses, predictors calling ``sde.discretize`` / ``sde.reverse``)
keeps working.  The per-step arithmetic of the recognised sampler
combinations does not run through these tensors methods at all — the native
engine consumes the scalar schedules exported by :meth:`SDE.schedule` — but
every method is kept functional for the generic (any model / any predictor)
host loop in :mod:`sampling`.
"""
import abc
import math

import numpy as np
import torch


def _bcast(v, x):
  """Reshape a per-sample vector ``v[B]`` so it broadcasts against ``x[B,...]``."""
  return v.reshape((v.shape[0],) + (1,) * (x.dim() - 1))


class SDE(abc.ABC):
  """Abstract forward SDE ``dx = f(x,t) dt + g(t) dw`` on ``t in [0, T]``."""

  def __init__(self, N):
    super().__init__()
    self.N = N

  @property
  @abc.abstractmethod
  def T(self):
    """End time of the SDE."""

  @abc.abstractmethod
  def sde(self, x, t):
    """Return ``(drift[B,...], diffusion[B])`` at state ``x`` and times ``t[B]``."""

  @abc.abstractmethod
  def marginal_prob(self, x, t):
    """Return ``(mean, std[B])`` of the perturbation kernel ``p_t(x(t)|x(0)=x)``."""

  @abc.abstractmethod
  def prior_sampling(self, shape):
    """Draw from the terminal distribution ``p_T`` (on the CPU generator, as the
    reference does — ``sde_lib.py:147-148,198-199,238-239``)."""

  @abc.abstractmethod
  def prior_logp(self, z):
    """Log density of ``p_T`` at ``z``."""

  def discretize(self, x, t):
    """One-step discretisation ``x_{i+1} = x_i + f_i(x_i) + G_i z_i``.
    Default: Euler–Maruyama with ``dt = 1/N`` (``sde_lib.py:52-69``)."""
    dt = 1.0 / self.N
    drift, diffusion = self.sde(x, t)
    return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

  def reverse(self, score_fn, probability_flow=False):
    """Reverse-time SDE (or probability-flow ODE) driven by ``score_fn``.

    Mirrors ``sde_lib.py:71-109``: the returned object is an instance of a
    subclass of ``type(self)`` whose ``sde`` / ``discretize`` subtract
    ``g^2 * score`` (halved for the ODE) and zero the diffusion for the ODE.
    """
    fwd = self
    half = 0.5 if probability_flow else 1.0

    class RSDE(fwd.__class__):
      def __init__(self):  # deliberately does not call the forward ctor
        self.N = fwd.N
        self.probability_flow = probability_flow

      @property
      def T(self):
        return fwd.T

      def sde(self, x, t):
        drift, diffusion = fwd.sde(x, t)
        drift = drift - _bcast(diffusion, x) ** 2 * score_fn(x, t) * half
        return drift, (0. if probability_flow else diffusion)

      def discretize(self, x, t):
        f, G = fwd.discretize(x, t)
        rev_f = f - _bcast(G, x) ** 2 * score_fn(x, t) * half
        return rev_f, (torch.zeros_like(G) if probability_flow else G)

    return RSDE()

  # ---- scalar schedules consumed by the native engine -----------------------
  def schedule(self, eps, device='cpu'):
    """Per-step scalars of the PC loop for ``timesteps = linspace(T, eps, N)``.

    Returns a dict of float32 CPU tensors of length ``N`` built with the *same
    torch ops* the reference evaluates inside its loop (``sampling.py:401-405``
    + the SDE methods), so table entries are bit-equal to what the reference
    computes per step for a batch whose times are all equal.
    """
    raise NotImplementedError(f"SDE class {self.__class__.__name__} has no native schedule.")


class VPSDE(SDE):
  """Variance-preserving SDE, ``beta(t) = beta_0 + t (beta_1 - beta_0)``."""

  def __init__(self, beta_min=0.1, beta_max=20, N=1000):
    super().__init__(N)
    self.beta_0 = beta_min
    self.beta_1 = beta_max
    self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
    self.alphas = 1. - self.discrete_betas
    self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
    self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
    self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

  @property
  def T(self):
    return 1

  def _beta(self, t):
    return self.beta_0 + t * (self.beta_1 - self.beta_0)

  def _log_mean_coeff(self, t):
    return -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0

  def sde(self, x, t):
    beta_t = self._beta(t)
    return -0.5 * _bcast(beta_t, x) * x, torch.sqrt(beta_t)

  def marginal_prob(self, x, t):
    lmc = self._log_mean_coeff(t)
    return torch.exp(_bcast(lmc, x)) * x, torch.sqrt(1. - torch.exp(2. * lmc))

  def prior_sampling(self, shape):
    return torch.randn(*shape)

  def prior_logp(self, z):
    n = np.prod(z.shape[1:])
    return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

  def discretize(self, x, t):
    """DDPM ancestral discretisation (``sde_lib.py:155-164``)."""
    idx = (t * (self.N - 1) / self.T).long()
    beta = self.discrete_betas.to(x.device)[idx]
    alpha = self.alphas.to(x.device)[idx]
    return _bcast(torch.sqrt(alpha), x) * x - x, torch.sqrt(beta)

  def schedule(self, eps, device='cpu'):
    t = torch.linspace(self.T, eps, self.N)
    idx = (t * (self.N - 1) / self.T).long()
    beta_t = self._beta(t)
    lmc = self._log_mean_coeff(t)
    return dict(t=t, index=idx, beta_t=beta_t, diffusion=torch.sqrt(beta_t),
                std=torch.sqrt(1. - torch.exp(2. * lmc)),
                alpha=self.alphas[idx], beta_disc=self.discrete_betas[idx])


class subVPSDE(SDE):
  """Sub-VP SDE (``sde_lib.py:167-204``)."""

  def __init__(self, beta_min=0.1, beta_max=20, N=1000):
    super().__init__(N)
    self.beta_0 = beta_min
    self.beta_1 = beta_max

  @property
  def T(self):
    return 1

  def sde(self, x, t):
    beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
    discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
    return -0.5 * _bcast(beta_t, x) * x, torch.sqrt(beta_t * discount)

  def marginal_prob(self, x, t):
    lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
    return _bcast(torch.exp(lmc), x) * x, 1 - torch.exp(2. * lmc)

  def prior_sampling(self, shape):
    return torch.randn(*shape)

  def prior_logp(self, z):
    n = np.prod(z.shape[1:])
    return -n / 2. * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2, 3)) / 2.

  def schedule(self, eps, device='cpu'):
    t = torch.linspace(self.T, eps, self.N)
    beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
    discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
    lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
    return dict(t=t, index=(t * (self.N - 1) / self.T).long(), beta_t=beta_t,
                diffusion=torch.sqrt(beta_t * discount), std=1 - torch.exp(2. * lmc))


class VESDE(SDE):
  """Variance-exploding SDE, ``sigma(t) = sigma_min (sigma_max/sigma_min)^t``."""

  def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
    super().__init__(N)
    self.sigma_min = sigma_min
    self.sigma_max = sigma_max
    self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

  @property
  def T(self):
    return 1

  def _sigma(self, t):
    return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

  def sde(self, x, t):
    g = self._sigma(t) * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)),
                                                 device=t.device))
    return torch.zeros_like(x), g

  def marginal_prob(self, x, t):
    return x, self._sigma(t)

  def prior_sampling(self, shape):
    return torch.randn(*shape) * self.sigma_max

  def prior_logp(self, z):
    n = np.prod(z.shape[1:])
    return (-n / 2. * np.log(2 * np.pi * self.sigma_max ** 2)
            - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2))

  def discretize(self, x, t):
    """SMLD discretisation (``sde_lib.py:246-254``).  The sigma table is moved to
    ``t.device`` before indexing (the reference indexes a CPU table with a device
    index at ``:251``, which current PyTorch rejects on CUDA)."""
    idx = (t * (self.N - 1) / self.T).long()
    table = self.discrete_sigmas.to(t.device)
    sigma = table[idx]
    adjacent = torch.where(idx == 0, torch.zeros_like(t), table[idx - 1])
    return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)

  def schedule(self, eps, device='cpu'):
    t = torch.linspace(self.T, eps, self.N)
    idx = (t * (self.N - 1) / self.T).long()
    sigma = self.discrete_sigmas[idx]
    adjacent = torch.where(idx == 0, torch.zeros_like(t), self.discrete_sigmas[idx - 1])
    G = torch.sqrt(sigma ** 2 - adjacent ** 2)
    return dict(t=t, index=idx, sigma_t=self._sigma(t), G=G, G2=G ** 2,
                sigma_disc=sigma, adjacent_sigma=adjacent)
