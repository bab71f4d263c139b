"""ORACLE (test infrastructure, not product code): plain PyTorch restatement of
the reference's predictor–corrector sampling loop and SDE scalar math.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Pinned against the real
reference by ``tools/make_golden.py`` → ``tests/golden/pc_*.npz``.

The loop is written flat (no registries, no closures rebuilt per step), with
each line citing the reference statement it restates.  ``model(x, labels)`` is
any callable returning the network output in NCHW.
"""
import numpy as np
import torch


class VE:
  """sde_lib.py:207-254."""
  kind = 've'

  def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
    self.sigma_min, self.sigma_max, self.N, self.T = sigma_min, sigma_max, N, 1
    self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))  # :219

  def prior_sampling(self, shape):
    return torch.randn(*shape) * self.sigma_max                                               # :238-239

  def sigma(self, t):
    return self.sigma_min * (self.sigma_max / self.sigma_min) ** t                            # :233-236

  def sde(self, x, t):                                                                        # :224-231
    g = self.sigma(t) * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)),
                                                device=t.device))
    return torch.zeros_like(x), g

  def discretize(self, x, t):                                                                 # :246-254
    idx = (t * (self.N - 1) / self.T).long()
    tab = self.discrete_sigmas.to(t.device)
    sigma = tab[idx]
    adj = torch.where(idx == 0, torch.zeros_like(t), tab[idx - 1])
    return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adj ** 2)

  def score(self, model, x, t, continuous=True):                                              # models/utils.py:163-173
    if continuous:
      labels = self.sigma(t)
    else:
      labels = torch.round((self.T - t) * (self.N - 1)).long()
    return model(x, labels)

  def langevin_alpha(self, t):
    return torch.ones_like(t)                                                                 # sampling.py:270-271


class VP:
  """sde_lib.py:112-164."""
  kind = 'vp'

  def __init__(self, beta_min=0.1, beta_max=20, N=1000):
    self.beta_0, self.beta_1, self.N, self.T = beta_min, beta_max, N, 1
    self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
    self.alphas = 1. - self.discrete_betas
    self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - torch.cumprod(self.alphas, dim=0))

  def prior_sampling(self, shape):
    return torch.randn(*shape)                                                                # :147-148

  def sde(self, x, t):                                                                        # :135-139
    beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
    return -0.5 * beta_t[:, None, None, None] * x, torch.sqrt(beta_t)

  def std(self, t):                                                                           # :141-145
    lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
    return torch.sqrt(1. - torch.exp(2. * lmc))

  def discretize(self, x, t):                                                                 # :155-164
    idx = (t * (self.N - 1) / self.T).long()
    beta = self.discrete_betas.to(x.device)[idx]
    alpha = self.alphas.to(x.device)[idx]
    return torch.sqrt(alpha)[:, None, None, None] * x - x, torch.sqrt(beta)

  def score(self, model, x, t, continuous=True):                                              # models/utils.py:144-161
    if continuous:
      out = model(x, t * 999)
      std = self.std(t)
    else:
      labels = t * (self.N - 1)
      out = model(x, labels)
      std = self.sqrt_1m_alphas_cumprod.to(labels.device)[labels.long()]
    return -out / std[:, None, None, None]

  def langevin_alpha(self, t):                                                                # sampling.py:267-269
    idx = (t * (self.N - 1) / self.T).long()
    return self.alphas.to(t.device)[idx]


class SubVP(VP):
  """sde_lib.py:167-204: same drift and mean as VP; diffusion sqrt(beta_t (1 - e^{-2 beta_0 t - (beta_1-beta_0) t^2})),
  marginal std 1 - e^{2 log_mean_coeff} (no square root).  No `alphas` table: the reference's Langevin corrector
  has no sub-VP branch that works (sampling.py:267-271 reads sde.alphas only for VPSDE, alpha = 1 otherwise)."""
  kind = 'subvp'

  def sde(self, x, t):                                                                        # :183-188
    beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
    discount = 1. - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
    return -0.5 * beta_t[:, None, None, None] * x, torch.sqrt(beta_t * discount)

  def std(self, t):                                                                           # :190-194
    lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
    return 1 - torch.exp(2. * lmc)

  def discretize(self, x, t):                                                                 # base class, sde_lib.py:52-69
    dt = 1 / self.N
    drift, diffusion = self.sde(x, t)
    return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

  def langevin_alpha(self, t):
    return torch.ones_like(t)                                                                 # sampling.py:270-271


def langevin_step(sde, model, x, t, snr, n_steps, continuous=True):
  """sampling.py:262-282."""
  alpha = sde.langevin_alpha(t)
  x_mean = x
  for _ in range(n_steps):
    grad = sde.score(model, x, t, continuous)
    noise = torch.randn_like(x)
    grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
    step_size = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
    x_mean = x + step_size[:, None, None, None] * grad
    x = x_mean + torch.sqrt(step_size * 2)[:, None, None, None] * noise
  return x, x_mean


def reverse_diffusion_step(sde, model, x, t, continuous=True, probability_flow=False):
  """sampling.py:195-200 with sde_lib.py:102-107."""
  f, G = sde.discretize(x, t)
  rev_f = f - G[:, None, None, None] ** 2 * sde.score(model, x, t, continuous) * (0.5 if probability_flow else 1.)
  rev_G = torch.zeros_like(G) if probability_flow else G
  z = torch.randn_like(x)
  x_mean = x - rev_f
  return x_mean + rev_G[:, None, None, None] * z, x_mean


def euler_maruyama_step(sde, model, x, t, continuous=True, probability_flow=False):
  """sampling.py:181-187 with sde_lib.py:93-100."""
  dt = -1. / sde.N
  z = torch.randn_like(x)
  drift, diffusion = sde.sde(x, t)
  drift = drift - diffusion[:, None, None, None] ** 2 * sde.score(model, x, t, continuous) * (0.5 if probability_flow else 1.)
  diffusion = 0. if probability_flow else diffusion
  x_mean = x + drift * dt
  if probability_flow:
    return x_mean, x_mean
  return x_mean + diffusion[:, None, None, None] * np.sqrt(-dt) * z, x_mean


def ancestral_step(sde, model, x, t, continuous=True):
  """AncestralSamplingPredictor, sampling.py:204-239 (VE :213-223, VP :225-232; no sub-VP branch in the reference)."""
  idx = (t * (sde.N - 1) / sde.T).long()
  score = sde.score(model, x, t, continuous)
  if sde.kind == 've':
    tab = sde.discrete_sigmas.to(t.device)
    sigma = tab[idx]
    adj = torch.where(idx == 0, torch.zeros_like(t), tab[idx - 1])
    x_mean = x + score * (sigma ** 2 - adj ** 2)[:, None, None, None]
    std = torch.sqrt((adj ** 2 * (sigma ** 2 - adj ** 2)) / (sigma ** 2))
    return x_mean + std[:, None, None, None] * torch.randn_like(x), x_mean
  beta = sde.discrete_betas.to(t.device)[idx]
  x_mean = (x + beta[:, None, None, None] * score) / torch.sqrt(1. - beta)[:, None, None, None]
  return x_mean + torch.sqrt(beta)[:, None, None, None] * torch.randn_like(x), x_mean


def ald_step(sde, model, x, t, snr, n_steps, continuous=True):
  """AnnealedLangevinDynamics, sampling.py:286-319: the step size comes from the marginal std, not from norms.
  (sub-VP reaches `sde.alphas`, which subVPSDE does not have: the reference raises there, so does this.)"""
  if sde.kind == 've':
    alpha = torch.ones_like(t)
    std = sde.sigma(t)
  else:
    alpha = sde.alphas.to(t.device)[(t * (sde.N - 1) / sde.T).long()]
    std = sde.std(t)
  x_mean = x
  for _ in range(n_steps):
    grad = sde.score(model, x, t, continuous)
    noise = torch.randn_like(x)
    step_size = (snr * std) ** 2 * 2 * alpha
    x_mean = x + step_size[:, None, None, None] * grad
    x = x_mean + noise * torch.sqrt(step_size * 2)[:, None, None, None]
  return x, x_mean


def pc_sample(sde, model, shape, predictor='reverse_diffusion', corrector='langevin', snr=0.16,
              n_steps=1, eps=1e-5, continuous=True, denoise=True, device='cpu', num_iters=None,
              x_init=None, trace=None):
  """sampling.py:390-409.  ``num_iters`` truncates the loop after that many of the
  ``sde.N`` iterations (for K-step parity tests); ``x_init`` replaces the prior draw;
  ``trace`` (a list) receives ``x`` after every iteration."""
  with torch.no_grad():
    x = (sde.prior_sampling(shape) if x_init is None else x_init).to(device)
    timesteps = torch.linspace(sde.T, eps, sde.N, device=device)
    x_mean = x
    for i in range(sde.N if num_iters is None else num_iters):
      vec_t = torch.ones(shape[0], device=device) * timesteps[i]
      if corrector == 'langevin':
        x, x_mean = langevin_step(sde, model, x, vec_t, snr, n_steps, continuous)
      elif corrector == 'ald':
        x, x_mean = ald_step(sde, model, x, vec_t, snr, n_steps, continuous)
      if predictor == 'ancestral_sampling':
        x, x_mean = ancestral_step(sde, model, x, vec_t, continuous)
      elif predictor == 'reverse_diffusion':
        x, x_mean = reverse_diffusion_step(sde, model, x, vec_t, continuous)
      elif predictor == 'euler_maruyama':
        x, x_mean = euler_maruyama_step(sde, model, x, vec_t, continuous)
      else:
        x_mean = x                                   # NonePredictor.update_fn returns (x, x), sampling.py:249-250
      if trace is not None:
        trace.append(x.clone())
    return (x_mean if denoise else x), sde.N * (n_steps + 1)


def ode_sample(sde, model, shape, z=None, denoise=False, rtol=1e-5, atol=1e-5, method='RK45', eps=1e-3, device='cpu'):
  """sampling.py:414-485 (get_ode_sampler / ode_sampler): the probability-flow ODE integrated with scipy's solve_ivp on
  a flattened float64 numpy state; the right-hand side (``drift_fn``, :443-447) is ``rsde.sde(x, t)[0]`` with
  ``probability_flow=True`` (sde_lib.py:93-100) on the float32 state; optional one-step denoise (:433-441) with the
  reverse-diffusion predictor at ``t = eps``.  Returns ``(x, nfev)``."""
  from scipy import integrate
  with torch.no_grad():
    x = (sde.prior_sampling(shape) if z is None else z).to(device)

    def ode_func(t, flat):
      xt = torch.from_numpy(flat.reshape(shape)).to(device).type(torch.float32)                 # :460-461
      vec_t = torch.ones(shape[0], device=xt.device) * t                                        # :462
      drift, diffusion = sde.sde(xt, vec_t)
      drift = drift - diffusion[:, None, None, None] ** 2 * sde.score(model, xt, vec_t, True) * 0.5
      return drift.detach().cpu().numpy().reshape((-1,))                                        # :464

    sol = integrate.solve_ivp(ode_func, (sde.T, eps), x.detach().cpu().numpy().reshape((-1,)), rtol=rtol, atol=atol, method=method)
    x = torch.tensor(sol.y[:, -1]).reshape(shape).to(device).type(torch.float32)                # :469
    if denoise:                                                                                 # :472-473
      vec_eps = torch.ones(shape[0], device=x.device) * eps
      _, x = reverse_diffusion_step(sde, model, x, vec_eps, continuous=True, probability_flow=False)
    return x, sol.nfev


# ---- controllable generation (controllable_generation.py) ------------------------------------------------------------
_COLOR_M = torch.tensor([[5.7735014e-01, -8.1649649e-01, 4.7008697e-08],
                         [5.7735026e-01, 4.0824834e-01, 7.0710671e-01],
                         [5.7735026e-01, 4.0824822e-01, -7.0710683e-01]])                       # :109-111


def _marginal(sde, x, t):
  """(mean, std) of p_t(x(t) | x(0)): sde_lib.py:233-236 (VE), :141-145 (VP), :190-194 (sub-VP)."""
  if sde.kind == 've':
    return x, sde.sigma(t)
  lmc = -0.25 * t ** 2 * (sde.beta_1 - sde.beta_0) - 0.5 * t * sde.beta_0
  return torch.exp(lmc)[:, None, None, None] * x, sde.std(t)


def _pc_update(sde, model, x, vec_t, which, name, snr, n_steps, continuous):
  if which == 'corrector':
    return langevin_step(sde, model, x, vec_t, snr, n_steps, continuous) if name == 'langevin' else (x, x)
  if name == 'reverse_diffusion':
    return reverse_diffusion_step(sde, model, x, vec_t, continuous)
  if name == 'euler_maruyama':
    return euler_maruyama_step(sde, model, x, vec_t, continuous)
  return x, x


def inpaint_sample(sde, model, data, mask, predictor='reverse_diffusion', corrector='langevin', snr=0.16, n_steps=1,
                   continuous=True, denoise=True, eps=1e-5):
  """controllable_generation.py:57-78 with the update of :43-52."""
  with torch.no_grad():
    x = data * mask + sde.prior_sampling(data.shape).to(data.device) * (1. - mask)                # :71
    timesteps = torch.linspace(sde.T, eps, sde.N)
    x_mean = x
    for i in range(sde.N):
      for which, name in (('corrector', corrector), ('predictor', predictor)):                    # :75-76
        vec_t = torch.ones(data.shape[0], device=data.device) * timesteps[i]
        x, x_mean = _pc_update(sde, model, x, vec_t, which, name, snr, n_steps, continuous)
        mean, std = _marginal(sde, data, vec_t)
        noisy = mean + torch.randn_like(x) * std[:, None, None, None]                             # :48
        x = x * (1. - mask) + noisy * mask
        x_mean = x * (1. - mask) + mean * mask                                                    # :50 (uses the blended x)
    return x_mean if denoise else x


def colorize_sample(sde, model, gray, predictor='reverse_diffusion', corrector='langevin', snr=0.16, n_steps=1,
                    continuous=True, denoise=True, eps=1e-5):
  """controllable_generation.py:172-196 with the update of :137-146."""
  M = _COLOR_M.to(gray.device)
  invM = torch.inverse(_COLOR_M).to(gray.device)
  dec = lambda v: torch.einsum('bihw,ij->bjhw', v, M)
  cou = lambda v: torch.einsum('bihw,ij->bjhw', v, invM)
  with torch.no_grad():
    mask = torch.cat([torch.ones_like(gray[:, :1]), torch.zeros_like(gray[:, 1:])], dim=1)        # :152-155
    x = cou(dec(gray) * mask + dec(sde.prior_sampling(gray.shape).to(gray.device) * (1. - mask)))  # :186-188
    timesteps = torch.linspace(sde.T, eps, sde.N)
    x_mean = x
    for i in range(sde.N):
      for which, name in (('corrector', corrector), ('predictor', predictor)):
        vec_t = torch.ones(gray.shape[0], device=gray.device) * timesteps[i]
        x, x_mean = _pc_update(sde, model, x, vec_t, which, name, snr, n_steps, continuous)
        mean, std = _marginal(sde, dec(gray), vec_t)
        noisy = mean + torch.randn_like(x) * std[:, None, None, None]
        x = cou(dec(x) * (1. - mask) + noisy * mask)
        x_mean = cou(dec(x) * (1. - mask) + mean * mask)
    return x_mean if denoise else x
