"""TEST INFRASTRUCTURE - CPU/GPU restatement of the reference's score-matching losses and of the gradients of its two
native ops.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package; the product
(score_sde_pytorch_b200/) never does.

Each function names the reference lines it follows.  Pinned to the real reference by tests/golden/f4_losses.npz and
tests/golden/f4_op_grads.npz (tools/make_golden_f4.py ran the reference's own losses.py / op/*.py on CPU)."""
import torch
import torch.nn.functional as F

from . import ncsnpp_oracle as NO
from . import sampling_oracle as SO


def _reduce(v, reduce_mean):
  """losses.py:71: torch.mean over the data dims, or half the sum."""
  v = v.reshape(v.shape[0], -1)
  return torch.mean(v, dim=-1) if reduce_mean else 0.5 * torch.sum(v, dim=-1)


def sde_loss(sde, score_fn, batch, t, z, reduce_mean=True, likelihood_weighting=True):
  """losses.py:85-99 for given draws ``t`` and ``z``.  ``sde``: an oracle SDE (sampling_oracle.VE / VP / SubVP);
  ``score_fn(x, t)``: the score (``lambda x, t: sde.score(model, x, t)``)."""
  mean, std = SO._marginal(sde, batch, t)
  perturbed = mean + std[:, None, None, None] * z
  score = score_fn(perturbed, t)
  if not likelihood_weighting:
    losses = _reduce(torch.square(score * std[:, None, None, None] + z), reduce_mean)
  else:
    g2 = sde.sde(torch.zeros_like(batch), t)[1] ** 2
    losses = _reduce(torch.square(score + z / std[:, None, None, None]), reduce_mean) * g2
  return torch.mean(losses), perturbed


def smld_loss(sigmas_descending, model_fn, batch, labels, z, reduce_mean=False):
  """losses.py:108-124; ``sigmas_descending`` = flip(vesde.discrete_sigmas)."""
  sigmas = sigmas_descending.to(batch.device)[labels]
  noise = z * sigmas[:, None, None, None]
  perturbed = noise + batch
  score = model_fn(perturbed, labels)
  target = -noise / (sigmas ** 2)[:, None, None, None]
  losses = _reduce(torch.square(score - target), reduce_mean) * sigmas ** 2
  return torch.mean(losses), perturbed


def ddpm_loss(sqrt_alphas_cumprod, sqrt_1m_alphas_cumprod, model_fn, batch, labels, z, reduce_mean=True):
  """losses.py:131-146."""
  a = sqrt_alphas_cumprod.to(batch.device)[labels, None, None, None]
  s = sqrt_1m_alphas_cumprod.to(batch.device)[labels, None, None, None]
  perturbed = a * batch + s * z
  score = model_fn(perturbed, labels)
  losses = _reduce(torch.square(score - z), reduce_mean)
  return torch.mean(losses), perturbed


def upfirdn2d_grads(x, k, up, down, pad, grad_out, v):
  """First and second derivative of the FIR resampling op through autograd of its pure-torch form
  (op/upfirdn2d.py:159-200, restated in ncsnpp_oracle.upfirdn2d_native): returns ``(y, grad_input, gradgrad_out)`` with
  grad_input = d<y, grad_out>/dx and gradgrad_out = d<grad_input, v>/d grad_out - what UpFirDn2d.backward
  (:127-141) and UpFirDn2dBackward.backward (:66-85) compute with the CUDA kernel."""
  x = x.detach().clone().requires_grad_(True)
  go = grad_out.detach().clone().requires_grad_(True)
  y = NO.upfirdn2d_native(x, k, up=up, down=down, pad=pad)
  gi, = torch.autograd.grad(y, x, go, create_graph=True)
  ggo, = torch.autograd.grad((gi * v).sum(), go)
  return y.detach(), gi.detach(), ggo.detach()


def fused_leaky_relu_grads(x, b, grad_out, vi, vb, negative_slope=0.2, scale=2 ** 0.5):
  """op/fused_act.py:20-72 through autograd of the CPU form (:85-93): ``(y, grad_input, grad_bias, gradgrad_out)``,
  gradgrad_out = d(<grad_input, vi> + <grad_bias, vb>)/d grad_out."""
  x = x.detach().clone().requires_grad_(True)
  b = b.detach().clone().requires_grad_(True)
  go = grad_out.detach().clone().requires_grad_(True)
  rest = [1] * (x.ndim - 2)
  y = F.leaky_relu(x + b.view(1, -1, *rest), negative_slope=negative_slope) * scale
  gi, gb = torch.autograd.grad(y, (x, b), go, create_graph=True)
  ggo, = torch.autograd.grad((gi * vi).sum() + (gb * vb).sum(), go)
  return y.detach(), gi.detach(), gb.detach(), ggo.detach()
