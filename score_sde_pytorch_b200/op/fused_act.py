"""``fused_leaky_relu`` / ``FusedLeakyReLU`` on the sm_100a library.

Mirror of ``op/fused_act.py:74-97``: ``y = leaky_relu(x + bias[c], negative_slope) * scale`` with the bias
broadcast over dim 1.  Forward only.  (No model in the reference calls it — it is compiled at import,
``op/__init__.py:1`` — but it is part of the native surface, so it is kept.)
"""
import torch
from torch import nn

from .. import _lib


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
  if not input.is_cuda:
    raise RuntimeError('fused_leaky_relu (score_sde_pytorch_b200): input must be a CUDA tensor (no CPU path)')
  x = input.detach().to(torch.float32).contiguous()
  b = bias.detach().to(device=input.device, dtype=torch.float32).contiguous()
  y = torch.empty_like(x)
  step_b = 1
  for d in x.shape[2:]:
    step_b *= d
  with torch.cuda.device(input.device):
    _lib.call('b200_fused_bias_act_f32', _lib.ptr(x), _lib.ptr(b) if b.numel() else None, None, _lib.ptr(y),
              x.numel(), step_b, max(b.numel(), 1), 3, 0, float(negative_slope), float(scale),
              _lib.stream_ptr(input.device))
  return y


class FusedLeakyReLU(nn.Module):
  def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
    super().__init__()
    self.bias = nn.Parameter(torch.zeros(channel))
    self.negative_slope = negative_slope
    self.scale = scale

  def forward(self, input):
    return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
