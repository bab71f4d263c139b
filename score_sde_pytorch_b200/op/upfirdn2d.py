"""``upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))`` — FIR resampling on the sm_100a library.

Same call surface and result as the reference op (``op/upfirdn2d.py:145-156``): ``input`` is NCHW,
``kernel`` a 2-D FIR, ``up``/``down`` integer factors, ``pad=(pad0, pad1)`` applied to both axes.
The reference reshapes to ``[N*C, H, W, 1]`` before its pybind call (``op/upfirdn2d.py:99``) and so
does this wrapper; the engine itself calls the same kernel on NHWC with ``minor = C``.
Forward only (the sampling path never differentiates through it); CPU tensors are rejected —
the reference's pure-torch ``upfirdn2d_native`` fallback (``:159-200``) is restated in ``oracle/``.
"""
import ctypes

import torch

from .. import _lib


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
  if not input.is_cuda:
    raise RuntimeError('upfirdn2d (score_sde_pytorch_b200): input must be a CUDA tensor (no CPU path)')
  n, c, in_h, in_w = input.shape
  k = kernel.detach().to('cpu', torch.float32).contiguous()
  kh, kw = k.shape
  x = input.detach().to(torch.float32).contiguous()
  out_h = (in_h * up + pad[0] + pad[1] - kh) // down + 1
  out_w = (in_w * up + pad[0] + pad[1] - kw) // down + 1
  y = torch.empty(n, c, out_h, out_w, dtype=torch.float32, device=input.device)
  karr = (ctypes.c_float * (kh * kw))(*k.reshape(-1).tolist())
  with torch.cuda.device(input.device):
    _lib.call('b200_upfirdn2d_f32', _lib.ptr(x), karr, _lib.ptr(y), n * c, in_h, in_w, 1, kh, kw,
              up, up, down, down, pad[0], pad[1], pad[0], pad[1], _lib.stream_ptr(input.device))
  return y


def upfirdn2d_nhwc(input, kernel, up=1, down=1, pad=(0, 0)):
  """Same op on a channels-last ``[N, H, W, C]`` tensor (the engine's internal layout)."""
  n, in_h, in_w, c = input.shape
  k = kernel.detach().to('cpu', torch.float32).contiguous()
  kh, kw = k.shape
  x = input.detach().to(torch.float32).contiguous()
  out_h = (in_h * up + pad[0] + pad[1] - kh) // down + 1
  out_w = (in_w * up + pad[0] + pad[1] - kw) // down + 1
  y = torch.empty(n, out_h, out_w, c, dtype=torch.float32, device=input.device)
  karr = (ctypes.c_float * (kh * kw))(*k.reshape(-1).tolist())
  with torch.cuda.device(input.device):
    _lib.call('b200_upfirdn2d_f32', _lib.ptr(x), karr, _lib.ptr(y), n, in_h, in_w, c, kh, kw,
              up, up, down, down, pad[0], pad[1], pad[0], pad[1], _lib.stream_ptr(input.device))
  return y
