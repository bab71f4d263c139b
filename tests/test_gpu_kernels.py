"""`-m gpu`: each HBM-bound device kernel and the strict-fp32 CUDA-core contraction, called through
the C ABI, against the oracle / plain torch fp32 on the same seeded inputs."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import golden
from oracle import ncsnpp_oracle as NO

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
  import gpu_util
  gpu_util.strict_fp32()
  return torch.device('cuda:0')


@pytest.mark.parametrize('name', ['down2', 'up2', 'pad22', 'generic'])
def test_upfirdn2d_matches_reference_golden(dev, name):
  from score_sde_pytorch_b200.op import upfirdn2d
  g = golden('upfirdn2d.npz')
  up, down, p0, p1 = [int(v) for v in g[name + '_p']]
  y = upfirdn2d(torch.from_numpy(g[name + '_x']).to(dev), torch.from_numpy(g[name + '_k']), up=up, down=down, pad=(p0, p1))
  ref = torch.from_numpy(g[name + '_y']).to(dev)
  assert y.shape == ref.shape
  assert (y - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('mode', ['down2', 'up2', 'pad22'])
@pytest.mark.parametrize('shape', [(2, 128, 32, 32), (3, 12, 5, 7), (1, 3, 33, 33)])
def test_upfirdn2d_nhwc_matches_oracle(dev, mode, shape):
  from score_sde_pytorch_b200.op.upfirdn2d import upfirdn2d_nhwc, upfirdn2d
  k = NO.setup_kernel([1, 3, 3, 1])
  up, down, pad = {'down2': (1, 2, (1, 1)), 'up2': (2, 1, (2, 1)), 'pad22': (1, 1, (2, 2))}[mode]
  kk = torch.tensor(k * (4 if mode == 'up2' else 1))
  torch.manual_seed(0)
  x = torch.randn(*shape)
  ref = NO.upfirdn2d_native(x, kk, up=up, down=down, pad=pad).to(dev)
  y = upfirdn2d_nhwc(x.permute(0, 2, 3, 1).contiguous().to(dev), kk, up=up, down=down, pad=pad).permute(0, 3, 1, 2)
  assert (y - ref).abs().max().item() < 1e-5
  y2 = upfirdn2d(x.to(dev), kk, up=up, down=down, pad=pad)
  assert (y2 - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize('up,down', [(1, 1), (1, 2), (2, 1)])
def test_upfirdn2d_planar_kernel_asymmetric_taps_and_ragged_shapes(dev, up, down):
  """The planar 4x4 path (the reference's [N*C, H, W, 1] layout, four outputs along W per thread) with a NON-symmetric FIR
  (a flipped or transposed tap order would show), every pad0 in 0..3, odd sizes, output widths that are not multiples of
  four and a 1-pixel-wide plane."""
  from score_sde_pytorch_b200.op import upfirdn2d
  g = torch.Generator().manual_seed(5)
  kk = torch.rand(4, 4, generator=g) - 0.3
  for shape in ((2, 3, 9, 13), (1, 5, 16, 6), (3, 1, 7, 1), (1, 2, 32, 34)):
    x = torch.randn(*shape, generator=g)
    for p0 in range(4):
      for p1 in (0, 1, 3):
        if (shape[2] * up + p0 + p1 - 4) // down + 1 <= 0 or (shape[3] * up + p0 + p1 - 4) // down + 1 <= 0:
          continue
        ref = NO.upfirdn2d_native(x, kk, up=up, down=down, pad=(p0, p1)).to(dev)
        y = upfirdn2d(x.to(dev), kk, up=up, down=down, pad=(p0, p1))
        assert y.shape == ref.shape, (shape, p0, p1)
        assert (y - ref).abs().max().item() < 1e-5, (shape, up, down, p0, p1)


def test_upfirdn2d_rejects_cpu_and_bad_kernel(dev):
  from score_sde_pytorch_b200.op import upfirdn2d
  with pytest.raises(RuntimeError):
    upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
  with pytest.raises(RuntimeError):
    upfirdn2d(torch.zeros(1, 1, 4, 4, device=dev), torch.ones(9, 9))   # > 64 taps


def test_fused_leaky_relu_matches_oracle(dev):
  from score_sde_pytorch_b200.op import fused_leaky_relu, FusedLeakyReLU
  torch.manual_seed(1)
  x = torch.randn(3, 5, 4, 6)
  b = torch.randn(5)
  ref = NO.fused_leaky_relu(x, b).to(dev)
  y = fused_leaky_relu(x.to(dev), b.to(dev))
  assert torch.allclose(y, ref, rtol=1e-6, atol=1e-6)
  m = FusedLeakyReLU(5).to(dev)
  assert m(x.to(dev)).shape == x.shape
  assert fused_leaky_relu(torch.zeros(0, 5, device=dev), b.to(dev)).numel() == 0   # empty input


@pytest.mark.parametrize('cfg', [(2, 128, 0, 32, 32, True), (3, 256, 128, 16, 16, True), (2, 256, 256, 4, 4, False),
                                 (1, 48, 0, 8, 8, True), (5, 16, 32, 3, 3, True)])
def test_groupnorm_silu_nhwc(dev, cfg):
  import gpu_util
  from score_sde_pytorch_b200 import _lib
  B, C1, C2, H, W, silu = cfg
  C = C1 + C2
  G = min(C // 4, 32)
  torch.manual_seed(2)
  x1 = torch.randn(B, C1, H, W, device=dev) * 3 + 1.5
  x2 = torch.randn(B, C2, H, W, device=dev) * 0.5 - 2 if C2 else None
  gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
  xc = x1 if x2 is None else torch.cat([x1, x2], 1)
  ref = F.group_norm(xc, G, gamma, beta, eps=1e-6)
  if silu:
    ref = F.silu(ref)
  y = torch.empty(B, H, W, C, device=dev)
  raw = torch.empty(B, H, W, C, device=dev)
  stats = torch.empty(B * C + 64, device=dev)      # fp64 quad sums of both sources: 16 B per (image, 4 channels)
  n1 = gpu_util.to_nhwc(x1)                       # keep the NHWC copies alive across the launches
  n2 = gpu_util.to_nhwc(x2) if C2 else None
  _lib.call('b200_groupnorm_nhwc_f32', _lib.ptr(n1), C1, _lib.ptr(n2), C2,
            _lib.ptr(gamma), _lib.ptr(beta), B, H * W, G, 1e-6, int(silu), 0, _lib.ptr(stats), _lib.ptr(y), _lib.ptr(raw),
            _lib.stream_ptr(dev))
  assert torch.allclose(gpu_util.to_nchw(y), ref, rtol=2e-5, atol=2e-5)
  assert torch.equal(gpu_util.to_nchw(raw), xc)
  # Operand modes: values on the TF32 grid (mode 1) / IEEE fp16 (mode 2), each within one rounding of the fp32
  # result (these modes use the approximate-SiLU path, ~1e-6 relative, two orders below the 2^-11 rounding)
  _lib.call('b200_groupnorm_nhwc_f32', _lib.ptr(n1), C1, _lib.ptr(n2), C2,
            _lib.ptr(gamma), _lib.ptr(beta), B, H * W, G, 1e-6, int(silu), 1, _lib.ptr(stats), _lib.ptr(raw), None,
            _lib.stream_ptr(dev))
  assert torch.equal(raw, gpu_util.round_tf32(raw))                       # on the grid
  assert ((raw - y).abs() <= y.abs() * 2.0 ** -11 + 2e-6).all()           # one rounding away
  yh = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
  rawh = torch.empty(B, H, W, C, device=dev, dtype=torch.float16)
  _lib.call('b200_groupnorm_nhwc_f32', _lib.ptr(n1), C1, _lib.ptr(n2), C2,
            _lib.ptr(gamma), _lib.ptr(beta), B, H * W, G, 1e-6, int(silu), 2, _lib.ptr(stats), _lib.ptr(yh), _lib.ptr(rawh),
            _lib.stream_ptr(dev))
  assert ((yh.float() - y).abs() <= y.abs() * 2.0 ** -11 + 2e-6).all()
  assert torch.equal(rawh, gpu_util.to_nhwc(xc).half())                   # the raw copy is exactly the RN fp16 of the input


@pytest.mark.parametrize('T', [16, 256, 100, 1024])
def test_softmax_rows(dev, T):
  from score_sde_pytorch_b200 import _lib
  torch.manual_seed(3)
  s = torch.randn(37, T, device=dev) * 20
  p = torch.empty_like(s)
  _lib.call('b200_softmax_rows_f32', _lib.ptr(s), _lib.ptr(p), 37, T, 1.0 / 16, 0, _lib.stream_ptr(dev))
  ref = torch.softmax(s * (1.0 / 16), dim=-1)
  assert torch.allclose(p, ref, rtol=1e-5, atol=1e-7)
  assert torch.allclose(p.sum(-1), torch.ones(37, device=dev), atol=1e-5)


@pytest.mark.parametrize('numel', [1, 5, 1024, 3 * 32 * 32 * 2, 303104 * 4 + 17, 1024 * 3 * 32 * 32])
def test_randn_equals_torch_cuda_generator(dev, numel):
  """The in-kernel Philox noise must be bit-identical to torch.randn on the CUDA generator (sampling.py:197,275)."""
  import gpu_util
  seed = 1234
  torch.cuda.manual_seed(seed)
  gen = torch.cuda.default_generators[0]
  warm = torch.randn(7, device=dev)          # move the offset off zero
  off0 = gen.get_offset()
  ref = torch.randn(numel, device=dev)
  off1 = gen.get_offset()
  mine, inc = gpu_util.randn_like_torch(numel, seed, off0, dev)
  assert inc == off1 - off0, (inc, off1 - off0)
  assert torch.equal(mine, ref)


@pytest.mark.parametrize('case', [
    dict(B=2, H=32, W=32, C1=3, C2=0, Cout=128, k=3),
    dict(B=2, H=16, W=16, C1=64, C2=32, Cout=48, k=3),
    dict(B=3, H=8, W=8, C1=32, C2=0, Cout=3, k=3),
    dict(B=2, H=4, W=4, C1=96, C2=0, Cout=64, k=1),
    dict(B=1, H=5, W=7, C1=10, C2=0, Cout=9, k=3),
])
def test_conv_fp32_cuda_cores_matches_torch(dev, case):
  import gpu_util
  B, H, W, C1, C2, Cout, k = (case[x] for x in ('B', 'H', 'W', 'C1', 'C2', 'Cout', 'k'))
  torch.manual_seed(4)
  x1 = torch.randn(B, C1, H, W, device=dev)
  x2 = torch.randn(B, C2, H, W, device=dev) if C2 else None
  w = torch.randn(Cout, C1 + C2, k, k, device=dev) / np.sqrt((C1 + C2) * k * k)
  bias = torch.randn(Cout, device=dev)
  rowvec = torch.randn(B, Cout, device=dev)
  res = torch.randn(B, Cout, H, W, device=dev)
  xc = x1 if x2 is None else torch.cat([x1, x2], 1)
  ref = (F.conv2d(xc, w, bias, padding=k // 2) + rowvec[:, :, None, None] + res) * 0.7071067690849304
  wp = gpu_util.pack_conv_weight(w)
  n1, n2, nres = gpu_util.to_nhwc(x1), (gpu_util.to_nhwc(x2) if C2 else None), gpu_util.to_nhwc(res)
  y = gpu_util.conv_nhwc(n1, n2, wp, bias, Cout, k, rowvec=rowvec, rowvec_ld=Cout, residual=nres,
                         scale=0.7071067690849304, impl=0)
  assert torch.allclose(gpu_util.to_nchw(y), ref, rtol=1e-4, atol=1e-4)


def test_gemm_nt_fp32_batched(dev):
  import gpu_util
  torch.manual_seed(5)
  nb, m, n, k = 3, 16, 16, 40
  a = torch.randn(nb * m, k, device=dev)
  w = torch.randn(nb * n, k, device=dev)
  out = gpu_util.gemm_nt(a, w, nb, m, n, k, impl=0)
  ref = torch.bmm(a.view(nb, m, k), w.view(nb, n, k).transpose(1, 2)).reshape(nb * m, n)
  assert torch.allclose(out, ref, rtol=1e-4, atol=1e-4)
  # shared A (weights as the row operand), batched W
  a2 = torch.randn(m, k, device=dev)
  out2 = gpu_util.gemm_nt(a2, w, nb, m, n, k, a_batch_rows=0, impl=0)
  ref2 = torch.einsum('mk,bnk->bmn', a2, w.view(nb, n, k)).reshape(nb * m, n)
  assert torch.allclose(out2, ref2, rtol=1e-4, atol=1e-4)
