"""Device-resident probability-flow ODE sampler (successor of ``sampling.py:414-485``'s host-side scipy loop).

``scipy.integrate.solve_ivp(method='RK45')`` keeps the state in a float64 numpy array and calls the right-hand side
with it: in the reference every function evaluation moves the whole batch host -> device -> host.  Here the float64
state, the seven Dormand-Prince stage derivatives and every stage / error sum live in HBM (``csrc/ode.cu`` behind
``b200_ode_*``); this module is scipy's step-size CONTROLLER restated on Python floats (IEEE doubles, like numpy's):
``select_initial_step`` (``scipy/integrate/_ivp/common.py``), ``RungeKutta._step_impl`` and ``rk_step``
(``_ivp/rk.py``), and ``solve_ivp``'s outer loop without events / dense output.  One double per attempted step
crosses PCIe (the sum of squares behind the error norm); the accepted/rejected decision is taken on the host exactly as
scipy takes it, so the trajectory of step sizes - and ``nfev`` - follow scipy's for the same right-hand side.

:class:`DormandPrince45` is arithmetic-agnostic: it drives an ``ops`` object (``CudaOdeOps`` below; the CPU tests
plug a numpy one in to compare the controller with scipy itself).
"""
import ctypes
import math

import torch

from . import _lib

# Dormand-Prince 5(4) tableau (scipy/integrate/_ivp/rk.py: class RK45)
C = (0.0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0)
A = ((),
     (1 / 5,),
     (3 / 40, 9 / 40),
     (44 / 45, -56 / 15, 32 / 9),
     (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
     (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
B = (35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84)
E = (-71 / 57600, 0.0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40)
ORDER, ERROR_ESTIMATOR_ORDER, N_STAGES = 5, 4, 6
SAFETY, MIN_FACTOR, MAX_FACTOR = 0.9, 0.2, 10.0     # _ivp/rk.py module constants


class DormandPrince45:
  """``ops`` provides (all state stays wherever ``ops`` keeps it; K[j] is stage-derivative slot j of 7):
      ops.n                                  number of state elements
      ops.rhs(t, coefs, h, slot, keep_y)     K[slot] = f(t, y + h * sum_j coefs[j] K[j]); keep_y: also store that state as y_new
      ops.error_sumsq(h, rtol, atol)         sum(((h * sum_j E[j] K[j]) / (atol + max(|y|, |y_new|) * rtol))**2)
      ops.scaled_sumsq(slot, minus, rtol, atol)   sum(((K[slot] - K[minus]) / (atol + |y| * rtol))**2), slot=-1: y itself
      ops.accept()                           y <- y_new, K[0] <- K[6]
  """

  def __init__(self, ops, t0, t_bound, rtol=1e-5, atol=1e-5, max_step=math.inf):
    self.ops, self.t, self.t_bound = ops, float(t0), float(t_bound)
    self.rtol, self.atol, self.max_step = float(rtol), float(atol), max_step
    self.direction = (1.0 if t_bound > t0 else -1.0) if t_bound != t0 else 1.0
    self.error_exponent = -1 / (ERROR_ESTIMATOR_ORDER + 1)
    self.nfev = 0
    self.n_accepted = self.n_rejected = 0
    self._rhs(self.t, (), 0.0, 0, False)                      # self.f = self.fun(self.t, self.y)
    self.h_abs = self._select_initial_step()

  def _rhs(self, t, coefs, h, slot, keep_y):
    self.nfev += 1
    self.ops.rhs(t, coefs, h, slot, keep_y)

  def _norm(self, sumsq):
    return math.sqrt(sumsq) / math.sqrt(self.ops.n)           # np.linalg.norm(x) / x.size ** 0.5

  def _select_initial_step(self):
    """scipy/integrate/_ivp/common.py:select_initial_step (order = error_estimator_order)."""
    interval_length = abs(self.t_bound - self.t)
    if interval_length == 0.0:
      return 0.0
    d0 = self._norm(self.ops.scaled_sumsq(-1, None, self.rtol, self.atol))
    d1 = self._norm(self.ops.scaled_sumsq(0, None, self.rtol, self.atol))
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval_length)
    self._rhs(self.t + h0 * self.direction, (1.0,), h0 * self.direction, 1, False)      # f1 = fun(t0 + h0*dir, y0 + h0*dir*f0)
    d2 = self._norm(self.ops.scaled_sumsq(1, 0, self.rtol, self.atol)) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
      h1 = max(1e-6, h0 * 1e-3)
    else:
      h1 = (0.01 / max(d1, d2)) ** (1 / (ERROR_ESTIMATOR_ORDER + 1))
    return min(100 * h0, h1, interval_length, self.max_step)

  def _step(self):
    """RungeKutta._step_impl + rk_step.  Returns False if the step size underflowed."""
    t = self.t
    min_step = 10 * abs(math.nextafter(t, self.direction * math.inf) - t)
    if self.h_abs > self.max_step:
      h_abs = self.max_step
    elif self.h_abs < min_step:
      h_abs = min_step
    else:
      h_abs = self.h_abs
    step_accepted = step_rejected = False
    while not step_accepted:
      if h_abs < min_step:
        return False
      h = h_abs * self.direction
      t_new = t + h
      if self.direction * (t_new - self.t_bound) > 0:
        t_new = self.t_bound
      h = t_new - t
      h_abs = abs(h)
      for s in range(1, N_STAGES):                             # rk_step: K[s] = fun(t + c*h, y + (K[:s].T @ a[:s]) * h)
        self._rhs(t + C[s] * h, A[s], h, s, False)
      self._rhs(t + h, B, h, N_STAGES, True)                   # y_new = y + h * K[:-1].T @ B; K[-1] = fun(t + h, y_new)
      error_norm = self._norm(self.ops.error_sumsq(h, self.rtol, self.atol))
      if error_norm < 1:
        factor = MAX_FACTOR if error_norm == 0 else min(MAX_FACTOR, SAFETY * error_norm ** self.error_exponent)
        if step_rejected:
          factor = min(1, factor)
        h_abs *= factor
        step_accepted = True
        self.n_accepted += 1
      else:
        h_abs *= max(MIN_FACTOR, SAFETY * error_norm ** self.error_exponent)
        step_rejected = True
        self.n_rejected += 1
    self.t = t_new
    self.ops.accept()
    self.h_abs = h_abs
    return True

  def solve(self):
    """solve_ivp's loop (no events, no dense output): step until t_bound.  Returns nfev."""
    while self.direction * (self.t - self.t_bound) < 0:
      if not self._step():
        raise RuntimeError('RK45: required step size is less than spacing between numbers')   # solve_ivp status -1
    return self.nfev


class CudaOdeOps:
  """The float64 state ``y`` / ``y_new``, the stage derivatives ``K[7][n]`` and the float32 network input on the device.
  ``drift(t, x32, k_out)`` (given by the sampler) evaluates the network on ``x32`` and writes the float64 drift."""

  def __init__(self, x0, drift):
    self.device = x0.device
    self.shape = tuple(x0.shape)
    self.n = x0.numel()
    self.y = x0.detach().to(torch.float64).reshape(-1).contiguous()
    self.y_new = torch.empty_like(self.y)
    self.K = torch.empty(N_STAGES + 1, self.n, dtype=torch.float64, device=self.device)
    self.x32 = torch.empty(self.shape, dtype=torch.float32, device=self.device)
    self.ws = torch.zeros(int(_lib.load().b200_ode_workspace_doubles()), dtype=torch.float64, device=self.device)
    self.drift = drift
    self.host_reads = 0

  def _coefs(self, coefs):
    arr = (ctypes.c_double * 8)()
    for j, c in enumerate(coefs):
      arr[j] = c
    return arr

  def rhs(self, t, coefs, h, slot, keep_y):
    st = _lib.stream_ptr(self.device)
    _lib.call('b200_ode_stage_f64', _lib.ptr(self.y), _lib.ptr(self.K), self.n, self._coefs(coefs), len(coefs), float(h),
              _lib.ptr(self.y_new) if keep_y else None, _lib.ptr(self.x32), st)
    self.drift(t, self.x32, self.K[slot])

  def _read(self):
    self.host_reads += 1
    return float(self.ws[0].item())          # the one device -> host scalar

  def error_sumsq(self, h, rtol, atol):
    _lib.call('b200_ode_error_sumsq_f64', _lib.ptr(self.y), _lib.ptr(self.y_new), _lib.ptr(self.K), self.n, self._coefs(E),
              len(E), float(h), float(rtol), float(atol), _lib.ptr(self.ws), _lib.stream_ptr(self.device))
    return self._read()

  def scaled_sumsq(self, slot, minus, rtol, atol):
    v = self.y if slot < 0 else self.K[slot]
    v2 = None if minus is None else self.K[minus]
    _lib.call('b200_ode_scaled_sumsq_f64', _lib.ptr(v), _lib.ptr(v2), _lib.ptr(self.y), self.n, float(rtol), float(atol),
              _lib.ptr(self.ws), _lib.stream_ptr(self.device))
    return self._read()

  def accept(self):
    self.y, self.y_new = self.y_new, self.y
    self.K[0].copy_(self.K[N_STAGES])         # FSAL: self.f = f_new

  def state_f32(self):
    return self.y.to(torch.float32).reshape(self.shape)


def engine_drift_fn(sde, model, batch, device):
  """Right-hand side of the probability-flow ODE for the engine-backed network and the stock VE / VP / sub-VP SDEs:
  ``rsde.sde(x, t)[0]`` with ``probability_flow=True`` (``sde_lib.py:93-100``) over ``get_score_fn(..., continuous=True)``
  (``models/utils.py:129-178``).  The per-evaluation scalars - network label, drift coefficient, g(t)^2, marginal std -
  come from the SDE's own torch ops on a one-element device tensor and stay on the device."""
  from . import sde_lib
  vp_like = isinstance(sde, (sde_lib.VPSDE, sde_lib.subVPSDE))
  one = torch.ones(1, 1, 1, 1, device=device)
  zero = torch.zeros(1, 1, 1, 1, device=device)
  scal = torch.zeros(3, dtype=torch.float32, device=device)

  def drift(t, x32, k_out):
    vec_t = torch.ones(1, device=device) * t                       # `torch.ones(shape[0]) * t`: float32
    f1, g = sde.sde(one, vec_t)                                    # drift of x = 1 (the drift is linear in x), diffusion
    std = sde.marginal_prob(zero, vec_t)[1]
    if vp_like:
      labels = vec_t * 999                                         # models/utils.py:150
      scal[2:3] = std
    else:
      labels = std                                                 # VE: labels = sigma(t) (:167), score = model output
      scal[2] = 0.0
    scal[0:1] = f1.reshape(1)
    scal[1:2] = g ** 2
    out = model(x32, labels.to(torch.float32).expand(batch).contiguous(), labels_uniform=True)
    _lib.call('b200_ode_drift_f64', _lib.ptr(x32), _lib.ptr(out), x32.numel(), _lib.ptr(scal), _lib.ptr(k_out),
              _lib.stream_ptr(device))

  return drift
