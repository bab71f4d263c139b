"""CPU test of the device ODE sampler's host logic: the step-size controller (score_sde_pytorch_b200/ode.py) is
scipy's, so driving it with a numpy `ops` must reproduce scipy.integrate.solve_ivp(method='RK45') itself - same number
of function evaluations, same accepted/rejected sequence, same final state to float64 round-off."""
import numpy as np
import pytest

from score_sde_pytorch_b200 import ode as O


class NumpyOps:
  def __init__(self, y0, fun):
    self.y = np.asarray(y0, dtype=np.float64).copy()
    self.n = self.y.size
    self.y_new = np.empty_like(self.y)
    self.K = np.zeros((O.N_STAGES + 1, self.n))
    self.fun = fun

  def rhs(self, t, coefs, h, slot, keep_y):
    s = len(coefs)
    ys = self.y + np.dot(self.K[:s].T, np.asarray(coefs)) * h if s else self.y
    if keep_y:
      self.y_new = ys.copy()
    self.K[slot] = self.fun(t, ys)

  def error_sumsq(self, h, rtol, atol):
    scale = atol + np.maximum(np.abs(self.y), np.abs(self.y_new)) * rtol
    return float(np.sum((np.dot(self.K.T, np.asarray(O.E)) * h / scale) ** 2))

  def scaled_sumsq(self, slot, minus, rtol, atol):
    v = self.y if slot < 0 else self.K[slot]
    if minus is not None:
      v = v - self.K[minus]
    return float(np.sum((v / (atol + np.abs(self.y) * rtol)) ** 2))

  def accept(self):
    self.y, self.y_new = self.y_new, self.y
    self.K[0] = self.K[O.N_STAGES]


@pytest.mark.parametrize('span', [(1.0, 1e-3), (0.0, 2.5)])
@pytest.mark.parametrize('tol', [1e-5, 1e-3])
def test_rk45_controller_reproduces_scipy_solve_ivp(span, tol):
  from scipy import integrate
  rng = np.random.default_rng(3)
  n = 257
  w = rng.normal(size=n)
  y0 = rng.normal(size=n).astype(np.float32)        # the reference hands solve_ivp a float32 array (to_flattened_numpy)

  def fun(t, y):                                    # float32 right-hand side widened to float64, like the reference's ode_func
    x = y.astype(np.float32)
    return (-(1.5 + np.float32(t)) * x + np.sin(3 * x + w.astype(np.float32)) * np.float32(4.0)).astype(np.float32).astype(np.float64)

  sol = integrate.solve_ivp(fun, span, y0, rtol=tol, atol=tol, method='RK45')
  assert sol.status == 0
  ops = NumpyOps(y0, fun)
  solver = O.DormandPrince45(ops, span[0], span[1], rtol=tol, atol=tol)
  nfev = solver.solve()
  assert nfev == sol.nfev
  assert solver.n_accepted == len(sol.t) - 1
  np.testing.assert_allclose(ops.y, sol.y[:, -1], rtol=1e-11, atol=1e-13)
