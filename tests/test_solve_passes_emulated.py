"""CPU: the control flow of DeviceModel.solve (residual-replacement passes) on a stub whose "Krylov solve" is an exact
dense solve with a prescribed relative error and whose "double-double residual" is numpy longdouble -- the Python
orchestration that decides how many passes run, with which tolerances, and what it does when a correction pass does
not converge.  (The kernels themselves are exercised by the -m gpu tests.)"""
import ctypes as C
import numpy as np
import pytest
import torch

from goldfish_b200 import _capi as capi
from goldfish_b200.device_model import DeviceModel


class _Lib:
    def __init__(self, owner):
        self.o = owner

    def gf_residual_dd(self, K, dist, x, b, r, stream):
        o = self.o
        res = o.b_np.astype(np.longdouble) - o.A.astype(np.longdouble) @ o.x_t.numpy().astype(np.longdouble)
        o._w_res.copy_(torch.from_numpy(np.asarray(res, dtype=np.float64)))
        return 0

    def gf_last_error(self):
        return b""


class Stub:
    """Just enough of DeviceModel for DeviceModel.solve(self, ...)."""

    def __init__(self, n=60, kappa=1e10, err=1e-6, polish=False, fail_corrections=False):
        rng = np.random.default_rng(0)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        self.A = (Q * np.logspace(0, np.log10(kappa), n)) @ Q.T
        self.err, self.fail = err, fail_corrections
        self.true_rtol, self.pass_rtol, self.max_refine, self.polish, self.krylov_rtol = 1e-8, 1e-6, 3, polish, 1e-11
        self._sw_factored, self.eager_refactor, self._fact_version, self._K_version = True, False, 0, 0
        self._w_res = self._w_cor = None
        self.auto_refresh_coarse = False
        self.lib = _Lib(self)
        self.calls = []
        self.K = type("K", (), {"c_struct": lambda s: C.c_int(0)})()

    def _dist_struct(self): return C.c_int(0)
    def _stream(self): return None
    def dot(self, a, b): return float((a * b).sum())
    def axpby(self, a, x, b, y): y.mul_(b).add_(x, alpha=a); return y

    def _krylov(self, b, x, rtol, max_it=None):
        self.calls.append(rtol)
        if self.fail and len(self.calls) > 1:
            raise capi.GoldfishNotConverged("gf_pcg: tolerance not reached")
        exact = np.linalg.solve(self.A, b.numpy())
        pert = np.random.default_rng(len(self.calls)).standard_normal(len(exact))
        x.copy_(torch.from_numpy(exact + rtol * np.linalg.norm(exact) * pert / np.linalg.norm(pert) * 0.1))
        return 10, rtol


def _solve(stub, b):
    stub.b_np = b.numpy().copy()
    stub.x_t = torch.zeros_like(b)
    return DeviceModel.solve(stub, b, stub.x_t)


def test_passes_reach_the_true_residual_target():
    s = Stub(kappa=1e6)
    b = torch.from_numpy(np.random.default_rng(1).standard_normal(60))
    x = _solve(s, b)
    assert s.last_true_relres <= 1e-8
    assert s.calls[0] == 1e-6 and 2 <= len(s.calls) <= 4 and all(1e-9 <= c <= 1e-1 for c in s.calls[1:])
    assert np.linalg.norm(s.A @ x.numpy() - b.numpy()) <= 1e-8 * np.linalg.norm(b.numpy())
    assert s.last_krylov_its == 10 * len(s.calls)


def test_small_systems_get_two_correction_passes():
    s = Stub(kappa=1e3, polish=True)
    _solve(s, torch.from_numpy(np.random.default_rng(2).standard_normal(60)))
    assert len(s.calls) == 3 and s.calls[2] <= 1e-2          # k = 0, 1 always run a correction on small systems


def test_a_correction_that_does_not_converge_keeps_the_iterate():
    s = Stub(kappa=1e6, fail_corrections=True)
    b = torch.from_numpy(np.random.default_rng(3).standard_normal(60))
    x = _solve(s, b)                                        # no exception: the pass-1 iterate is returned
    assert len(s.calls) == 2 and s.last_true_relres is not None and s.last_true_relres > 1e-8
    assert np.isfinite(x.numpy()).all()


def test_zero_right_hand_side_and_explicit_tolerance():
    s = Stub()
    x = _solve(s, torch.zeros(60, dtype=torch.float64))
    assert s.last_true_relres == 0.0 and float(x.abs().max()) == 0.0
    s2 = Stub()
    s2.b_np = np.ones(60); s2.x_t = torch.zeros(60, dtype=torch.float64)
    DeviceModel.solve(s2, torch.ones(60, dtype=torch.float64), s2.x_t, rtol=1e-4)     # explicit rtol: one plain pass
    assert s2.calls == [1e-4] and s2.last_true_relres is None


class NewtonStub:
    """DeviceModel.newton on a small cubic-spring system R(u) = A u + c u^3 - f with a noise floor in R."""

    def __init__(self, noise=0.0):
        rng = np.random.default_rng(0)
        n = 20
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        self.A = (Q * np.linspace(1, 50, n)) @ Q.T
        self.f = rng.standard_normal(n) * 40.0
        self.c, self.noise = 5.0, noise
        self.u = torch.zeros(n, dtype=torch.float64); self.R = torch.zeros(n, dtype=torch.float64)
        self._epochs, self.state_epoch = {}, 0
        self.last_krylov_its, self.last_true_relres, self.n_asm = 7, 1e-9, 0

    def touch(self): self.state_epoch += 1
    def dot(self, a, b): return float((a * b).sum())
    def axpby(self, a, x, b, y): y.mul_(b).add_(x, alpha=a); return y

    def assemble(self, residual=False, tangent=False, functionals=False):
        u = self.u.numpy(); self.n_asm += 1
        r = self.A @ u + self.c * u ** 3 - self.f
        r = r + self.noise * np.random.default_rng(self.n_asm).standard_normal(len(r))
        self.R.copy_(torch.from_numpy(r))
        self.Kd = self.A + np.diag(3 * self.c * u ** 2)

    def solve(self, rhs, du, refactor=None):
        du.copy_(torch.from_numpy(np.linalg.solve(self.Kd, rhs.numpy())))
        return du


def test_newton_loop_matches_the_reference_stopping_rule():
    s = NewtonStub()
    DeviceModel.newton(s, max_it=30, rtol=1e-3)
    h = s.newton_history
    assert h[0] == 1.0 and h[-1] < 1e-3 and all(v >= 1e-3 for v in h[1:-1])      # |R|/|R0| < rtol, checked before a solve
    assert len(s.newton_krylov_its) == len(h) - 1 and s.newton_stagnated is False
    with pytest.raises(capi.GoldfishNotConverged):
        DeviceModel.newton(NewtonStub(), max_it=1, rtol=1e-12)


def test_newton_accepts_the_fp64_floor_only_when_asked():
    s = NewtonStub(noise=2e-6)               # residual evaluation floor ~1e-8 |R0|
    DeviceModel.newton(s, max_it=30, rtol=1e-12, accept_stagnation=True)
    assert s.newton_stagnated and s.newton_history[-1] < 1e-5 and len(s.newton_history) < 12
    s2 = NewtonStub(noise=2e-6)
    with pytest.raises(capi.GoldfishNotConverged):
        DeviceModel.newton(s2, max_it=12, rtol=1e-12)
    assert len(s2.newton_history) == 13      # history kept on failure
