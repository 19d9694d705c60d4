"""CPU: the device point mathematics (goldfish_b200/csrc/kl_point.cuh), built
for the host with g++, against the oracle's jets: value, hand-derived first
variation and dual-number directional second derivatives."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from oracle.jet import Jet
from oracle.kl_shell import shell_energy_density
from oracle.penalty import penalty_point_energy

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pm") / "libpointmath.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "hostmath", "pointmath_host.cpp")])
    return C.CDLL(so)


dp = C.POINTER(C.c_double)
P = lambda a: a.ctypes.data_as(dp)


def test_shell_point(lib):
    rng = np.random.default_rng(0)
    n = 6
    GX = rng.standard_normal((n, 15)); Gu = 0.3 * rng.standard_normal((n, 15)); t = 0.1 + 0.1 * rng.random(n)
    E, nu = 1e3, 0.3
    V = Jet.variables(np.concatenate([GX, Gu, t[:, None]], 1))
    e, J, _, _ = shell_energy_density([V[3 * k:3 * k + 3] for k in range(5)], [V[15 + 3 * k:18 + 3 * k] for k in range(5)], V[30], E, nu)
    lib.gf_test_shell_point.argtypes = [dp, dp, C.c_double, C.c_double, C.c_double, dp] + [dp] * 6
    for q in range(n):
        d = rng.standard_normal(31)
        oe, oJ, og, ode, odJ, odg = np.zeros(1), np.zeros(1), np.zeros(15), np.zeros(1), np.zeros(1), np.zeros(15)
        lib.gf_test_shell_point(P(GX[q].copy()), P(Gu[q].copy()), t[q], E, nu, P(d), P(oe), P(oJ), P(og), P(ode), P(odJ), P(odg))
        assert abs(oe[0] - e.v[q]) < 1e-13 * abs(e.v[q])
        assert np.abs(og - e.g[q, 15:30]).max() < 1e-13 * np.abs(og).max()
        assert abs(ode[0] - e.g[q] @ d) < 1e-12 * max(abs(ode[0]), 1.0)
        assert np.abs(odg - e.h[q, 15:30, :] @ d).max() < 1e-12 * np.abs(odg).max()
        assert abs(odJ[0] - J.g[q] @ d) < 1e-13 * max(1.0, abs(odJ[0]))


def test_penalty_point(lib):
    rng = np.random.default_rng(1)
    n = 6
    uv = 0.2 * rng.standard_normal((n, 18)); Xv = rng.standard_normal((n, 18))
    tp = rng.standard_normal((n, 2)); tp /= np.linalg.norm(tp, axis=1)[:, None]
    ad = 3.0 + rng.random(n); ar = 2.0 + rng.random(n)
    V = Jet.variables(np.concatenate([uv, Xv], 1)); X = V[18:]
    e = penalty_point_energy(V[0:3], [V[3:6], V[6:9]], V[9:12], [V[12:15], V[15:18]], X[0:3], X[3:6],
                             [X[6:9], X[9:12]], [X[12:15], X[15:18]], tp, ad, ar)
    lib.gf_test_penalty_point.argtypes = [dp, dp, dp, C.c_double, C.c_double, dp] + [dp] * 4
    for q in range(n):
        d = rng.standard_normal(36)
        oe, og, ode, odg = np.zeros(1), np.zeros(18), np.zeros(1), np.zeros(18)
        lib.gf_test_penalty_point(P(uv[q].copy()), P(Xv[q].copy()), P(tp[q].copy()), ad[q], ar[q], P(d), P(oe), P(og), P(ode), P(odg))
        assert abs(oe[0] - e.v[q]) < 1e-13 * abs(e.v[q])
        assert np.abs(og - e.g[q, :18]).max() < 1e-13 * np.abs(og).max()
        assert np.abs(odg - e.h[q, :18, :] @ d).max() < 1e-12 * np.abs(odg).max()
