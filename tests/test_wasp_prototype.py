"""CPU prototype of the round-2 kernel candidate ("warp-autonomous" streaming pass)
against the oracle, bit for bit: lag-1 in-order half-sweeps on a private row window, recomputed
strip / band halos, epilogue one row behind the store.  Prototype only: nothing in the product uses it."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def wasp():
    src = os.path.join(HERE, "emu", "wasp_proto.cpp")
    out = os.path.join(HERE, "emu", "libwasp.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", out, src], check=True, env=env)
    L = C.CDLL(out)
    dp = C.c_void_p
    L.wasp_pass.argtypes = [C.c_long, dp, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_double), C.c_int, C.c_int,
                            C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_long]
    return L


def ptr(a):
    return None if a is None else a.ctypes.data


def fields(oracle, n, seed):
    rng = np.random.default_rng(seed)
    u0, v1, v2 = oracle.initial_conditions(n, 3.0)
    u = u0 + 0.01 * rng.standard_normal(u0.shape)
    u[0, :] = u[-1, :] = u[:, 0] = u[:, -1] = 0.0
    return u, v1, v2


@pytest.mark.parametrize("n,K,wp,hp,rb", [(32, 3, 5, 4, 7), (64, 3, 56, 4, 64), (64, 2, 9, 3, 10), (128, 1, 20, 2, 33), (64, 3, 8, 4, 3)])
def test_down_leg_matches_oracle(wasp, oracle, n, K, wp, hp, rb):
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    u, v1, v2 = fields(oracle, n, n + K)
    rhs = oracle.compute_rhs(u, n, v1, v2, dt, nu, dx)
    want = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, iters=K)
    res = oracle.residual(want, rhs, n, v1, v2, dt, nu, dx)
    want_c = oracle.restriction(res, n)
    got = np.full_like(u, np.nan)
    crhs = np.zeros((n // 2 + 1, n // 2 + 1))
    rc = wasp.wasp_pass(n, ptr(u), ptr(got), ptr(rhs), ptr(v1), ptr(v2), None, ptr(crhs), None, K, 1, dt, nu, dx, wp, hp, rb)
    assert rc == 0
    assert np.array_equal(got, want)
    assert np.array_equal(crhs[1:-1, 1:-1], want_c[1:-1, 1:-1])


@pytest.mark.parametrize("n,K,wp,hp,rb", [(32, 3, 6, 4, 9), (64, 3, 56, 4, 64), (128, 2, 11, 3, 17)])
def test_up_leg_matches_oracle(wasp, oracle, n, K, wp, hp, rb):
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    u, v1, v2 = fields(oracle, n, 7 * n + K)
    rhs = oracle.compute_rhs(u, n, v1, v2, dt, nu, dx)
    rng = np.random.default_rng(n)
    cu = 1e-3 * rng.standard_normal((n // 2 + 1, n // 2 + 1))
    cu[0, :] = cu[-1, :] = cu[:, 0] = cu[:, -1] = 0.0
    start = u + oracle.prolongation(cu, n // 2)                     # multigrid.cpp:81-83
    want = oracle.gauss_seidel(start.copy(), rhs, n, v1, v2, dt, nu, dx, iters=K)
    res = oracle.residual(want, rhs, n, v1, v2, dt, nu, dx)
    want_sumsq = float(np.sum(res[1:-1, 1:-1] ** 2))
    got = np.full_like(u, np.nan)
    sumsq = C.c_double(0.0)
    rc = wasp.wasp_pass(n, ptr(u), ptr(got), ptr(rhs), ptr(v1), ptr(v2), ptr(cu), None, C.byref(sumsq), K, 2, dt, nu, dx, wp, hp, rb)
    assert rc == 0
    assert np.array_equal(got, want)
    assert abs(sumsq.value - want_sumsq) <= 1e-12 * want_sumsq


def test_geometry_checks(wasp, oracle):
    n = 32; dx = 1.0 / n; dt = dx / 10
    u, v1, v2 = fields(oracle, n, 1)
    out = np.zeros_like(u)
    # halo columns too narrow for 2K half-sweeps + the epilogue's neighbour
    assert wasp.wasp_pass(n, ptr(u), ptr(out), ptr(u), ptr(v1), ptr(v2), None, None, None, 3, 0, dt, -4e-4, dx, 8, 3, 8) == -1
