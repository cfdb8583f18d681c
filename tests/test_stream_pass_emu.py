"""Host emulation of the fused streaming kernel (tests/emu/sp_emu.cpp compiles the SAME per-thread
step code as the CUDA kernel, csrc/stream_pass_body.cuh) against the oracle: ring/pipeline/halo
index logic, checked bit-for-bit, with NaN-poisoned shared memory and three thread orders."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "libspemu.so")
    src = [os.path.join(EMU_DIR, "sp_emu.cpp"),
           os.path.join(ROOT, "hpcclassmultigridproject_b200", "csrc", "stream_pass_body.cuh"),
           os.path.join(ROOT, "hpcclassmultigridproject_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                        "-I/usr/local/cuda/include", "-o", so, src[0]], check=True)
    lib = C.CDLL(so)
    lib.sp_emu_run.restype = C.c_long
    lib.sp_emu_run.argtypes = [C.c_long] * 5 + [_dp] * 8 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int] * 3 + [C.c_long] * 6
    return lib


def layout(n):
    odd = (n // 2 + 1 + 15) // 16 * 16 + 32          # split_layout() of common.cuh
    return 2 * odd, odd


def to_split(a, fill=np.nan):
    n = a.shape[0] - 1
    pitch, odd = layout(n)
    s = np.full((n + 1, pitch), fill)
    s[:, : n // 2 + 1] = a[:, 0::2]
    s[:, odd: odd + n // 2] = a[:, 1::2]
    return s


def from_split(s, n):
    pitch, odd = layout(n)
    a = np.empty((n + 1, n + 1))
    a[:, 0::2] = s[:, : n // 2 + 1]
    a[:, 1::2] = s[:, odd: odd + n // 2]
    return a


def ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def run(emu, n, u, rhs, v1, v2, K, post, arith, dt, nu, dx, cu=None, wk=0, nbands=0, order=0):
    pitch, odd = layout(n)
    cp, co = layout(n // 2)
    us = None if u is None else to_split(u)
    out = np.full((n + 1, pitch), np.nan)
    crhs = np.zeros((n // 2 + 1, cp))
    partials = np.zeros(4096)
    cus = None if cu is None else to_split(cu)
    nt = emu.sp_emu_run(n, pitch, odd, cp, co, ptr(us), ptr(out), ptr(to_split(rhs)), ptr(to_split(v1)),
                        ptr(to_split(v2)), ptr(cus), ptr(crhs), ptr(partials), K, post, arith, dt, nu, dx, wk, nbands,
                        order, 0, 0, 0, 0, 0, 0)
    assert nt > 0
    return from_split(out, n), from_split(crhs, n // 2), partials[:nt]


def fields(n, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((n + 1, n + 1)) for _ in range(4)]


GEOMS = [(0, 0), (24, 1), (24, 3), (40, 2), (120, 5)]      # (owned pairs per strip = SWK - 8, bands)


@pytest.mark.parametrize("n", [32, 64, 200, 256])
@pytest.mark.parametrize("K", [1, 2, 3])
def test_down_leg(emu, oracle, n, K):
    """K RB iterations + residual + injection in one pass == oracle operator sequence, bitwise"""
    u, rhs, v1, v2 = fields(n, 7 * n + K)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want_u = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, K)
    want_c = oracle.restriction(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n)
    for (wk, nb), order in zip(GEOMS, (0, 1, 2, 0, 2)):
        got_u, got_c, _ = run(emu, n, u, rhs, v1, v2, K, 1, 1, dt, nu, dx, wk=wk, nbands=nb, order=order)
        assert np.array_equal(got_u, want_u), (wk, nb)
        assert np.array_equal(got_c[1:-1, 1:-1], want_c[1:-1, 1:-1]), (wk, nb)
        assert not got_c[0, :].any() and not got_c[:, 0].any()       # coarse rhs boundary untouched


@pytest.mark.parametrize("n", [32, 64, 200])
@pytest.mark.parametrize("K", [0, 1, 3])
def test_up_leg(emu, oracle, n, K):
    """u += P(coarse) ; K RB iterations ; sum of squares of the residual"""
    u, rhs, v1, v2 = fields(n, 11 * n + K)
    cu = np.random.default_rng(n).standard_normal((n // 2 + 1, n // 2 + 1))
    cu[0, :] = cu[-1, :] = 0; cu[:, 0] = cu[:, -1] = 0           # a coarse correction has a zero boundary
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want_u = u + oracle.prolongation(cu, n // 2)
    want_u = oracle.gauss_seidel(want_u, rhs, n, v1, v2, dt, nu, dx, K)
    want_r2 = oracle.norm(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n) ** 2
    for (wk, nb), order in zip(GEOMS, (2, 0, 1, 2, 0)):
        got_u, _, parts = run(emu, n, u, rhs, v1, v2, K, 2, 1, dt, nu, dx, cu=cu, wk=wk, nbands=nb, order=order)
        assert np.array_equal(got_u, want_u), (wk, nb)
        assert abs(parts.sum() - want_r2) <= 1e-12 * want_r2


def test_zero_input_and_fast_arithmetic(emu, oracle):
    n = 128
    _, rhs, v1, v2 = fields(n, 5)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want = oracle.gauss_seidel(np.zeros((n + 1, n + 1)), rhs, n, v1, v2, dt, nu, dx, 3)
    got, _, _ = run(emu, n, None, rhs, v1, v2, 3, 1, 1, dt, nu, dx)          # u_in == NULL: u is zero
    assert np.array_equal(got, want)
    gotf, _, _ = run(emu, n, None, rhs, v1, v2, 3, 1, 0, dt, nu, dx, wk=40, nbands=3, order=2)
    assert np.linalg.norm(gotf - want) <= 1e-13 * np.linalg.norm(want)


def test_residual_only_pass(emu, oracle):
    """K = 0 and no prolongation: nothing is written to u_out, only the residual epilogue runs"""
    n = 64
    u, rhs, v1, v2 = fields(n, 9)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want = oracle.norm(oracle.residual(u, rhs, n, v1, v2, dt, nu, dx), n) ** 2
    got_u, _, parts = run(emu, n, u, rhs, v1, v2, 0, 2, 1, dt, nu, dx, wk=24, nbands=2)
    assert np.isnan(got_u).all()
    assert abs(parts.sum() - want) <= 1e-12 * want
