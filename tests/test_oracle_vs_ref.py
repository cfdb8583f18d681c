"""The oracle restatement against the compiled UNMODIFIED reference (oracle/_ref/*.so), live.
Skipped where the .so files are absent (they are built in the container that mounts
/root/reference and travel to the GPU box with the snapshot)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle.oracle import Oracle, OracleSolver, Towers, ref_available


@pytest.mark.parametrize("n", [5, 32, 48, 100])
def test_operators_random(oracle, ref, n):
    rng = np.random.default_rng(n)
    u, rhs, v1, v2 = (rng.standard_normal((n + 1, n + 1)) for _ in range(4))
    dx = 1.0 / n; dt = dx / 7; nu = -3e-3
    assert np.array_equal(oracle.compute_rhs(u, n, v1, v2, dt, nu, dx), ref.compute_rhs(u, n, v1, v2, dt, nu, dx))
    a = oracle.residual(u, rhs, n, v1, v2, dt, nu, dx); b = ref.residual(u, rhs, n, v1, v2, dt, nu, dx)
    assert np.array_equal(a, b) and oracle.norm(a, n) == ref.norm(b, n)
    assert np.array_equal(oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 2),
                          ref.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 2))
    assert np.array_equal(oracle.prolongation(u, n), ref.prolongation(u, n))
    if n % 2 == 0:
        assert np.array_equal(oracle.restriction(u, n), ref.restriction(u, n))


@pytest.mark.parametrize("n,shape,vscale,nu", [(64, 1, 1.0, -4e-4), (128, 2, 2.0, -4e-4), (256, 1, 5.0, -1e-5)])
def test_cycle_by_cycle(oracle, ref, n, shape, vscale, nu):
    """mg_inner of the reference vs orc_cycle: every level array after every cycle."""
    u0, v1, v2 = oracle.initial_conditions(n, vscale)
    dx = 1.0 / n; dt = dx / 10
    tw = Towers(ref, n, u0, v1, v2, nu, dt, dx, 1e-12, shape)
    s = OracleSolver(n, u0, v1, v2, nu, dt, dx, 1e-12, shape)
    for l in range(1, s.maxlvl):
        assert np.array_equal(tw.level(tw.v1, l), s.v1(l)) and np.array_equal(tw.level(tw.v2, l), s.v2(l))
    tw.form_rhs(); s.form_rhs()
    for _ in range(3):
        tw.cycle(); s.cycle()
        for l in range(s.maxlvl):
            assert np.array_equal(tw.level(tw.u, l), s.u(l))
            ri, rj = tw.level(tw.rhs, l), s.rhs(l)
            assert np.array_equal(ri[1:-1, 1:-1], rj[1:-1, 1:-1])
    s.close()


def test_omp_team_is_bitwise_serial(ref):
    """multigrid.cpp:249-266: the OMP-task run equals the serial run exactly."""
    n = 128
    u0, v1, v2 = Oracle().initial_conditions(n)
    dx = 1.0 / n; dt = dx / 10
    a = Towers(ref, n, u0, v1, v2, -4e-4, dt, dx, 1e-8); a.form_rhs(); ia, ha = a.solve(threads=1)
    b = Towers(ref, n, u0, v1, v2, -4e-4, dt, dx, 1e-8); b.form_rhs(); ib, hb = b.solve(threads=4)
    assert ia == ib and np.array_equal(ha, hb) and np.array_equal(a.u[0], b.u[0])


FRESH = os.path.join(ROOT, "oracle", "_ref", "ref_fresh")


@pytest.mark.skipif(not os.path.exists(FRESH), reason="oracle/_ref/ref_fresh not built")
@pytest.mark.parametrize("cfg", ["256 3 -4e-4 1 1e-6 1", "128 2 -4e-4 2 1e-8 2", "512 1 -1e-5 3 1e-10 1"])
def test_zero_filled_towers_are_what_a_fresh_reference_process_computes(oracle, ref, cfg, tmp_path):
    """oracle/ref_prelude.h (malloc -> calloc) and the oracle's calloc'ed towers claim that the
    reference's uninitialised coarse-velocity tails are zero in a real run.
    ref_fresh = the reference timestepper with NO prelude, called once in a fresh process."""
    out = tmp_path / "uT.bin"
    subprocess.run([FRESH, *cfg.split(), str(out)], check=True)
    N, steps, nu, vs, tol, shape = cfg.split()
    N, steps, shape = int(N), int(steps), int(shape)
    fresh = np.fromfile(out).reshape(N + 1, N + 1)
    u0, v1, v2 = oracle.initial_conditions(N, float(vs))
    dx = 1.0 / N; dt = dx / 10
    assert np.array_equal(fresh, ref.timestepper(u0, v1, v2, float(nu), N, dt, steps, dx, float(tol), shape))
    assert np.array_equal(fresh, oracle.timestepper(u0, v1, v2, float(nu), N, dt, steps, dx, float(tol), shape))
