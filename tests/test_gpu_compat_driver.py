"""Acceptance test of the drop-in boundary: the reference's UNMODIFIED CUDA
driver multigrid.cu (its own mg_inner / mg_outer / timestepper / main) compiled against
include/compat/gscu.h and linked with libmgb200.so (oracle/Makefile target `compat`, built where
/root/reference is mounted) runs on the B200 and writes the same uTcuda.txt as the CPU oracle
computes for the reference's parameters (N=256, 100 steps, multigrid.cu:210-256)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "multigrid_cu_compat")


@pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/multigrid_cu_compat not built (needs /root/reference)")
def test_unmodified_reference_cuda_driver_on_libmgb200(oracle, tmp_path):
    r = subprocess.run([EXE], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    tab = np.loadtxt(tmp_path / "uTcuda.txt")               # "%d\t%d\t%f" (multigrid.cu:262)
    n = 256
    got = tab[:, 2].reshape(n + 1, n + 1)
    dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n)
    want = oracle.timestepper(u0, v1, v2, -4e-4, n, dt, 100, dx, 1e-6)
    assert np.abs(got - want).max() <= 5.01e-7               # %f prints 6 decimals
    mid = float(r.stdout.split()[0])                         # printf("%g", uT[N/2][N/2]) (multigrid.cu:258)
    assert abs(mid - want[n // 2, n // 2]) <= 1e-5 * abs(want[n // 2, n // 2])


TOOL = os.path.join(ROOT, "hpcclassmultigridproject_b200", "multigrid_b200")


@pytest.mark.skipif(not os.path.exists(TOOL), reason="multigrid_b200 not built")
def test_command_line_front_end(oracle, tmp_path):
    """tools/multigrid_b200.cpp: the reference executables' parameters as flags, the same uT.txt format
    (multigrid.cpp:269-275) and a raw dump"""
    r = subprocess.run([TOOL, "--N", "128", "--steps", "5", "--vscale", "2", "--tol", "1e-8", "--exact", "--out", "uT.txt", "--bin", "uT.f64"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    n = 128; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n, 2.0)
    want = oracle.timestepper(u0, v1, v2, -4e-4, n, dt, 5, dx, 1e-8)
    raw = np.fromfile(tmp_path / "uT.f64").reshape(n + 1, n + 1)
    assert np.linalg.norm(raw - want) <= 1e-13 * np.linalg.norm(want)      # device initial conditions differ by ulps
    tab = np.loadtxt(tmp_path / "uT.txt")
    assert tab.shape == ((n + 1) ** 2, 3) and np.array_equal(tab[:, 0], np.repeat(np.arange(n + 1), n + 1))
    assert np.abs(tab[:, 2].reshape(n + 1, n + 1) - want).max() <= 5.01e-7
