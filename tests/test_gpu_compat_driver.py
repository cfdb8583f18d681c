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


EXE_CPP = os.path.join(ROOT, "oracle", "_ref", "multigrid_cpp_compat")


@pytest.mark.skipif(not os.path.exists(EXE_CPP), reason="oracle/_ref/multigrid_cpp_compat not built (needs /root/reference)")
def test_unmodified_reference_cpu_driver_on_libmgb200(oracle, tmp_path):
    """the reference's multigrid.cpp, byte for byte, over include/compat/gs.h: its serial run and its run inside an
    OpenMP team (multigrid.cpp:245-258) both execute every operator on the GPU, agree with each other exactly as the
    reference's do (multigrid.cpp:261-266 prints 0) and write the reference's uT.txt"""
    r = subprocess.run([EXE_CPP], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Error (compared to the referenced solution) = 0.000000e+00" in r.stdout, r.stdout
    n = 256; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n)
    want = oracle.timestepper(u0, v1, v2, -4e-4, n, dt, 100, dx, 1e-6)
    for name in ("uT.txt", "uTomp.txt"):
        tab = np.loadtxt(tmp_path / name)                   # "%d\t%d\t%f" (multigrid.cpp:272,281)
        assert np.abs(tab[:, 2].reshape(n + 1, n + 1) - want).max() <= 5.01e-7


def test_mg_inner_and_mg_outer_on_caller_owned_towers(oracle):
    """mgb200_mg_inner / mgb200_mg_outer with the reference argument lists (multigrid.cu:17-21,101-103) on towers the
    CALLER built the reference's way (multigrid.cpp:138-160: every coarse level (N/2+1)^2 doubles, velocities by the
    never-halved restriction): bit-identical to the oracle's driver, level by level"""
    import ctypes as C
    import torch
    import hpcclassmultigridproject_b200 as mg
    from oracle.oracle import OracleSolver
    L = mg.lib()
    pp = C.POINTER(C.c_void_p)
    L.mgb200_mg_inner.argtypes = [pp, pp, pp, pp, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                  C.POINTER(mg.Options), C.c_void_p]
    L.mgb200_mg_outer.argtypes = [pp, pp, pp, pp, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(mg.Options), C.c_void_p, C.POINTER(mg.SolveInfo)]
    n = 128; dx = 1.0 / n; dt = dx / 10; nu = -4e-4; tol = 1e-6
    maxlvl = mg.maxlvl_for(n)
    u0, v1, v2 = oracle.initial_conditions(n, 2.0)
    o = OracleSolver(n, u0, v1, v2, nu, dt, dx, tol, 1)
    o.form_rhs()
    # the caller's towers: level l > 0 gets (n/2+1)^2 doubles (multigrid.cpp:150-153) holding the oracle's level arrays
    # (which reproduce the reference's velocity towers) at the level's true stride, zero tail
    def tower(get):
        out = []
        for l in range(maxlvl):
            nl = n >> l
            a = torch.zeros((n + 1) ** 2 if l == 0 else (n // 2 + 1) ** 2, dtype=torch.float64, device="cuda")
            a[: (nl + 1) ** 2] = torch.from_numpy(np.ascontiguousarray(get(l))).reshape(-1).cuda()
            out.append(a)
        return out
    U, F, V1, V2 = tower(o.u), tower(o.rhs), tower(o.v1), tower(o.v2)
    tmp = torch.zeros((n + 1) ** 2, dtype=torch.float64, device="cuda")
    arr = lambda ts: (C.c_void_p * maxlvl)(*[t.data_ptr() for t in ts])
    opt = mg.default_options(arith=mg.ARITH_EXACT)
    # one mg_inner == one oracle cycle, on every level
    assert L.mgb200_mg_inner(arr(U), arr(F), arr(V1), arr(V2), tmp.data_ptr(), dx, n, 0, maxlvl, 1, dt, nu, C.byref(opt), None) == 0, \
        L.mgb200_last_error()
    o.cycle()
    for l in range(maxlvl):
        nl = n >> l
        assert np.array_equal(U[l][: (nl + 1) ** 2].cpu().numpy().reshape(nl + 1, nl + 1), o.u(l)), l
    # mg_outer from there: same number of further cycles and the same iterate as the oracle's loop
    info = mg.SolveInfo()
    assert L.mgb200_mg_outer(arr(U), arr(V1), arr(V2), arr(F), tmp.data_ptr(), nu, maxlvl, n, dt, dx, tol, 1, C.byref(opt), None,
                             C.byref(info)) == 0, L.mgb200_last_error()
    it, hist = o.solve()
    assert info.cycles == it and bool(info.converged) == bool(hist[-1] / hist[0] <= tol)
    assert np.allclose(info.history(), hist, rtol=1e-9, atol=0)
    assert np.array_equal(U[0].cpu().numpy().reshape(n + 1, n + 1), o.u(0))
    o.close()


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


@pytest.mark.skipif(not os.path.exists(TOOL), reason="multigrid_b200 not built")
def test_command_line_front_end_flags_and_gpus(oracle, tmp_path):
    """the literals of multigrid.cpp:41,60,94 as flags (--niter --coarse-tol --coarse-maxit --max-cycle), --correct-towers,
    and --gpus 2 (one forked process per GPU, row slabs): the two-GPU dump equals the one-GPU dump byte for byte"""
    import torch
    base = ["--N", "1024", "--steps", "2", "--tol", "1e-9", "--exact", "--out", "-"]
    r = subprocess.run([TOOL] + base + ["--bin", "one.f64", "--coarse-tol", "1e-5", "--coarse-maxit", "1000", "--max-cycle", "50", "--niter", "3"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    one = np.fromfile(tmp_path / "one.f64")
    n = 1024; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n)
    want = oracle.timestepper(u0, v1, v2, -4e-4, n, dt, 2, dx, 1e-9)
    assert np.linalg.norm(one.reshape(n + 1, n + 1) - want) <= 1e-13 * np.linalg.norm(want)
    r = subprocess.run([TOOL] + base + ["--bin", "ct.f64", "--correct-towers", "--max-cycle", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert not np.array_equal(np.fromfile(tmp_path / "ct.f64"), one)
    if torch.cuda.device_count() >= 2:
        r = subprocess.run([TOOL] + base + ["--bin", "two.f64", "--gpus", "2"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert np.array_equal(np.fromfile(tmp_path / "two.f64"), one)
