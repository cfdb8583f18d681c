"""EXPERIMENTAL round-2 kernel candidate (csrc/experimental/wasp_body.cuh, "warp-autonomous" streaming
pass; not part of libmgb200.so):
  * its per-thread code compiled for the host (tests/emu/wasp_emu.cpp: 32 threads per warp, shuffles as
    rendezvous) against the oracle, bit for bit;
  * it compiles for sm_100a without spilling its register window;
  * GPU parity and a first timing -- only with MGB200_TEST_EXPERIMENTAL=1 (never part of the default
    GPU suite: the kernel has not run on hardware yet)."""
import ctypes as C
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_stream_pass_emu import fields, from_split, layout, ptr, to_split

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EXP_DIR = os.path.join(ROOT, "hpcclassmultigridproject_b200", "csrc", "experimental")
_dp = C.POINTER(C.c_double)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-ffp-contract=off,-O2", "-shared", "-cudart", "static"]


def _env():
    return {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}


@pytest.fixture(scope="module")
def wemu():
    so = os.path.join(EMU_DIR, "libwaspemu.so")
    src = [os.path.join(EMU_DIR, "wasp_emu.cpp"), os.path.join(EXP_DIR, "wasp_body.cuh"),
           os.path.join(ROOT, "hpcclassmultigridproject_b200", "csrc", "common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-pthread",
                        "-I/usr/local/cuda/include", "-o", so, src[0]], check=True, env=_env())
    lib = C.CDLL(so)
    lib.wasp_emu_run.restype = C.c_long
    lib.wasp_emu_run.argtypes = [C.c_long] * 5 + [_dp] * 8 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_long] * 5
    return lib


def run(wemu, n, u, rhs, v1, v2, K, post, arith, dt, nu, dx, cu=None, rb=16):
    pitch, odd = layout(n)
    cp, co = layout(n // 2)
    us = None if u is None else to_split(u)
    out = np.full((n + 1, pitch), np.nan)
    crhs = np.zeros((n // 2 + 1, cp))
    partials = np.zeros(4096)
    cus = None if cu is None else to_split(cu)
    nt = wemu.wasp_emu_run(n, pitch, odd, cp, co, ptr(us), ptr(out), ptr(to_split(rhs)), ptr(to_split(v1)), ptr(to_split(v2)),
                           ptr(cus), ptr(crhs), ptr(partials), K, post, arith, dt, nu, dx, rb, 0, -1, 0, 0)
    assert 0 < nt <= 4096
    return from_split(out, n), from_split(crhs, n // 2), partials[:nt]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("P", [2, 4])
def test_row_slabs(wemu, oracle, P):
    """the same pass on row slabs (own rows + 8 halo rows held, like the sharded solver's windows): every
    slab is computed from its own window only and the pieces equal the unsharded oracle result"""
    n, K, HALO = 128, 3, 8
    u, rhs, v1, v2 = fields(n, 99)
    cu = np.random.default_rng(3).standard_normal((n // 2 + 1, n // 2 + 1))
    cu[0, :] = cu[-1, :] = 0; cu[:, 0] = cu[:, -1] = 0
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    start = u + oracle.prolongation(cu, n // 2)
    want_u = oracle.gauss_seidel(start.copy(), rhs, n, v1, v2, dt, nu, dx, K)
    want_c = oracle.restriction(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n)
    pitch, odd = layout(n); cp, co = layout(n // 2)
    S = [to_split(a) for a in (u, rhs, v1, v2)]
    CU = to_split(cu)
    got_u = np.full((n + 1, n + 1), np.nan)
    got_c = np.zeros((n // 2 + 1, n // 2 + 1))
    per = n // P
    for r in range(P):
        own_lo, own_hi = r * per, (r + 1) * per - 1 + (1 if r == P - 1 else 0)
        mem_lo, mem_hi = max(0, own_lo - HALO), min(n, own_hi + HALO)
        clo, chi = max(0, (own_lo + 1) // 2 - HALO), min(n // 2, own_hi // 2 + HALO)
        win = [np.ascontiguousarray(a[mem_lo:mem_hi + 1]) for a in S]
        cwin = np.ascontiguousarray(CU[clo:chi + 1])
        out = np.full((mem_hi - mem_lo + 1, pitch), np.nan)
        crhs = np.zeros((chi - clo + 1, cp))
        parts = np.zeros(4096)
        nt = wemu.wasp_emu_run(n, pitch, odd, cp, co, ptr(win[0]), ptr(out), ptr(win[1]), ptr(win[2]), ptr(win[3]), ptr(cwin),
                               ptr(crhs), ptr(parts), K, 1, 1, dt, nu, dx, 24, own_lo, own_hi, mem_lo, clo)
        assert nt > 0
        full = np.full((n + 1, pitch), np.nan); full[mem_lo:mem_hi + 1] = out
        got_u[own_lo:own_hi + 1] = from_split(full, n)[own_lo:own_hi + 1]
        cfull = np.zeros((n // 2 + 1, cp)); cfull[clo:chi + 1] = crhs
        c = from_split(cfull, n // 2)
        ilo, ihi = (own_lo + 1) // 2, own_hi // 2
        got_c[ilo:ihi + 1] = c[ilo:ihi + 1]
    assert np.array_equal(got_u, want_u)
    assert np.array_equal(got_c[1:-1, 1:-1], want_c[1:-1, 1:-1])


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n,K,rb", [(32, 3, 16), (64, 1, 9), (64, 3, 65), (128, 2, 40), (200, 3, 64)])
def test_down_leg(wemu, oracle, n, K, rb):
    """K RB iterations + residual + injection == the oracle's operator sequence, bitwise (n = 200: two strips)"""
    u, rhs, v1, v2 = fields(n, 7 * n + K)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want_u = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, K)
    want_c = oracle.restriction(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n)
    got_u, got_c, _ = run(wemu, n, u, rhs, v1, v2, K, 1, 1, dt, nu, dx, rb=rb)
    assert np.array_equal(got_u, want_u)
    assert np.array_equal(got_c[1:-1, 1:-1], want_c[1:-1, 1:-1])
    assert not got_c[0, :].any() and not got_c[:, 0].any()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("n,K,rb", [(32, 0, 8), (64, 3, 16), (128, 1, 129), (200, 3, 50)])
def test_up_leg(wemu, oracle, n, K, rb):
    """u += P(coarse) ; K RB iterations ; sum of squares of the residual"""
    u, rhs, v1, v2 = fields(n, 11 * n + K)
    cu = np.random.default_rng(n).standard_normal((n // 2 + 1, n // 2 + 1))
    cu[0, :] = cu[-1, :] = 0; cu[:, 0] = cu[:, -1] = 0
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want_u = u + oracle.prolongation(cu, n // 2)
    want_u = oracle.gauss_seidel(want_u, rhs, n, v1, v2, dt, nu, dx, K)
    want_r2 = oracle.norm(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n) ** 2
    got_u, _, parts = run(wemu, n, u, rhs, v1, v2, K, 2, 1, dt, nu, dx, cu=cu, rb=rb)
    assert np.array_equal(got_u, want_u)
    assert abs(parts.sum() - want_r2) <= 1e-12 * want_r2


@pytest.mark.timeout(600)
def test_zero_input_fast_arithmetic_and_residual_only(wemu, oracle):
    n = 64
    u, rhs, v1, v2 = fields(n, 5)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want = oracle.gauss_seidel(np.zeros((n + 1, n + 1)), rhs, n, v1, v2, dt, nu, dx, 3)
    got, _, _ = run(wemu, n, None, rhs, v1, v2, 3, 1, 1, dt, nu, dx)              # u_in == NULL: u is zero
    assert np.array_equal(got, want)
    gotf, _, _ = run(wemu, n, None, rhs, v1, v2, 3, 1, 0, dt, nu, dx, rb=20)
    assert np.linalg.norm(gotf - want) <= 1e-13 * np.linalg.norm(want)
    want_r2 = oracle.norm(oracle.residual(u, rhs, n, v1, v2, dt, nu, dx), n) ** 2
    got_u, _, parts = run(wemu, n, u, rhs, v1, v2, 0, 2, 1, dt, nu, dx)            # K = 0, no prolongation: nothing stored
    assert np.isnan(got_u).all()
    assert abs(parts.sum() - want_r2) <= 1e-12 * want_r2


def _build_x():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    so = os.path.join(EXP_DIR, "libmgb200x.so")
    r = subprocess.run([nvcc, *NVCC_FLAGS, "-Xptxas", "-v", "-o", so, os.path.join(EXP_DIR, "wasp_pass.cu")],
                       capture_output=True, text=True, env=_env())
    assert r.returncode == 0, r.stderr[-3000:]
    return so, r.stderr


def test_compiles_for_sm100a_without_spilling_the_window():
    so, log = _build_x()
    spills = [int(x) for x in re.findall(r"(\d+) bytes spill stores", log)]
    stacks = [int(x) for x in re.findall(r"(\d+) bytes stack frame", log)]
    regs = [int(x) for x in re.findall(r"Used (\d+) registers", log)]
    assert regs and max(regs) <= 255, log            # window of 2K+3 rows in registers (255 x 256 threads still fit one SM)
    assert spills and max(spills) == 0 and max(stacks) == 0, log


@pytest.mark.gpu
@pytest.mark.skipif(not os.environ.get("MGB200_TEST_EXPERIMENTAL"), reason="experimental kernel: set MGB200_TEST_EXPERIMENTAL=1")
def test_gpu_parity_and_first_timing(oracle):
    torch = pytest.importorskip("torch")
    so, _ = _build_x()
    L = C.CDLL(so)
    vp = C.c_void_p
    L.mgb200x_wasp_pass.argtypes = [C.c_long] + [vp] * 8 + [C.c_int] * 3 + [C.c_double] * 3 + [C.c_long, vp]
    L.mgb200x_wasp_tiles.restype = C.c_long
    L.mgb200x_wasp_tiles.argtypes = [C.c_long, C.c_long]
    L.mgb200x_last_error.restype = C.c_char_p

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    for n, K in ((64, 3), (256, 3), (1024, 2)):
        u, rhs, v1, v2 = fields(n, n + K)
        dx = 1.0 / n; dt = dx / 10; nu = -4e-4
        want_u = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, K)
        want_c = oracle.restriction(oracle.residual(want_u, rhs, n, v1, v2, dt, nu, dx), n)
        pitch, _ = layout(n); cp, _ = layout(n // 2)
        d = [dev(to_split(a, 0.0)) for a in (u, rhs, v1, v2)]
        out = torch.full((n + 1, pitch), float("nan"), dtype=torch.float64, device="cuda")
        crhs = torch.zeros((n // 2 + 1, cp), dtype=torch.float64, device="cuda")
        rc = L.mgb200x_wasp_pass(n, d[0].data_ptr(), out.data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), None,
                                 crhs.data_ptr(), None, K, 1, 1, dt, nu, dx, 0, None)
        torch.cuda.synchronize()
        assert rc == 0, L.mgb200x_last_error()
        assert np.array_equal(from_split(out.cpu().numpy(), n), want_u)
        assert np.array_equal(from_split(crhs.cpu().numpy(), n // 2)[1:-1, 1:-1], want_c[1:-1, 1:-1])
    # first timing at the production size (fast arithmetic), printed for the experimenter
    n = 16384; dx = 1.0 / n; dt = dx / 10
    pitch, _ = layout(n); cp, _ = layout(n // 2)
    a = [torch.rand((n + 1, pitch), dtype=torch.float64, device="cuda") for _ in range(4)]
    out = torch.empty((n + 1, pitch), dtype=torch.float64, device="cuda")
    crhs = torch.zeros((n // 2 + 1, cp), dtype=torch.float64, device="cuda")
    for rb in (256, 512, 1024):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(4):
            if rep == 1:
                ev[0].record()
            assert L.mgb200x_wasp_pass(n, a[0].data_ptr(), out.data_ptr(), a[1].data_ptr(), a[2].data_ptr(), a[3].data_ptr(), None,
                                       crhs.data_ptr(), None, 3, 1, 0, dt, -4e-4, dx, rb, None) == 0
        ev[1].record(); ev[1].synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        print(f"wasp down leg N={n} rows/band={rb}: {ms:.3f} ms = {42.0 * (n + 1) ** 2 / ms / 1e6:.0f} GB/s")
