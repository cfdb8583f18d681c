#!/usr/bin/env python
"""Generate tests/golden/*.npz from the compiled, UNMODIFIED reference.

Run in the build container (needs /root/reference and oracle/_ref/libmgref_O0.so,
built by `make -C oracle`):   python tests/golden/make_golden.py

Fixtures:
  ops_n{32,64}.npz        seeded random inputs + outputs of every gs.h operator (gs.h:3-17)
  solve_n{64,128,256}.npz reference-IC run: per-step V-cycle counts, residual history of
                          every step (mg_outer loop, multigrid.cpp:104-114) and final field
  solve_cfg_*.npz         other parameter corners (W-cycle, vscale, nu, tol)
The two *_stdout.txt files are the stdout of the reference's own print tests
(resnormtest.cpp, prolrestest.cpp compiled as-is with g++ -fopenmp -O0), timing line removed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Towers  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ref = Oracle("O0")


def ops(n, seed):
    rng = np.random.default_rng(seed)
    u, rhs, v1, v2 = (rng.standard_normal((n + 1, n + 1)) for _ in range(4))
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    out = dict(n=n, seed=seed, dx=dx, dt=dt, nu=nu, u=u, rhs=rhs, v1=v1, v2=v2)
    out["compute_rhs"] = ref.compute_rhs(u, n, v1, v2, dt, nu, dx)
    out["residual"] = ref.residual(u, rhs, n, v1, v2, dt, nu, dx)
    out["norm"] = ref.norm(out["residual"], n)
    out["gs1"] = ref.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 1)
    out["gs3"] = ref.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 3)
    out["restriction"] = ref.restriction(u, n)
    out["prolongation"] = ref.prolongation(u, n)
    np.savez_compressed(os.path.join(OUT, f"ops_n{n}.npz"), **out)


def solve(tag, n, steps, nu=-4e-4, vscale=1.0, tol=1e-6, shape=1, keep_u=True):
    u0, v1, v2 = ref.initial_conditions(n, vscale)
    dx = 1.0 / n; dt = dx / 10
    tw = Towers(ref, n, u0, v1, v2, nu, dt, dx, tol, shape)
    cycles, hists = [], []
    for _ in range(steps):
        tw.form_rhs()
        it, h = tw.solve()
        cycles.append(it)
        hh = np.full(52, np.nan); hh[: len(h)] = h
        hists.append(hh)
    uT = tw.level(tw.u, 0).copy()
    # the reference's own timestepper must give the same field
    uT2 = ref.timestepper(u0, v1, v2, nu, n, dt, steps, dx, tol, shape)
    assert np.array_equal(uT, uT2), tag
    out = dict(n=n, steps=steps, nu=nu, vscale=vscale, tol=tol, shape=shape, dx=dx, dt=dt,
               cycles=np.array(cycles), hist=np.array(hists),
               norm_uT=np.linalg.norm(uT), mid=uT[n // 2, n // 2])
    if keep_u:
        out["uT"] = uT
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print(tag, "cycles", cycles, "norm", out["norm_uT"])


def refmain():
    """Run the reference's own main() (multigrid.cpp:188-293: N=256, 100 steps, its own inline
    ICs) in a scratch directory and keep what it writes to uT.txt ("%d\\t%d\\t%f", 6 decimals).
    This pins the initial conditions and the whole pipeline against the shipped executable."""
    import subprocess, tempfile
    # `main` is renamed ref_main by -Dmain=ref_main and, having no return statement
    # (multigrid.cpp:293), ends in a g++ trap once it is no longer `main`; by then both files are
    # written and closed.  Run it in a child process and ignore how the child ends.
    code = ("import ctypes; L = ctypes.CDLL(%r); getattr(L, '_Z8ref_mainv')()"
            % os.path.join(ROOT, "oracle", "_ref", "libmgref_O0.so"))
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([sys.executable, "-c", code], cwd=d, stdout=subprocess.DEVNULL)
        tab = np.loadtxt(os.path.join(d, "uT.txt"))
        tab2 = np.loadtxt(os.path.join(d, "uTomp.txt"))
    n = 256
    uT = tab[:, 2].reshape(n + 1, n + 1)
    assert np.array_equal(tab, tab2)      # serial == OMP, the reference's own self-check (:261-266)
    np.savez_compressed(os.path.join(OUT, "refmain_uT_n256_100steps.npz"), uT=uT.astype(np.float32), n=n)
    print("refmain: max", uT.max())


if __name__ == "__main__":
    refmain()
    ops(32, 1)
    ops(64, 2)
    solve("solve_n32", 32, 4)
    solve("solve_n64", 64, 6)
    solve("solve_n128", 128, 4)
    solve("solve_n256", 256, 10)                       # BASELINE.json configs[0]
    solve("solve_cfg_wcycle_n128", 128, 3, shape=2)
    solve("solve_cfg_tight_n256", 256, 2, tol=1e-10)
    solve("solve_cfg_adv_n256", 256, 2, nu=-1e-6, vscale=3.0, tol=1e-10)
    solve("solve_cfg_diff_n256", 256, 2, nu=-1e-2, vscale=0.01, tol=1e-10)
    solve("solve_cfg_n512", 512, 2, tol=1e-10, keep_u=False)
    solve("solve_cfg_n1024", 1024, 1, tol=1e-10, keep_u=False)
