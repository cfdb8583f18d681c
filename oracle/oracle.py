"""ctypes front-end to the CPU checker (TEST INFRASTRUCTURE ONLY).

Two libraries, one numpy-level interface:

* ``Oracle()``            -- oracle/liboracle.so, our plain-C restatement (mg_oracle.c)
* ``Oracle(ref="O0")``    -- oracle/_ref/libmgref_O0.so, the UNMODIFIED reference sources
  (``"O3"`` is the optimised build of the same sources; see oracle/Makefile)

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py may import this module.  The product (hpcclassmultigridproject_b200)
never does: it has no CPU path at all.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_dpp = C.POINTER(_dp)


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is mounted, oracle/_ref/*.so."""
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["make", "-C", HERE, "all"], check=True, env=env,
                   stdout=subprocess.DEVNULL if quiet else None,
                   stderr=subprocess.DEVNULL if quiet else None)


def ref_available(kind: str = "O0") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libmgref_{kind}.so"))


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def maxlvl_for(n: int) -> int:
    """multigrid.cpp:193 -- int(log2(N)) - 4, coarsest level has n = 32."""
    return int(np.log2(n)) - 4


class Oracle:
    def __init__(self, ref: str | None = None):
        self.ref = ref
        if ref is None:
            path = os.path.join(HERE, "liboracle.so")
            if not os.path.exists(path):
                build()
            self.lib = C.CDLL(path)
            self.pfx = "orc_"
        else:
            path = os.path.join(HERE, "_ref", f"libmgref_{ref}.so")
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            self.lib = C.CDLL(path)
            self.pfx = "ref_"
        L, d, l, i = self.lib, C.c_double, C.c_long, C.c_int
        if ref is None:
            L.orc_compute_rhs.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.orc_residual.argtypes = [_dp, _dp, _dp, l, _dp, _dp, d, d, d]
            L.orc_norm.argtypes = [_dp, l]; L.orc_norm.restype = d
            L.orc_gauss_seidel.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.orc_prolongation.argtypes = [_dp, _dp, l]
            L.orc_restriction.argtypes = [_dp, _dp, l]
            L.orc_restriction_fw.argtypes = [_dp, _dp, l]
            L.orc_timestepper.argtypes = [_dp, _dp, _dp, _dp, d, i, i, d, d, d, d, i]
            L.orc_initial_conditions.argtypes = [_dp, _dp, _dp, l, d]
            L.orc_create.argtypes = [l, i, _dp, _dp, _dp, d, d, d, d, i]; L.orc_create.restype = C.c_void_p
            L.orc_destroy.argtypes = [C.c_void_p]
            L.orc_correct_towers.argtypes = [C.c_void_p]
            L.orc_set_coarse_exact.argtypes = [C.c_void_p, i]
            L.orc_coarse_lu_solve.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.orc_cycle.argtypes = [C.c_void_p, i]
            L.orc_form_rhs.argtypes = [C.c_void_p]
            L.orc_residual_norm.argtypes = [C.c_void_p]; L.orc_residual_norm.restype = d
            L.orc_solve.argtypes = [C.c_void_p, _dp]; L.orc_solve.restype = i
            L.orc_advance.argtypes = [C.c_void_p, i, C.POINTER(i)]
            for nm in ("u", "rhs", "v1", "v2"):
                f = getattr(L, f"orc_level_{nm}"); f.argtypes = [C.c_void_p, i]; f.restype = _dp
            L.orc_tmp.argtypes = [C.c_void_p]; L.orc_tmp.restype = _dp
        else:
            L.ref_compute_rhs.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.ref_residual.argtypes = [_dp, _dp, _dp, l, _dp, _dp, d, d, d]
            L.ref_compute_norm.argtypes = [_dp, l]; L.ref_compute_norm.restype = d
            L.ref_gauss_seidel.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.ref_gauss_seidel2.argtypes = [_dp, _dp, l, _dp, _dp, d, d, d]
            L.ref_prolongation.argtypes = [_dp, _dp, i]
            L.ref_restriction.argtypes = [_dp, _dp, i]
            L.ref_timestepper.argtypes = [_dp, _dp, _dp, _dp, d, i, i, d, d, d, d, i]
            L.ref_mg_inner.argtypes = [_dpp, _dpp, _dpp, _dpp, _dp, d, i, i, i, i, d, d]
            L.ref_mg_outer.argtypes = [_dpp, _dpp, _dpp, _dpp, _dp, d, i, i, d, d, d, i]
            L.ref_mg_outer_hist.argtypes = [_dpp, _dpp, _dpp, _dpp, _dp, d, i, i, d, d, d, i, _dp]
            L.ref_mg_outer_hist.restype = i
            L.ref_timestepper_omp.argtypes = [i, _dp, _dp, _dp, _dp, d, i, i, d, d, d, d, i]
            L.ref_mg_outer_hist_omp.argtypes = [i, _dpp, _dpp, _dpp, _dpp, _dp, d, i, i, d, d, d, i, _dp]
            L.ref_mg_outer_hist_omp.restype = i
            L.ref_cycle_and_norm_omp.argtypes = [i, _dpp, _dpp, _dpp, _dpp, _dp, d, i, i, d, d, i]
            L.ref_cycle_and_norm_omp.restype = d

    # ---- operators: same argument meaning as gs.h:3-17 -----------------------
    def compute_rhs(self, u, n, v1, v2, dt, nu, dx, rhs=None):
        rhs = np.zeros_like(u) if rhs is None else rhs
        getattr(self.lib, self.pfx + "compute_rhs")(_p(rhs), _p(u), n, _p(v1), _p(v2), dt, nu, dx)
        return rhs

    def residual(self, u, rhs, n, v1, v2, dt, nu, dx, res=None):
        res = np.zeros_like(u) if res is None else res
        getattr(self.lib, self.pfx + "residual")(_p(res), _p(u), _p(rhs), n, _p(v1), _p(v2), dt, nu, dx)
        return res

    def norm(self, res, n) -> float:
        f = self.lib.orc_norm if self.ref is None else self.lib.ref_compute_norm
        return float(f(_p(res), n))

    def gauss_seidel(self, u, rhs, n, v1, v2, dt, nu, dx, iters: int = 1):
        """in place on u"""
        for _ in range(iters):
            getattr(self.lib, self.pfx + "gauss_seidel")(_p(u), _p(rhs), n, _p(v1), _p(v2), dt, nu, dx)
        return u

    def prolongation(self, coarse, nc):
        fine = np.zeros((2 * nc + 1, 2 * nc + 1))
        getattr(self.lib, self.pfx + "prolongation")(_p(fine), _p(coarse), nc)
        return fine

    def restriction(self, fine, nf):
        coarse = np.zeros((nf // 2 + 1, nf // 2 + 1))
        getattr(self.lib, self.pfx + "restriction")(_p(coarse), _p(fine), nf)
        return coarse

    def restriction_fw(self, fine, nf):
        """full weighting of gs.cpp:277-280 (commented out in the reference: restatement only, no _ref twin)"""
        coarse = np.zeros((nf // 2 + 1, nf // 2 + 1))
        self.lib.orc_restriction_fw(_p(coarse), _p(np.ascontiguousarray(fine)), nf)
        return coarse

    # ---- drivers ---------------------------------------------------------------
    def timestepper(self, u0, v1, v2, nu, n, dt, nsteps, dx, tol, shape=1, maxlvl=None):
        """multigrid.cpp:124 with T = nsteps*dt (int(T/dt) == nsteps for the sizes used)."""
        maxlvl = maxlvl_for(n) if maxlvl is None else maxlvl
        uT = np.zeros_like(u0)
        T = nsteps * dt
        assert int(T / dt) == nsteps
        getattr(self.lib, self.pfx + "timestepper")(_p(uT), _p(u0.copy()), _p(v1.copy()), _p(v2.copy()),
                                                    nu, maxlvl, n, dt, T, dx, tol, shape)
        return uT

    def initial_conditions(self, n, vscale=1.0):
        """multigrid.cpp:206-233 (own restatement; the reference has them inline in main)."""
        lib = self.lib if self.ref is None else Oracle().lib
        u0 = np.zeros((n + 1, n + 1)); v1 = np.zeros_like(u0); v2 = np.zeros_like(u0)
        lib.orc_initial_conditions(_p(u0), _p(v1), _p(v2), n, vscale)
        return u0, v1, v2


class Towers:
    """The level towers of timestepper (multigrid.cpp:131-162) as numpy arrays, for
    driving the reference's mg_inner / mg_outer directly (ref builds) or the oracle's
    orc_cycle / orc_solve through the same calls."""

    def __init__(self, orc: Oracle, n, u0, v1, v2, nu, dt, dx, tol, shape=1, maxlvl=None):
        self.o, self.n, self.nu, self.dt, self.dx, self.tol, self.shape = orc, n, nu, dt, dx, tol, shape
        self.maxlvl = maxlvl_for(n) if maxlvl is None else maxlvl
        nh = n // 2
        self.u = [u0.copy()] + [np.zeros((nh + 1) ** 2) for _ in range(1, self.maxlvl)]
        self.rhs = [np.zeros_like(u0)] + [np.zeros((nh + 1) ** 2) for _ in range(1, self.maxlvl)]
        self.v1 = [v1.copy()] + [np.zeros((nh + 1) ** 2) for _ in range(1, self.maxlvl)]
        self.v2 = [v2.copy()] + [np.zeros((nh + 1) ** 2) for _ in range(1, self.maxlvl)]
        o = Oracle()  # restriction of the velocity towers exactly as multigrid.cpp:155,157
        for l in range(1, self.maxlvl):
            for tow in (self.v1, self.v2):
                src = tow[l - 1].reshape(-1)
                dst = tow[l]
                o.lib.orc_restriction(_p(dst), _p(src), nh)
        self.tmp = np.zeros((n + 1) ** 2)

    def _pp(self, tow):
        # one extra slot: mg_inner reads u[lvl+1] at the coarsest level (multigrid.cpp:44)
        arr = (_dp * (self.maxlvl + 1))(*[_p(a.reshape(-1)) for a in tow], None)
        return C.cast(arr, _dpp)

    def level(self, tow, l):
        nl = self.n >> l
        return tow[l].reshape(-1)[: (nl + 1) ** 2].reshape(nl + 1, nl + 1)

    def form_rhs(self):
        self.o.compute_rhs(self.u[0], self.n, self.v1[0], self.v2[0], self.dt, self.nu, self.dx, rhs=self.rhs[0])

    def cycle(self):
        assert self.o.ref is not None
        self.o.lib.ref_mg_inner(self._pp(self.u), self._pp(self.rhs), self._pp(self.v1), self._pp(self.v2),
                                _p(self.tmp), self.dx, self.n, 0, self.maxlvl, self.shape, self.dt, self.nu)

    def solve(self, threads: int = 1):
        """mg_outer with the residual history: returns (cycles, [res0, res1, ...]).
        threads > 1 runs it inside omp parallel/single as multigrid.cpp:252-258 does."""
        assert self.o.ref is not None
        hist = np.zeros(64)
        it = self.o.lib.ref_mg_outer_hist_omp(threads, self._pp(self.u), self._pp(self.v1), self._pp(self.v2),
                                              self._pp(self.rhs), _p(self.tmp), self.nu, self.maxlvl, self.n,
                                              self.dt, self.dx, self.tol, self.shape, _p(hist))
        return it, hist[: it + 1].copy()

    def cycle_and_norm(self, threads: int = 1) -> float:
        """one V-cycle + convergence check (multigrid.cpp:110-113) -- the unit bench.py times"""
        assert self.o.ref is not None
        return float(self.o.lib.ref_cycle_and_norm_omp(threads, self._pp(self.u), self._pp(self.v1),
                                                       self._pp(self.v2), self._pp(self.rhs), _p(self.tmp),
                                                       self.nu, self.maxlvl, self.n, self.dt, self.dx,
                                                       self.shape))


class OracleSolver:
    """orc_solver handle (mg_oracle.h) -- the restatement's driver with histories."""

    def __init__(self, n, u0, v1, v2, nu, dt, dx, tol, shape=1, maxlvl=None, correct_towers=False, coarse_exact=False):
        self.o = Oracle()
        self.n = n
        self.maxlvl = maxlvl_for(n) if maxlvl is None else maxlvl
        self.h = self.o.lib.orc_create(n, self.maxlvl, _p(u0), _p(v1), _p(v2), nu, dt, dx, tol, shape)
        assert self.h
        if correct_towers:                # opt-in: true injection of the velocities (no reference twin)
            self.o.lib.orc_correct_towers(self.h)
        if coarse_exact:                  # opt-in: direct solve of the coarsest level (no reference twin)
            self.o.lib.orc_set_coarse_exact(self.h, 1)

    def close(self):
        if self.h:
            self.o.lib.orc_destroy(self.h); self.h = None

    def __del__(self):
        self.close()

    def _lvl(self, name, l):
        nl = self.n >> l
        ptr = getattr(self.o.lib, f"orc_level_{name}")(self.h, l)
        return np.ctypeslib.as_array(ptr, shape=((nl + 1) ** 2,)).reshape(nl + 1, nl + 1)

    def u(self, l=0): return self._lvl("u", l)
    def rhs(self, l=0): return self._lvl("rhs", l)
    def v1(self, l=0): return self._lvl("v1", l)
    def v2(self, l=0): return self._lvl("v2", l)

    def form_rhs(self): self.o.lib.orc_form_rhs(self.h)
    def cycle(self, lvl=0): self.o.lib.orc_cycle(self.h, lvl)
    def residual_norm(self): return float(self.o.lib.orc_residual_norm(self.h))

    def solve(self):
        hist = np.zeros(64)
        it = self.o.lib.orc_solve(self.h, _p(hist))
        return it, hist[: it + 1].copy()

    def advance(self, nsteps):
        cyc = (C.c_int * nsteps)()
        self.o.lib.orc_advance(self.h, nsteps, cyc)
        return list(cyc)
