#!/usr/bin/env python
"""Target program for compute-sanitizer: the fused plan at small sizes through every pass flavour
-- both arithmetics, V and W cycles (an up-leg chunk with injection only exists in W-cycles), niter = 3 and 4
(a second, epilogue-free chunk), direct launches (use_graph = 0: the sanitizer instruments kernel launches) --
checked against the one-operator plan so that a run that "passes" the sanitizer also computed the right thing.

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_target.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hpcclassmultigridproject_b200 as mg  # noqa: E402


def ic(n, vscale):
    x = np.arange(n + 1) / n
    X, Y = np.meshgrid(x, x, indexing="ij")
    u0 = np.exp(-100.0 * ((X - 0.2) ** 2 + (Y - 0.4) ** 2))
    u0[0, :] = u0[-1, :] = 0.0; u0[:, 0] = u0[:, -1] = 0.0
    v1 = -vscale * np.pi * np.sin(np.pi * X) * np.cos(np.pi * Y)
    v2 = vscale * np.pi * np.cos(np.pi * X) * np.sin(np.pi * Y)
    return u0, v1, v2


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 256]
    runs = 0
    for n in sizes:
        dx = 1.0 / n; dt = dx / 10
        u0, v1, v2 = ic(n, 2.0)
        for shape, niter in ((1, 3), (2, 3), (1, 4)):
            res = {}
            for plan in (mg.PLAN_UNFUSED, mg.PLAN_FUSED):
                for arith in (mg.ARITH_EXACT, mg.ARITH_FAST):
                    if plan == mg.PLAN_UNFUSED and arith == mg.ARITH_FAST:
                        continue
                    with mg.Solver(n, -4e-4, dt, dx, 1e-10, shape=shape, niter=niter, arith=arith, plan=plan, use_graph=0) as s:
                        s.set_fields_host(u0, v1, v2)
                        s.timestep(2)
                        res[(plan, arith)] = s.get_u_host()
                    runs += 1
            exact, unfused, fast = res[(mg.PLAN_FUSED, mg.ARITH_EXACT)], res[(mg.PLAN_UNFUSED, mg.ARITH_EXACT)], res[(mg.PLAN_FUSED, mg.ARITH_FAST)]
            assert np.array_equal(exact, unfused), (n, shape, niter)
            assert np.linalg.norm(fast - exact) <= 1e-10 * np.linalg.norm(exact), (n, shape, niter)
    print(f"sanitize_target: {runs} solver runs at n = {sizes} agree (fused == one-operator plan bit for bit)")


if __name__ == "__main__":
    main()
