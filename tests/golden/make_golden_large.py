#!/usr/bin/env python
"""Large-size golden SCALARS + strided samples from the compiled reference (-O3 build of the
unmodified sources, bit-identical to -O0: no FMA, no fast-math), run with an OpenMP team exactly
as multigrid.cpp:252-258 does.  Needs ~40 GB of host RAM for N=16384.

    python tests/golden/make_golden_large.py [tag ...]

Each fixture holds: per-cycle residual norms (hist), cycle count, ||uT||_2, uT[N/2,N/2] and
uT[::N/128, ::N/128] (129x129 samples) -- enough for full-size parity tests on the GPU box,
where the reference cannot be run at these sizes inside a test budget.
"""
import gc
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, Towers  # noqa: E402
from hpcclassmultigridproject_b200.digest import combine, slab_digests  # noqa: E402  (hashlib only, no library load)

OUT = os.path.dirname(os.path.abspath(__file__))
CFG = {
    # BASELINE.json configs[1]: diffusion-dominated, small v
    "large_c2_n4096": dict(n=4096, nu=-4e-4, vscale=0.01, tol=1e-10),
    "large_n2048_v6": dict(n=2048, nu=-4e-4, vscale=6.0, tol=1e-10),
    "large_n8192": dict(n=8192, nu=-4e-4, vscale=1.0, tol=1e-10),
    # configs[2]: the headline size
    "large_c3_n16384": dict(n=16384, nu=-4e-4, vscale=1.0, tol=1e-10),
    # configs[2] at the tolerance bench.py runs (tol = 1e-6: 3 cycles); carries the slab digests of the
    # whole final field (hpcclassmultigridproject_b200/digest.py) for the bench line's parity block
    "large_c3_n16384_tol1e-6": dict(n=16384, nu=-4e-4, vscale=1.0, tol=1e-6),
    # configs[3]: advection-dominated
    "large_c4_n16384": dict(n=16384, nu=-1e-6, vscale=1.0, tol=1e-10),
}


def run(tag, n, nu, vscale, tol, threads=8):
    ref = Oracle("O3")
    t0 = time.time()
    u0, v1, v2 = ref.initial_conditions(n, vscale)
    dx = 1.0 / n; dt = dx / 10
    tw = Towers(ref, n, u0, v1, v2, nu, dt, dx, tol, 1)
    del u0, v1, v2; gc.collect()
    tw.form_rhs()
    it, h = tw.solve(threads=threads)
    uT = tw.level(tw.u, 0)
    st = n // 128
    out = dict(n=n, nu=nu, vscale=vscale, tol=tol, dx=dx, dt=dt, steps=1, cycles=it, hist=h,
               norm_uT=float(np.sqrt(np.sum(uT * uT))), mid=uT[n // 2, n // 2],
               sample=uT[::st, ::st].copy(), stride=st)
    dense = np.ascontiguousarray(uT)
    dig = slab_digests(dense, n)
    out["slab_sha256"] = np.array([dig[g].hex() for g in sorted(dig)])
    out["u_sha256"] = combine(dig)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, "cycles", it, "hist", h / h[0], "norm", out["norm_uT"], "mid", out["mid"],
          "%.1fs" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    for tag in (sys.argv[1:] or CFG):
        run(tag, **CFG[tag])
        gc.collect()
