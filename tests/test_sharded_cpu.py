"""Host-side logic of the N > 1 path, on CPU:

* the slab planner of the C ABI (mgb200_slab_plan: pure arithmetic, no GPU) partitions every level,
  cuts at even rows and keeps parent/child windows consistent;
* two `gloo` ranks run the REAL per-thread streaming-pass code (host emulation, tests/emu) on their
  row slabs, swap halo rows with torch.distributed exactly as solver.cu::exchange_halo does, and the
  stitched result equals the unsharded oracle bit for bit (down leg + injection, halo swap, up leg
  with prolongation and the all-reduced residual norm)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

HALO = 8


def test_slab_plan_partitions_every_level():
    import hpcclassmultigridproject_b200 as mg
    for n in (256, 4096, 16384, 65536):
        maxlvl = mg.maxlvl_for(n)
        for P in (1, 2, 4, 8):
            for l in range(maxlvl):
                nl = n >> l
                w = [mg.slab_plan(n, maxlvl, P, r, l) for r in range(P)]
                if w[0]["sharded"]:
                    assert all(x["sharded"] and x["present"] for x in w)
                    assert w[0]["own_lo"] == 0 and w[-1]["own_hi"] == nl
                    for a, b in zip(w, w[1:]):
                        assert b["own_lo"] == a["own_hi"] + 1 and b["own_lo"] % 2 == 0     # even cuts: injection stays local
                    for x in w:
                        assert x["mem_lo"] == max(0, x["own_lo"] - HALO) and x["mem_hi"] == min(nl, x["own_hi"] + HALO)
                        assert x["own_hi"] - x["own_lo"] + 1 >= 2 * HALO
                    assert l < maxlvl - 1                                                  # the coarsest level is never cut
                else:
                    assert w[0]["present"] and (w[0]["own_lo"], w[0]["own_hi"]) == (0, nl)  # whole on rank 0
                    parent_sharded = l > 0 and mg.slab_plan(n, maxlvl, P, 0, l - 1)["sharded"]
                    for r in range(1, P):
                        if parent_sharded:       # slab of the first agglomerated level = the rows under the parent's slab
                            par = mg.slab_plan(n, maxlvl, P, r, l - 1)
                            assert w[r]["present"]
                            assert w[r]["own_lo"] == (par["own_lo"] + 1) // 2 and w[r]["own_hi"] == par["own_hi"] // 2
                        else:
                            assert not w[r]["present"]
            if P > 1 and n >= 4096:
                assert mg.slab_plan(n, maxlvl, P, 0, 0)["sharded"]


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import hpcclassmultigridproject_b200 as mg
    import emu_util as T
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    emu = T.load_model()

    maxlvl = 3
    w = mg.slab_plan(n, maxlvl, world, rank, 0, 16)
    wc = mg.slab_plan(n, maxlvl, world, rank, 1, 16)
    assert w["sharded"] and wc["sharded"]
    rng = np.random.default_rng(123)                    # every rank draws the same full fields
    u, rhs, v1, v2 = (rng.standard_normal((n + 1, n + 1)) for _ in range(4))
    cu = rng.standard_normal((n // 2 + 1, n // 2 + 1)); cu[0, :] = cu[-1, :] = 0; cu[:, 0] = cu[:, -1] = 0
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    pitch, odd = T.layout(n); cp, co = T.layout(n // 2)
    rows = slice(w["mem_lo"], w["mem_hi"] + 1); crows = slice(wc["mem_lo"], wc["mem_hi"] + 1)

    def window(full):          # split layout, rows of this rank's window only
        return np.ascontiguousarray(T.to_split(full)[rows])

    def run(u_in, coarse_u, K, post):
        out = np.full((w["mem_hi"] - w["mem_lo"] + 1, pitch), np.nan)
        crhs = np.zeros((wc["mem_hi"] - wc["mem_lo"] + 1, cp))
        parts = np.zeros(4096)
        nt = emu.syst_emu_run(n, pitch, odd, cp, co, T.ptr(u_in), T.ptr(out), T.ptr(window(rhs)), T.ptr(window(v1)),
                            T.ptr(window(v2)), T.ptr(coarse_u), T.ptr(crhs), T.ptr(parts), K, post, 1, dt, nu, dx, 0, 0, rank,
                            w["own_lo"], w["own_hi"], w["mem_lo"], out.shape[0], wc["mem_lo"], crhs.shape[0])
        assert nt > 0
        return out, crhs, parts[:nt]

    def exchange(arr, win):    # solver.cu::exchange_halo: HALO boundary rows to / from the slab neighbours
        t = torch.from_numpy(arr)
        lo, hi, m0 = win["own_lo"], win["own_hi"], win["mem_lo"]
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(t[lo - m0: lo - m0 + HALO].clone(), rank - 1))
            up = torch.empty(HALO, arr.shape[1], dtype=torch.float64); reqs.append(dist.irecv(up, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(t[hi - HALO + 1 - m0: hi + 1 - m0].clone(), rank + 1))
            dn = torch.empty(HALO, arr.shape[1], dtype=torch.float64); reqs.append(dist.irecv(dn, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            t[lo - HALO - m0: lo - m0] = up
        if rank < world - 1:
            t[hi + 1 - m0: hi + 1 + HALO - m0] = dn

    # down leg: 3 RB iterations + residual + injection on the slab
    u1, crhs, _ = run(window(u), None, 3, 1)
    exchange(u1, w)                                     # the up leg reads the new iterate's halo rows
    # up leg: u += P(coarse u), 3 RB iterations, sum of squares of the residual; coarse window from the full array
    cu_win = np.ascontiguousarray(T.to_split(cu)[crows])
    u2, _, parts = run(u1, cu_win, 3, 2)
    tot = torch.tensor([parts.sum()], dtype=torch.float64)
    dist.all_reduce(tot)                                # solver.cu: comm_allreduce_sum of the norm
    own = slice(w["own_lo"] - w["mem_lo"], w["own_hi"] - w["mem_lo"] + 1)
    cown = slice(wc["own_lo"] - wc["mem_lo"], wc["own_hi"] - wc["mem_lo"] + 1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), u1=u1[own], u2=u2[own], crhs=crhs[cown], norm2=tot.numpy(),
             own=np.array([w["own_lo"], w["own_hi"], wc["own_lo"], wc["own_hi"]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ranks_reproduce_the_unsharded_pass_gloo(oracle, tmp_path, world):
    import torch.multiprocessing as mp
    import emu_util as T
    T.load_model()                      # built once here, loaded by the ranks
    n = 256
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    u, rhs, v1, v2 = (rng.standard_normal((n + 1, n + 1)) for _ in range(4))
    cu = rng.standard_normal((n // 2 + 1, n // 2 + 1)); cu[0, :] = cu[-1, :] = 0; cu[:, 0] = cu[:, -1] = 0
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want1 = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 3)
    wantc = oracle.restriction(oracle.residual(want1, rhs, n, v1, v2, dt, nu, dx), n)
    want2 = oracle.gauss_seidel(want1 + oracle.prolongation(cu, n // 2), rhs, n, v1, v2, dt, nu, dx, 3)
    wantn = oracle.norm(oracle.residual(want2, rhs, n, v1, v2, dt, nu, dx), n) ** 2
    got1 = np.full((n + 1, n + 1), np.nan); got2 = got1.copy(); gotc = np.zeros((n // 2 + 1, n // 2 + 1))
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi, clo, chi = (int(x) for x in z["own"])
        got1[lo: hi + 1] = T.from_split(np.vstack([np.zeros((lo, z["u1"].shape[1])), z["u1"], np.zeros((n - hi, z["u1"].shape[1]))]), n)[lo: hi + 1]
        got2[lo: hi + 1] = T.from_split(np.vstack([np.zeros((lo, z["u2"].shape[1])), z["u2"], np.zeros((n - hi, z["u2"].shape[1]))]), n)[lo: hi + 1]
        nc = n // 2
        gotc[clo: chi + 1] = T.from_split(np.vstack([np.zeros((clo, z["crhs"].shape[1])), z["crhs"], np.zeros((nc - chi, z["crhs"].shape[1]))]), nc)[clo: chi + 1]
        assert abs(float(z["norm2"][0]) - wantn) <= 1e-12 * wantn
    assert np.array_equal(got1, want1)
    assert np.array_equal(got2, want2)
    assert np.array_equal(gotc[1:-1, 1:-1], wantc[1:-1, 1:-1])
