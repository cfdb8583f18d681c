"""The V/W-cycle driver (mgb200_solver: mg_inner / mg_outer / timestepper of multigrid.cpp:17-186)
on a B200 against the CPU oracle and the golden vectors of the compiled reference.

Bars (BASELINE.json north_star): final u within 1e-10 relative L2 with the SAME cycle counts.
EXACT arithmetic is held to bit-for-bit equality of u on every level after every cycle."""
import os

import numpy as np
import pytest

from conftest import golden
from oracle.oracle import OracleSolver

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def mg():
    import hpcclassmultigridproject_b200 as m
    m.lib()
    assert torch.cuda.is_available()
    return m


PLANS = os.environ.get("MGB200_TEST_PLANS", "unfused,fused").split(",")


def plan_id(mg, name):
    return mg.PLAN_UNFUSED if name == "unfused" else mg.PLAN_FUSED


def rel_l2(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("n,shape,vscale,nu", [(32, 1, 1.0, -4e-4), (64, 1, 1.0, -4e-4), (128, 2, 2.0, -4e-4),
                                               (256, 1, 5.0, -1e-5), (512, 1, 1.0, -4e-4), (1024, 2, 1.0, -4e-4)])
def test_cycle_by_cycle_exact(mg, oracle, plan, n, shape, vscale, nu):
    """every level array after every cycle, bit-for-bit (EXACT arithmetic)"""
    u0, v1, v2 = oracle.initial_conditions(n, vscale)
    dx = 1.0 / n; dt = dx / 10
    o = OracleSolver(n, u0, v1, v2, nu, dt, dx, 1e-12, shape)
    with mg.Solver(n, nu, dt, dx, 1e-12, shape=shape, arith=mg.ARITH_EXACT, plan=plan_id(mg, plan)) as s:
        s.set_fields_host(u0, v1, v2)
        for l in range(s.maxlvl):                       # coarse velocity towers incl. the P1 quirk
            assert np.array_equal(s.level(l, "v1"), o.v1(l)) and np.array_equal(s.level(l, "v2"), o.v2(l))
        for step in range(2):
            o.form_rhs()
            r0 = s.form_rhs()
            assert np.array_equal(s.level(0, "rhs")[1:-1, 1:-1], o.rhs(0)[1:-1, 1:-1])
            assert abs(r0 - o.residual_norm()) <= 1e-12 * r0
            for _ in range(2):
                o.cycle()
                r = s.cycle()
                for l in range(s.maxlvl):
                    assert np.array_equal(s.level(l, "u"), o.u(l)), (step, l)
                    if l > 0:
                        assert np.array_equal(s.level(l, "rhs")[1:-1, 1:-1], o.rhs(l)[1:-1, 1:-1]), (step, l)
                ro = o.residual_norm()
                assert abs(r - ro) <= 1e-9 * ro + 1e-300
    o.close()


SOLVES = ["solve_n32", "solve_n64", "solve_n128", "solve_n256", "solve_cfg_wcycle_n128", "solve_cfg_tight_n256",
          "solve_cfg_adv_n256", "solve_cfg_diff_n256", "solve_cfg_n512", "solve_cfg_n1024"]


@pytest.mark.parametrize("plan", PLANS)
@pytest.mark.parametrize("arith", ["exact", "fast"])
@pytest.mark.parametrize("tag", SOLVES)
def test_time_steps_against_golden(mg, oracle, tag, arith, plan):
    """golden vectors of the compiled reference: cycle counts per step, residual histories, final field"""
    g = golden(tag + ".npz")
    n, steps = int(g["n"]), int(g["steps"])
    u0, v1, v2 = oracle.initial_conditions(n, float(g["vscale"]))
    A = mg.ARITH_EXACT if arith == "exact" else mg.ARITH_FAST
    with mg.Solver(n, float(g["nu"]), float(g["dt"]), float(g["dx"]), float(g["tol"]), shape=int(g["shape"]),
                   arith=A, plan=plan_id(mg, plan)) as s:
        s.set_fields_host(u0, v1, v2)
        infos = s.timestep(steps)
        for k, info in enumerate(infos):
            assert info.cycles == int(g["cycles"][k]), (k, info.history())
            want = g["hist"][k][: info.cycles + 1]
            got = np.array(info.history())
            # exact: the norms differ only by summation order.  fast: u moves by a few ulp, which
            # shifts an already tiny residual by ~eps*||u|| in absolute terms
            floor = 0.0 if arith == "exact" else 1e-12 * want[0]
            assert np.all(np.abs(got - want) <= 1e-9 * want + floor), (got, want)
        uT = s.get_u_host()
    if "uT" in g:
        if arith == "exact":
            assert np.array_equal(uT, g["uT"])
        else:
            assert rel_l2(uT, g["uT"]) <= 1e-10
    assert abs(np.linalg.norm(uT) - float(g["norm_uT"])) <= 1e-10 * float(g["norm_uT"])
    assert abs(uT[n // 2, n // 2] - float(g["mid"])) <= 1e-10 * abs(float(g["mid"]))


@pytest.mark.parametrize("plan", PLANS)
def test_timestepper_entry_point(mg, oracle, plan):
    """mgb200_timestepper_host = timestepper(uT,u0,v1,v2,nu,maxlvl,n,dt,T,dx,tol,shape), multigrid.cpp:124"""
    g = golden("solve_n256.npz")        # BASELINE.json configs[0]: N=256, 10 implicit steps
    n = 256; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n)
    uT = np.zeros_like(u0)
    info = mg.timestepper_host(uT, u0, v1, v2, -4e-4, mg.maxlvl_for(n), n, dt, 10 * dt, dx, 1e-6, 1,
                               arith=mg.ARITH_EXACT, plan=plan_id(mg, plan))
    assert info.cycles == 1 and np.array_equal(uT, g["uT"])
    uT2 = np.zeros_like(u0)
    mg.timestepper_host(uT2, u0, v1, v2, -4e-4, mg.maxlvl_for(n), n, dt, 10 * dt, dx, 1e-6, 1, plan=plan_id(mg, plan))
    assert rel_l2(uT2, g["uT"]) <= 1e-10
    # inputs are copied, not aliased or modified (multigrid.cpp:143-145)
    u0b, _, _ = oracle.initial_conditions(n)
    assert np.array_equal(u0, u0b)


def test_timestepper_frees_its_towers_on_request(mg, oracle, monkeypatch):
    """MGB200_TIMESTEPPER_CACHE=0: the one-call driver keeps no handle alive (multigrid.cpp:177-185 frees everything)"""
    import torch
    n = 1024; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n)
    uT = np.zeros_like(u0)
    mg.release_cached()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    mg.timestepper_host(uT, u0, v1, v2, -4e-4, mg.maxlvl_for(n), n, dt, 2 * dt, dx, 1e-6, 1)
    kept = free0 - torch.cuda.mem_get_info()[0]
    assert kept > 5 * 8 * (n + 1) ** 2            # the towers of the cached handle
    mg.release_cached()
    monkeypatch.setenv("MGB200_TIMESTEPPER_CACHE", "0")
    uT2 = np.zeros_like(u0)
    mg.timestepper_host(uT2, u0, v1, v2, -4e-4, mg.maxlvl_for(n), n, dt, 2 * dt, dx, 1e-6, 1)
    assert free0 - torch.cuda.mem_get_info()[0] < kept // 4
    assert np.array_equal(uT, uT2)


@pytest.mark.parametrize("plan", PLANS)
def test_graph_replay_equals_direct_launch(mg, oracle, plan):
    n = 256; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n, 3.0)
    outs = []
    for use_graph in (1, 0):
        with mg.Solver(n, -4e-4, dt, dx, 1e-10, arith=mg.ARITH_EXACT, plan=plan_id(mg, plan), use_graph=use_graph) as s:
            s.set_fields_host(u0, v1, v2)
            infos = s.timestep(3)
            outs.append((s.get_u_host(), [i.cycles for i in infos], s.kernel_launches))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    # the graph path runs mg_outer's loop on the device: one k_loop_begin per solve and one
    # k_loop_check per cycle on top of the direct path's kernels
    extra = len(outs[0][1]) + sum(outs[0][1])
    assert outs[0][2] == outs[1][2] + extra and outs[1][2] > 0


@pytest.mark.skipif("fused" not in PLANS or "unfused" not in PLANS, reason="needs both plans")
def test_fused_equals_unfused_bitwise(mg, oracle):
    """the streaming kernels do the same arithmetic per node as the one-operator kernels"""
    for n, shape in ((64, 1), (256, 2), (2048, 1)):
        dx = 1.0 / n; dt = dx / 10
        res = []
        for plan in (mg.PLAN_UNFUSED, mg.PLAN_FUSED):
            with mg.Solver(n, -4e-4, dt, dx, 1e-12, shape=shape, arith=mg.ARITH_EXACT, plan=plan) as s:
                s.set_fields_reference_ic(2.0)
                s.form_rhs()
                norms = [s.cycle() for _ in range(3)]
                res.append(([s.level(l, "u") for l in range(s.maxlvl)], norms))
        for a, b in zip(res[0][0], res[1][0]):
            assert np.array_equal(a, b)
        assert np.allclose(res[0][1], res[1][1], rtol=1e-9, atol=0)


@pytest.mark.skipif("fused" not in PLANS or "unfused" not in PLANS, reason="needs both plans")
@pytest.mark.parametrize("niter", [0, 1, 2, 4, 5, 7])
def test_smoothing_counts_other_than_the_reference_default(mg, niter):
    """options.niter != 3 (the reference hard-codes NITER = 3, multigrid.cpp:41): the fused plan cuts the sweeps into
    passes of at most three iterations (4 = 3+1, 7 = 3+3+1) and must equal the one-operator plan bit for bit;
    niter = 0 has nothing to fuse and runs the one-operator plan"""
    n = 256; dx = 1.0 / n; dt = dx / 10
    res = []
    for plan in (mg.PLAN_UNFUSED, mg.PLAN_FUSED):
        with mg.Solver(n, -4e-4, dt, dx, 1e-12, shape=1, arith=mg.ARITH_EXACT, plan=plan, niter=niter) as s:
            s.set_fields_reference_ic(2.0)
            s.form_rhs()
            norms = [s.cycle() for _ in range(2)]
            res.append(([s.level(l, "u") for l in range(s.maxlvl)], norms))
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a, b)
    assert np.allclose(res[0][1], res[1][1], rtol=1e-9, atol=0)


@pytest.mark.parametrize("plan", PLANS)
def test_corrected_velocity_towers_option(mg, oracle, plan):
    """options.correct_towers = 1 (opt-in; the reference's towers, multigrid.cpp:148-160, never halve n): every coarse
    velocity level is the plain subsampling of level 0, and the cycle equals the oracle's with the same towers bit for
    bit; with the vortex of the reference the two tower variants give different coarse operators, hence different
    iterates (so the option is not a no-op) but the same converged solution"""
    n = 256; dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    u0, v1, v2 = oracle.initial_conditions(n, 3.0)
    o = OracleSolver(n, u0, v1, v2, nu, dt, dx, 1e-12, 1, correct_towers=True)
    with mg.Solver(n, nu, dt, dx, 1e-12, arith=mg.ARITH_EXACT, plan=plan_id(mg, plan), correct_towers=1) as s:
        s.set_fields_host(u0, v1, v2)
        for l in range(s.maxlvl):
            assert np.array_equal(s.level(l, "v1"), v1[:: 1 << l, :: 1 << l]) and np.array_equal(s.level(l, "v2"), v2[:: 1 << l, :: 1 << l])
            assert np.array_equal(o.v1(l), v1[:: 1 << l, :: 1 << l])
        o.form_rhs(); s.form_rhs()
        for _ in range(2):
            o.cycle(); s.cycle()
            for l in range(s.maxlvl):
                assert np.array_equal(s.level(l, "u"), o.u(l)), l
        corrected = s.get_u_host()
    with mg.Solver(n, nu, dt, dx, 1e-12, arith=mg.ARITH_EXACT, plan=plan_id(mg, plan)) as s:
        s.set_fields_host(u0, v1, v2)
        s.form_rhs()
        for _ in range(2):
            s.cycle()
        reference_towers = s.get_u_host()
    assert not np.array_equal(corrected, reference_towers)
    assert rel_l2(corrected, reference_towers) <= 1e-6          # two convergent iterations for the same fine-grid system
    o.close()


@pytest.mark.parametrize("n,shape", [(64, 1), (256, 2), (1024, 1)])
def test_exact_coarse_solve_option(mg, oracle, n, shape):
    """options.coarse_exact = 1 (opt-in; what the reference's unfinished exact_solve.cpp:1-55 set out to do): the
    coarsest level is solved directly by a banded LU without pivoting.  Bit-identical to the oracle's restatement of
    the same factorisation on every level; the coarsest residual drops to rounding level; the outer iteration
    converges to the same solution as the reference's iterated coarse solve in no more cycles."""
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    u0, v1, v2 = oracle.initial_conditions(n, 2.0)
    o = OracleSolver(n, u0, v1, v2, nu, dt, dx, 1e-12, shape, coarse_exact=True)
    with mg.Solver(n, nu, dt, dx, 1e-12, shape=shape, arith=mg.ARITH_EXACT, coarse_exact=1) as s:
        s.set_fields_host(u0, v1, v2)
        o.form_rhs(); s.form_rhs()
        for _ in range(2):
            o.cycle(); s.cycle()
            for l in range(s.maxlvl):
                assert np.array_equal(s.level(l, "u"), o.u(l)), l
        lc = s.maxlvl - 1
        uc, fc, w1, w2 = (s.level(lc, k) for k in ("u", "rhs", "v1", "v2"))
        nc = n >> lc
        res = oracle.residual(uc, fc, nc, w1, w2, dt, nu, dx * (1 << lc))
        assert oracle.norm(res, nc) <= 1e-13 * max(1e-300, np.abs(fc).max()) * nc
    o.close()
    outs = []
    for exact in (0, 1):
        with mg.Solver(n, nu, dt, dx, 1e-9, shape=shape, coarse_exact=exact) as s:
            s.set_fields_host(u0, v1, v2)
            infos = s.timestep(2)
            outs.append((s.get_u_host(), [i.cycles for i in infos]))
    assert rel_l2(outs[1][0], outs[0][0]) <= 1e-8
    assert all(a <= b for a, b in zip(outs[1][1], outs[0][1]))


def test_full_weighting_option(mg, oracle):
    """options.restriction = 1 (opt-in, UNFUSED plan): converges to the same solution as the reference's
    injection; rejected on the fused plan"""
    n = 256; dx = 1.0 / n; dt = dx / 10
    u0, v1, v2 = oracle.initial_conditions(n, 2.0)
    out = []
    for restriction in (0, 1):
        with mg.Solver(n, -4e-4, dt, dx, 1e-10, plan=mg.PLAN_UNFUSED, restriction=restriction) as s:
            s.set_fields_host(u0, v1, v2)
            infos = s.timestep(3)
            assert all(i.converged for i in infos)
            out.append(s.get_u_host())
    assert rel_l2(out[1], out[0]) <= 1e-8          # two convergent iterations for the same linear system
    with pytest.raises(Exception):
        mg.Solver(n, -4e-4, dt, dx, 1e-10, plan=mg.PLAN_FUSED, restriction=1)


LARGE = ["large_c2_n4096", "large_n2048_v6", "large_n8192", "large_c3_n16384", "large_c4_n16384"]


@pytest.mark.parametrize("arith", ["fast", "exact"])
@pytest.mark.parametrize("tag", LARGE)
def test_full_size_configs(mg, tag, arith):
    """BASELINE.json configs[1..3] at full size against scalars + strided samples of the compiled
    reference (tests/golden/make_golden_large.py).  ICs are generated on the device (libdevice
    exp/sin/cos differ from glibc by <= 2 ulp, far inside the 1e-10 bar)."""
    g = golden(tag + ".npz")
    n = int(g["n"])
    A = mg.ARITH_EXACT if arith == "exact" else mg.ARITH_FAST
    with mg.Solver(n, float(g["nu"]), float(g["dt"]), float(g["dx"]), float(g["tol"]), arith=A,
                   plan=plan_id(mg, PLANS[-1])) as s:
        s.set_fields_reference_ic(float(g["vscale"]))
        info = s.timestep(1)[0]
        want = g["hist"]
        assert info.cycles == int(g["cycles"]), info.history()
        got = np.array(info.history())
        assert np.all(np.abs(got - want) <= 1e-3 * want + 3e-12 * want[0]), (got, want)
        out = torch.empty(n + 1, n + 1, dtype=torch.float64, device="cuda")
        s.get_u_device(out)
        st = int(g["stride"])
        sample = out[::st, ::st].cpu().numpy()
        assert rel_l2(sample, g["sample"]) <= 1e-10
        assert abs(float(torch.linalg.vector_norm(out)) - float(g["norm_uT"])) <= 1e-10 * float(g["norm_uT"])
        assert abs(out[n // 2, n // 2].item() - float(g["mid"])) <= 1e-10 * abs(float(g["mid"]))
        # size-independent properties: the boundary is untouched and one more cycle is a contraction
        assert out[0, :].abs().max().item() == 0 and out[:, n].abs().max().item() == 0
        r_before = info.res
        r_after = s.cycle()
        assert r_after <= max(r_before, 1e-11 * info.res0)


def test_linearity_of_the_cycle(mg):
    """size-independent property: with a zero initial guess the V-cycle is linear in the rhs.
    Scaling u0 by 2 scales rhs = B u0, every iterate and every norm by exactly 2 (power of two)."""
    n = 1024; dx = 1.0 / n; dt = dx / 10
    base = None
    for scale in (1.0, 2.0):
        # fixed coarsest iteration count: its ABSOLUTE stopping test (multigrid.cpp:60) is not linear
        with mg.Solver(n, -4e-4, dt, dx, 1e-12, arith=mg.ARITH_EXACT, coarse_tol=0.0, coarse_maxit=4,
                       plan=plan_id(mg, PLANS[-1])) as s:
            d = [torch.zeros(n + 1, n + 1, dtype=torch.float64, device="cuda") for _ in range(3)]
            mg.ops.initial_conditions(*d, n, 1.0)
            d[0] *= scale
            s.set_fields_device(*d)
            s.form_rhs()
            s.cycle()
            u = s.get_u_host()
        if base is None:
            base = u
        else:
            assert np.array_equal(u, 2.0 * base)


@pytest.mark.parametrize("swk", [32, 64, 96])
def test_every_strip_width_gives_the_same_field(swk, tmp_path):
    """MGB200_SWK pins the strip width of the streaming pass (one warp per role for <= 64 pairs, a partly filled second
    warp at 96): the result does not depend on it, bit for bit (exact arithmetic, two time steps at N=1024)"""
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import hpcclassmultigridproject_b200 as mg\n"
        "from oracle.oracle import Oracle\n"
        "n = 1024; dx = 1.0 / n; dt = dx / 10\n"
        "u0, v1, v2 = Oracle().initial_conditions(n)\n"
        "with mg.Solver(n, -4e-4, dt, dx, 1e-10, arith=mg.ARITH_EXACT, plan=mg.PLAN_FUSED) as s:\n"
        "    s.set_fields_host(u0, v1, v2)\n"
        "    infos = s.timestep(2)\n"
        "    print(hashlib.sha256(s.get_u_host().tobytes()).hexdigest(), [i.cycles for i in infos])\n" % ROOT)
    outs = []
    for env_swk in (None, str(swk)):
        env = dict(os.environ)
        env.pop("MGB200_SWK", None)
        if env_swk:
            env["MGB200_SWK"] = env_swk
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip().splitlines()[-1])
    assert outs[0] == outs[1]
