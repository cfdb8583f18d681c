"""C-ABI grid operators (include/mgb200.h, replacing gs.h:3-17 / gscu.h:3-16) on a B200 against the
CPU oracle, same seeded inputs.  EXACT arithmetic: bit-for-bit.  FAST arithmetic (FMA + reciprocal
diagonal): relative error <= 1e-13 per operator application."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def mg():
    import hpcclassmultigridproject_b200 as m
    m.lib()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return m


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def fields(n, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((n + 1, n + 1)) for _ in range(4)]


SIZES = [5, 32, 64, 100, 257, 512]


@pytest.mark.parametrize("n", SIZES)
def test_gauss_seidel_exact_bitwise(mg, oracle, n):
    u, rhs, v1, v2 = fields(n, n)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    for iters in (1, 3):
        want = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, iters)
        got = mg.ops.gauss_seidel(dev(u), dev(rhs), n, dev(v1), dev(v2), dt, nu, dx, iters, mg.ARITH_EXACT)
        assert np.array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("n", SIZES)
def test_gauss_seidel_fast_close(mg, oracle, n):
    u, rhs, v1, v2 = fields(n, n + 1)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 3)
    got = mg.ops.gauss_seidel(dev(u), dev(rhs), n, dev(v1), dev(v2), dt, nu, dx, 3, mg.ARITH_FAST).cpu().numpy()
    assert np.linalg.norm(got - want) <= 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize("n", SIZES)
def test_residual_rhs_norm(mg, oracle, n):
    u, rhs, v1, v2 = fields(n, 3 * n)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want_rhs = oracle.compute_rhs(u, n, v1, v2, dt, nu, dx)
    want_res = oracle.residual(u, rhs, n, v1, v2, dt, nu, dx)
    want_norm = oracle.norm(want_res, n)
    du, dr, d1, d2 = dev(u), dev(rhs), dev(v1), dev(v2)
    for arith, exact in ((mg.ARITH_EXACT, True), (mg.ARITH_FAST, False)):
        got_rhs = mg.ops.compute_rhs(torch.zeros_like(du), du, n, d1, d2, dt, nu, dx, arith).cpu().numpy()
        got_res = mg.ops.residual(torch.zeros_like(du), du, dr, n, d1, d2, dt, nu, dx, arith)
        if exact:
            assert np.array_equal(got_rhs, want_rhs)
            assert np.array_equal(got_res.cpu().numpy(), want_res)
        else:
            assert np.linalg.norm(got_rhs - want_rhs) <= 1e-14 * np.linalg.norm(want_rhs)
            assert np.linalg.norm(got_res.cpu().numpy() - want_res) <= 1e-14 * np.linalg.norm(want_res)
        # norm: interior only, like gs.cpp:86-107; summation order differs from the serial sum
        got_res[0, :] = 7.0; got_res[:, 0] = -3.0      # boundary garbage must be ignored
        assert abs(mg.ops.compute_norm(got_res, n) - want_norm) <= 1e-13 * want_norm
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        mg.ops.residual_norm2(None, du, dr, n, d1, d2, dt, nu, dx, out, arith)
        assert abs(float(out.cpu()[0]) ** 0.5 - want_norm) <= 1e-13 * want_norm


@pytest.mark.parametrize("nc", [1, 5, 16, 33, 128])
def test_transfer_operators_bitwise(mg, oracle, nc):
    rng = np.random.default_rng(nc)
    c = rng.standard_normal((nc + 1, nc + 1))
    want = oracle.prolongation(c, nc)
    got = mg.ops.prolongation(torch.zeros(2 * nc + 1, 2 * nc + 1, dtype=torch.float64, device="cuda"), dev(c), nc)
    assert np.array_equal(got.cpu().numpy(), want)
    base = rng.standard_normal((2 * nc + 1, 2 * nc + 1))
    got2 = mg.ops.prolong_correct(dev(base), dev(c), nc).cpu().numpy()
    assert np.array_equal(got2, base + want)          # multigrid.cpp:81-83
    back = mg.ops.restriction(torch.zeros(nc + 1, nc + 1, dtype=torch.float64, device="cuda"), dev(want), 2 * nc)
    assert np.array_equal(back.cpu().numpy(), c)      # restriction o prolongation = identity (prolrestest.cpp)
    assert np.array_equal(back.cpu().numpy(), oracle.restriction(want, 2 * nc))
    s = mg.ops.vecadd(torch.zeros_like(dev(base)), dev(base), dev(want), 2 * nc).cpu().numpy()
    assert np.array_equal(s, base + want)


@pytest.mark.parametrize("nf", [2, 4, 32, 130, 512])
def test_full_weighting_restriction_opt_in(mg, oracle, nf):
    """gs.cpp:277-280 (commented out in the reference): the opt-in operator against its C restatement"""
    rng = np.random.default_rng(nf)
    fine = rng.standard_normal((nf + 1, nf + 1))
    got = mg.ops.restriction_fw(torch.zeros(nf // 2 + 1, nf // 2 + 1, dtype=torch.float64, device="cuda"), dev(fine), nf)
    assert np.array_equal(got.cpu().numpy(), oracle.restriction_fw(fine, nf))


def test_known_answer_ramp(mg):
    """prolrestest.cpp:76-118"""
    n = 5
    up = np.add.outer(np.arange(n + 1.0), np.arange(n + 1.0))
    fine = mg.ops.prolongation(torch.zeros(2 * n + 1, 2 * n + 1, dtype=torch.float64, device="cuda"), dev(up), n)
    assert np.array_equal(fine.cpu().numpy(), np.add.outer(np.arange(2 * n + 1.0), np.arange(2 * n + 1.0)) / 2)


@pytest.mark.parametrize("name", ["ops_n32.npz", "ops_n64.npz"])
def test_golden_vectors(mg, name):
    g = golden(name)
    n = int(g["n"]); dx, dt, nu = float(g["dx"]), float(g["dt"]), float(g["nu"])
    u, rhs, v1, v2 = (dev(g[k]) for k in ("u", "rhs", "v1", "v2"))
    E = mg.ARITH_EXACT
    assert np.array_equal(mg.ops.compute_rhs(torch.zeros_like(u), u, n, v1, v2, dt, nu, dx, E).cpu().numpy(), g["compute_rhs"])
    assert np.array_equal(mg.ops.residual(torch.zeros_like(u), u, rhs, n, v1, v2, dt, nu, dx, E).cpu().numpy(), g["residual"])
    assert np.array_equal(mg.ops.gauss_seidel(u.clone(), rhs, n, v1, v2, dt, nu, dx, 1, E).cpu().numpy(), g["gs1"])
    assert np.array_equal(mg.ops.gauss_seidel(u.clone(), rhs, n, v1, v2, dt, nu, dx, 3, E).cpu().numpy(), g["gs3"])
    assert np.array_equal(mg.ops.restriction(torch.zeros(n // 2 + 1, n // 2 + 1, dtype=torch.float64, device="cuda"), u, n).cpu().numpy(), g["restriction"])
    assert np.array_equal(mg.ops.prolongation(torch.zeros(2 * n + 1, 2 * n + 1, dtype=torch.float64, device="cuda"), u, n).cpu().numpy(), g["prolongation"])
    assert abs(mg.ops.compute_norm(dev(g["residual"]), n) - float(g["norm"])) <= 1e-13 * float(g["norm"])


def test_padded_leading_dimension(mg, oracle):
    """explicit ld > n+1: the operators must honour the row stride"""
    n = 64
    u, rhs, v1, v2 = fields(n, 11)
    dx = 1.0 / n; dt = dx / 10; nu = -4e-4
    want = oracle.gauss_seidel(u.copy(), rhs, n, v1, v2, dt, nu, dx, 2)

    def padded(a):
        t = torch.full((n + 1, n + 16), float("nan"), dtype=torch.float64, device="cuda")
        t[:, : n + 1] = dev(a)
        return t[:, : n + 1]

    got = mg.ops.gauss_seidel(padded(u), padded(rhs), n, padded(v1), padded(v2), dt, nu, dx, 2, mg.ARITH_EXACT)
    assert np.array_equal(got.cpu().numpy(), want)


def test_device_initial_conditions(mg, oracle):
    n = 128
    u0, v1, v2 = oracle.initial_conditions(n, 2.0)
    d = [torch.zeros(n + 1, n + 1, dtype=torch.float64, device="cuda") for _ in range(3)]
    mg.ops.initial_conditions(*d, n, 2.0)
    for got, want in zip(d, (u0, v1, v2)):
        assert np.abs(got.cpu().numpy() - want).max() <= 4e-15 * max(1.0, np.abs(want).max())
    assert d[0][n, 0].item() > 0 and d[0][0, n].item() == 0      # multigrid.cpp:227-233 quirk


def test_bad_arguments_fail_loudly(mg):
    t = torch.zeros(9, 9, dtype=torch.float64, device="cuda")
    with pytest.raises(mg.MgError):
        mg.ops.restriction(t, t, 7)          # odd fine n
    with pytest.raises(mg.MgError):
        mg.Solver(48, -4e-4, 1e-3, 1 / 48, 1e-6, maxlvl=1)   # not a power of two
