"""The CUDA path at BASELINE.json's full sizes: one direct comparison with the portable oracle at the bench size, and
size-independent properties:
linearity of Mult at the bench size, exact curl of polynomial fields, the fused RK4 step against four Mult calls,
and the analytic PEC-cavity mode the reference's own solver tests use as behavioural pin
(test/maxwell/solver/Solver3DTest.cpp:57-102: tets, order 3, PEC box, returns to the analytic state)."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu

ORDER = 3


@pytest.fixture(scope="module")
def dg():
    import dgtd_b200
    return dgtd_b200


def _box(dg, cubes, alpha=1.0, order=ORDER):
    mesh = dg.Mesh.cartesian3d(cubes)
    return dg.Evolution(mesh, order=order, alpha=alpha, bdr={a: dg.BC_PEC for a in range(1, 7)})


def test_mult_is_linear_at_the_bench_size(dg):
    """config 5 per-GPU size: 32^3 cubes x 6 tets, order 3, 23.6 M DOFs.  Mult(a x + b y) = a Mult(x) + b Mult(y)."""
    ev = _box(dg, 32)
    assert 6 * ev.N == 23592960
    rng = np.random.default_rng(11)
    x = rng.standard_normal(6 * ev.N)
    y = rng.standard_normal(6 * ev.N)
    a, b = 0.75, -1.5
    ev.SetTime(0.0)
    kx = ev.Mult(x)
    ky = ev.Mult(y)
    x *= a
    x += b * y
    kl = ev.Mult(x)
    kx *= a
    kx += b * ky
    assert rel_l2(kl, kx) < 1e-13
    assert np.isfinite(kl).all() and np.linalg.norm(kl) > 0
    assert "stage_wg_kernel" in ev.kernel_info()
    ev.close()


def test_bench_size_mult_and_rk4_step_match_the_oracle(dg):
    """The bench workload itself (32^3 cubes x 6 tets, order 3, 23.6 M DOFs) against the portable oracle on the FULL vectors:
    one Mult and one fused RK4 step (the numpy oracle needs ~15 s of setup and ~4 s per Mult at this size)."""
    from oracle.dgtd_oracle import PEC, HesthavenOracle, Problem
    mesh = dg.Mesh.cartesian3d(32)
    v, e, ea, b, ba = mesh.arrays()
    O = HesthavenOracle(Problem(v, e.astype(np.int64), ea, b.astype(np.int64), ba, ORDER, 1.0, {a: PEC for a in range(1, 7)}))
    ev = dg.Evolution(mesh, order=ORDER, alpha=1.0, bdr={a: dg.BC_PEC for a in range(1, 7)})
    assert 6 * ev.N == 23592960 == 6 * O.N
    x = np.random.default_rng(5).standard_normal(6 * ev.N)
    ev.SetTime(0.0)
    assert rel_l2(ev.Mult(x), O.mult(0.0, x)) < 1e-12
    dt = 0.05 / 32 / 9
    ev.set_state(x)
    ev.Step(0.0, dt)
    assert rel_l2(ev.get_state(), O.rk4_step(x, 0.0, dt)) < 1e-12      # north_star: 1e-10 per step
    ev.close()


@pytest.mark.parametrize("order", [2, 3, 4])
def test_curl_of_polynomial_fields_is_exact_in_the_interior(dg, order):
    """Globally continuous polynomial fields of degree <= p have no jumps, so away from the PEC walls Mult must return
    (curl H, -curl E) at every node (config 3 size at order 3: 16^3 cubes x 6 tets, 2.9 M DOFs)."""
    cubes = 16 if order <= 3 else 10
    ev = _box(dg, cubes, order=order)
    N, Np = ev.N, ev.Np
    X = ev.node_coords()
    x, y, z = X[:, 0], X[:, 1], X[:, 2]
    p = order
    # E = (y^p + x z, z^p - x y, x^p + y z^(p-1)),  H = (z^p + x, x^(p-1) y, y^p - z x)
    E = [y ** p + x * z, z ** p - x * y, x ** p + y * z ** (p - 1)]
    H = [z ** p + x, x ** (p - 1) * y, y ** p - z * x]
    curlE = [z ** (p - 1) - p * z ** (p - 1), x - p * x ** (p - 1), -y - p * y ** (p - 1)]
    curlH = [p * y ** (p - 1) - 0.0, p * z ** (p - 1) + z, (p - 1) * x ** (p - 2) * y - 0.0]
    # d/dy Ez - d/dz Ey = z^(p-1) - p z^(p-1);  d/dz Ex - d/dx Ez = x - p x^(p-1);  d/dx Ey - d/dy Ex = -y - p y^(p-1)
    # d/dy Hz - d/dz Hy = p y^(p-1);  d/dz Hx - d/dx Hz = p z^(p-1) + z;  d/dx Hy - d/dy Hx = (p-1) x^(p-2) y
    u = np.concatenate(E + H)
    ev.SetTime(0.0)
    k = ev.Mult(u).reshape(6, N)
    want = np.stack(curlH + [-c for c in curlE])
    on_wall = ((np.abs(X) < 1e-12) | (np.abs(X - 1.0) < 1e-12)).any(axis=1).reshape(-1, Np).any(axis=1)
    interior = np.repeat(~on_wall, Np)
    assert interior.sum() > 0.5 * N
    err = np.linalg.norm((k - want)[:, interior]) / np.linalg.norm(want[:, interior])
    assert err < 1e-11, err
    ev.close()


def test_fused_rk4_step_equals_four_mults(dg):
    """mfem::RK4Solver::Step (ode.cpp:109-136) restated on the host with the product's own Mult, against the fused
    4-launch step on the resident state (16^3 cubes, 2.9 M DOFs)."""
    ev = _box(dg, 16)
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal(6 * ev.N) * 1e-2
    t, dt = 0.3, 2.0e-4
    f = lambda tt, v: (ev.SetTime(tt), ev.Mult(v))[1]
    k1 = f(t, x0)
    k2 = f(t + dt / 2, x0 + dt / 2 * k1)
    k3 = f(t + dt / 2, x0 + dt / 2 * k2)
    k4 = f(t + dt, x0 + dt * k3)
    want = x0 + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    ev.set_state(x0)
    ev.Step(t, dt)
    assert rel_l2(ev.get_state(), want) < 1e-14
    # run(n) == n x Step
    ev.set_state(x0)
    ev.run(t, dt, 3)
    a = ev.get_state()
    ev.set_state(x0)
    tt = t
    for _ in range(3):
        tt = ev.Step(tt, dt)
    assert np.array_equal(a, ev.get_state())
    ev.close()


@pytest.mark.parametrize("alpha", [1.0, 0.0])
def test_pec_cavity_mode_follows_the_analytic_solution(dg, alpha):
    """TM110 mode of the unit PEC box, E_z = sin(pi x) sin(pi y) cos(w t), w = pi sqrt(2): after 1/4 period the state must
    be the analytic one to discretisation accuracy (h = 1/16, order 3), upwind and centred flux."""
    ev = _box(dg, 16, alpha=alpha)
    N = ev.N
    X = ev.node_coords()
    x, y = X[:, 0], X[:, 1]
    w = np.pi * np.sqrt(2.0)
    u0 = np.zeros((6, N))
    u0[2] = np.sin(np.pi * x) * np.sin(np.pi * y)
    T = 0.25 * 2 * np.pi / w
    nsteps = 800
    dt = T / nsteps
    ev.set_state(u0.ravel())
    ev.run(0.0, dt, nsteps)
    u = ev.get_state().reshape(6, N)
    want = np.zeros((6, N))
    want[2] = np.sin(np.pi * x) * np.sin(np.pi * y) * np.cos(w * T)
    want[3] = -(np.pi / w) * np.sin(np.pi * x) * np.cos(np.pi * y) * np.sin(w * T)
    want[4] = (np.pi / w) * np.cos(np.pi * x) * np.sin(np.pi * y) * np.sin(w * T)
    err = np.linalg.norm(u - want) / np.linalg.norm(want)
    # the numpy oracle gives 2.7e-2 / 1.7e-3 (upwind) and 5.6e-2 / 5.2e-3 (centred) at 2^3 / 4^3 cubes: order h^4 and ~h^3.5
    assert err < (2e-5 if alpha == 1.0 else 2e-4), err
    ev.close()


def test_run_until_follows_the_reference_time_loop(dg):
    """Solver::run / Solver::step (Solver.cpp:497-551): steps of min(dt, T - t) while t <= T - 1e-8 dt, stability test on
    the state norm."""
    ev = _box(dg, 4)
    rng = np.random.default_rng(2)
    x0 = rng.standard_normal(6 * ev.N) * 1e-2
    dt, T = 1.0e-3, 0.0105                      # 10 full steps and one of half the size
    ev.set_state(x0)
    t, n, bad = ev.run_until(0.0, dt, T, check_every=1)
    assert n == 11 and not bad and abs(t - T) < 1e-15
    a = ev.get_state()
    ev.set_state(x0)
    tt = 0.0
    for _ in range(10):
        tt = ev.Step(tt, dt)
    ev.Step(tt, T - tt)
    assert np.array_equal(a, ev.get_state())
    # far beyond the RK4 stability limit: the norm test must fire and stop the loop
    ev.set_state(x0)
    t, n, bad = ev.run_until(0.0, 0.5, 200.0, check_every=1)
    assert bad and n < 400
    ev.close()
