"""Parity of the CUDA path (through the C ABI) with the committed reference vectors and the numpy oracle."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, product_mesh_and_kwargs, rel_l2

pytestmark = pytest.mark.gpu

TOL_MULT = 1e-12     # north_star: 1e-10 relative L2 per step; we hold two more digits
TOL_RUN = 1e-12


@pytest.fixture(scope="module")
def dg():
    import dgtd_b200
    return dgtd_b200


@pytest.mark.parametrize("name", golden_cases())
def test_mult_and_rk4_match_reference_vectors(dg, name):
    pb, dat = load_golden(name)
    meta = dat["meta"]
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    assert ev.Height() == dat["x0_f64"].size
    ev.SetTime(meta["t0"])
    k = ev.Mult(dat["x0_f64"])
    assert rel_l2(k, dat["k0_f64"]) < TOL_MULT
    ev.set_state(dat["x0_f64"])
    t = meta["t0"]
    for _ in range(meta["steps"]):
        t = ev.Step(t, meta["dt"])
    assert rel_l2(ev.get_state(), dat["x_final_f64"]) < TOL_RUN
    # the same through the batched loop entry point
    ev.set_state(dat["x0_f64"])
    ev.run(meta["t0"], meta["dt"], meta["steps"])
    assert rel_l2(ev.get_state(), dat["x_final_f64"]) < TOL_RUN
    assert ev.launch_count() > 0
    ev.close()


@pytest.mark.parametrize("name", ["box3d_p3_pec_upwind", "tfsf3d_p2_modulated", "box3d_p2_materials"])
def test_matches_numpy_oracle_on_seeded_inputs(dg, name):
    from oracle.dgtd_oracle import HesthavenOracle
    pb, dat = load_golden(name)
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    rng = np.random.default_rng(1234)
    for t in (0.0, 0.37):
        x = rng.standard_normal(6 * O.N)
        ev.SetTime(t)
        assert rel_l2(ev.Mult(x), O.mult(t, x)) < TOL_MULT
    x = rng.standard_normal(6 * O.N)
    ev.set_state(x)
    xo, t = x.copy(), 0.1
    for _ in range(3):
        xo = O.rk4_step(xo, t, 1e-3)
        t = ev.Step(t, 1e-3)
    assert rel_l2(ev.get_state(), xo) < TOL_RUN
    ev.close()


@pytest.mark.parametrize("name,nsteps", [("box3d_p3_pec_upwind", 1000), ("tfsf3d_p2_on", 300), ("box3d_p4_sma_partial", 500), ("config2_2d_pec_p3", 1000)])
def test_full_run_stays_within_the_run_tolerance(dg, name, nsteps):
    """north_star: fields within 1e-8 relative L2 of the reference after the FULL run.  Hundreds of fused RK4 steps on the
    device against the same number of oracle steps (round-off must not accumulate beyond 1e-10 here)."""
    from oracle.dgtd_oracle import HesthavenOracle
    pb, dat = load_golden(name)
    meta = dat["meta"]
    O = HesthavenOracle(pb)
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    x0, t0, dt = dat["x0_f64"], meta["t0"], meta["dt"]
    ev.set_state(x0)
    t_end, n, bad = ev.run_until(t0, dt, t0 + nsteps * dt - 1e-12, check_every=100)
    assert n == nsteps and not bad
    xo, t = x0.copy(), t0
    for _ in range(nsteps):
        xo = O.rk4_step(xo, t, dt)
        t += dt
    assert np.isfinite(xo).all() and np.linalg.norm(xo) > 0
    assert rel_l2(ev.get_state(), xo) < 1e-10
    ev.close()


def test_argument_errors_are_reported_not_executed(dg):
    """Wrong sizes, bad time steps and out-of-range probe dofs come back as DgtdError (the reference throws std::runtime_error)."""
    pb, dat = load_golden("box3d_p3_pec_upwind")
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    with pytest.raises(dg.DgtdError):
        ev.Mult(np.zeros(6 * ev.N - 1))
    with pytest.raises(dg.DgtdError):
        ev.set_state(np.zeros(5))
    with pytest.raises(dg.DgtdError):
        dg.Gather(ev, [ev.N])
    with pytest.raises(dg.DgtdError):
        ev.run_until(0.0, -1.0, 1.0)
    g = dg.Gather(ev, [])                      # empty probe list is legal and a no-op
    assert g.n_local == 0
    g.launch(np.zeros((6, 0)))
    g.wait()
    ev.close()
