"""BASELINE.json configs 3 and 4 on the reference's own meshes: the CUDA path against the sampled reference vectors
(tests/golden/*.cfg.npz, produced by the reference-based oracle) and against the portable oracle on the full vectors."""
import numpy as np
import pytest

from conftest import config_cases, initial_state, load_config_case, product_mesh_and_kwargs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order_override", [None, 4])
@pytest.mark.parametrize("name", config_cases())
def test_baseline_config_matches_reference_and_oracle(name, order_override):
    import dgtd_b200 as dg
    from oracle.dgtd_oracle import HesthavenOracle
    pb, meta, smp = load_config_case(name)
    if order_override is not None:
        if not meta.get("tfsf"):
            pytest.skip("order override only for the RCS case (BASELINE config 4 is quoted at order 4)")
        pb.order = order_override
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    x0 = initial_state(meta, ev.node_coords())
    st, t0, dt = meta["stride"], meta["t0"], meta["dt"] * (0.5 if order_override else 1.0)
    # north_star: 1e-10 relative L2 per step.  We hold 1e-12 at the fixtures' order; at order 4 the gmsh sphere mesh's worst
    # tetrahedra put 114 of 15 886 elements at 1e-9 absolute between two FP64 evaluation orders (measured 2.2e-12 overall)
    tol = 1e-12 if order_override is None else 1e-11
    O = HesthavenOracle(pb)
    ev.SetTime(t0)
    k = ev.Mult(x0)
    assert rel_l2(k, O.mult(t0, x0)) < tol
    ev.set_state(x0)
    ev.run(t0, dt, meta["steps"])
    x = ev.get_state()
    xo, t = x0, t0
    for _ in range(meta["steps"]):
        xo = O.rk4_step(xo, t, dt)
        t += dt
    assert rel_l2(x, xo) < tol
    if order_override is None:      # the stored samples are at the fixture's own order
        assert rel_l2(k[::st], smp["k0_sample_f64"]) < 1e-12
        assert rel_l2(x[::st], smp["x_final_sample_f64"]) < 1e-12
        assert abs(np.linalg.norm(x) / meta["x_final_norm"] - 1.0) < 1e-12
    ev.close()
