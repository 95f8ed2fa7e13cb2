"""Asynchronous probe / surface-export gather (dgtd_gather_*): snapshots of a dof list taken while the time loop runs."""
import numpy as np
import pytest

from conftest import load_golden, product_mesh_and_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dg():
    import dgtd_b200
    return dgtd_b200


@pytest.mark.parametrize("name", ["box3d_p3_pec_upwind", "tri2d_p3_mixed", "seg1d_config1_pec", "tfsf3d_p2_on"])
def test_gather_snapshots_equal_the_state_at_launch_time(dg, name):
    import torch
    pb, dat = load_golden(name)
    meta = dat["meta"]
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    N = ev.N
    rng = np.random.default_rng(3)
    dofs = rng.choice(N, size=min(N, 257), replace=False)          # unsorted, ragged count
    g = dg.Gather(ev, dofs)
    assert g.n_local == len(dofs) and np.array_equal(g.dofs, dofs)
    snaps = [torch.empty(6 * g.n_local, dtype=torch.float64).pin_memory().numpy().reshape(6, -1) for _ in range(3)]
    ev.set_state(dat["x0_f64"])
    t = meta["t0"]
    states = []
    for s in snaps:
        states.append(ev.get_state().reshape(6, N)[:, dofs])
        g.launch(s)                      # snapshot now ...
        t = ev.Step(t, meta["dt"])       # ... while the loop goes on
        t = ev.Step(t, meta["dt"])
        g.wait()
        assert np.array_equal(s, states[-1])
    assert not np.array_equal(snaps[0], snaps[2])
    g.close()
    ev.close()


def test_surface_export_gather_of_the_tfsf_box(dg):
    """All dofs of the elements on the inner side of a tagged interior surface (RCSSurfaceExporter's sub-mesh fields,
    RCSSurfaceExporter.cpp:71-79), gathered every step of a short run and compared with full-state downloads."""
    pb, dat = load_golden("tfsf3d_p2_on")
    meta = dat["meta"]
    mesh, kw = product_mesh_and_kwargs(pb)
    ev = dg.Evolution(mesh, **kw)
    N, Np = ev.N, ev.Np
    pairs = mesh.boundary_elements(pb.tfsf_tags)
    inner = np.unique(pairs[0::2, 0])                             # Elem1 of every face
    dofs = (inner[:, None] * Np + np.arange(Np)[None, :]).ravel()
    g = dg.Gather(ev, dofs)
    out = np.zeros((6, g.n_local))
    ev.set_state(dat["x0_f64"])
    t = meta["t0"]
    for _ in range(meta["steps"]):
        t = ev.Step(t, meta["dt"])
        g.launch(out)
        g.wait()
        assert np.array_equal(out, ev.get_state().reshape(6, N)[:, dofs])
    assert np.abs(out).max() > 0
    ev.close()
