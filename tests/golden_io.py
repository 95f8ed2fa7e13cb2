"""Oracle-free readers of the committed fixtures (tests/golden/*.npz, *.cfg.npz): mesh arrays, the keyword arguments of the
product's `Evolution`, the stored reference vectors.  Nothing here imports `oracle/`, so bench.py's multi-GPU parity gate
and the product-side tests can use the reference vectors without executing the checker.

Fixtures are written by tests/golden/make_golden.py and make_config_fixtures.py from `oracle/_ref/dgtd_ref gen` (the
reference's own MFEM fork + DG integrators)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BC_CODES = {"none": 0, "pec": 1, "pmc": 2, "sma": 3}
STRIDE_KEY = "stride"


def smooth_state(xyz):
    """`--init smooth` of oracle/ref/dgtd_ref.cpp: u_c = sin(1.3 x + 0.7 c + 0.2) cos(0.9 y - 0.4 c) + 0.5 sin(1.1 z + c)."""
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    return np.concatenate([np.sin(1.3 * x + 0.7 * c + 0.2) * np.cos(0.9 * y - 0.4 * c) + 0.5 * np.sin(1.1 * z + c) for c in range(6)])


def initial_state(meta, xyz):
    """The fixture's initial condition, rebuilt from node coordinates [N][3]."""
    if meta["init"] == "smooth":
        return smooth_state(xyz)
    kind, comp, modes = meta["init"].split(":")
    assert kind == "resonant"
    x0 = np.zeros((6, len(xyz)))
    v = np.ones(len(xyz))
    for k, m in enumerate(modes.split(",")):
        v = v * np.sin(float(m) * np.pi * xyz[:, k])
    x0[int(comp)] = v
    return x0.ravel()


def read_fixture(name):
    """-> (arrays, meta): mesh arrays + whatever vectors the fixture stores (full `x0/k0/x_final` or `*_sample` every
    meta['stride']-th entry)."""
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        path = os.path.join(GOLDEN, name + ".cfg.npz")
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return {k: z[k] for k in z.files if k != "meta"}, meta


def product_problem(arr, meta, order=None):
    """Mesh + Evolution keyword arguments of a fixture for the product's C ABI."""
    import dgtd_b200 as dg

    dim = meta["dim"]
    mesh = dg.Mesh.from_arrays(dim, arr["verts_f64"].reshape(-1, 3), arr["elems_i32"].reshape(-1, dim + 1), arr["elem_attr_i32"],
                               arr["bdr_i32"].reshape(-1, dim), arr["bdr_attr_i32"])
    pw = None
    if meta.get("pw", {}).get("on"):
        w = meta["pw"]
        pw = dg.PlaneWave(w["spread"], w["mean1d"], tuple(w["pol"]), tuple(w["dir"]), w.get("freq", 0.0), w.get("fieldtype", 0))
    kw = dict(order=order or meta["order"], alpha=meta["alpha"], bdr={int(k): BC_CODES[v] for k, v in meta.get("bdr", {}).items()},
              tfsf=tuple(meta.get("tfsf", ())), materials={int(k): tuple(v) for k, v in meta.get("mat", {}).items()}, planewave=pw)
    return mesh, kw


def write_mfem_mesh(path, arr, meta):
    """The fixture's mesh as an "MFEM mesh v1.0" file (what the reference-based CPU arm loads): element / boundary order and
    vertex numbering are kept, so dof numbering equals the product's."""
    dim = meta["dim"]
    v = arr["verts_f64"].reshape(-1, 3)
    e = arr["elems_i32"].reshape(-1, dim + 1)
    b = arr["bdr_i32"].reshape(-1, dim)
    geom_e, geom_b = {1: 1, 2: 2, 3: 4}[dim], {1: 0, 2: 1, 3: 2}[dim]
    with open(path, "w") as f:
        f.write(f"MFEM mesh v1.0\n\ndimension\n{dim}\n\nelements\n{len(e)}\n")
        for a, row in zip(arr["elem_attr_i32"], e):
            f.write(f"{a} {geom_e} " + " ".join(str(int(x)) for x in row) + "\n")
        f.write(f"\nboundary\n{len(b)}\n")
        for a, row in zip(arr["bdr_attr_i32"], b):
            f.write(f"{a} {geom_b} " + " ".join(str(int(x)) for x in row) + "\n")
        f.write(f"\nvertices\n{len(v)}\n{dim}\n")
        for row in v:
            f.write(" ".join(repr(float(x)) for x in row[:dim]) + "\n")
