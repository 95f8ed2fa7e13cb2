import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if not p.endswith(".cfg.npz"))


def config_cases():
    """Sampled reference vectors of BASELINE.json's configs on the reference's own meshes (make_config_fixtures.py)."""
    return sorted(os.path.basename(p)[:-len(".cfg.npz")] for p in glob.glob(os.path.join(GOLDEN, "*.cfg.npz")))


from golden_io import initial_state, smooth_state  # noqa: E402,F401  (oracle-free fixture helpers, shared with bench.py)


def load_config_case(name):
    """-> (oracle Problem, meta, dict of sampled arrays) for a *.cfg.npz fixture; the initial state comes from initial_state()."""
    import json
    from oracle.dgtd_oracle import BC_NAMES, PlaneWave, Problem

    z = np.load(os.path.join(GOLDEN, name + ".cfg.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    pb = Problem(verts=z["verts_f64"].reshape(-1, 3), elems=z["elems_i32"].reshape(-1, 4).astype(np.int64), elem_attr=z["elem_attr_i32"],
                 bdr=z["bdr_i32"].reshape(-1, 3).astype(np.int64), bdr_attr=z["bdr_attr_i32"], order=meta["order"], alpha=meta["alpha"])
    pb.bdr_cond = {int(k): BC_NAMES[v] for k, v in meta.get("bdr", {}).items()}
    pb.tfsf_tags = tuple(meta.get("tfsf", ()))
    if meta.get("pw", {}).get("on"):
        w = meta["pw"]
        pb.planewave = PlaneWave(w["spread"], w["mean1d"], w["pol"], w["dir"], w["freq"])
    return pb, meta, {k: z[k] for k in ("x0_sample_f64", "k0_sample_f64", "x_final_sample_f64")}


def load_golden(name):
    """-> (oracle Problem, dict of arrays + meta) for a committed fixture."""
    from oracle.dgtd_oracle import BC_NAMES, load_case

    pb, dat = load_case(os.path.join(GOLDEN, name + ".npz"))
    meta = dat["meta"]
    pb.bdr_cond = {int(k): BC_NAMES[v] for k, v in meta.get("bdr", {}).items()}
    pb.tfsf_tags = tuple(meta.get("tfsf", ()))
    pb.materials = {int(k): tuple(v) for k, v in meta.get("mat", {}).items()}
    return pb, dat


def product_mesh_and_kwargs(pb):
    """The same problem expressed for the product's C ABI (dgtd_b200.Mesh / Evolution kwargs)."""
    import dgtd_b200 as dg

    dim = pb.elems.shape[1] - 1
    mesh = dg.Mesh.from_arrays(dim, pb.verts, pb.elems, pb.elem_attr, pb.bdr, pb.bdr_attr)
    pw = None
    if pb.planewave is not None:
        w = pb.planewave
        pw = dg.PlaneWave(w.spread, w.mean1d, tuple(w.pol), tuple(w.dir), w.freq, w.fieldtype)
    kw = dict(order=pb.order, alpha=pb.alpha, bdr=dict(pb.bdr_cond), tfsf=tuple(pb.tfsf_tags),
              materials=dict(pb.materials), planewave=pw)
    return mesh, kw


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
