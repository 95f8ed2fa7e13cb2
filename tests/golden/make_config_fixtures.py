#!/usr/bin/env python3
"""Reference vectors for BASELINE.json configs 3 and 4 on the reference's OWN meshes, produced by oracle/_ref/dgtd_ref
(the reference's MFEM fork + DG integrators, assembled `global` operator + mfem::RK4Solver).

The full vectors are 15-30 MB each, so a fixture keeps: the mesh arrays as MFEM sees them after loading (and refining),
the problem description, and every STRIDE-th entry of k0 = Mult(t0, x0) and of the state after `steps` RK4 steps, plus
the vectors' norms.  The initial state is reproducible from the node coordinates ("smooth": conftest.smooth_state,
"resonant": prod sin(m_k pi x_k) on E_z, the config's own initial condition), so it is not stored.
Needs the build container (/root/reference + oracle/_ref).  ~3 min and ~12 GB per order-3 case.
"""
import json, os, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref")
DATA = "/root/reference/testData/maxwellInputs"
STRIDE = 61

CASES = {
    # 3D_Resonant_Box_TM55_H2_P3.json: refinement 2, order 3, upwind, all PEC, E_z = sin(5 pi x) sin(5 pi y), dt 1e-4
    "config3_resonant_box_p3": dict(
        args=f"--mesh {DATA}/3D_Resonant_Box_TM55/3D_Resonant_Box_TM55_H2_P3/3D_Resonant_Box_TM55_H2_P3.msh --refine 2 --order 3 --alpha 1.0 "
             "--bdr-all pec --init resonant:2:5,5 --t0 0.0 --dt 1e-4 --steps 2",
        extra={"bdr": {str(a): "pec" for a in range(1, 7)}, "init": "resonant:2:5,5"}),
    # 3D_RCS_PEC_1m.json: PEC sphere (tag 1), SMA outer sphere (tag 2), TF/SF box (tags 3-8), x-polarised plane wave along z,
    # spread 0.3, automatic delay (driver.cpp:576-589); t0 puts the pulse on the TF/SF surface
    "config4_rcs_pec_p3": dict(
        args=f"--mesh {DATA}/3D_RCS_PEC_1m/3D_RCS_PEC_1m.msh --order 3 --alpha 1.0 --bdr 1:pec,2:sma --tfsf 3,4,5,6,7,8 "
             "--pw 0.3:auto:0:1,0,0:0,0,1 --init smooth --t0 2.4 --dt 0.005 --steps 2",
        extra={"bdr": {"1": "pec", "2": "sma"}, "tfsf": [3, 4, 5, 6, 7, 8], "init": "smooth"}),
    "config4_rcs_pec_p4": dict(
        args=f"--mesh {DATA}/3D_RCS_PEC_1m/3D_RCS_PEC_1m.msh --order 4 --alpha 1.0 --bdr 1:pec,2:sma --tfsf 3,4,5,6,7,8 "
             "--pw 0.3:auto:0:1,0,0:0,0,1 --init smooth --t0 2.4 --dt 0.0025 --steps 2",
        extra={"bdr": {"1": "pec", "2": "sma"}, "tfsf": [3, 4, 5, 6, 7, 8], "init": "smooth"}),
}


def run(name, args, extra):
    with tempfile.TemporaryDirectory(dir="/tmp") as d:
        out = subprocess.run([REF, "gen", "--out", d] + args.split(), check=True, capture_output=True, text=True).stdout
        meta = json.loads(out.strip().splitlines()[-1])
        meta["cmd"] = "dgtd_ref gen " + args.replace(DATA, "<reference>/testData/maxwellInputs")
        meta.update(extra)
        meta["stride"] = STRIDE
        arr = {}
        for f in ("verts.f64", "elems.i32", "elem_attr.i32", "bdr.i32", "bdr_attr.i32"):
            arr[f.replace(".", "_")] = np.fromfile(os.path.join(d, f), np.float64 if f.endswith("f64") else np.int32)
        for v in ("k0", "x_final"):
            full = np.fromfile(os.path.join(d, v + ".f64"), np.float64)
            arr[v + "_sample_f64"] = full[::STRIDE].copy()
            meta[v + "_norm"] = float(np.linalg.norm(full))
        x0 = np.fromfile(os.path.join(d, "x0.f64"), np.float64)
        meta["x0_norm"] = float(np.linalg.norm(x0))
        arr["x0_sample_f64"] = x0[::STRIDE].copy()
        arr["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(HERE, name + ".cfg.npz"), **arr)
        print(name, {k: meta[k] for k in ("order", "ne", "n", "nnz", "tfsf_applied", "tfsf_skipped", "assemble_s")}, flush=True)


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref first: make -C oracle/ref")
    for name in sys.argv[1:] or ["config3_resonant_box_p3", "config4_rcs_pec_p3"]:
        run(name, **CASES[name])
