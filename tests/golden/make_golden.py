#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/*.npz with the reference-based oracle.

Runs oracle/_ref/dgtd_ref (the reference's own MFEM fork + DG integrators compiled from /root/reference,
see oracle/ref/Makefile) on small cases and stores mesh arrays, the initial state x0, k0 = Mult(t0, x0) and
the state after `steps` RK4 steps.  Needs the build container (oracle/_ref + /root/reference); the .npz files
are committed so that the tests run anywhere.
"""
import json, os, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref")


def run(name, args, extra=None):
    with tempfile.TemporaryDirectory() as d:
        cmd = [REF, "gen", "--out", d] + args
        out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout
        meta = json.loads(out.strip().splitlines()[-1])
        meta["cmd"] = " ".join(["dgtd_ref", "gen"] + args)
        if extra:
            meta.update(extra)
        arr = {}
        for f in os.listdir(d):
            if f.endswith(".f64"):
                arr[f.replace(".f64", "_f64")] = np.fromfile(os.path.join(d, f), np.float64)
            elif f.endswith(".i32"):
                arr[f.replace(".i32", "_i32")] = np.fromfile(os.path.join(d, f), np.int32)
        arr["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arr)
        print(name, {k: meta[k] for k in ("dim", "order", "ne", "n", "nnz", "tfsf_applied", "tfsf_skipped")})


def tfsf_box_mesh(path, n=4, lo=0.25, hi=0.75):
    """n^3 Kuhn box with the faces on the surface of [lo,hi]^3 listed as interior boundary elements (attribute 7)."""
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i
    verts = [(i / n, j / n, k / n) for k in range(n + 1) for j in range(n + 1) for i in range(n + 1)]
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                for pm in perms:
                    c = [0, 0, 0]; v = [vid(i, j, k)]
                    for s in pm:
                        c[s] = 1; v.append(vid(i + c[0], j + c[1], k + c[2]))
                    tets.append(v)
    faces = {}
    for t in tets:
        for f in range(4):
            key = tuple(sorted(t[:f] + t[f + 1:]))
            faces.setdefault(key, []).append(t)
    V = np.array(verts)
    bdr = []
    for key, ts in faces.items():
        P = V[list(key)]
        if len(ts) == 1:
            for ax in range(3):
                if np.all(P[:, ax] == 0.0): bdr.append((1 + 2 * ax, key))
                elif np.all(P[:, ax] == 1.0): bdr.append((2 + 2 * ax, key))
        else:
            inside = np.all((P >= lo - 1e-12) & (P <= hi + 1e-12))
            for ax in range(3):
                if inside and (np.all(np.abs(P[:, ax] - lo) < 1e-12) or np.all(np.abs(P[:, ax] - hi) < 1e-12)):
                    bdr.append((7, key)); break
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\ndimension\n3\n\nelements\n%d\n" % len(tets))
        for t in tets: f.write("1 4 %d %d %d %d\n" % tuple(t))
        f.write("\nboundary\n%d\n" % len(bdr))
        for a, k in bdr: f.write("%d 2 %d %d %d\n" % ((a,) + k))
        f.write("\nvertices\n%d\n3\n" % len(verts))
        for v in verts: f.write("%.17g %.17g %.17g\n" % v)


def sheet_mesh(path, n=4):
    """n^3 Kuhn box with three open sheets of interior faces listed as boundary elements: attribute 7 on the plane x = 1/2
    (y <= 1/2), 8 on y = 1/2 (z >= 1/2), 9 on z = 1/4 (x >= 1/2) — interior PEC / PMC / SMA boundaries."""
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i
    verts = [(i / n, j / n, k / n) for k in range(n + 1) for j in range(n + 1) for i in range(n + 1)]
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets, faces = [], {}
    for k in range(n):
        for j in range(n):
            for i in range(n):
                for pm in perms:
                    c = [0, 0, 0]; v = [vid(i, j, k)]
                    for s in pm:
                        c[s] = 1; v.append(vid(i + c[0], j + c[1], k + c[2]))
                    tets.append(v)
    for t in tets:
        for f in range(4):
            faces.setdefault(tuple(sorted(t[:f] + t[f + 1:])), []).append(t)
    V = np.array(verts)
    bdr = []
    for key, ts in faces.items():
        P = V[list(key)]
        if len(ts) == 1:
            for ax in range(3):
                if np.all(P[:, ax] == 0.0): bdr.append((1 + 2 * ax, key))
                elif np.all(P[:, ax] == 1.0): bdr.append((2 + 2 * ax, key))
        elif np.all(P[:, 0] == 0.5) and np.all(P[:, 1] <= 0.5): bdr.append((7, key))
        elif np.all(P[:, 1] == 0.5) and np.all(P[:, 2] >= 0.5): bdr.append((8, key))
        elif np.all(P[:, 2] == 0.25) and np.all(P[:, 0] >= 0.5): bdr.append((9, key))
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\ndimension\n3\n\nelements\n%d\n" % len(tets))
        for t in tets: f.write("1 4 %d %d %d %d\n" % tuple(t))
        f.write("\nboundary\n%d\n" % len(bdr))
        for a, k in bdr: f.write("%d 2 %d %d %d\n" % ((a,) + k))
        f.write("\nvertices\n%d\n3\n" % len(verts))
        for v in verts: f.write("%.17g %.17g %.17g\n" % v)


def sheet_mesh_2d(path, nx=4, ny=3):
    """nx x ny squares split into triangles with the interior edges on x = 1/2 (lower half) tagged 5 and on y = 1/3 tagged 6."""
    vid = lambda i, j: j * (nx + 1) + i
    verts = [(i / nx, j / ny) for j in range(ny + 1) for i in range(nx + 1)]
    tris = []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
            tris += [(a, b, c), (a, c, d)]
    edges = {}
    for t in tris:
        for f in range(3):
            edges.setdefault(tuple(sorted(t[:f] + t[f + 1:])), []).append(t)
    V = np.array(verts)
    bdr = []
    for key, ts in edges.items():
        P = V[list(key)]
        if len(ts) == 1:
            if np.all(P[:, 1] == 0.0): bdr.append((1, key))
            elif np.all(P[:, 0] == 1.0): bdr.append((2, key))
            elif np.all(P[:, 1] == 1.0): bdr.append((3, key))
            else: bdr.append((4, key))
        elif np.all(P[:, 0] == 0.5) and np.all(P[:, 1] <= 2.0 / 3 + 1e-12): bdr.append((5, key))
        elif np.all(np.abs(P[:, 1] - 1.0 / 3) < 1e-12) and np.all(P[:, 0] >= 0.5): bdr.append((6, key))
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\ndimension\n2\n\nelements\n%d\n" % len(tris))
        for t in tris: f.write("1 2 %d %d %d\n" % t)
        f.write("\nboundary\n%d\n" % len(bdr))
        for a, k in bdr: f.write("%d 1 %d %d\n" % ((a,) + k))
        f.write("\nvertices\n%d\n2\n" % len(verts))
        for v in verts: f.write("%.17g %.17g\n" % v)


def interior_cases():
    """Interior PEC / PMC / SMA boundaries (DGOperatorFactory.h:575-675): sheets and a closed box inside the mesh."""
    with tempfile.TemporaryDirectory() as d:
        mp = os.path.join(d, "sheets.mesh"); sheet_mesh(mp)
        run("ibc3d_p3_mixed_sheets", f"--mesh {mp} --order 3 --alpha 0.6 --bdr 1:pec,2:sma,3:pmc,4:pec,5:sma,6:pec,7:pec,8:pmc,9:sma --init random:11 --dt 1e-3 --steps 2".split(),
            {"bdr": {"1": "pec", "2": "sma", "3": "pmc", "4": "pec", "5": "sma", "6": "pec", "7": "pec", "8": "pmc", "9": "sma"}})
        mb = os.path.join(d, "box.mesh"); tfsf_box_mesh(mb)
        run("ibc3d_p2_pec_box", f"--mesh {mb} --order 2 --alpha 1.0 --bdr 1:sma,2:sma,3:sma,4:sma,5:sma,6:sma,7:pec --init random:12 --dt 2e-3 --steps 3".split(),
            {"bdr": {"1": "sma", "2": "sma", "3": "sma", "4": "sma", "5": "sma", "6": "sma", "7": "pec"}})
        m2 = os.path.join(d, "sheets2d.mesh"); sheet_mesh_2d(m2)
        run("ibc2d_p3_sheets", f"--mesh {m2} --order 3 --alpha 1.0 --bdr 1:pec,2:sma,3:pmc,4:pec,5:pec,6:sma --init random:13 --dt 1e-3 --steps 2".split(),
            {"bdr": {"1": "pec", "2": "sma", "3": "pmc", "4": "pec", "5": "pec", "6": "sma"}})


def two_material_mesh(path):
    """2x2x2 Kuhn box, elements with barycentre x > 0.5 get attribute 2."""
    n = 2
    vid = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i
    verts = [(i / n, j / n, k / n) for k in range(n + 1) for j in range(n + 1) for i in range(n + 1)]
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets, faces = [], {}
    for k in range(n):
        for j in range(n):
            for i in range(n):
                for pm in perms:
                    c = [0, 0, 0]; v = [vid(i, j, k)]
                    for s in pm:
                        c[s] = 1; v.append(vid(i + c[0], j + c[1], k + c[2]))
                    tets.append((2 if i >= 1 else 1, v))
    V = np.array(verts)
    for a, t in tets:
        for f in range(4):
            faces.setdefault(tuple(sorted(t[:f] + t[f + 1:])), []).append(t)
    bdr = []
    for key, ts in faces.items():
        if len(ts) == 1:
            P = V[list(key)]
            for ax in range(3):
                if np.all(P[:, ax] == 0.0): bdr.append((1 + 2 * ax, key))
                elif np.all(P[:, ax] == 1.0): bdr.append((2 + 2 * ax, key))
    with open(path, "w") as f:
        f.write("MFEM mesh v1.0\n\ndimension\n3\n\nelements\n%d\n" % len(tets))
        for a, t in tets: f.write("%d 4 %d %d %d %d\n" % ((a,) + tuple(t)))
        f.write("\nboundary\n%d\n" % len(bdr))
        for a, k in bdr: f.write("%d 2 %d %d %d\n" % ((a,) + k))
        f.write("\nvertices\n%d\n3\n" % len(verts))
        for v in verts: f.write("%.17g %.17g %.17g\n" % v)


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref first: make -C oracle/ref")
    if len(sys.argv) > 1 and sys.argv[1] == "tfsf12":
        # the reference's own 1-D and 2-D TF/SF cases (test/cases/CasesTest.cpp:230-454 `1D_TFSF`, ExtensiveCasesTest.cpp `2D_TFSF`):
        # meshes and source parameters of testData/maxwellInputs/{1D_TFSF,2D_TFSF}/*.json, started when the pulse sits on the
        # TF/SF points / line (auto delay of driver.cpp:576-589: the pulse centre reaches them at t = 5 sqrt(2) spread)
        R = "/root/reference/testData/maxwellInputs"
        run("tfsf1d_ref_p3", f"--mesh {R}/1D_TFSF/1D_TFSF.msh --order 3 --alpha 1.0 --bdr 1:pec,2:pec --tfsf 3,4 --pw 0.6:auto:0:0,1,0:1,0,0 --init random:21 --t0 4.0 --dt 0.01 --steps 40".split(),
            {"bdr": {"1": "pec", "2": "pec"}, "tfsf": [3, 4]})
        run("tfsf2d_ref_p3", f"--mesh {R}/2D_TFSF/2D_TFSF.msh --order 3 --alpha 1.0 --bdr 1:pec,3:pec,5:pec,7:pec,4:pmc,6:pmc --tfsf 2 --pw 0.4:auto:0:0,1,0:1,0,0 --init random:22 --t0 2.7 --dt 0.01 --steps 20".split(),
            {"bdr": {"1": "pec", "3": "pec", "5": "pec", "7": "pec", "4": "pmc", "6": "pmc"}, "tfsf": [2]})
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "interior":
        interior_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "config2":
        # config 2 of BASELINE.json on the reference's own gmsh triangle mesh: 2D_PEC.json (246 triangles, order 3, upwind,
        # global operator, PEC on tags 2,4 and PMC on 1,3, Gaussian E_z of spread 0.12 varying along x, dt 1e-3)
        run("config2_2d_pec_p3", "--mesh /root/reference/testData/maxwellInputs/2D_PEC/2D_PEC.msh --order 3 --alpha 1.0 --bdr 1:pmc,2:pec,3:pmc,4:pec "
            "--init gauss:E:2:0.12:1:0.5,0.5 --dt 1e-3 --steps 20".split(), {"bdr": {"1": "pmc", "2": "pec", "3": "pmc", "4": "pec"}})
        sys.exit(0)
    run("box3d_p3_pec_upwind", "--mesh cart3d:2 --order 3 --alpha 1.0 --bdr-all pec --init random:1 --dt 1e-3 --steps 2".split(),
        {"bdr": {str(a): "pec" for a in range(1, 7)}})
    run("box3d_p2_mixed_centered", "--mesh cart3d:2 --order 2 --alpha 0.0 --bdr 1:pec,2:pmc,3:pec,4:pmc,5:pec,6:pmc --init random:4 --dt 1e-3 --steps 2".split(),
        {"bdr": {"1": "pec", "2": "pmc", "3": "pec", "4": "pmc", "5": "pec", "6": "pmc"}})
    run("box3d_p4_sma_partial", "--mesh cart3d:1 --order 4 --alpha 0.7 --bdr 1:sma,2:pec,3:sma,4:pmc,6:sma --init random:5 --dt 5e-4 --steps 2".split(),
        {"bdr": {"1": "sma", "2": "pec", "3": "sma", "4": "pmc", "6": "sma"}})
    run("box3d_p1_gauss", "--mesh cart3d:3 --order 1 --alpha 1.0 --bdr-all pec --init gauss:E:2:0.2:3:0.5,0.5,0.5 --dt 2e-3 --steps 3".split(),
        {"bdr": {str(a): "pec" for a in range(1, 7)}})
    run("tri2d_p3_mixed", "--mesh cart2d:3:2 --order 3 --alpha 1.0 --bdr 1:pec,2:pmc,3:sma,4:pec --init random:2 --dt 1e-3 --steps 2".split(),
        {"bdr": {"1": "pec", "2": "pmc", "3": "sma", "4": "pec"}})
    run("seg1d_p3_pec_sma", "--mesh cart1d:5 --order 3 --alpha 0.5 --bdr 1:pec,2:sma --init random:3 --dt 1e-3 --steps 2".split(),
        {"bdr": {"1": "pec", "2": "sma"}})
    # config 1 of BASELINE.json in miniature: 1D_PEC (20 segments on [0,1], order 3, upwind, Gaussian E_y, dt 5e-3)
    run("seg1d_config1_pec", "--mesh cart1d:20 --order 3 --alpha 1.0 --bdr 1:pec,2:pec --init gauss:E:1:0.1:1:0.5 --dt 5e-3 --steps 40".split(),
        {"bdr": {"1": "pec", "2": "pec"}})
    with tempfile.TemporaryDirectory() as d:
        mp = os.path.join(d, "tfsf_box.mesh"); tfsf_box_mesh(mp)
        common = f"--mesh {mp} --order 2 --alpha 1.0 --bdr 1:sma,2:sma,3:sma,4:sma,5:sma,6:pec --tfsf 7".split()
        bdr = {"1": "sma", "2": "sma", "3": "sma", "4": "sma", "5": "sma", "6": "pec"}
        # pulse on the TF/SF surface: injection active in every Mult
        run("tfsf3d_p2_on", common + "--pw 0.15:0.0:0:1,0,0:0,0,1 --init random:6 --t0 0.4 --dt 2e-3 --steps 3".split(),
            {"bdr": bdr, "tfsf": [7]})
        # pulse far upstream: ||s|| < 1e-8, the `global` operator skips the injection
        run("tfsf3d_p2_skipped", common + "--pw 0.05:-3.0:0:0,1,0:1,0,0 --init random:7 --t0 0.0 --dt 2e-3 --steps 2".split(),
            {"bdr": bdr, "tfsf": [7]})
        # modulated Gaussian, oblique incidence, H-field given
        run("tfsf3d_p2_modulated", common + "--pw 0.2:0.1:2.5:0,1,-1:1,1,1 --init zero --t0 0.3 --dt 2e-3 --steps 3".split(),
            {"bdr": bdr, "tfsf": [7]})
        mp2 = os.path.join(d, "two_mat.mesh"); two_material_mesh(mp2)
        run("box3d_p2_materials", f"--mesh {mp2} --order 2 --alpha 1.0 --bdr-all pec --mat 1:1.0:1.0:0.0,2:2.5:1.3:0.8 --init random:8 --dt 1e-3 --steps 2".split(),
            {"bdr": {str(a): "pec" for a in range(1, 7)}, "mat": {"1": [1.0, 1.0, 0.0], "2": [2.5, 1.3, 0.8]}})
