"""`python -m dgtd_b200.launcher -i case.json [-d b200]` — the reference launcher's command line
(src/launcher/launcher.cpp:24-90: `opensemba_dgtd -i case.json -d cpu|omp|cuda`) with the device string `b200`: reads the
reference's JSON case format for the keys the evolution hot path needs, builds the operator through the C ABI, runs
`Solver::run` on the GPU (`dgtd_run_until`: final short step, stability test) and writes the reference's statistics file.

JSON keys honoured (src/driver/driver.cpp):
  solver_options  order, upwind_alpha, time_step, final_time, evolution_operator ("global" | "hesthaven": TF/SF gate on | off)   :644-705
  model           filename (Gmsh 2.2 / MFEM v1.0, relative to the JSON), materials [{tags, type vacuum | relative_permittivity,
                  relative_permeability, bulk_conductivity}], boundaries [{tags, type PEC | PMC | SMA}]                           :1242-1479
  sources         {type initial, field_type, center, polarization, dimension, magnitude {type gaussian, spread | resonant, modes}}
                  {type planewave, polarization, propagation, tags, magnitude {spread [, mean] [, frequency]}}                    :531-641
Without `time_step` (or with 0) the step is the reference's estimate (estimate_time_step below, Solver.cpp:175-351; `cfl` honoured).
Mesh refinement, probes / exporters, SGBC and the implicit integrators stay with the reference's host code (out of the hot path).
Multi-GPU: launch one process per GPU with torch.distributed.run; the mesh is partitioned with METIS (driver.cpp:1269).
Output: <out>/SimulationStats/statistics_rank<r>.dat (the keys of Solver::writeSimulationStatistics, Solver.cpp:404-445) and
<out>/final_state_rank<r>.npy (this rank's owned dofs, [6][n_local], with element ids in final_elements_rank<r>.npy).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

C_SI = 299792458.0     # physicalConstants::speedOfLight_SI (the reference's times are in metres of light travel)


def _vec3(v):
    out = np.zeros(3)
    out[:len(v)] = v
    return out


def gauss_lobatto_01(order):
    """The order+1 Gauss-Lobatto points on [0, 1] (the nodes of mfem's L2 GaussLobatto segment, in dof order)."""
    if order == 0:
        return np.array([0.5])
    inner = np.polynomial.legendre.Legendre.basis(order).deriv().roots() if order > 1 else np.zeros(0)
    return 0.5 * (np.concatenate([[-1.0], np.sort(inner.real), [1.0]]) + 1.0)


def estimate_time_step(verts, elems, dim, order, cfl=1.0, operator="global"):
    """estimateTimeStep of the reference (src/solver/Solver.cpp:175-351) for segment / triangle / tetrahedron meshes, in the
    reference's normalised units (c = 1): used when solver_options.time_step is absent or 0 (Solver.cpp:110-115).
      1-D: cfl * (smallest distance between two nodes of an element) / order^1.5                                :311-320
      2-D, 3-D: cfl * 0.75 * min_e(volume / (sum of face measures / 2)) * rmin * 2/3, rmin = distance between the
           first two Gauss-Lobatto nodes of a segment of length 2                                       :206-262, 290-300
      3-D `global`: that / 0.8;  3-D `hesthaven`: cfl / (max fscale * order^2), fscale = 2 |J_face| / |J_elem|  :326-348"""
    v = np.asarray(verts, float)[:, :3]
    x = v[np.asarray(elems)]                                             # [NE][dim+1][3]
    gl = gauss_lobatto_01(order)
    if dim == 1:
        h = np.abs(x[:, 1, 0] - x[:, 0, 0]).min()
        dmin = h if order == 0 else h * np.diff(gl).min()
        return cfl * (dmin if order == 0 else dmin / order ** 1.5)
    rmin = 2.0 * (gl[1] - gl[0]) if order > 0 else 2.0
    if dim == 2:
        a, b, c = x[:, 0, :2], x[:, 1, :2], x[:, 2, :2]
        vol = 0.5 * np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
        faces = np.linalg.norm(b - a, axis=1) + np.linalg.norm(c - b, axis=1) + np.linalg.norm(a - c, axis=1)
        return cfl * 0.75 * (vol / (faces / 2.0)).min() * rmin * 2.0 / 3.0
    e1, e2, e3 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]
    vol = np.abs(np.einsum("ij,ij->i", np.cross(e1, e2), e3)) / 6.0
    area = np.zeros((len(x), 4))
    for f in range(4):
        p = [q for q in range(4) if q != f]
        area[:, f] = 0.5 * np.linalg.norm(np.cross(x[:, p[1]] - x[:, p[0]], x[:, p[2]] - x[:, p[0]]), axis=1)
    if operator == "hesthaven":
        fscale = 2.0 * (2.0 * area) / (6.0 * vol)[:, None]
        return cfl / (fscale.max() * order * order)
    return cfl * 0.75 * (vol / (area.sum(axis=1) / 2.0)).min() * rmin * 2.0 / 3.0 / 0.8


def build_case(case, base_dir, dg):
    """-> (mesh, Evolution kwargs, dt, final_time, initial-state function of node coordinates [N][3])."""
    so = case.get("solver_options", {})
    model = case["model"]
    mesh = dg.Mesh.load(os.path.join(base_dir, model["filename"]))
    if model.get("refinement", 0):
        raise SystemExit("launcher: model.refinement is applied by the reference's driver (Mesh::UniformRefinement); refine the mesh file instead")
    bdr = {}
    for b in model.get("boundaries", []):
        code = {"PEC": dg.BC_PEC, "PMC": dg.BC_PMC, "SMA": dg.BC_SMA}.get(b["type"])
        if code is None:
            raise SystemExit(f"launcher: boundary type {b['type']} is not part of the evolution hot path")
        for t in b["tags"]:
            bdr[int(t)] = code
    materials = {}
    for m in model.get("materials", []):
        if m.get("type", "vacuum") == "vacuum":
            continue
        for t in m["tags"]:
            materials[int(t)] = (float(m.get("relative_permittivity", 1.0)), float(m.get("relative_permeability", 1.0)), float(m.get("bulk_conductivity", 0.0)))
    planewave, tfsf, inits = None, (), []
    v, e, ea, b, ba = mesh.arrays()
    for s in case.get("sources", []):
        if s["type"] == "planewave":
            mag = s["magnitude"]
            pol, dirv = _vec3(s["polarization"]), _vec3(s["propagation"])
            dhat = dirv / np.linalg.norm(dirv)
            tfsf = tuple(int(t) for t in s["tags"])
            if "mean" in mag:
                mean1d = float(_vec3(mag["mean"]) @ dhat)
            else:       # auto delay (driver.cpp:406-408, 576-589): the pulse sits 5 sigma sqrt(2) upstream of the TF/SF surface at t = 0
                sel = np.isin(ba, tfsf)
                phase = (v[b[sel]].mean(axis=1) @ dhat) if sel.any() else np.zeros(1)
                mean1d = float(phase.min() - 5.0 * mag["spread"] * np.sqrt(2.0))
            freq = float(mag.get("frequency", 0.0)) / (C_SI if "frequency" in mag else 1.0)
            planewave = dg.PlaneWave(float(mag["spread"]), mean1d, tuple(pol), tuple(dirv), freq, 0)
        elif s["type"] == "initial":
            inits.append(s)
        else:
            raise SystemExit(f"launcher: source type {s['type']} is not part of the evolution hot path")

    def initial_state(xyz):
        x0 = np.zeros((6, len(xyz)))
        for s in inits:
            f = 0 if s.get("field_type", "electric").lower().startswith("e") else 1
            pol = _vec3(s["polarization"])
            mag = s["magnitude"]
            if mag["type"] == "gaussian":       # InitialField::eval (Sources.cpp:38-58) x Gaussian of `dimension` (Function.h:71-94)
                dim = int(s.get("dimension", 1))
                ctr = _vec3(s.get("center", [0.0]))
                r2 = ((xyz[:, :dim] - ctr[:dim]) ** 2).sum(axis=1)
                g = np.exp(-r2 / (2.0 * float(mag["spread"]) ** 2))
            elif mag["type"] == "resonant":     # SinusoidalMode: prod sin(m_k pi x_k)
                g = np.ones(len(xyz))
                for k, m in enumerate(mag["modes"]):
                    g = g * np.sin(float(m) * np.pi * xyz[:, k])
            else:
                raise SystemExit(f"launcher: initial magnitude {mag['type']} is not supported")
            for d in range(3):
                x0[3 * f + d] += pol[d] * g
        return x0.ravel()

    kw = dict(order=int(so.get("order", 2)), alpha=float(so.get("upwind_alpha", 1.0)), bdr=bdr, tfsf=tfsf, materials=materials,
              planewave=planewave, tfsf_gate=so.get("evolution_operator", "global") != "hesthaven")
    dt = float(so.get("time_step", 0.0))
    if dt == 0.0:                      # automatic time step (driver.cpp:714-726, Solver.cpp:110-115)
        dt = float(estimate_time_step(v, e, e.shape[1] - 1, kw["order"], float(so.get("cfl", 1.0)), so.get("evolution_operator", "global")))
    return mesh, kw, dt, float(so.get("final_time", 2.0)), initial_state


def write_statistics(path, run_s, final_time, dt, ne_local, avg_h, n_local, device_bytes):
    """The keys of Solver::writeSimulationStatistics (Solver.cpp:404-445); the operator-size lines of the assembled `global`
    matrix have no counterpart in a matrix-free operator and are replaced by its device memory."""
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "a") as f:
        f.write(f"Simulation Run Time: {run_s:.5e} (s)\n")
        f.write(f"Final Time: {final_time / C_SI * 1e9:g} (ns)\n")
        f.write(f"Time Step: {dt / C_SI * 1e9:g} (ns)\n")
        f.write(f"Number of Mesh Elements: {ne_local}\n")
        f.write(f"Average Element Size in Mesh: {avg_h:g}\n")
        f.write(f"Number of Local Degrees of Freedom: {n_local}\n")
        f.write(f"Temporal Evolution Memory Consumption (B): {device_bytes}\n")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="dgtd_b200.launcher", description=__doc__.split("\n\n")[0])
    ap.add_argument("-i", dest="input", required=True, help="case .json (the reference's format)")
    ap.add_argument("-d", "--device", default="b200", help='"b200" (there is no CPU path)')
    ap.add_argument("-o", "--out", default=None, help="output directory (default: Exports/b200-<ranks>/<case>/ next to the CWD, like the reference)")
    ap.add_argument("--check-every", type=int, default=1, help="stability test every that many steps (the reference: every step)")
    args = ap.parse_args(argv)
    if args.device != "b200":
        raise SystemExit('Available device string is "b200" (the reference launcher serves "cpu", "omp" and "cuda")')
    if not args.input.endswith(".json"):
        print("Input File is not a .json file.", file=sys.stderr)
        return 3
    import torch
    import dgtd_b200 as dg
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    case = json.load(open(args.input))
    name = os.path.splitext(os.path.basename(args.input))[0]
    out = args.out or os.path.join("Exports", f"b200-{world}", name)
    t_init = time.perf_counter()
    mesh, kw, dt, t_final, initial_state = build_case(case, os.path.dirname(os.path.abspath(args.input)), dg)
    torch.cuda.set_device(local)
    part = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        part = mesh.partition(world, "metis")
    ev = dg.Evolution(mesh, device=local, rank=rank, nranks=world, partitioning=part, **kw)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dg.Evolution.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ev.comm_init(bytes(idt.cpu().numpy().tobytes()))
    gid = ev.local_elements()
    xyz = ev.node_coords().reshape(-1, ev.Np, 3)[gid].reshape(-1, 3)
    ev.set_state_local(initial_state(xyz))
    t_init = time.perf_counter() - t_init
    t0 = time.perf_counter()
    t, nsteps, unstable = ev.run_until(0.0, dt, t_final, check_every=args.check_every)
    ev.synchronize()
    run_s = time.perf_counter() - t0
    x = ev.get_state_local()
    os.makedirs(out, exist_ok=True)
    np.save(os.path.join(out, f"final_state_rank{rank}.npy"), x.reshape(6, -1))
    np.save(os.path.join(out, f"final_elements_rank{rank}.npy"), gid)
    v, e, _, _, _ = mesh.arrays()
    ext = v[e[gid]].max(axis=1) - v[e[gid]].min(axis=1)
    avg_h = float(np.linalg.norm(ext, axis=1).mean())
    write_statistics(os.path.join(out, "SimulationStats", f"statistics_rank{rank}.dat"), run_s, t_final, dt, len(gid), avg_h, ev.n_local,
                     4 * 6 * 8 * ev.n_local)
    if rank == 0:
        if unstable:
            print("WARNING: the state norm left the stable range (Solver.cpp:500-516)")
        print(f"[{name}] {nsteps} RK4 steps to t = {t:g} in {run_s:.3f} s ({6 * ev.N * 4 * nsteps / max(run_s, 1e-12) / 1e9:.2f} G DOF-updates/s), "
              f"set-up {t_init:.2f} s, kernel: {ev.kernel_info()[:60]}")
        print("Solver has finished performing its operations.")
    ev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
