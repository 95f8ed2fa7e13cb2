#!/usr/bin/env python3
"""bench.py — DG DOF-updates/s (RHS + RK stage, FP64) of the fused RK4 hot path, and its reference arm.

Default workload = BASELINE.json configs[4] ("c5", the configuration the 1/2/4/8-GPU metric is quoted on): order-3
tetrahedral PEC box, MakeCartesian3D-style Kuhn mesh, 32^3 cubes (196 608 tets = 23.6 M DOFs) PER GPU; the box doubles along
x, y, z in turn as the GPU count doubles (8 GPUs: 64^3 cubes = 188.7 M DOFs ~ "200 M"), partitioned with METIS k-way on the
element dual graph like the reference's MPI build (driver.cpp:1269).  `--scaling strong --cubes 65` is the fixed
MakeCartesian3D(65,...) box (197.7 M DOFs) over 1/2/4/8 GPUs; `--partition rcb --shape bar` the slab geometry of round 1.
Other workloads (side measurements): `--workload c3` / `c4` = BASELINE configs 3 and 4 on the reference's own meshes
(tests/golden/*.cfg.npz).  A "step" is one classical RK4 step = 4 fused stage launches; DOF-updates = 6N * 4 per step
(SURVEY.md 8d).  The state (4 vectors x 189 MB per GPU on c5) is far larger than L2, so no flush is needed between steps.

  python bench.py [--gpus N --steps K --warmup W]           one JSON line (rank 0)
  python bench.py --impl reference [...]                    the reference's CPU algorithm (assembled CSR `global`
                                                            operator + mfem::RK4Solver, OpenMP) on the host cores

`value`    : device-resident throughput over K steps, CUDA events on the launching stream, max over ranks.
`sustained`: the same over a >= 2 s window (thousands of steps) with the SM clock and board power sampled through NVML.
`e2e`      : the step through the host-facing call a drop-in ODESolver makes (B200RK4Solver::Step on a HOST vector):
             H2D of the state from pinned memory + fused step + D2H of the new state, every step.
`roofline` : algorithmic bytes (41.67 B per DOF-update at order 3) per stage launch / average launch time.
`parity`   : N > 1 only.  Before anything is timed every rank (i) runs the small reference fixtures partitioned over the N
             ranks against the committed reference vectors and aborts above 1e-12, and (ii) compares its owned DOFs after
             `--parity-steps` steps of THIS workload with a single-GPU run of the same global mesh on rank 0
             (mirror of the reference's Scaling2D test, test/mfem/mpi_FiniteElementSpaceTest.cpp:171-244).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "DG DOF-updates/s (RHS+RK stage, FP64)"
UNIT = "DOF-updates/s"
CUBES_PER_GPU = 32           # 32^3 cubes x 6 tets per GPU
PARITY_TOL = 1e-12


def np_of(p):
    return (p + 1) * (p + 2) * (p + 3) // 6


def b_alg(p):
    """Algorithmic bytes per DOF-update: 40 B of state traffic + 200 B of geometry per element."""
    return 40.0 + 200.0 / (6 * np_of(p))


def box_shape(cubes, n_gpus, scaling, shape):
    """Cubes per axis of the GLOBAL Kuhn box."""
    if scaling == "strong":
        return (cubes, cubes, cubes)
    if shape == "bar":
        return (cubes * n_gpus, cubes, cubes)
    f = [1, 1, 1]
    k, ax = n_gpus, 0
    while k > 1:                      # double x, y, z in turn: 2 -> (2,1,1), 4 -> (2,2,1), 8 -> (2,2,2)
        if k % 2:
            f[ax % 3] *= k
            break
        f[ax % 3] *= 2
        k //= 2
        ax += 1
    return (cubes * f[0], cubes * f[1], cubes * f[2])


def workload_config(args, n_gpus):
    """The `config` object: identical in the b200 and the reference arm (the reference arm runs it, or a bounded sample
    of it that its cpu_baseline.sample describes)."""
    p = args.order
    if args.workload == "c5":
        sx, sy, sz = box_shape(args.cubes, n_gpus, args.scaling, args.shape)
        tets = sx * sy * sz * 6
        name = (f"config 5: 3D tet PEC box (Kuhn Cartesian {sx}x{sy}x{sz} cubes x6 tets), order {p}, upwind alpha=1, "
                f"{tets} tets, {tets * np_of(p) * 6} DOFs, classical RK4")
    elif args.workload == "c3":
        name = ("config 3: 3D_Resonant_Box_TM55_H2_P3 (the reference's gmsh box refined twice, 22400 tets), "
                f"order {p}, all PEC, upwind alpha=1, {22400 * np_of(p) * 6} DOFs, TM55 resonant initial field, dt 1e-4, classical RK4")
    else:
        name = ("config 4: 3D_RCS_PEC_1m (the reference's gmsh mesh, 15886 tets: PEC sphere, SMA outer boundary, TF/SF box), "
                f"order {p}, upwind alpha=1, {15886 * np_of(p) * 6} DOFs, Gaussian plane wave on the TF/SF surface, classical RK4")
    return {"workload": name, "order": p, "scaling": args.scaling if args.workload == "c5" else "replicas",
            "partition": ("none" if n_gpus == 1 else args.partition), "l2": "state vectors are larger than L2 (c5); no flush between steps"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_info, dofs):
    """dram__bytes_read.sum + dram__bytes_write.sum per stage launch from the committed `ncu --set full` capture of this
    kernel on this workload (profiles/traffic.json: one record per kernel and size with the four launches of a step);
    None when there is no capture of this kernel at this size.  -> (mean over the step's launches, per-launch dict)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    for rec in json.load(open(p)):
        if kernel_info.startswith(rec["kernel_prefix"]) and rec["dofs_per_gpu"] == dofs:
            per = rec.get("dram_bytes_per_stage")
            if per:
                return sum(per.values()) / len(per), per
            return rec.get("dram_bytes_per_launch"), None
    return None, None


class ClockSampler(threading.Thread):
    """SM clock, board power and throttle reasons sampled DURING the timed regions: NVML queries every 2 ms
    (nvidia_ml_py), nvidia-smi every 100 ms as fallback."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.power, self.stamps, self.reasons, self.stop_flag = index, [], [], [], set(), False
        self.max_mhz, self.how = None, "nvml"
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a list of ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(v) for v in vis.split(",") if v.strip().isdigit()]
            phys = ids[index] if index < len(ids) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml, self.how = None, "nvidia-smi"

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            while not self.stop_flag:
                try:
                    self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                    self.stamps.append(time.perf_counter())
                    try:
                        self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
                    except Exception:
                        self.power.append(float("nan"))
                    try:
                        mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    for name, bit in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
                time.sleep(0.002)
            return
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1]); self.power.append(float(out[2]))
                self.stamps.append(time.perf_counter())
                for nme, v in zip(names, out[3:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.1)

    def window(self, w0, w1):
        """Median SM clock / max power of the samples taken inside [w0, w1] (perf_counter)."""
        ins = [(c, p) for c, p, t in zip(self.samples, self.power, self.stamps) if w0 <= t <= w1]
        cl = sorted(c for c, _ in ins)
        pw = [p for _, p in ins if p == p]
        return {"sm_mhz": cl[len(cl) // 2] if cl else None, "sm_min_mhz": cl[0] if cl else None,
                "power_w_max": max(pw) if pw else None, "power_w_mean": (sum(pw) / len(pw)) if pw else None, "samples": len(cl)}

    def result(self, w0, w1):
        allp = sorted(self.samples)
        r = self.window(w0, w1)
        if r["sm_mhz"] is None and allp:          # region shorter than one NVML query: the samples of the loaded phase
            r["sm_mhz"], r["sm_min_mhz"] = allp[len(allp) // 2], allp[0]
        r.update({"sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples_under_load": len(allp),
                  "sm_mhz_under_load": allp[len(allp) // 2] if allp else None, "how": self.how})
        return r


# ---- the reference's CPU algorithm (oracle/_ref/dgtd_ref, else the numpy port) -------------------------------------------
def cpu_reference(args, steps, warmup, repeats, cubes=None, same_mesh=False, threads=None):
    """Time the reference's CPU algorithm: MFEM SparseMatrix::Mult (OpenMP row loop, sparsemat.cpp:921-931) on the assembled
    `global` CSR + mfem::RK4Solver, compiled from the reference sources (oracle/_ref/dgtd_ref, `-d omp`).  c5: a bounded
    sample of the workload, the same box family at `cubes`^3 cubes; c3 / c4 (same_mesh): the workload's own mesh.
    BASELINE.md 5: assembly excluded, median of `repeats` windows of `steps` steps, SpMV-only share, cores stated."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref")
    threads = threads or os.cpu_count() or 1
    p = args.order
    if os.path.exists(exe):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
        tmp = None
        if same_mesh:
            from golden_io import read_fixture, write_mfem_mesh
            fx = {"c3": "config3_resonant_box_p3", "c4": "config4_rcs_pec_p3"}[args.workload]
            arr, meta = read_fixture(fx)
            tmp = tempfile.mkdtemp(prefix="dgtd_bench_")
            mpath = os.path.join(tmp, fx + ".mesh")
            write_mfem_mesh(mpath, arr, meta)
            cmd = [exe, "bench", "--device", "omp", "--mesh", mpath, "--order", str(p), "--alpha", "1.0"]
            if args.workload == "c3":
                cmd += ["--bdr-all", "pec", "--init", meta["init"], "--dt", "1e-4"]
            else:
                w = meta["pw"]
                cmd += ["--bdr", "1:pec,2:sma", "--tfsf", ",".join(str(t) for t in meta["tfsf"]),
                        "--pw", f"{w['spread']}:{w['mean1d']!r}:{w['freq']}:" + ",".join(str(v) for v in w["pol"]) + ":" + ",".join(str(v) for v in w["dir"]),
                        "--init", "smooth", "--t0", str(meta["t0"]), "--dt", repr(meta["dt"] * (9.0 / (p * p)))]
            sample = f"the workload itself (same mesh, {meta['ne']} tets, {meta['ne'] * np_of(p) * 6} DOFs)"
        else:
            cmd = [exe, "bench", "--device", "omp", "--mesh", f"cart3d:{cubes}", "--order", str(p), "--alpha", "1.0", "--bdr-all", "pec",
                   "--init", "random:1", "--dt", "1e-4"]
            sample = f"same box family at {cubes}^3 cubes ({cubes ** 3 * 6} tets, {cubes ** 3 * 6 * np_of(p) * 6} DOFs)"
        cmd += ["--steps", str(steps), "--warmup", str(warmup), "--repeats", str(repeats), "--spmv-share", "1"]
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True).stdout.strip().splitlines()[-1]
        if tmp:
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
        d = json.loads(out)
        sample += (f", {steps} RK4 steps after {warmup} warm-up, median of {repeats} windows; CSR assembly ({d['assemble_s']:.0f} s, "
                   f"{d['nnz']} non-zeros = {d['nnz'] * 12 / 1e9:.1f} GB) excluded; SpMV share {d['spmv_share']:.2f}; OpenMP threads only (no MPI in the image)")
        return {"value": d["dof_updates_per_s"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                "ms_per_step": 1e3 * d["run_s"] / max(1, steps), "nnz": d["nnz"], "assemble_s": d["assemble_s"], "dofs": 6 * d["n"],
                "runs_s": d.get("runs_s"), "spmv_share": d.get("spmv_share")}
    # the portable numpy restatement (matrix-free, single thread): only when the reference build did not travel
    import numpy as np
    from oracle.dgtd_oracle import PEC, HesthavenOracle, Problem
    import dgtd_b200 as dg
    cubes = min(cubes or 6, 6)
    m = dg.Mesh.cartesian3d(cubes)
    v, e, ea, b, ba = m.arrays()
    O = HesthavenOracle(Problem(v, e.astype(np.int64), ea, b.astype(np.int64), ba, p, 1.0, {a: PEC for a in range(1, 7)}))
    x = np.random.default_rng(1).standard_normal(6 * O.N)
    for _ in range(warmup):
        x = O.rk4_step(x, 0.0, 1e-4)
    runs = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        for _ in range(steps):
            x = O.rk4_step(x, 0.0, 1e-4)
        runs.append(time.perf_counter() - t0)
    dt = sorted(runs)[len(runs) // 2]
    return {"value": 6 * O.N * 4 * steps / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"same box family at {cubes}^3 cubes ({6 * O.N} DOFs), numpy port of the matrix-free algorithm, {steps} steps, median of {repeats}",
            "ms_per_step": 1e3 * dt / steps, "dofs": 6 * O.N, "runs_s": runs, "spmv_share": None}


# ---- workloads for the GPU arm -----------------------------------------------------------------------------------------
def build_workload(args, dg, n_gpus):
    """-> (mesh, Evolution kwargs, t0, dt, init(xyz) -> [6 n] state)."""
    import numpy as np
    from golden_io import initial_state, product_problem, read_fixture, smooth_state
    p = args.order
    if args.workload == "c5":
        sx, sy, sz = box_shape(args.cubes, n_gpus, args.scaling, args.shape)
        c = args.cubes
        mesh = dg.Mesh.cartesian3d(sx, sy, sz, sx=sx / c, sy=sy / c, sz=sz / c)
        kw = dict(order=p, alpha=1.0, bdr={a: dg.BC_PEC for a in range(1, 7)})
        dt = 0.05 * (1.0 / c) / (p * p)          # well inside the RK4 stability region; the rate does not depend on it
        return mesh, kw, 0.0, dt, lambda xyz: 1e-3 * smooth_state(xyz)
    fx = {"c3": "config3_resonant_box_p3", "c4": "config4_rcs_pec_p3"}[args.workload]
    arr, meta = read_fixture(fx)
    mesh, kw = product_problem(arr, meta, order=p)
    dt = meta["dt"] * (9.0 / (p * p))            # the fixtures' step is for order 3
    return mesh, kw, meta["t0"], dt, lambda xyz: initial_state(meta, xyz)


def fresh_unique_id(dg, dist, rank):
    import torch
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(dg.Evolution.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().numpy().tobytes())


def owned_mask(ev):
    import numpy as np
    mine = np.zeros(ev.N, bool)
    mine.reshape(-1, ev.Np)[ev.local_elements()] = True
    return np.tile(mine, 6)


def parity_gate_fixtures(dg, dist, rank, world, local):
    """(i) of the parity gate: the small reference fixtures, partitioned over all ranks (METIS parts; RCB for the 48-element
    box), Mult and the fixture's RK4 run against the committed reference vectors on every rank's owned DOFs."""
    import numpy as np
    import torch
    from golden_io import initial_state, product_problem, read_fixture
    worst, lines = 0.0, []
    for name, method in (("box3d_p3_pec_upwind", "rcb"), ("tfsf3d_p2_on", "metis"), ("config4_rcs_pec_p3", "metis")):
        arr, meta = read_fixture(name)
        mesh, kw = product_problem(arr, meta)
        if mesh.ne < 4 * world:
            continue
        uid = fresh_unique_id(dg, dist, rank)
        ev = dg.Evolution(mesh, device=local, rank=rank, nranks=world, partitioning=mesh.partition(world, method), **kw)
        ev.comm_init(uid)
        mask = owned_mask(ev)
        sampled = "x0_f64" not in arr
        st = meta.get("stride", 1) if sampled else 1
        x0 = initial_state(meta, ev.node_coords()) if sampled else arr["x0_f64"]
        kref, xref = (arr["k0_sample_f64"], arr["x_final_sample_f64"]) if sampled else (arr["k0_f64"], arr["x_final_f64"])
        ev.SetTime(meta["t0"])
        k = ev.Mult(x0)
        ev.set_state(x0)
        ev.run(meta["t0"], meta["dt"], meta["steps"])
        x = ev.get_state(np.zeros(6 * ev.N))
        m = mask[::st]
        num = torch.tensor([np.sum((k[::st][m] - kref[m]) ** 2), np.sum((x[::st][m] - xref[m]) ** 2)], dtype=torch.float64, device="cuda")
        dist.all_reduce(num)
        e_mult = math.sqrt(float(num[0])) / np.linalg.norm(kref)
        e_run = math.sqrt(float(num[1])) / np.linalg.norm(xref)
        lines.append({"fixture": name, "partition": method, "mult_rel_l2": e_mult, "run_rel_l2": e_run, "halo_mode": ev.halo_mode()})
        worst = max(worst, e_mult, e_run)
        ev.close()
    return worst, lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=["c5", "c3", "c4"])
    ap.add_argument("--cubes", type=int, default=None, help="c5: cubes per axis per GPU (weak, default 32) or of the whole box (strong, default 65)")
    ap.add_argument("--order", type=int, default=None, help="polynomial order (default 3; 4 for c4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--partition", default="metis", choices=["metis", "rcb"])
    ap.add_argument("--shape", default="cube", choices=["cube", "bar"], help="weak scaling: grow the box as a cube (default) or as a bar along x (slabs)")
    ap.add_argument("--cpu-cubes", type=int, default=None, help="box size of the bounded CPU sample (c5)")
    ap.add_argument("--cpu-same", action="store_true", help="c3/c4: time the CPU reference on the same mesh inside the b200 arm too")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained window in seconds (0: skip)")
    ap.add_argument("--parity-steps", type=int, default=5, help="N > 1: steps of the N-rank vs 1-rank comparison (0: skip)")
    ap.add_argument("--no-gate", action="store_true", help="N > 1: skip the small-fixture parity gate")
    ap.add_argument("--parity-report-only", action="store_true", help="N > 1: report the bench-size parity instead of aborting on it (debugging)")
    args = ap.parse_args()
    if args.order is None:
        args.order = 4 if args.workload == "c4" else 3
    if args.cubes is None:
        args.cubes = 65 if args.scaling == "strong" else CUBES_PER_GPU
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = world if world > 1 else 1
    warm = max(3, args.warmup) if args.impl == "b200" else args.warmup
    scaling = args.scaling if args.workload == "c5" else "weak"      # c3 / c4 are single-GPU side measurements

    if args.impl == "reference":
        if rank != 0:
            return 0
        same = args.workload != "c5"
        cb = cpu_reference(args, max(1, args.steps), max(0, args.warmup), 3, cubes=args.cpu_cubes or 16, same_mesh=same)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
                "warmup": max(0, args.warmup), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(args, args.gpus),
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "runs_s": cb.get("runs_s"), "spmv_share": cb.get("spmv_share"), "sample_dofs": cb.get("dofs")}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import dgtd_b200 as dg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dgtd_b200 has no CPU fallback)")
    if args.gpus != n_gpus and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        tv = torch.tensor([v], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        return float(tv.item())

    # ---- parity gate (i): small reference fixtures over all ranks ----------------------------------------------------
    parity = None
    if world > 1 and not args.no_gate:
        worst, lines = parity_gate_fixtures(dg, dist, rank, world, local_rank)
        parity = {"fixtures": lines, "fixtures_worst_rel_l2": worst, "tolerance": PARITY_TOL}
        if not (worst <= PARITY_TOL):
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity gate failed", "parity": parity}))
            dist.destroy_process_group()
            return 3

    mesh, kw, t_start, dt, init = build_workload(args, dg, n_gpus)
    part = None
    if n_gpus > 1:
        part = mesh.partition(n_gpus, args.partition)
    # a non-default torch stream: the kernels and the torch.cuda.Event timers share it (the legacy default stream has
    # handle 0, which dgtd_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    def make_evolution():
        t0 = time.perf_counter()
        e = dg.Evolution(mesh, device=local_rank, rank=rank, nranks=n_gpus, partitioning=part, **kw)
        dt_setup = time.perf_counter() - t0
        e.set_stream(stream.cuda_stream)
        if world > 1:
            e.comm_init(fresh_unique_id(dg, dist, rank))
        return e, dt_setup

    ev, t_setup = make_evolution()
    N, nloc, Np = ev.N, ev.n_local, ev.Np
    # deterministic initial state from the node coordinates of the owned elements (local layout [6][n_local], pinned)
    gid = ev.local_elements()
    xyz = ev.node_coords()
    host = torch.empty(6 * nloc, dtype=torch.float64).pin_memory()
    hx = host.numpy()
    hx[:] = init(xyz.reshape(-1, Np, 3)[gid].reshape(-1, 3))
    x0_local = hx.copy()

    # ---- parity gate (ii): this workload, N ranks against one GPU ----------------------------------------------------
    if world > 1 and args.parity_steps > 0:
        refl = None

        def bench_size_parity(e):
            nonlocal refl
            e.set_state_local(hx)
            e.run(t_start, dt, args.parity_steps)
            mine = e.get_state_local().reshape(6, -1, Np)
            if refl is None:                                   # the single-GPU run of the same global mesh, once
                ref = torch.empty(6 * N, dtype=torch.float64, device="cuda")
                if rank == 0:
                    ev1 = dg.Evolution(mesh, device=local_rank, **kw)
                    ev1.set_state(init(xyz))
                    ev1.run(t_start, dt, args.parity_steps)
                    ref.copy_(torch.from_numpy(ev1.get_state()))
                    ev1.close()
                dist.broadcast(ref, 0)
                refl = ref.view(6, -1, Np)[:, torch.from_numpy(gid.astype(np.int64)).cuda()].cpu().numpy()
                del ref
                torch.cuda.empty_cache()
            num, den = float(np.sum((mine - refl) ** 2)), float(np.sum(refl ** 2))
            bad = np.abs(mine - refl).max(axis=(0, 2)) > 1e-13 * max(1e-300, float(np.abs(refl).max()))      # per owned element
            nbad = torch.tensor([int(bad.sum()), len(bad)], dtype=torch.float64, device="cuda")
            dist.all_reduce(nbad)
            rel_rank = math.sqrt(num / den) if den > 0 else math.sqrt(num)
            tot = torch.tensor([num, den], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot)
            return {"bench_size_rel_l2_max_over_ranks": allmax(rel_rank), "bench_size_rel_l2_global": math.sqrt(float(tot[0]) / float(tot[1])),
                    "bench_size_steps": args.parity_steps, "bench_size_dofs": 6 * N, "elements_differing": int(nbad[0].item()), "elements": int(nbad[1].item()),
                    "halo": {0: "none", 1: "nccl send/recv", 2: "peer-memory stores fused into the stage kernel"}[e.halo_mode()],
                    "how": "owned DOFs of every rank after the same steps of the same global mesh on rank 0's GPU alone"}

        parity = parity or {}
        res = bench_size_parity(ev)
        if not (res["bench_size_rel_l2_max_over_ranks"] <= PARITY_TOL) and ev.halo_mode() == 2 and not args.parity_report_only:
            # the fused peer-memory exchange disagrees with the single-GPU run: measure on the NCCL send/recv path instead of
            # reporting a number whose results are not the reference's, and say so
            parity["peer_memory_halo_rejected"] = res
            ev.close()
            os.environ["DGTD_B200_HALO"] = "nccl"
            ev, t_setup = make_evolution()
            res = bench_size_parity(ev)
        parity.update(res)
        if not (parity["bench_size_rel_l2_max_over_ranks"] <= PARITY_TOL) and not args.parity_report_only:
            if rank == 0:
                print(json.dumps({"error": "multi-GPU parity at the bench size failed", "parity": parity}))
            ev.close()
            dist.destroy_process_group()
            return 3
    xyz = None

    ev.set_state_local(x0_local)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t = t_start
    for _ in range(warm):
        t = ev.Step(t, dt)
    barrier()

    def timed(steps):
        nonlocal t
        l0 = ev.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        w0 = time.perf_counter()
        e0.record(stream)
        t = ev.run(t, dt, steps)
        e1.record(stream)
        barrier()
        w1 = time.perf_counter()
        return allmax(e0.elapsed_time(e1)), ev.launch_count() - l0, (w0, w1)

    ms, launches, win = timed(args.steps)
    value = 6.0 * N * 4 * args.steps / (ms * 1e-3)

    # ---- sustained window: thousands of steps back to back (Solver::run is a seconds-to-hours loop) ----------------------
    sustained = None
    if args.sustain_s > 0:
        ev.set_state_local(x0_local)       # the upwind operator is dissipative; restart from the same amplitude anyway
        t = t_start
        ns = int(math.ceil(args.sustain_s / (ms * 1e-3 / args.steps)))
        ms_s, _, win_s = timed(ns)
        sustained = {"value": 6.0 * N * 4 * ns / (ms_s * 1e-3), "unit": UNIT, "steps": ns, "seconds": ms_s * 1e-3, "ms_per_step": ms_s / ns}
        sustained["ratio_to_burst"] = sustained["value"] / value
        if rank == 0:
            sustained.update({k: v for k, v in sampler.window(*win_s).items()})

    # ---- end to end through host buffers: H2D state, fused step, D2H state, every step --------------------------------
    e2e_steps = max(1, args.e2e_steps)
    hx[:] = x0_local
    ev.set_state_local(hx)
    barrier()
    t0 = time.perf_counter()
    te = t_start
    # N = 1: exactly the three calls B200RK4Solver::Step makes on a host vector in the reference's numbering
    # (dgtd_set_state / dgtd_rk4_step / dgtd_get_state); N > 1: each rank moves its own partition (dgtd_*_state_local)
    put, get = (ev.set_state, ev.get_state) if n_gpus == 1 else (ev.set_state_local, ev.get_state_local)
    for _ in range(e2e_steps):
        put(hx)                     # host -> device (the ODESolver::Step(x, t, dt) contract: x lives on the host)
        te = ev.Step(te, dt)
        get(hx)                     # device -> host (synchronises)
    barrier()
    e2e_value = 6.0 * N * 4 * e2e_steps / allmax(time.perf_counter() - t0)

    # ---- end to end with the state kept resident (B200Evolution::Upload / RunUntil / B200Gather): one H2D, `steps` fused
    # steps with an asynchronous probe snapshot every 10 steps, one D2H of the final state ------------------------------
    res_steps = max(10, args.steps)
    probe_dofs = (gid[:: max(1, len(gid) // 256)][:256].astype(np.int64) * Np)        # first node of ~256 owned elements
    gth = dg.Gather(ev, probe_dofs)
    pout = torch.empty(6 * max(1, gth.n_local), dtype=torch.float64).pin_memory().numpy()[:6 * gth.n_local]
    hx[:] = x0_local
    barrier()
    t0 = time.perf_counter()
    ev.set_state_local(hx)
    tr = t_start
    for s in range(0, res_steps, 10):
        tr = ev.run(tr, dt, min(10, res_steps - s))
        if gth.n_local:
            gth.launch(pout)
    gth.wait()
    ev.get_state_local(hx)
    barrier()
    res_value = 6.0 * N * 4 * res_steps / allmax(time.perf_counter() - t0)
    gth.close()

    if rank == 0:
        sampler.stop_flag = True
        hbm, how = peaks()
        launch_ms = ms / (4 * args.steps)                    # the step is 4 back-to-back stage launches
        nloc_max = nloc                                      # rank 0's share (METIS parts differ by < 0.2 %)
        alg_bytes = b_alg(args.order) * 6.0 * nloc_max       # per launch, per GPU
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        traffic, per_stage = ncu_traffic(ev.kernel_info(), 6 * nloc)
        cfg = workload_config(args, n_gpus)
        if parity is not None and "bench_size_rel_l2_max_over_ranks" in parity:
            cfg["parity"] = parity["bench_size_rel_l2_max_over_ranks"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "run": {"dofs_total": 6 * N, "dofs_rank0": 6 * nloc, "dt": dt, "setup_s": t_setup, "halo_bytes_per_rhs_rank0": ev.halo_bytes(),
                        "halo": {0: "none", 1: "nccl send/recv", 2: "peer-memory stores fused into the stage kernel"}[ev.halo_mode()]},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": traffic, "traffic_per_stage": per_stage, "peak_source": how,
                             "alg_bytes_per_dof_update": b_alg(args.order), "alg_bytes_per_launch": alg_bytes,
                             "kernel": ev.kernel_info(), "avg_launch_ms": launch_ms},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 6 * N * 8, "d2h_bytes_per_step": 6 * N * 8,
                        "steps": e2e_steps, "how": ("dgtd_set_state(host) + dgtd_rk4_step + dgtd_get_state(host) per step = B200RK4Solver::Step, pinned host memory" if n_gpus == 1 else
                                "dgtd_set_state_local(host) + dgtd_rk4_step + dgtd_get_state_local(host) per step and rank, pinned host memory")},
                "e2e_resident": {"value": res_value, "unit": UNIT, "steps": res_steps, "h2d_bytes_total": 6 * N * 8, "d2h_bytes_total": 6 * N * 8 + (res_steps // 10) * 6 * 8 * len(probe_dofs) * n_gpus,
                                 "how": "B200Evolution::Upload (one H2D) + dgtd_rk4_run + asynchronous probe gather every 10 steps + one D2H of the final state"},
                "gpu_launches": int(launches),
                "clocks": sampler.result(*win)}
        if sustained is not None:
            line["sustained"] = sustained
        if parity is not None:
            line["parity"] = parity
        if not args.no_cpu and n_gpus == 1 and (args.workload == "c5" or args.cpu_same):
            try:
                same = args.workload != "c5"
                cb = cpu_reference(args, 5, 1, 3, cubes=args.cpu_cubes or 8, same_mesh=same)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:   # the GPU number must not be lost to a CPU-side failure
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    ev.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
