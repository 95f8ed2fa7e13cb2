#!/usr/bin/env python3
"""bench.py — DG DOF-updates/s (RHS + RK stage, FP64) of the fused RK4 hot path, and its reference arm.

Workload (BASELINE.json configs[4], the configuration the 1/2/4/8-GPU metric is quoted on): order-3 tetrahedral
PEC box, MakeCartesian3D-style Kuhn mesh with 32^3 cubes (196 608 tets = 23.6 M DOFs) PER GPU, the box growing along x
with the GPU count (weak scaling; 8 GPUs = 188.7 M DOFs ~ "200 M").  A "step" is one classical RK4 step = 4 fused
stage launches; DOF-updates = 6N * 4 per step.  Inputs are far larger than L2 (4 vectors x 189 MB per
GPU), so no L2 flush is needed between iterations.

  python bench.py [--gpus N --steps K --warmup W]           one JSON line (rank 0)
  python bench.py --impl reference [...]                    the reference's CPU algorithm (assembled CSR `global`
                                                            operator + mfem::RK4Solver, OpenMP) on the host cores

`value`  : device-resident throughput, CUDA events on the launching stream, max over ranks.
`e2e`    : the same step through the host-facing call a drop-in ODESolver makes (B200RK4Solver::Step on a HOST vector):
           H2D of the state from pinned memory + fused step + D2H of the new state, every step.
`roofline`: algorithmic bytes (42 B per DOF-update at order 3) per stage launch / average launch time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DG DOF-updates/s (RHS+RK stage, FP64)"
UNIT = "DOF-updates/s"
ORDER = 3
CUBES_PER_GPU = 32           # 32^3 cubes x 6 tets per GPU
B_ALG = {1: 40.0 + 200.0 / (6 * 4), 2: 40.0 + 200.0 / (6 * 10), 3: 40.0 + 200.0 / (6 * 20), 4: 40.0 + 200.0 / (6 * 35)}


def workload_name(n_gpus, cubes):
    return (f"3D tet PEC box (Kuhn Cartesian {cubes * n_gpus}x{cubes}x{cubes} cubes x6 tets), order {ORDER}, upwind alpha=1, "
            f"{cubes ** 3 * 6 * n_gpus} tets, {cubes ** 3 * 6 * n_gpus * (ORDER + 1) * (ORDER + 2) * (ORDER + 3)} DOFs, RCB slabs, classical RK4")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_info, dofs):
    """dram__bytes_read.sum + dram__bytes_write.sum per stage launch from the committed `ncu --set full` capture of this
    kernel on this workload (profiles/traffic.json, written from the .ncu-rep by hand); None when there is no capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    for rec in json.load(open(p)):
        if kernel_info.startswith(rec["kernel_prefix"]) and rec["dofs_per_gpu"] == dofs:
            return rec["dram_bytes_per_launch"]
    return None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML queries every 2 ms (nvidia_ml_py), nvidia-smi
    every 100 ms as fallback (one nvidia-smi call takes longer than a whole 20-step timed region)."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stamps, self.reasons, self.stop_flag = index, [], [], set(), False
        self.window = None            # (t0, t1) of the timed region, perf_counter
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
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                self.stamps.append(time.perf_counter())
                for nme, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        """Median SM clock of the samples taken inside the timed region; when the region is shorter than a few NVML
        queries the samples of the whole loaded phase (warm-up .. end of the timed region) are reported beside it."""
        allp = sorted(self.samples)
        inw = sorted(v for v, t in zip(self.samples, self.stamps) if self.window and self.window[0] <= t <= self.window[1])
        use = inw or allp
        return {"sm_mhz": use[len(use) // 2] if use else None, "sm_min_mhz": use[0] if use else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(inw), "samples_under_load": len(allp),
                "sm_mhz_under_load": allp[len(allp) // 2] if allp else None, "how": self.how}


def cpu_reference(cubes, steps, warmup, threads=None):
    """Time the reference's CPU algorithm (oracle/_ref/dgtd_ref: MFEM SparseMatrix::Mult + mfem::RK4Solver compiled from
    the reference sources, `-d omp`) on a bounded sample of the workload: the same box family at `cubes`^3 cubes."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dgtd_ref")
    threads = threads or os.cpu_count() or 1
    sample = f"same box family at {cubes}^3 cubes ({cubes ** 3 * 6} tets, {cubes ** 3 * 720} DOFs), {steps} RK4 steps after {warmup} warm-up; CSR assembly excluded"
    if os.path.exists(exe):
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_PROC_BIND="close")
        cmd = [exe, "bench", "--device", "omp", "--mesh", f"cart3d:{cubes}", "--order", str(ORDER), "--alpha", "1.0", "--bdr-all", "pec",
               "--init", "random:1", "--dt", "1e-4", "--steps", str(steps), "--warmup", str(warmup)]
        out = subprocess.run(cmd, capture_output=True, text=True, env=env, check=True).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        return {"value": d["dof_updates_per_s"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample,
                "ms_per_step": 1e3 * d["run_s"] / max(1, steps), "nnz": d["nnz"], "assemble_s": d["assemble_s"]}
    # the portable numpy restatement (matrix-free, single thread)
    import numpy as np
    from oracle.dgtd_oracle import PEC, HesthavenOracle, Problem
    import dgtd_b200 as dg
    m = dg.Mesh.cartesian3d(cubes)
    v, e, ea, b, ba = m.arrays()
    O = HesthavenOracle(Problem(v, e.astype(np.int64), ea, b.astype(np.int64), ba, ORDER, 1.0, {a: PEC for a in range(1, 7)}))
    x = np.random.default_rng(1).standard_normal(6 * O.N)
    for _ in range(warmup):
        x = O.rk4_step(x, 0.0, 1e-4)
    t0 = time.perf_counter()
    for _ in range(steps):
        x = O.rk4_step(x, 0.0, 1e-4)
    dt = time.perf_counter() - t0
    return {"value": 6 * O.N * 4 * steps / dt, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample + " (numpy port)",
            "ms_per_step": 1e3 * dt / steps}


def main():
    global ORDER
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cubes", type=int, default=CUBES_PER_GPU, help="cubes per axis per GPU (default 32)")
    ap.add_argument("--order", type=int, default=ORDER, help="polynomial order (default 3 = the headline workload; others are side measurements)")
    ap.add_argument("--cpu-cubes", type=int, default=6, help="box size of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-steps", type=int, default=3)
    args = ap.parse_args()
    ORDER = args.order
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warm = max(3, args.warmup) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 20))
        cb = cpu_reference(args.cpu_cubes, steps, min(args.warmup, 2))
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 2), "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.gpus, args.cubes), "sample": cb["sample"]},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import dgtd_b200 as dg

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (dgtd_b200 has no CPU fallback)")
    n_gpus = world if world > 1 else 1
    if args.gpus != n_gpus and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    if world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cubes = args.cubes
    mesh = dg.Mesh.cartesian3d(cubes * n_gpus, cubes, cubes, sx=float(n_gpus), sy=1.0, sz=1.0)
    bdr = {a: dg.BC_PEC for a in range(1, 7)}
    ev = dg.Evolution(mesh, order=ORDER, alpha=1.0, bdr=bdr, device=local_rank, rank=rank, nranks=n_gpus)
    # a non-default torch stream: the kernels and the torch.cuda.Event timers share it (the legacy default stream has
    # handle 0, which dgtd_set_stream reads as "use the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ev.set_stream(stream.cuda_stream)
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(dg.Evolution.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        ev.comm_init(bytes(idt.cpu().numpy().tobytes()))
    N, nloc = ev.N, ev.n_local
    # initial state: small random field on the owned dofs (local layout [6][n_local], pinned host memory)
    host = torch.empty(6 * nloc, dtype=torch.float64).pin_memory()
    hx = host.numpy()
    rng = np.random.default_rng(7 + rank)
    hx[:] = rng.standard_normal(6 * nloc) * 1e-3
    ev.set_state_local(hx)
    h = 1.0 / cubes
    dt = 0.05 * h / (ORDER * ORDER)            # well inside the RK4 stability region; the rate does not depend on it

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t = 0.0
    for _ in range(warm):
        t = ev.Step(t, dt)
    barrier()
    l0 = ev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.perf_counter()
    e0.record(stream)
    ev.run(t, dt, args.steps)
    e1.record(stream)
    barrier()
    sampler.window = (w0, time.perf_counter())
    if rank == 0:
        sampler.stop_flag = True
    ms = e0.elapsed_time(e1)
    launches = ev.launch_count() - l0
    if rank == 0:
        sampler.stop_flag = True
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = 6.0 * N * 4 * args.steps / (ms * 1e-3)

    # ---- end to end through host buffers: H2D state, fused step, D2H state, every step --------------------------------
    e2e_steps = max(1, args.e2e_steps)
    ev.set_state_local(hx)
    barrier()
    t0 = time.perf_counter()
    te = 0.0
    # N = 1: exactly the three calls B200RK4Solver::Step makes on a host vector in the reference's numbering
    # (dgtd_set_state / dgtd_rk4_step / dgtd_get_state); N > 1: each rank moves its own partition (dgtd_*_state_local)
    put, get = (ev.set_state, ev.get_state) if n_gpus == 1 else (ev.set_state_local, ev.get_state_local)
    for _ in range(e2e_steps):
        put(hx)                     # host -> device (the ODESolver::Step(x, t, dt) contract: x lives on the host)
        te = ev.Step(te, dt)
        get(hx)                     # device -> host (synchronises)
    barrier()
    e2e_s = time.perf_counter() - t0
    te2 = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te2, op=dist.ReduceOp.MAX)
    e2e_value = 6.0 * N * 4 * e2e_steps / float(te2.item())

    norm2 = torch.tensor([ev.norm2_local()], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(norm2)
    if rank == 0:
        hbm, how = peaks()
        stage_launches = 4 * args.steps
        launch_ms = ms / stage_launches                      # the step is 4 back-to-back stage launches (+ halo pack at N>1)
        alg_bytes = B_ALG[ORDER] * 6.0 * nloc                # per launch, per GPU
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(n_gpus, cubes), "dofs_per_gpu": 6 * nloc, "dt": dt,
                           "l2": "inputs (4 x %.0f MB per GPU) are larger than L2, no flush" % (6 * nloc * 8 / 1e6),
                           "halo_bytes_per_rhs": ev.halo_bytes(), "halo": {0: "none", 1: "nccl send/recv", 2: "peer-memory stores fused into the stage kernel"}[ev.halo_mode()], "state_norm": float(norm2.sqrt().item())},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             "traffic": ncu_traffic(ev.kernel_info(), 6 * nloc), "peak_source": how,
                             "alg_bytes_per_dof_update": B_ALG[ORDER], "alg_bytes_per_launch": alg_bytes,
                             "kernel": ev.kernel_info(), "avg_launch_ms": launch_ms},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 6 * nloc * 8 * n_gpus, "d2h_bytes_per_step": 6 * nloc * 8 * n_gpus,
                        "steps": e2e_steps, "how": ("dgtd_set_state(host) + dgtd_rk4_step + dgtd_get_state(host) per step = B200RK4Solver::Step, pinned host memory" if n_gpus == 1 else
                                "dgtd_set_state_local(host) + dgtd_rk4_step + dgtd_get_state_local(host) per step and rank, pinned host memory")},
                "gpu_launches": int(launches),
                "clocks": sampler.result()}
        if not args.no_cpu and n_gpus == 1:
            try:
                cb = cpu_reference(args.cpu_cubes, 5, 1)
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
