#!/usr/bin/env python
"""bench.py -- D-equation PCG/DIC hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--impl reference]

A STEP is one segregated D-equation solve, i.e. what one DEqn.solve() of unsTotalLagrangianStress::evolve()
(unsTotalLagrangianStress.C:908) hands to the lduMatrix::solver boundary: three scalar PCG/DIC component
solves (Dx, Dy, Dz) that share the laplacian coefficients, controls `tolerance 1e-9; relTol 0.01`
(run/stressFoam/plateHole/plateHole/system/fvSolution:18-27), x0 = 0.

Workload at N=1: BASELINE.json configs[1], the synthetic 3-D hex cantilever with 1 M cells (200x50x100),
steel, fixedDisplacement at x=0, traction on x=L.  N>1: configs[2], the 64 M-cell (400^3... see --cells) box
decomposed into one sub-domain per GPU with processor-patch halos and global reductions over NCCL.

metric  : PCG cell-updates/s = cells x PCG iterations / time  (whole job, all ranks); solves/s beside it.
The three component solves of a step go through the component-batched entry points (gpupcg_set_matrix_vector /
gpupcg_solve_vector: fvMatrix<vector>::solve's component loop in one launch sequence, every kernel advances the three
systems together); --scalar-solves runs them one after the other through gpupcg_set_matrix / gpupcg_solve instead.
value   : matrix, x0, b resident in HBM (gpupcg_set_matrix_vector once, gpupcg_solve_vector_device per step).
e2e     : the same step through the reference-facing C-ABI calls with HOST buffers: gpupcg_set_matrix_vector(diag per
          component, upper) + gpupcg_solve_vector(x, b), pinned host memory, H2D/D2H inside the timed region.
roofline: dominant kernel of the step by measured time, algorithmic bytes / CUDA-event time,
          against MEASURED_PEAKS.json hbm_gbs; the Amul figure BASELINE.json names is in roofline_amul.
cpu_baseline: the CPU oracle (port of the reference's foam-extend PCG/DIC loops, oracle/ldu_oracle.c) on a
          bounded sample of the same system, serial.
--impl reference: the oracle with every host core (one decomposePar-style sub-domain per thread).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# The contract is ONE JSON line on stdout.  Libraries write banners there on their own (NCCL prints its version line
# to fd 1 at communicator creation whatever NCCL_DEBUG_FILE says), so fd 1 is pointed at stderr for the whole run and
# the JSON line is written to the saved, real stdout at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the batched kernels (three systems per launch) from the
# committed `ncu --set full` capture of this workload (profiles/r1d_vector_solve.md)
NCU_TRAFFIC_1M_VECTOR = {"amul": 113.2e6 + 12.9e6, "dic": (140.6e6 + 15.5e6) + (173.3e6 + 38.2e6), "p_update": 64.0e6 + 0.5e6}

TOL, RELTOL = 1e-9, 0.01
TRACTION = (0.5e4, 0.25e4, -1.0e4)          # all three components get a non-zero source
METRIC = "D-eqn PCG cell-updates/s (cells x PCG iterations / s; 3 component solves per step)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", default=None, help="nx,ny,nz override (default: 200,50,100 at N=1; 400,400,400 at N>1)")
    ap.add_argument("--cpu-maxiter", type=int, default=0,
                    help="iteration cap of the bounded CPU sample (default: ~4e8 cell-updates per solve, 4..40)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scalar-solves", action="store_true",
                    help="three sequential scalar component solves per step (gpupcg_solve) instead of the batched vector solve")
    ap.add_argument("--no-evolve", action="store_true", help="skip the device-resident evolve() outer-iteration measurement")
    ap.add_argument("--no-64m", action="store_true",
                    help="skip the single-GPU run of the 64 M-cell case (scale_base_64M, the N = 1 point of the strong-scaling curve)")
    ap.add_argument("--no-8m", action="store_true",
                    help="skip the extra 8.1 M-cell kernel measurements (Amul / DIC / assembly roofline fractions the north star quotes)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_system(nx, ny, nz):
    from fluidstructureinteraction_b200 import cases
    s = cases.cantilever_system(nx, ny, nz, tz=1.0)       # unit traction; scaled per component below
    return s


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle with all host threads (decomposed, one sub-domain per thread)
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, nx, ny, nz):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from fluidstructureinteraction_b200 import decompose
    cores = os.cpu_count() or 1
    threads = cores                                 # every host core: one decomposePar sub-domain per thread
    s = build_system(nx, ny, nz)
    n = s["nCells"]
    px, py, pz = host_proc_grid(threads, nx, ny, nz)
    cellProc = decompose.simple_cell_proc(nx, ny, nz, px, py, pz)
    oracle.set_num_threads(threads)
    times, iters = [], []
    doms = decompose.decompose_ldu(s["l"], s["u"], s["diag"], s["upper"], s["b"], np.zeros(n), cellProc, threads)
    bunit = [d["b"] for d in doms]
    for it in range(args.warmup + args.steps):
        dt, its = 0.0, 0
        for tr in TRACTION:                         # the three component solves of one DEqn.solve(), as in the GPU arm
            for d, bu in zip(doms, bunit):
                d["b"] = bu*tr; d["x"] = np.zeros(bu.shape[0])
            t0 = time.perf_counter()
            perf, _ = oracle.pcg_solve_multi(doms, TOL, RELTOL, 0, args.cpu_maxiter)
            dt += time.perf_counter() - t0
            its += perf["nIterations"]
        if it >= args.warmup:
            times.append(dt); iters.append(its)
    tot = float(np.sum(times))
    value = n*float(np.sum(iters))/tot
    sample = ("the three component solves (Dx, Dy, Dz) of the same %dx%dx%d system per step, each capped at %d PCG iterations, "
              "decomposed simple(%d %d %d) over all %d host cores, one sub-domain per OpenMP thread (stand-in for mpirun -np %d)"
              % (nx, ny, nz, args.cpu_maxiter, px, py, pz, threads, threads))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3*tot/args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(nx, ny, nz, args.gpus), "cells": n, "tolerance": TOL, "relTol": RELTOL},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "iterations_per_step": float(np.mean(iters)), "host_cores": cores,
    }
    emit(line)


def host_proc_grid(threads, nx, ny, nz):
    """`simple (px py pz)` with px*py*pz = threads: factors of two peeled off the longest remaining block edge."""
    p = [1, 1, 1]
    n = [float(nx), float(ny), float(nz)]
    t = threads
    f = 2
    while t > 1:
        while t % f:
            f += 1
        k = int(np.argmax([n[i]/p[i] for i in range(3)]))
        p[k] *= f
        t //= f
    return tuple(p)


def workload_name(nx, ny, nz, gpus):
    if gpus == 1:
        return "configs[1]: synthetic 3-D hex cantilever %dx%dx%d = %d cells, linear-elastic D equation" % (nx, ny, nz, nx*ny*nz)
    return "configs[2]: synthetic hex cantilever %dx%dx%d = %d cells decomposed simple over %d B200" % (nx, ny, nz, nx*ny*nz, gpus)


# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(s, n, maxiter):
    import oracle
    oracle.set_num_threads(1)
    x = np.zeros(n)
    t0 = time.perf_counter()
    perf, _ = oracle.pcg_solve(s["l"], s["u"], s["diag"], s["upper"], x, s["b"]*TRACTION[2], TOL, RELTOL, 0, maxiter)
    dt = time.perf_counter() - t0
    return {"value": n*perf["nIterations"]/dt, "unit": "cell-updates/s", "cores": 1, "kind": "port",
            "sample": "one Dz component solve of the same system, first %d PCG iterations (%.1f s), oracle/ldu_oracle.c serial"
                      % (perf["nIterations"], dt),
            "ms_per_iteration": 1e3*dt/max(1, perf["nIterations"])}


def kernel_bytes(n, nF, nsys):
    """Algorithmic bytes per launch.  `systems`: the reference's bytes for nsys independent systems;
    `shared`: the minimum when the nsys systems of one fvMatrix<vector>::solve share upper()/lduAddr() (what the batched
    kernels can at best move) -- Amul nsys*24N + 16F, DIC nsys*24N + 32F, updates nsys*24N / nsys*48N."""
    per = {"amul": (24*n, 16*nF), "dic": (24*n, 32*nF), "p_update": (24*n, 0), "xr_update": (48*n, 0)}
    return {k: {"systems": nsys*(a + b), "shared": nsys*a + b} for k, (a, b) in per.items()}


def kernel_entry(t, nsys, kb, peak):
    e = {"ms": t, "systems_per_launch": nsys, "algorithmic_bytes": kb["shared"], "achieved_gbs": kb["shared"]/t/1e6,
         "frac": kb["shared"]/t/1e6/peak}
    if nsys > 1:
        e["algorithmic_bytes_as_independent_systems"] = kb["systems"]
        e["frac_as_independent_systems"] = kb["systems"]/t/1e6/peak
    return e


def ncu_traffic(nx, ny, nz, batched):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/r2_ncu_traffic.json: kernel -> bytes, with the capture's file and commit)."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    if not os.path.exists(p):
        return {}, None
    with open(p) as f:
        t = json.load(f)
    key = "%dx%dx%d_%s" % (nx, ny, nz, "batched3" if batched else "scalar")
    return t.get(key, {}), t.get("source")


def measure_evolve(nx, ny, nz, nOuter=4):
    """One evolve() call (unsTotalLagrangianStress.C:769-1009) resident on the device: ms per OUTER iteration, split
    into assembly | DEqn.solve() | correctBoundaryConditions + relax + least-squares vol->point + grad/fGrad + residual."""
    from fluidstructureinteraction_b200 import geometry, meshes, lsq
    from fluidstructureinteraction_b200.solid import GpuSolidCase
    t0 = time.perf_counter()
    m = meshes.hex_box(nx, ny, nz)
    geom = geometry.compute(m)
    pt = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
    L = lsq.build(m, geom, pt)
    t_setup = time.perf_counter() - t0
    mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
    nBF, nIF = m.nFaces - m.nInternalFaces, m.nInternalFaces
    tr = np.zeros((nBF, 3)); p = m.patch("loaded")
    tr[p["start"] - nIF:p["start"] - nIF + p["size"]] = TRACTION
    c = GpuSolidCase(m, geom, pt, mu, lam, cases.STEEL["rho"], traction=tr, lsq=L)
    c.evolve(1, 0.0, tolerance=TOL, relTol=RELTOL)                      # warm-up outer iteration
    t1 = time.perf_counter()
    r = c.evolve(nOuter, 0.0, tolerance=TOL, relTol=RELTOL)             # convergenceTolerance 0: exactly nOuter iterations
    wall = time.perf_counter() - t1
    tvp = c.gm.vol_to_point(reps=10)
    c.close()
    k = r["nOuterIterations"]
    nP = m.nPoints
    return {"workload": "hex cantilever %dx%dx%d, %d outer iterations of evolve() resident on the device" % (nx, ny, nz, k),
            "outer_iterations": k, "pcg_iterations": int(r["totalPcgIterations"]),
            "ms_per_outer_iteration": 1e3*wall/k, "assemble_ms": r["assemble_ms"]/k, "solve_ms": r["solve_ms"]/k,
            "update_ms": r["update_ms"]/k, "host_bytes_per_outer_iteration": 48,
            "vol_to_point": {"ms": tvp, "kernel": "k_lsq_interpolate", "points": nP,
                             "algorithmic_bytes": int(L["entry"].shape[0]*28 + nP*(24 + 24 + 24 + 4 + 1) + 24*m.nCells + 24*nBF)},
            "lsq_setup_s": t_setup}


def measure_assembly(nx, ny, nz, peak):
    """GPU assembly of the D equation and the vertex-based gradients on the same mesh (one outer iteration's worth of
    evolve() around DEqn.solve()): device time of the kernels on resident data, algorithmic bytes per SURVEY.md 8d."""
    import fluidstructureinteraction_b200 as pkg
    from fluidstructureinteraction_b200 import geometry, meshes
    m = meshes.hex_box(nx, ny, nz)
    geom = geometry.compute(m)
    patches = {"fixed": "fixedDisplacement", "loaded": "tractionDisplacement", "free": "tractionDisplacement"}
    gm = pkg.GpuMesh(m, geom, patches)
    rng = np.random.default_rng(0)
    nC, nF, nBF = m.nCells, m.nFaces, m.nFaces - m.nInternalFaces
    mu, lam = cases.lame(cases.STEEL["E"], cases.STEEL["nu"])
    D = 1e-5*rng.standard_normal((nC, 3)); pD = 1e-5*rng.standard_normal((m.nPoints, 3))
    G = 1e-5*rng.standard_normal((nC, 9)); Gf = 1e-5*rng.standard_normal((nF, 9))
    trac = np.zeros((nBF, 3)); pres = np.zeros(nBF); fixed = np.zeros((nBF, 3)); pgrad = np.zeros((nBF, 3))
    # one untimed pass (first-launch cost, clocks ramping up after the host-side mesh build), then the best of three
    # timed batches of 10 back-to-back assemblies (CUDA events on the launching stream; the working set is >> L2)
    args_a = (np.full(nC, mu), np.full(nC, lam), np.full(nC, cases.STEEL["rho"]), np.zeros(3), D, G, Gf, trac, pres, fixed, 1.0)
    gm.assemble_D(*args_a, reps=3)
    a = min((gm.assemble_D(*args_a, reps=10) for _ in range(3)), key=lambda r: r["kernel_ms"])
    gm.gradients(D, pD, G, fixed, pgrad, 1.0, reps=2)
    gms = min(gm.gradients(D, pD, G, fixed, pgrad, 1.0, reps=10)[3] for _ in range(3))
    gm.close()
    b_asm = (104 + 392 + 360)*nC          # laplacian + explicit flux/div + skew correction (hex)
    b_grad = (200 + 384)*nC               # cell gradient + face gradient
    return {"assemble_D": {"ms": a["kernel_ms"], "algorithmic_bytes": b_asm, "achieved_gbs": b_asm/a["kernel_ms"]/1e6,
                           "frac": b_asm/a["kernel_ms"]/1e6/peak,
                           "kernels": "k_asm_faces + k_asm_cells + k_asm_patches"},
            "gradients": {"ms": gms, "algorithmic_bytes": b_grad, "achieved_gbs": b_grad/gms/1e6, "frac": b_grad/gms/1e6/peak,
                          "kernels": "k_grad_cells + k_grad_patches + k_fgrad_faces"}}


from fluidstructureinteraction_b200 import cases  # noqa: E402


def measure_8m(peak):
    """BASELINE.json's target is quoted on >= 8 M-cell meshes on one B200: the same kernels on the 400x100x200 hex
    cantilever (8.0 M cells), device-resident, L2 flushed between launches, CUDA events on the launching stream."""
    import fluidstructureinteraction_b200 as pkg
    nx, ny, nz = 400, 100, 200
    s = build_system(nx, ny, nz)
    n, nF = s["nCells"], s["l"].shape[0]
    g = pkg.GpuPCG(s["l"], s["u"], n)
    g.set_matrix(s["diag"], s["upper"])
    x = np.zeros(n)
    g.solve(x, s["b"]*-1e4, 1e-9, 0.01, 0, 4)
    out = {"workload": "synthetic hex cantilever %dx%dx%d = %d cells" % (nx, ny, nz, n), "cells": n}
    kb1, kb3 = kernel_bytes(n, nF, 1), kernel_bytes(n, nF, 3)
    for k in ("amul", "dic", "p_update", "xr_update"):
        out[k] = kernel_entry(g.time_kernel(k, reps=10, flushL2=True), 1, kb1[k], peak)
    # the batched kernels (what a D-equation solve runs): one launch advances the three component systems
    g.set_matrix_vector(np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1)), s["upper"])
    if g.vector_mode(3) in (1, 2):          # 2: the step runs concurrent scalar solves; the batched kernels are timed all the same
        xv = np.zeros((n, 3))
        g.solve_vector(xv, np.ascontiguousarray(np.stack([s["b"]*t for t in TRACTION], axis=1)), 1e-9, 0.01, 0, 4, history=False)
        for k in ("amul", "dic", "p_update", "xr_update"):
            out[k + "_batched3"] = kernel_entry(g.time_kernel_vector(3, k, reps=10, flushL2=True), 3, kb3[k], peak)
    g.close()
    del s
    out.update(measure_assembly(nx, ny, nz, peak))
    return out


def run_ours(args, nx, ny, nz):
    import torch
    import torch.distributed as dist
    import fluidstructureinteraction_b200 as pkg
    from fluidstructureinteraction_b200 import cases, decompose

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg.load_library()

    nGlobal = nx*ny*nz
    nFglobal = (nx - 1)*ny*nz + nx*(ny - 1)*nz + nx*ny*(nz - 1)
    if world > 1:
        # one decomposePar `simple` sub-domain per rank, written down directly (no rank builds the global system)
        px, py, pz = decompose.proc_grid(world)
        s = cases.cantilever_subdomain(nx, ny, nz, px, py, pz, rank, tz=1.0)
        g = pkg.GpuPCG(s["l"], s["u"], s["nCells"], s["patchFaceCells"], s["patchNbrDomain"], device=local)
        uid = [pkg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(rank, world, uid[0])
        diag, upper, bunit, bou = s["diag"], s["upper"], s["b"], s["patchBouCoeffs"]
        n = s["nCells"]
    else:
        s = build_system(nx, ny, nz)
        g = pkg.GpuPCG(s["l"], s["u"], nGlobal, device=local)
        diag, upper, bunit, bou = s["diag"], s["upper"], s["b"], []
        n = nGlobal
    nF = upper.shape[0]
    info = g.plan_info()
    stream = torch.cuda.current_stream()
    g.set_stream(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------------------------------
    vector = not args.scalar_solves
    diagC = np.ascontiguousarray(np.stack([diag]*3, axis=1))                # addBoundaryDiag per component: equal here
    bC = np.ascontiguousarray(np.stack([bunit*t for t in TRACTION], axis=1))
    bouC = [np.ascontiguousarray(np.stack([bb]*3, axis=1)) for bb in bou]
    if vector:
        g.set_matrix_vector(diagC, upper, patchBouCoeffsCmpt=bouC)
        vmode = g.vector_mode(3)
        d_bv = torch.from_numpy(bC).cuda()
        d_xv = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    else:
        g.set_matrix(diag, upper, patchBouCoeffs=bou)
        vmode = None
        d_b = [torch.from_numpy(bunit*t).cuda() for t in TRACTION]
        d_x = [torch.zeros(n, dtype=torch.float64, device="cuda") for _ in TRACTION]

    def step_device():
        if vector:
            d_xv.zero_()
            perfs, _ = g.solve_vector_device(3, d_xv.data_ptr(), d_bv.data_ptr(), TOL, RELTOL, 0, 1000)
            return sum(p["nIterations"] for p in perfs)
        its = 0
        for c in range(3):
            d_x[c].zero_()
            perf, _ = g.solve_device(d_x[c].data_ptr(), d_b[c].data_ptr(), TOL, RELTOL, 0, 1000)
            its += perf["nIterations"]
        return its

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = g.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 0
    for _ in range(args.steps):
        iters += step_device()
    e1.record()
    barrier()
    launches = g.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_dev = float(ms.item())

    # ---- A/B inside the same run: the component-batched kernels on one stream (what round 1 / 2a measured) -----------
    batched_ab = None
    if vector and vmode == 2:
        os.environ["GPUPCG_VECTOR_CONCURRENT"] = "0"
        step_device()
        barrier()
        e0.record()
        itb = 0
        for _ in range(args.steps):
            itb += step_device()
        e1.record()
        barrier()
        del os.environ["GPUPCG_VECTOR_CONCURRENT"]
        msb = float(e0.elapsed_time(e1))
        batched_ab = {"path": "GPUPCG_VECTOR_CONCURRENT=0: three systems per launch, one stream", "ms_per_step": msb/args.steps,
                      "value": nGlobal*itb/(msb*1e-3), "unit": "cell-updates/s", "iterations_per_step": itb/args.steps}

    # ---- end-to-end arm: host buffers through gpupcg_set_matrix + gpupcg_solve ------------------------------------
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    keep = []
    _, h_diag = pinned(diag); keep.append(_)
    _, h_upper = pinned(upper); keep.append(_)
    _, h_diagC = pinned(diagC); keep.append(_)
    _, h_bC = pinned(bC); keep.append(_)
    _, h_xC = pinned(np.zeros((n, 3))); keep.append(_)
    h_b, h_x = [], []
    for t in TRACTION:
        tb, nb = pinned(bunit*t); keep.append(tb); h_b.append(nb)
        tx, nxx = pinned(np.zeros(n)); keep.append(tx); h_x.append(nxx)

    def step_e2e():
        if vector:
            g.set_matrix_vector(h_diagC, h_upper, patchBouCoeffsCmpt=bouC)
            h_xC[:] = 0.0
            perfs, _ = g.solve_vector(h_xC, h_bC, TOL, RELTOL, 0, 1000, history=False)
            return sum(p["nIterations"] for p in perfs)
        its = 0
        for c in range(3):
            g.set_matrix(h_diag, h_upper if c == 0 else None, patchBouCoeffs=bou)
            h_x[c][:] = 0.0
            perf, _ = g.solve(h_x[c], h_b[c], TOL, RELTOL, 0, 1000, history=False)
            its += perf["nIterations"]
        return its

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    e0.record()
    iters_e2e = 0
    for _ in range(args.steps):
        iters_e2e += step_e2e()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = float(ms2.item())
    h2d = 3*8*n + 8*nF + 3*16*n          # diag per component + upper + x0 and b per component (either path)
    d2h = 3*8*n

    # ---- the drop-in path an unmodified foam-extend reaches with `solver gpuPCG;` alone: fvMatrix<vector>::solve's own
    #      component loop, i.e. 3 x (gpupcg_set_matrix + gpupcg_solve) with host buffers
    dropin = None
    if vector:
        def step_scalar_e2e():
            its = 0
            for c in range(3):
                g.set_matrix(h_diag, h_upper if c == 0 else None, patchBouCoeffs=bou)
                h_x[c][:] = 0.0
                perf, _ = g.solve(h_x[c], h_b[c], TOL, RELTOL, 0, 1000, history=False)
                its += perf["nIterations"]
            return its
        step_scalar_e2e()
        barrier()
        e0.record()
        its_s = 0
        nS = max(1, min(args.steps, 3))
        for _ in range(nS):
            its_s += step_scalar_e2e()
        e1.record()
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        dropin = {"path": "3 x (gpupcg_set_matrix + gpupcg_solve), host buffers: what `solver gpuPCG;` in fvSolution gives without "
                          "touching fvMatrix<Type>::solve", "value": nGlobal*its_s/(float(ms3.item())*1e-3), "unit": "cell-updates/s",
                  "ms_per_step": float(ms3.item())/nS, "steps": nS, "iterations_per_step": its_s/nS}
        g.set_matrix_vector(diagC, upper, patchBouCoeffsCmpt=bouC)      # back to the batched matrix for the kernel timings

    # ---- per-kernel roofline (rank 0, live CUDA-event timing on the launching stream, L2 flushed) --------------
    line = None
    if rank == 0:
        peak, peak_src = peaks()
        reps = 20
        batched = bool(vector and vmode == 1)
        concurrent = bool(vector and vmode == 2)    # three scalar solves in flight at once: the step's kernels are the scalar ones
        nsys = 3 if batched else 1          # systems one launch advances
        kb = kernel_bytes(n, nF, nsys)
        kern_batched = None
        if concurrent:
            # for comparison: the component-batched kernels (the other way to run the step: GPUPCG_VECTOR_CONCURRENT=0)
            kb3 = kernel_bytes(n, nF, 3)
            kern_batched = {k: kernel_entry(g.time_kernel_vector(3, k, reps=reps, flushL2=True), 3, kb3[k], peak)
                            for k in ("amul", "dic", "p_update", "xr_update")}
        in_flight = None
        if concurrent and world == 1:
            # what the device does DURING the step: the same kernel of the three systems in flight at once (three handles, three
            # streams, three host threads -- as the step runs them); per-launch duration under that load and the aggregate rate
            import threading
            hs = []
            for k in range(3):
                hk = pkg.GpuPCG(s["l"], s["u"], nGlobal, device=local)
                hk.set_matrix(diag, upper)
                hs.append(hk)
            in_flight = {}
            kb1 = kernel_bytes(n, nF, 1)
            for kname in ("dic", "amul"):
                ts = [None]*3

                def work(k, kname=kname, ts=ts):
                    torch.cuda.set_device(local)
                    ts[k] = hs[k].time_kernel(kname, reps=30, flushL2=False)
                th = [threading.Thread(target=work, args=(k,)) for k in range(3)]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()
                if all(t_ is not None for t_ in ts):
                    tm = sum(ts)/3.0
                    in_flight[kname] = {"systems_in_flight": 3, "ms_per_launch": tm, "algorithmic_bytes_per_launch": kb1[kname]["shared"],
                                        "aggregate_achieved_gbs": 3*kb1[kname]["shared"]/tm/1e6,
                                        "aggregate_frac": 3*kb1[kname]["shared"]/tm/1e6/peak,
                                        "l2": "no flush: the three systems' working sets (3 x %.0f MB) exceed the 126 MB L2" % (kb1[kname]["shared"]/1e6)}
            for hk in hs:
                hk.close()
        if vector and not batched:
            g.set_matrix(diag, upper, patchBouCoeffs=bou)
        kern = {}
        for k in ("amul", "dic", "p_update", "xr_update"):
            t = g.time_kernel_vector(3, k, reps=reps, flushL2=True) if batched else g.time_kernel(k, reps=reps, flushL2=True)
            kern[k] = kernel_entry(t, nsys, kb[k], peak)
        pencil = info["nTiles"] > 0 and info["tileRowsMax"] > 4096 and info["maxRounds"]*32 == info["tileRowsMax"]
        # on the tile sweeps of small sub-domains the x/r update rides in the staging of the next forward sweep (no
        # launch of its own); the pencil sweeps and large sub-domains run k_xr_update as a kernel of the iteration
        fused_xr = ((not pencil) and os.environ.get("GPUPCG_FUSE_XR", "1") != "0" and n <= 3000000) or \
                   (pencil and not batched and world == 1 and os.environ.get("GPUPCG_FUSE_XR_PENCIL", "1") != "0")
        tot = sum(v["ms"] for k, v in kern.items() if not (fused_xr and k == "xr_update"))
        for k, v in kern.items():
            v["share_of_iteration"] = None if (fused_xr and k == "xr_update") else v["ms"]/tot
        dom_k = max(kern, key=lambda k: kern[k]["ms"])
        sfx = "_v<3>" if batched else ""
        sweep = ("k_dic_pencil<%d,0|2> forward + backward" % nsys) if pencil else ("k_dic_tile%s forward + backward" % sfx)
        if pencil and fused_xr:
            sweep += "; in the solve the forward sweep also carries the x/r update of the previous iteration (update warp), timed here without it"
        names = {"amul": "k_amul_tile%s (lduMatrix::Amul + wApA)" % sfx,
                 "dic": sweep + " (DIC precondition + wArA)",
                 "p_update": "k_p_update%s" % sfx, "xr_update": "k_xr_update%s (+ sum|rA|)" % sfx}
        traffic, traffic_src = ncu_traffic(nx, ny, nz, batched) if world == 1 else ({}, None)

        def roof(k):
            r = {"kernel": names[k], "bound": "hbm", "achieved": kern[k]["achieved_gbs"], "peak": peak,
                 "unit": "GB/s", "frac": kern[k]["frac"], "traffic": traffic.get(k),
                 "traffic_source": traffic_src if traffic.get(k) is not None else None, "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": kern[k]["algorithmic_bytes"],
                 "algorithmic_bytes_definition": ("%d systems sharing upper()/lduAddr(): SURVEY.md 8d per-system vector bytes x %d + the "
                                                  "coefficient and label bytes once" % (nsys, nsys)) if nsys > 1 else "SURVEY.md 8d",
                 "systems_per_launch": nsys, "ms_per_launch": kern[k]["ms"], "share_of_iteration": kern[k]["share_of_iteration"]}
            if nsys > 1:
                r["frac_as_independent_systems"] = kern[k]["frac_as_independent_systems"]
            if in_flight and k in in_flight:
                # `frac` above is the kernel of ONE system timed alone; in the step three of them are in flight
                r["in_flight"] = in_flight[k]
            return r

        value = nGlobal*iters/(ms_dev*1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev/args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(nx, ny, nz, world), "cells": nGlobal, "internal_faces": int(nFglobal),
                       "tolerance": TOL, "relTol": RELTOL, "solver": "gpuPCG", "preconditioner": "DIC",
                       "component_solves": ("batched: gpupcg_solve_vector, 3 systems per launch" if batched else
                                            "concurrent: gpupcg_solve_vector runs the 3 systems as 3 scalar solves in flight at once "
                                            "(one stream per component; kernel times below are of one system's kernel alone, "
                                            "share_of_iteration is of one system's serial chain)" if concurrent else
                                            "sequential: 3 x gpupcg_solve"),
                       "sweeps": "streaming pencils (kernels_pencil.cuh)" if pencil else "tiles (kernels.cuh / kernels_vec.cuh)",
                       "parallelism": "1 sub-domain/GPU x %d" % world,
                       "scaling_workload": ("N = 1 runs configs[1] (1 M cells, the config the metric is quoted on); N >= 2 run "
                                            "configs[2] (64 M cells, strong scaling).  The 1 -> N curve of the 64 M-cell case "
                                            "starts at scale_base_64M of the N = 1 line, not at its `value`."),
                       "l2": "working set (%.0f MB) vs 126 MB L2: inputs larger than L2 at >= 1M cells; kernel timings flush L2" %
                             (cases.pcg_iteration_bytes(n, nF)/1e6)},
            "solves_per_s": 3*args.steps/(ms_dev*1e-3), "iterations_per_step": iters/args.steps,
            "cells_per_gpu": n,
            "e2e": {"value": nGlobal*iters_e2e/(ms_e2e*1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e/args.steps,
                    "solves_per_s": 3*args.steps/(ms_e2e*1e-3),
                    "path": "gpupcg_set_matrix_vector + gpupcg_solve_vector (pinned host buffers)" if vector else
                            "3 x (gpupcg_set_matrix + gpupcg_solve)"},
            "dropin_scalar_e2e": dropin, "batched_mode_same_run": batched_ab,
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": roof(dom_k), "roofline_amul": roof("amul"), "kernels": kern, "kernels_batched_mode": kern_batched,
            "plan": {k: info[k] for k in ("nRows", "nLevels", "maxLevelRows", "sweepGrid", "streamGrid", "nTiles")},
        }
        if world == 1 and nGlobal <= 4000000:
            line["assembly"] = measure_assembly(nx, ny, nz, peak)
        if not args.no_cpu_baseline:
            if world == 1:
                line["cpu_baseline"] = cpu_baseline(s, nGlobal, args.cpu_maxiter)
            else:
                # N > 1: the same serial port on the 1 M-cell cantilever (a rate; the 64 M-cell system does not fit a bounded sample)
                s1 = build_system(200, 50, 100)
                line["cpu_baseline"] = cpu_baseline(s1, s1["nCells"], 20)
                line["cpu_baseline"]["sample"] += " -- 200x50x100 cantilever (rate comparison for the N > 1 lines)"
        else:
            line["cpu_baseline"] = None
    g.close()
    del g
    if world == 1 and rank == 0 and not args.no_8m and not args.cells:
        line["targets_8M"] = measure_8m(peak)
    if world == 1 and rank == 0 and not args.no_evolve and nGlobal <= 4000000:
        line["evolve"] = measure_evolve(nx, ny, nz)
    if world == 1 and rank == 0 and not args.no_64m and not args.cells:
        line["scale_base_64M"] = measure_64m(args)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)


def measure_64m(args):
    """configs[2] on ONE GPU: the 400^3 = 64 M-cell cantilever, batched three-component solve, device-resident --
    the N = 1 point of the 64 M-cell strong-scaling curve (the N >= 2 lines run the same system decomposed)."""
    import torch
    import fluidstructureinteraction_b200 as pkg
    nx, ny, nz = 400, 400, 400
    s = build_system(nx, ny, nz)
    n = s["nCells"]
    g = pkg.GpuPCG(s["l"], s["u"], n)
    g.set_stream(torch.cuda.current_stream().cuda_stream)
    diagC = np.ascontiguousarray(np.stack([s["diag"]]*3, axis=1))
    g.set_matrix_vector(diagC, s["upper"])
    del diagC
    d_b = torch.from_numpy(np.ascontiguousarray(np.stack([s["b"]*t for t in TRACTION], axis=1))).cuda()
    d_x = torch.zeros((n, 3), dtype=torch.float64, device="cuda")
    del s
    steps = 3

    def step():
        d_x.zero_()
        perfs, _ = g.solve_vector_device(3, d_x.data_ptr(), d_b.data_ptr(), TOL, RELTOL, 0, 1000)
        return sum(p["nIterations"] for p in perfs)
    step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    its = sum(step() for _ in range(steps))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    g.close()
    return {"workload": workload_name(nx, ny, nz, 1).replace("configs[1]", "configs[2] on one GPU"), "cells": n, "n_gpus": 1,
            "steps": steps, "warmup": 1, "value": n*its/(ms*1e-3), "unit": "cell-updates/s", "ms_per_step": ms/steps,
            "iterations_per_step": its/steps}


def main():
    args = parse_args()
    if args.cells:
        nx, ny, nz = [int(v) for v in args.cells.split(",")]
    elif args.gpus == 1:
        nx, ny, nz = 200, 50, 100
    else:
        nx, ny, nz = 400, 400, 400
    if args.cpu_maxiter <= 0:
        args.cpu_maxiter = int(min(40, max(4, 4e8/(nx*ny*nz))))
    if args.impl == "reference":
        run_reference(args, nx, ny, nz)
    else:
        run_ours(args, nx, ny, nz)


if __name__ == "__main__":
    main()
