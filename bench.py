#!/usr/bin/env python
"""bench.py -- headline benchmark of the MG-PCG hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one complete multigrid-preconditioned CG solve (K x = f to ||r||/||f|| < 1e-10) of the
reference's own single-solve benchmark (python/CoarseningLevelBenchmark.py) on the 3D 256^3 Q1 grid
(BASELINE.json configs[2]): cantilever BC, rho = 0.5, E_min = 1e-5, nu = 0.3, 5 coarsening levels,
FMG-preconditioned PCG with 1 smoothing sweep, zero initial guess.
metric = DOF*iterations / s = N * numNodes * PCG iterations / solve time.

With --gpus W > 1 (torchrun, one rank per GPU) ONE solve is partitioned over the W GPUs: the grid is (256 W) x 256 x 256
(domain W x 1 x 1, weak scaling: a 256^3 slab per GPU), cut into slabs along axis 0 with NCCL ghost-plane exchange and
all-reduced PCG scalars (BASELINE.json configs[3] style); value = global DOF * iterations / max-over-ranks time.

Emits ONE JSON line (rank 0).  Timing: CUDA events on the solver's stream, max over ranks; inputs are
re-zeroed on the device before every step; the working set (>4 GB) is far larger than L2.
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
sys.path.insert(0, ROOT)

MATERIAL = dict(E=1.0, nu=0.3)
WORKLOADS = {
    # name: (grid, domain max, bc file, levels)
    "C3_pcg_256^3": ((256, 256, 256), (1.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 5),
    "C3_pcg_128^3": ((128, 128, 128), (1.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 4),
    "C3_pcg_64^3": ((64, 64, 64), (1.0, 1.0, 1.0), "3D/cantilever_flexion_E.bc", 3),
    "C1_pcg_256x128": ((256, 128), (2.0, 1.0), "mbb_N.bc", 3),
}
PCG = dict(max_iter=100, tol=1e-10, mg_iterations=1, mg_smoothing=1, fmg=True)
# PCG iterations of the ORACLE's full solve of each workload (tests/test_gpu_baseline_configs.py runs oracle and GPU side by side on
# these grids; log: profiles/r03f_pytest_baseline_configs.log): the CPU arm times a bounded sample and scales it to the full solve
ORACLE_ITERATIONS = {"C3_pcg_256^3": 14, "C3_pcg_128^3": 14, "C1_pcg_256x128": 16}
# executed FP64 instructions per unit of the level-0 kernels: the FP64-pipe view of the roofline
FP64_OPS = {"gs_l0": 414.0, "apply_l0": 185.0, "residual_l0": 188.0}   # gs_l0: neighbour form (331 DFMA + 59 DADD + 24 DMUL per node, SASS of k_gs3_nbt)

# algorithmic bytes per unit of work (SURVEY.md section 8d / DESIGN.md): fp64, 3D Q1
ALG_BYTES = {"gs_l0": 80.0, "apply_l0": 56.0, "residual_l0": 80.0, "gs_stencil": 80.0 + 27 * 9 * 8.0,
             "apply_stencil": 56.0 + 27 * 9 * 8.0, "residual_stencil": 80.0 + 27 * 9 * 8.0}


def setup(mod_sim, mod_mg, workload, data_dir):
    ne, dmax, bc, levels = WORKLOADS[workload]
    ne = np.array(ne)
    s = mod_sim(ne, np.zeros(len(ne)), np.array(dmax))
    s.set_isotropic(MATERIAL["E"], MATERIAL["nu"])
    s.set_interp(0, 1.0, 1e-5, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", bc))
    s.set_uniform_density(0.5)
    mg = mod_mg(s, levels)
    return s, mg


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_sample(workload, steps, warmup, threads=None, n_full=None):
    """Times the CPU restatement of the reference algorithm (oracle) on a bounded sample of the workload and scales it to the full
    solve with the SAME accounting as the GPU arm: one solve = one coarse-hierarchy rebuild (the reference's PCG rebuilds it on
    every call, MultigridSolver.hh:1104-1107) + n_full PCG iterations.  The rebuild is timed once, a step is the PCG capped at
    `cap` iterations on the prebuilt hierarchy, and value = DOF * n_full / (t_rebuild + n_full * t_iteration)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    L = oracle.lib()
    if threads:
        L.vfo_set_num_threads(threads)
    cores = L.vfo_num_threads()
    data_dir = os.path.join(ROOT, "voxelfem_b200", "data")
    s, mg = setup(oracle.OracleSim, oracle.OracleMG, workload, data_dir)
    f = s.build_load()
    N = s.N
    cap = 2
    n_full = n_full or ORACLE_ITERATIONS.get(workload, 14)
    t0 = time.perf_counter()
    mg.update_stiffness()
    t_fixed = time.perf_counter() - t0
    mg.set_stiffness_prebuilt(True)
    times, iters = [], 0
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        _, it, _ = mg.pcg(np.zeros_like(f), f, cap, PCG["tol"], PCG["mg_iterations"], PCG["mg_smoothing"], PCG["fmg"])
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt); iters = it
    t_step = float(np.mean(times))
    t_iter = t_step / max(iters, 1)           # the preamble (one residual, two norms) is charged to the iterations
    t_full = t_fixed + n_full * t_iter
    val = N * s.num_nodes * n_full / t_full
    return {"value": val, "unit": "DOF*iters/s", "cores": cores, "kind": "port",
            "sample": "%s: hierarchy rebuild timed once (%.2f s) + PCG capped at %d iterations per step on the prebuilt hierarchy (%.2f s per iteration, %d step(s)); "
                      "value = DOF * %d / (rebuild + %d * per-iteration time), the full solve's iteration count (same accounting as the GPU arm: one rebuild per solve)"
                      % (workload, t_fixed, cap, t_iter, steps, n_full, n_full),
            "ms_per_step": t_full * 1e3, "iters": n_full, "rebuild_s": t_fixed, "per_iteration_s": t_iter}


def setup_slab(capi, rank, world, data_dir):
    """This rank's slab of the weak-scaling grid (256 * world) x 256 x 256."""
    ne = np.array([256 * world, 256, 256])
    dmax = np.array([float(world), 1.0, 1.0])
    levels = 5 if world <= 2 else 6          # keeps the replicated coarsest grid at <= 4,131 DOF
    # first replicated level: the levels below it are windowed per GPU (ghost-plane exchange after every colour pass), the rest is
    # replicated on every rank (no communication, replayed as a CUDA graph).  Replicating level 3 as well pays while the replicated
    # grid is small (2 GPUs: 240 vs 251 ms, 4 GPUs: 231 vs 242 ms per solve); at 8 slabs level 3 has 280k nodes and stays windowed.
    first_rep = int(os.environ.get("VF_BENCH_FIRST_REP", "3" if world <= 4 else "4"))
    a, b = capi.slab_ranges(int(ne[0]), world, 2 ** first_rep)[rank]
    s = capi.SlabSim(ne, np.zeros(3), dmax, a, b)
    s.set_isotropic(MATERIAL["E"], MATERIAL["nu"])
    s.set_interp(0, 1.0, 1e-5, 3.0, 3.0)
    s.apply_bc_file(os.path.join(data_dir, "bcs", "3D/cantilever_flexion_E.bc"))
    s.set_uniform_density(0.5)
    mg = capi.SlabMG(s, levels, first_rep)
    return s, mg, ne, levels, first_rep


def extra_single_gpu(capi):
    """The other single-GPU configurations of BASELINE.json, recorded next to the headline: C1 (2D MBB solve), C2 (topopt iterations/s --
    the second half of BASELINE's metric) and C5 (layer-by-layer, layers/s)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_lbl
    import bench_topopt
    out = {}
    s1, m1 = setup(capi.Sim, capi.MG, "C1_pcg_256x128", capi.DATA_DIR)
    f1 = s1.build_load()
    m1.pcg(np.zeros_like(f1), f1, **PCG)
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps):
        m1.update_stiffness()
        _, it1, _ = m1.pcg(np.zeros_like(f1), f1, **PCG)
    dt = (time.perf_counter() - t0) / reps
    out["C1_pcg_256x128"] = {"pcg_iterations": int(it1), "pcg_iterations_oracle": ORACLE_ITERATIONS["C1_pcg_256x128"], "ms_per_solve_host_buffers": dt * 1e3,
                             "dof_iterations_per_s": 2 * s1.num_nodes * it1 / dt}
    r2 = bench_topopt.run(iters=50, profile=False)
    out["C2_topopt_128x64x64"] = {k: r2[k] for k in ("iterations", "topopt_iterations_per_s", "ms_per_iteration", "update_stiffness_ms", "compliance", "volume_constraint")}
    out["C2_topopt_128x64x64"]["pcg_iterations_mean"] = float(np.mean(r2["pcg_iterations"]))
    # a host-bound chain of 256 small solves whose wall time follows the host's load (30-95 layers/s between runs of the same code on the
    # same box): two runs, the faster one reported, both times listed
    runs5 = [bench_lbl.run() for _ in range(2)]
    r5 = min(runs5, key=lambda r: r["seconds"])
    out["C5_lbl_128x256x128"] = {k: r5[k] for k in ("layers", "layers_per_s", "seconds", "pcg_iterations_total", "dof_iterations_per_s", "objective")}
    out["C5_lbl_128x256x128"]["seconds_of_both_runs"] = [r["seconds"] for r in runs5]
    return out


def extra_c4(capi, torch, dist, rank, world, peak, fp64_peak):
    """BASELINE.json configs[3] (north_star's target): 512 x 256 x 256 compliance topopt partitioned into slabs over the N GPUs of this
    run (strong scaling), OC iterations timed on the device (max over ranks), with the roofline of the level-0 kernels of rank 0."""
    ne, dom, levels = (512, 256, 256), (2.0, 1.0, 1.0), 5
    first_rep = 4 if world > 2 else 3
    V, filters = 0.3, [("smooth", 3, 1), ("project", 1.0)]
    a, b = capi.slab_ranges(int(ne[0]), world, 2 ** first_rep)[rank]
    s = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b)
    s.set_isotropic(MATERIAL["E"], MATERIAL["nu"]); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0)
    s.apply_bc_file(os.path.join(capi.DATA_DIR, "bcs", "3D/cantilever_flexion_E.bc")); s.set_uniform_density(1.0)
    mg = capi.SlabMG(s, levels, first_rep)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=bytes(uid.cpu().numpy().tobytes()))
    top = capi.SlabProblem([(s, mg)], grp, filters, V)
    top.set_solver(100, 1e-5, 1, 2, True, False)
    top.set_vars(np.full(int(np.prod(ne)), 0.5 + np.arctanh((2 * V - 1) * np.tanh(0.5))))
    top.oc_step()
    iters = 5
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.ExternalStream(top.stream_handle))
    its = []
    for _ in range(iters):
        top.oc_step(); its.append(int(top.last_pcg_iters))
    e1.record(torch.cuda.ExternalStream(top.stream_handle))
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # one more iteration with CUDA events around every launch: per-kernel durations on rank 0
    mg.prof_reset(); mg.prof_enable(True)
    top.oc_step()
    dist.barrier(); torch.cuda.synchronize()
    mg.prof_enable(False)
    prof = mg.prof_report()
    kern = {}
    for k in ("gs_l0", "apply_l0", "residual_l0"):
        if k in prof and prof[k]["launches"]:
            p = prof[k]
            gbs = ALG_BYTES[k] * p["units"] / (p["ms"] * 1e-3) / 1e9
            kern[k] = {"launches": p["launches"], "ms": p["ms"], "achieved_GBs": gbs, "hbm_frac": gbs / peak,
                       "fp64_frac": (FP64_OPS[k] * p["units"] / (p["ms"] * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None}
    out = {"workload": "C4_topopt_512x256x256_slabs%d" % world, "scaling": "strong", "n_gpus": world, "iterations": iters, "ms_per_iteration": ms.item() / iters,
           "topopt_iterations_per_s": iters / (ms.item() * 1e-3), "pcg_iterations": its, "compliance": top.compliance(), "volume_constraint": top.constraint(),
           "levels": levels, "first_replicated_level": first_rep, "rank0_level0_kernels": kern}
    grp.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--cpu-workload", default=None, help="workload of the cpu_baseline leg (default: the headline workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel-family device-time breakdown to stderr")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations (C1, C2, C5 at one GPU; C4 at N > 1)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = args.workload or "C3_pcg_256^3"

    if args.impl == "reference":
        # The reference's own TBB/Eigen/CHOLMOD build is impossible in this image; this arm times
        # the CPU restatement of its algorithm (oracle/) with all host threads.  Rank 0 only.
        if rank != 0:
            return
        # same workload as the GPU arm (256^3); each step is a bounded sample of it: the PCG capped at 2 iterations on the prebuilt
        # hierarchy (a few seconds per step on the box's host cores), scaled to the full solve with one hierarchy rebuild per solve --
        # the GPU arm's accounting (cpu_sample); warm-up capped at one step so that the default --steps/--warmup ends within minutes
        wl = args.workload or workload
        # all host threads, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers
        cb = cpu_sample(wl, max(1, args.steps), max(0, min(args.warmup, 1)), threads=os.cpu_count())
        line = {"impl": "reference", "metric": "MG-PCG DOF*iterations per second (3D Q1, FMG-PCG, tol 1e-10)", "value": cb["value"], "unit": "DOF*iters/s",
                "n_gpus": 0, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "pcg_iterations_per_solve": cb["iters"], "step": "one hierarchy rebuild + full-solve iterations, scaled from a bounded sample",
                           "note": "CPU restatement of the reference algorithm (reference not buildable here: Eigen/TBB/CHOLMOD absent); bounded sample"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "DOF*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import ctypes as C
    import torch
    import torch.distributed as dist
    from voxelfem_b200 import capi
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    capi._check(capi.lib().vf_set_device(local_rank))
    L = capi.lib()
    grp = None
    if world > 1:
        s, mg, gne, levels, first_rep = setup_slab(capi, rank, world, capi.DATA_DIR)
        workload = "C3w_pcg_%dx256x256_slabs%d" % (256 * world, world)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=bytes(uid.cpu().numpy().tobytes()))
        global_dof = 3 * int(np.prod(gne + 1))
    else:
        s, mg = setup(capi.Sim, capi.MG, workload, capi.DATA_DIR)
    N = s.N
    ndof = N * s.num_nodes
    if world == 1:
        global_dof = ndof
    x = capi.DeviceArray(ndof)
    b = capi.DeviceArray(ndof)
    capi._check(L.vf_sim_build_load_vector_dev(s.h, b.ptr))
    stream = torch.cuda.ExternalStream(mg.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The reference's PCG rebuilds the coarse hierarchy on every call (MultigridSolver.hh:1104-1107): so does a timed step.  One GPU: the
    # solver is put into that mode and rebuilds in its first PCG iteration, exactly where the reference does; a slab group rebuilds
    # explicitly before the solve.
    if grp is None:
        mg.set_rebuild_every_solve(True)

    def step():
        x.zero()
        if grp is not None:
            mg.update_stiffness()
            return grp.pcg_dev([x], [b], **PCG)
        return mg.pcg_dev(x, b, **PCG)

    for _ in range(args.warmup):
        it, res = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.vf_reset_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    iters = 0
    for _ in range(args.steps):
        it, res = step()
        iters += it
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.vf_kernel_launch_count()
    # second pass over the same steps with a CUDA-event pair around every launch (per-kernel durations for the roofline
    # line; the preconditioner then runs as individual launches instead of the captured graph)
    mg.prof_reset(); mg.prof_enable(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        step()
    p1.record(stream)
    barrier()
    ms_prof = p0.elapsed_time(p1)
    mg.prof_enable(False)
    prof = mg.prof_report()
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        sl = s.build_load().reshape((s.plane_hi - s.plane_lo + 1, -1))
        bnorm2 = torch.tensor([float((sl[s.own_lo - s.plane_lo:s.own_hi - s.plane_lo + 1] ** 2).sum())], dtype=torch.float64, device="cuda")
        dist.all_reduce(bnorm2)
        relres = float(res[-1] / np.sqrt(bnorm2.item())) if len(res) else None
    else:
        relres = float(res[-1] / np.linalg.norm(s.build_load())) if len(res) else None

    # end-to-end through the host-pointer C ABI call (what the reference's binding does: copy u and f in, x out)
    xh = torch.zeros(ndof, dtype=torch.float64).pin_memory().numpy()
    bh = torch.zeros(ndof, dtype=torch.float64).pin_memory().numpy()
    u0h = torch.zeros(ndof, dtype=torch.float64).pin_memory().numpy()   # zero initial guess, as in python/CoarseningLevelBenchmark.py
    bh[:] = b.download()
    def e2e_step():
        if grp is not None:  # this rank's window: host -> device, partitioned solve, device -> host
            capi._check(L.vf_dev_upload(x.ptr, u0h, ndof)); capi._check(L.vf_dev_upload(b.ptr, bh, ndof))
            mg.update_stiffness()
            it_, _ = grp.pcg_dev([x], [b], **PCG)
            capi._check(L.vf_dev_download(xh, x.ptr, ndof))
            return it_
        itc = C.c_int(0)
        # the hierarchy rebuild happens inside the call (rebuild-every-solve mode), overlapped with the input copies
        # the reference binding's call shape: initial guess and load in (pinned host arrays), solution out (pinned host array)
        capi._check(L.vf_mg_pcg_io(mg.h, u0h, bh, xh, PCG["max_iter"], PCG["tol"], PCG["mg_iterations"], PCG["mg_smoothing"], int(PCG["fmg"]), 0, C.byref(itc), np.zeros(PCG["max_iter"] + 1), capi.PCG_CALLBACK(), None))
        return itc.value
    e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    e2e_iters = 0
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        e2e_iters += e2e_step()
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    # one partitioned solve: every rank ran the same iterations over the GLOBAL grid
    tot = torch.tensor([float(iters) * global_dof, float(e2e_iters) * global_dof], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    work, work_e2e = tot.tolist()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (sustained: kernel timed inside a long step)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    # FP64 FMA peak of this device, measured now (register-resident independent DFMA chains on every SM; vf_measure_fp64_peak)
    fp64_peak = C.c_double(0.0)
    capi._check(L.vf_measure_fp64_peak(C.byref(fp64_peak)))
    fp64_peak = fp64_peak.value
    extra = {}
    if not args.no_extra:
        try:
            if world > 1:
                extra["C4"] = extra_c4(capi, torch, dist, rank, world, peak, fp64_peak)
            else:
                extra.update(extra_single_gpu(capi))
        except Exception as e:   # the headline line must still be printed
            extra["error"] = "%s: %s" % (type(e).__name__, e)
    if rank != 0:
        grp.close()
        dist.destroy_process_group()
        return
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        ncu_traffic = {}

    def kernel_roofline(k):
        p = prof[k]
        sec_per_launch = p["ms"] * 1e-3 / p["launches"]
        units_per_launch = p["units"] / p["launches"]
        achieved = ALG_BYTES[k] * units_per_launch / sec_per_launch / 1e9
        r = {"bound": "hbm", "kernel": k, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
             "traffic": ncu_traffic.get(k, {}).get("bytes_per_launch"), "algorithmic_bytes_per_launch": ALG_BYTES[k] * units_per_launch,
             "launches": p["launches"], "avg_launch_us": sec_per_launch * 1e6, "share_of_step": p["ms"] / ms_prof}
        if k in FP64_OPS:
            tdfma = FP64_OPS[k] * units_per_launch / sec_per_launch / 1e12
            r["fp64"] = {"achieved": tdfma, "peak": fp64_peak, "unit": "T DFMA/s", "frac": tdfma / fp64_peak if fp64_peak else None,
                         "ops_per_unit": FP64_OPS[k], "peak_source": "measured live (vf_measure_fp64_peak)"}
        return r
    # Headline: the kernels BASELINE.json's north_star names -- the level-0 smoother (the reference's top-1 cost) first, then the
    # level-0 stiffness apply / residual.  The stored-stencil families of the coarse levels stay in `families`; their bytes are the
    # traffic of this implementation's stored-stencil design (1944 B of stencil per node), not SURVEY.md 8(d)'s matrix-free bytes.
    north_star = [k for k in ("gs_l0", "apply_l0", "residual_l0") if k in prof and prof[k]["launches"]]
    roofline = None
    if north_star:
        dom = max(north_star, key=lambda k: prof[k]["ms"])
        roofline = kernel_roofline(dom)
        roofline.update({
            "peak_source": peak_src,
            "north_star_kernels": {k: kernel_roofline(k) for k in north_star},
            "note": "algorithmic bytes per node (SURVEY.md 8d): gs_l0 80 per sweep, apply_l0 56, residual_l0 80; fp64.ops_per_unit = FP64 instructions this implementation executes per node (DESIGN.md section 3)",
            "families": {k: {"ms_per_launch": v["ms"] / v["launches"], "launches": v["launches"], "share": v["ms"] / ms_prof,
                             "achieved_GBs": (ALG_BYTES[k] * v["units"] / (v["ms"] * 1e-3) / 1e9) if k in ALG_BYTES else None,
                             "bytes_model": ("SURVEY.md 8(d) algorithmic bytes" if k.endswith("_l0") else "stored-stencil design traffic (1944 B stencil + field bytes per node)") if k in ALG_BYTES else None}
                         for k, v in prof.items() if v["launches"]}})
    if args.profile:
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            print("  %-18s launches %6d  ms %9.3f  share %5.1f%%  ms/launch %8.4f" % (k, v["launches"], v["ms"], 100 * v["ms"] / ms_prof, v["ms"] / v["launches"]), file=sys.stderr)
        print("  timed pass %.3f ms, instrumented pass %.3f ms (%d steps each)" % (ms, ms_prof, args.steps), file=sys.stderr)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:      # reported at N = 1 only
        cpu_baseline = cpu_sample(args.cpu_workload or workload, 1, 1, threads=os.cpu_count(), n_full=int(round(iters / args.steps)))
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    value = work / (ms * 1e-3)
    line = {
        "metric": "MG-PCG DOF*iterations per second (3D Q1, FMG-PCG, tol 1e-10)", "value": value, "unit": "DOF*iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "grid": [int(v) for v in gne] if world > 1 else list(WORKLOADS[workload][0]),
                   "levels": levels if world > 1 else WORKLOADS[workload][3], "pcg": PCG,
                   "pcg_iterations_per_solve": iters / args.steps, "pcg_iterations_oracle": ORACLE_ITERATIONS.get(workload), "final_relative_residual": relres,
                   "step": "x = 0; coarse-hierarchy rebuild (the reference's PCG rebuilds it on every call, MultigridSolver.hh:1104-1107); MG-PCG solve to tol",
                   "extra": extra,
                   "solve_time_s": ms * 1e-3 / args.steps,
                   "parallelism": ("one solve partitioned into %d slabs along axis 0 (256 element layers per GPU), levels 0-%d windowed with NCCL ghost-plane exchange, coarser levels replicated, PCG scalars all-reduced" % (world, first_rep - 1)) if world > 1 else "single GPU",
                   "timing": "value/ms_per_step: uninstrumented pass (preconditioner replayed as a captured CUDA graph); roofline: second pass of the same steps with CUDA events around every launch",
                   "l2": "working set >> 126 MB L2 (x,b,r,d,Ad = 5 x 407 MB + 4.7 GB of coarse stencils); inputs re-zeroed every step"},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": work_e2e / (ms_e2e * 1e-3), "unit": "DOF*iters/s", "h2d_bytes_per_step": 2 * ndof * 8 * world, "d2h_bytes_per_step": ndof * 8 * world,
                "ms_per_step": ms_e2e / n_e2e, "steps": n_e2e},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        grp.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
