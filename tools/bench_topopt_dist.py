"""Config 4 (BASELINE.json configs[3]): 3D cantilever 512x256x256 SIMP compliance topology optimization partitioned into slabs
along axis 0, one rank per GPU (torchrun), NCCL ghost-plane / filter-halo exchange and all-reduced scalars:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_topopt_dist.py [iters] [small]

First a reduced-size problem is run both partitioned and undivided (on every rank) and compared; then the C4 iterations are
timed (device time, max over ranks).  Same problem definition as tools/bench_topopt.py (C2) at the larger grid, 5 levels."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelfem_b200 import capi  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
capi._check(capi.lib().vf_set_device(local))
BC = os.path.join(capi.DATA_DIR, "bcs", "3D", "cantilever_flexion_E.bc")
V = 0.3
FILTERS = [("smooth", 3, 1), ("project", 1.0)]


def make_part(ne, dom, levels, first_rep):
    a, b = capi.slab_ranges(int(ne[0]), world, 2 ** first_rep)[rank]
    s = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b)
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0); s.apply_bc_file(BC)
    s.set_uniform_density(1.0)
    mg = capi.SlabMG(s, levels, first_rep)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=bytes(uid.cpu().numpy().tobytes()))
    return s, mg, grp


def x_start(ne):
    x0 = np.full(int(np.prod(ne)), 0.5 + np.arctanh((2 * V - 1) * np.tanh(0.5)))     # ProjectionFilter(1).invert(V)
    return x0


def parity_check():
    ne, dom, levels, first_rep = (16 * max(world, 2) * 2, 16, 16), None, 3, 2
    dom = (ne[0] / 16.0, 1.0, 1.0)
    x0 = x_start(ne) * (1 + 0.05 * np.sin(np.arange(int(np.prod(ne)))))
    s, mg, grp = make_part(ne, dom, levels, first_rep)
    top = capi.SlabProblem([(s, mg)], grp, FILTERS, V)
    top.set_solver(200, 1e-12, 1, 2, True, False)
    top.set_vars(x0)
    rs = capi.Sim(np.array(ne), np.zeros(3), np.array(dom))
    rs.set_isotropic(1.0, 0.3); rs.set_interp(0, 1.0, 1e-4, 3.0, 3.0); rs.apply_bc_file(BC); rs.set_uniform_density(1.0)
    ref = capi.Problem(capi.MG(rs, levels), FILTERS, V)
    ref.set_solver(200, 1e-12, 1, 2, True, False)
    ref.set_vars(x0)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    ok = True
    for it in range(2):
        e = dict(compliance=abs(top.compliance() - ref.compliance()) / abs(ref.compliance()), constraint=abs(top.constraint() - ref.constraint()),
                 gradient=rel(top.objective_gradient(), ref.objective_gradient()))
        na, nb = top.oc_step(), ref.oc_step()
        e["design_vars"] = rel(top.design_vars(), ref.design_vars())
        good = e["compliance"] < 1e-8 and e["constraint"] < 1e-10 and e["gradient"] < 1e-7 and e["design_vars"] < 1e-6 and na == nb
        ok = ok and good
        print("rank %d parity grid %s iteration %d: %s  OC evaluations %d / %d -> %s" % (rank, ne, it, {k: "%.2e" % v for k, v in e.items()}, na, nb, "OK" if good else "MISMATCH"), flush=True)
    grp.close()
    return ok


def bench(ne, dom, levels, first_rep, iters):
    s, mg, grp = make_part(ne, dom, levels, first_rep)
    top = capi.SlabProblem([(s, mg)], grp, FILTERS, V)
    top.set_solver(100, 1e-5, 1, 2, True, False)
    top.set_vars(x_start(ne))
    top.oc_step()                                   # warm-up iteration
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    capi.lib().vf_reset_kernel_launch_count()
    t0 = time.perf_counter()
    e0.record(torch.cuda.ExternalStream(top.stream_handle))
    its, evals = [], []
    for _ in range(iters):
        evals.append(top.oc_step()); its.append(top.last_pcg_iters)
    e1.record(torch.cuda.ExternalStream(top.stream_handle))
    dist.barrier(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    c, v = top.compliance(), top.constraint()
    if rank == 0:
        print(json.dumps({"workload": "C4_topopt_%dx%dx%d_slabs%d" % (ne + (world,)), "n_gpus": world, "iterations": iters, "ms_per_iteration": ms[0].item() / iters,
                          "topopt_iterations_per_s": iters / (ms[0].item() * 1e-3), "wall_ms_per_iteration": ms[1].item() / iters, "pcg_iterations": its,
                          "oc_constraint_evals": evals, "compliance": c, "volume_constraint": v, "gpu_launches_rank0": int(capi.lib().vf_kernel_launch_count()),
                          "levels": levels, "first_replicated_level": first_rep, "filter_halo_planes": top.R}), flush=True)
    grp.close()


if __name__ == "__main__":
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    ok = parity_check()
    if "small" in sys.argv[2:]:
        bench((128, 64, 64), (2.0, 1.0, 1.0), 3, 2, iters)
    else:
        bench((512, 256, 256), (2.0, 1.0, 1.0), 5, 4, iters)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)
