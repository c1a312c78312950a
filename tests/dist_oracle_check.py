"""torchrun script (one rank per GPU): the NCCL slab group's solve against the CPU ORACLE's undivided solve (tests/test_gpu_nccl.py
launches it when at least two devices are visible).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_oracle_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
from oracle import OracleLBL, OracleMG, OracleSim  # noqa: E402
from voxelfem_b200 import capi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
capi._check(capi.lib().vf_set_device(local))
ok = True
for ne, levels, first_rep in [((64, 32, 32), 3, 2), ((32 * world, 32, 32), 3, 3)]:
    dom = (ne[0] / 32.0, ne[1] / 32.0, ne[2] / 32.0)
    bc = os.path.join(capi.DATA_DIR, "bcs", "3D", "cantilever_flexion_E.bc")
    rho = np.random.default_rng(5).uniform(0.2, 1.0, int(np.prod(ne)))
    so = OracleSim(np.array(ne), np.zeros(3), np.array(dom))
    so.set_isotropic(1.0, 0.3); so.set_interp(0, 1.0, 1e-4, 3.0, 3.0); so.apply_bc_file(bc); so.set_densities(rho)
    f = so.build_load()
    u_ref, it_ref, res_ref = OracleMG(so, levels).pcg(np.zeros_like(f), f, 100, 1e-10, 1, 1, True)
    a, b = capi.slab_ranges(ne[0], world, 2 ** first_rep)[rank]
    s = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b)
    s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0); s.apply_bc_file(bc); s.set_densities(s.window_of_elements(rho))
    mg = capi.SlabMG(s, levels, first_rep)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=uid.cpu().numpy().tobytes())
    n = s.num_nodes * 3
    x, bd = capi.DeviceArray(n), capi.DeviceArray(n)
    bd.upload(capi.to_soa(s.window_of_nodal(f)))
    it, res = grp.pcg_dev([x], [bd], 100, 1e-10, 1, 1, True)
    w = capi.from_soa(x.download(), 3)
    wref = s.window_of_nodal(u_ref)
    err = np.linalg.norm(w - wref) / np.linalg.norm(wref)
    fw = s.window_of_nodal(f)
    good = abs(it - it_ref) <= 1 and err <= 1e-6
    print("rank %d/%d grid %s: PCG iterations nccl=%d oracle=%d, window rel-L2(u) vs oracle %.3e, window f.u nccl=%.12e oracle=%.12e -> %s"
          % (rank, world, ne, it, it_ref, err, float((fw * w).sum()), float((fw * wref).sum()), "OK" if good else "FAIL"), flush=True)
    ok = ok and good
    grp.close()
# layer-by-layer evaluator over NCCL (vf_group_lbl_*) against the oracle's undivided evaluator
ne = np.array([16 * world, 8, 8]); dom = ne.astype(float) / ne[0]
rho = np.random.default_rng(3).uniform(0.3, 1.0, int(np.prod(ne)))


def prep(s, r):
    s.set_isotropic(1.0, 0.3); s.set_interp(1, 1.0, 1e-4, 3.0, 3.0)
    s.add_dirichlet([0, 0, 0], [-1, -1e-9, -1], [100, 1e-9, 100], 7)
    s.set_gravity(np.array([0.0, -1.0, 0.0])); s.set_densities(r)


so = OracleSim(ne, np.zeros(3), dom); prep(so, rho)
oe = OracleLBL(OracleMG(so, 2)); oe.select_init_method("N=3")
oi, oc = oe.run(True, 1, 200, 1e-11, 1, 1, False)
a, b = capi.slab_ranges(int(ne[0]), world, 2)[rank]
s = capi.SlabSim(ne, np.zeros(3), dom, a, b); prep(s, s.window_of_elements(rho))
mg = capi.SlabMG(s, 2, 1)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=uid.cpu().numpy().tobytes())
ge = capi.SlabLBL(grp, ne); ge.select_init_method("N=3")
gi, gc = ge.run(True, 1, 200, 1e-11, 1, 1, False)
eobj = abs(ge.objective() - oe.objective()) / abs(oe.objective())
egrad = np.abs(ge.gradient() - oe.gradient()).max() / np.abs(oe.gradient()).max()
good = len(gi) == len(oi) and np.abs(gi.astype(int) - oi.astype(int)).max() <= 1 and np.abs(gc - oc).max() < 1e-8 * np.abs(oc).max() and eobj < 1e-8 and egrad < 1e-8
print("rank %d/%d layer-by-layer grid %s: %d layers, PCG iterations nccl=%d oracle=%d, objective rel err %.2e, gradient rel err %.2e -> %s"
      % (rank, world, tuple(int(v) for v in ne), len(gi), int(gi.sum()), int(oi.sum()), eobj, egrad, "OK" if good else "FAIL"), flush=True)
ok = ok and good
ge.close(); grp.close()
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
