"""Run under torchrun (one rank per GPU): the NCCL slab group must reproduce the undivided solve computed on each rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from voxelfem_b200 import capi  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
capi._check(capi.lib().vf_set_device(local))
ok = True
for ne, levels, first_rep, fmg in [((32, 16, 16), 3, 2, True), ((16 * world, 8, 8), 2, 1, False), ((64, 32, 16), 3, 3, True)]:
    dom = (ne[0] / 8.0, ne[1] / 8.0, ne[2] / 8.0)
    bc = os.path.join(capi.DATA_DIR, "bcs", "3D", "cantilever_flexion_E.bc")
    rho = np.random.default_rng(5).uniform(0.2, 1.0, int(np.prod(ne)))

    def prep(s, r):
        s.set_isotropic(1.0, 0.3); s.set_interp(0, 1.0, 1e-4, 3.0, 3.0); s.apply_bc_file(bc); s.set_densities(r)
    sref = capi.Sim(np.array(ne), np.zeros(3), np.array(dom)); prep(sref, rho)
    mref = capi.MG(sref, levels)
    f = sref.build_load()
    u_ref, it_ref, res_ref = mref.pcg(np.zeros_like(f), f, 60, 1e-10, 1, 1, fmg)
    a, b = capi.slab_ranges(ne[0], world, 2 ** first_rep)[rank]
    s = capi.SlabSim(np.array(ne), np.zeros(3), np.array(dom), a, b); prep(s, s.window_of_elements(rho))
    mg = capi.SlabMG(s, levels, first_rep)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(capi.SlabGroup.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    grp = capi.SlabGroup([mg], rank=rank, world=world, unique_id=uid.cpu().numpy().tobytes())
    n = s.num_nodes * 3
    x, bd = capi.DeviceArray(n), capi.DeviceArray(n)
    bd.upload(capi.to_soa(s.window_of_nodal(f)))
    it, res = grp.pcg_dev([x], [bd], 60, 1e-10, 1, 1, fmg)
    w = capi.from_soa(x.download(), 3)
    err = np.linalg.norm(w - s.window_of_nodal(u_ref)) / np.linalg.norm(u_ref)
    good = it == it_ref and err < 1e-9 and np.allclose(res, res_ref, rtol=1e-4)   # the residual-emitting sweep adds its contributions in no fixed order: rounding-level differences, amplified in the last entries of a history falling by ten orders
    print("rank %d grid %s levels %d first_rep %d fmg %s: iterations %d (undivided %d), window rel err %.2e -> %s" % (rank, ne, levels, first_rep, fmg, it, it_ref, err, "OK" if good else "FAIL"), flush=True)
    ok = ok and good
    grp.close()
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if flag.item() else 0)
