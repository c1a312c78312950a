"""CPU tests of the host-side slab logic incl. a world_size-2 gloo run: the slabs partition the grid, every
node plane is owned exactly once, windows carry one ghost plane per neighbour, and owned-node reductions all-reduce to the
global value (this is what the PCG scalars rely on)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402


@pytest.mark.parametrize("ne0,nparts,align", [(16, 2, 2), (512, 8, 16), (48, 3, 4), (2048, 8, 16), (32, 4, 8)])
def test_slab_ranges_partition(ne0, nparts, align):
    from voxelfem_b200 import capi
    r = capi.slab_ranges(ne0, nparts, align)
    assert r[0][0] == 0 and r[-1][1] == ne0
    owned = np.zeros(ne0 + 1, dtype=int)
    for i, (a, b) in enumerate(r):
        assert a % align == 0 and b % align == 0 and b > a
        if i:
            assert a == r[i - 1][1]
        lo, hi, olo, ohi = capi.slab_window(ne0, a, b)
        assert lo == max(a - 1, 0) and hi == min(b + 1, ne0) and olo == a
        owned[olo:ohi + 1] += 1
    assert (owned == 1).all()


def test_slab_ranges_reject_indivisible():
    from voxelfem_b200 import capi
    with pytest.raises(AssertionError):
        capi.slab_ranges(24, 4, 16)


def _worker(rank, world, port, ne, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voxelfem_b200 import capi
    nn = np.array(ne) + 1
    field = np.random.default_rng(11).normal(size=(int(np.prod(nn)), 3))           # same global field on every rank
    a, b = capi.slab_ranges(ne[0], world, 4)[rank]
    lo, hi, olo, ohi = capi.slab_window(ne[0], a, b)
    win = field.reshape(tuple(nn) + (3,))[lo:hi + 1]
    own = win[olo - lo:ohi - lo + 1]
    t = torch.tensor([float((own * own).sum()), float(own.shape[0])], dtype=torch.float64)
    dist.all_reduce(t)
    # ghost planes: what rank r would receive from its neighbours equals the global field's planes
    # a part sends the owned planes next to its shared planes: global a + 1 to the left neighbour, b - 1 to the right one
    planes = [torch.from_numpy(np.ascontiguousarray(win[a + 1 - lo])), torch.from_numpy(np.ascontiguousarray(win[b - 1 - lo]))]
    ok = abs(t[0].item() - float((field * field).sum())) < 1e-9 * t[0].item() and int(t[1].item()) == nn[0]
    if world == 2:
        if rank == 0:
            recv = torch.empty_like(planes[0]); dist.send(planes[1], 1); dist.recv(recv, 1)
            ok = ok and np.array_equal(recv.numpy(), field.reshape(tuple(nn) + (3,))[b + 1])
        else:
            recv = torch.empty_like(planes[0]); dist.recv(recv, 0); dist.send(planes[0], 0)
            ok = ok and np.array_equal(recv.numpy(), field.reshape(tuple(nn) + (3,))[a - 1])
    out[rank] = 1 if ok else 0
    dist.destroy_process_group()


def test_owned_reduction_and_ghost_exchange_gloo_world2():
    world, ne = 2, (16, 4, 6)
    out = mp.get_context("spawn").Manager().dict()
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ne, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def _halo_worker(rank, world, port, ne, R, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from voxelfem_b200 import capi
    field = np.random.default_rng(5).normal(size=ne)                                # same global element field on every rank
    sb, se = capi.slab_ranges(ne[0], world, 4)[rank]
    ext = slab_ref.slab_halo_exchange([torch.from_numpy(np.ascontiguousarray(field[sb:se]))], [(sb, se)], ne[0], R, dist)[0]
    elo, ehi = capi.slab_halo_range(sb, se, ne[0], R)
    out[rank] = 1 if np.array_equal(ext.numpy(), field[elo:ehi]) else 0
    dist.destroy_process_group()


@pytest.mark.parametrize("world,R", [(2, 3), (3, 4)])
def test_filter_halo_exchange_gloo(world, R):
    """The filter halos of the slab-partitioned topology optimization (vf_group_top_*, vf_api.cu): after the exchange every rank holds
    exactly the global field's layers [sb - R, se + R) clipped at the grid -- over torch.distributed (gloo here, NCCL on GPUs)."""
    ne = (8 * world, 3, 5)
    out = mp.get_context("spawn").Manager().dict()
    port = 31000 + (os.getpid() % 2000) + world
    mp.spawn(_halo_worker, args=(world, port, ne, R, out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}


def test_filter_halo_exchange_local_parts():
    from voxelfem_b200 import capi
    ne, R = (24, 2, 3), 5
    field = np.random.default_rng(6).normal(size=ne)
    slabs = capi.slab_ranges(ne[0], 3, 8)
    ext = slab_ref.slab_halo_exchange([torch.from_numpy(np.ascontiguousarray(field[a:b])) for a, b in slabs], slabs, ne[0], R, None)
    for (a, b), e in zip(slabs, ext):
        lo, hi = capi.slab_halo_range(a, b, ne[0], R)
        assert np.array_equal(e.numpy(), field[lo:hi])
