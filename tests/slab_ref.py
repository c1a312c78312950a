"""Host-side statement of the filter-halo layout of the slab-partitioned topology optimization (vf_group_top_*, vf_api.cu):
which element layers a slab holds once its halo is attached, and what they must contain after the exchange.  Test infrastructure:
drives the same layout over torch.distributed (gloo on CPU) so that the N > 1 logic is covered without GPUs."""
from voxelfem_b200.capi import slab_halo_range


def slab_halo_exchange(owned, slabs, ne0, R, dist=None):
    """Arrays over the slab + halo layers of every local part: owned layers from `owned` (torch tensors, layer axis first, any
    device), halo layers from the neighbouring slabs' owned layers.  `slabs`: the (sb, se) element-layer ranges of the local parts.
    dist is None: all slabs of the grid are local and ordered (device copies); else the single local part is rank dist.get_rank()
    of an initialised torch.distributed group (NCCL on GPUs, gloo in the CPU tests) and halos travel as batched isend / irecv."""
    import torch
    ext = []
    for (sb, se), o in zip(slabs, owned):
        elo, ehi = slab_halo_range(sb, se, ne0, R)
        e = torch.zeros((ehi - elo,) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device)
        e[sb - elo:se - elo] = o
        ext.append(e)
    if dist is None:
        for i, (sb, se) in enumerate(slabs):
            elo, ehi = slab_halo_range(sb, se, ne0, R)
            if i > 0:
                k = sb - elo; ext[i][:k] = owned[i - 1][owned[i - 1].shape[0] - k:]
            if i + 1 < len(slabs):
                k = ehi - se; ext[i][se - elo:] = owned[i + 1][:k]
        return ext
    (sb, se), o, e = slabs[0], owned[0], ext[0]
    elo, ehi = slab_halo_range(sb, se, ne0, R)
    rank, world = dist.get_rank(), dist.get_world_size()
    ops, kl, kr = [], sb - elo, ehi - se
    assert o.shape[0] >= R, "a slab must hold at least R element layers"
    lo_send = o[:R].contiguous(); hi_send = o[o.shape[0] - R:].contiguous()
    lo_recv = torch.empty((kl,) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device) if kl else None
    hi_recv = torch.empty((kr,) + tuple(o.shape[1:]), dtype=o.dtype, device=o.device) if kr else None
    if rank > 0: ops += [dist.P2POp(dist.isend, lo_send, rank - 1), dist.P2POp(dist.irecv, lo_recv, rank - 1)]
    if rank + 1 < world: ops += [dist.P2POp(dist.isend, hi_send, rank + 1), dist.P2POp(dist.irecv, hi_recv, rank + 1)]
    if ops:
        for r in dist.batch_isend_irecv(ops): r.wait()
    if kl: e[:kl] = lo_recv
    if kr: e[se - elo:] = hi_recv
    return ext
