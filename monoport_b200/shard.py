"""z-slab sharding of the query grid over GPUs: one process per GPU (torch.distributed / NCCL),
every rank evaluates its contiguous slab of z planes with the fused kernel, then ONE all-gather of the occupancy
volume over NVLink so every rank holds the full [R,R,R] volume for marching cubes.  No other collective."""
import torch


def slab_bounds(R, world_size):
    """Contiguous plane ranges [(z0, nz)] -- the first R % world_size ranks get one extra plane (257 = 8*32+1)."""
    base, extra = divmod(int(R), int(world_size))
    out, z = [], 0
    for r in range(world_size):
        nz = base + (1 if r < extra else 0)
        out.append((z, nz))
        z += nz
    return out


def max_slab(R, world_size):
    return -(-int(R) // int(world_size))


def gather_slabs(slab, R, rank, world_size, group=None, out=None):
    """All-gather unequal z slabs into the full [R,R,R] volume.  Slabs are padded to max_slab planes so a single
    `all_gather_into_tensor` moves everything (67.9 MB at 257^3: latency-bound, one collective is the point)."""
    import torch.distributed as dist
    bounds = slab_bounds(R, world_size)
    pad = max_slab(R, world_size)
    dev = slab.device
    send = slab
    if slab.shape[0] != pad:
        send = torch.zeros((pad, R, R), dtype=slab.dtype, device=dev)
        send[:slab.shape[0]] = slab
    recv = torch.empty((world_size * pad, R, R), dtype=slab.dtype, device=dev)
    if world_size == 1:
        recv.copy_(send)
    else:
        dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if out is None:
        out = torch.empty((R, R, R), dtype=slab.dtype, device=dev)
    if all(nz == pad for _, nz in bounds):
        out.copy_(recv)
    else:
        for r, (z0, nz) in enumerate(bounds):
            out[z0:z0 + nz] = recv[r * pad:r * pad + nz]
    return out


def query_grid_sharded(net, feat, calibs, R, b_min, b_max, rank, world_size, group=None, gather=True):
    """Dense R^3 occupancy volume evaluated slab-wise across ranks; returns the full volume (gather=True) or the
    local slab."""
    z0, nz = slab_bounds(R, world_size)[rank]
    slab = net.query_grid(feat, calibs, R, b_min, b_max, z0=z0, nz=nz)
    if not gather:
        return slab
    return gather_slabs(slab, R, rank, world_size, group)
