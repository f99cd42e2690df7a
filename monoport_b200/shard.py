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


# ------------------------------------------------------------------------------------------------------------
# Fused slab exchange (EXPERIMENTAL, opt-in; validated on multi-GPU boxes only): instead of the all-gather every rank's
# kernel stores its slab straight into the full volumes of ALL ranks over NVLink peer memory (mp_query_grid_peers), so
# compute and transfer are one kernel; what is left of the collective is a barrier.
# ------------------------------------------------------------------------------------------------------------
class _DeviceArray:
    """Zero-copy torch view of a raw device pointer (via __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerVolumes:
    """`n_sets` (default 2, alternating per frame) full [R,R,R] float32 volumes per rank, each mapped into every other
    rank of the node.  Collective constructor: call it on all ranks of `group` (one process per GPU)."""

    def __init__(self, R, rank, world_size, device, group=None, n_sets=2):
        import ctypes
        import torch.distributed as dist
        from . import _lib
        lib = _lib.load()
        if world_size > 8:
            raise ValueError("at most 8 peers (one NVSwitch node)")
        self.R, self.rank, self.world, self.device, self.group = int(R), rank, world_size, torch.device(device), group
        self._lib, self._own, self._mapped, self.sets, self._frame = lib, [], [], [], 0
        nbytes = self.R ** 3 * 4
        with _lib.device_guard(self.device):
            for _ in range(n_sets):
                ptr = ctypes.c_void_p()
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.mp_ipc_alloc(nbytes, ctypes.byref(ptr), handle), "mp_ipc_alloc")
                self._own.append(ptr)
                handles = [None] * world_size
                if world_size > 1:
                    dist.all_gather_object(handles, handle.raw, group=group)
                ptrs = []
                for r in range(world_size):
                    if r == rank:
                        ptrs.append(ptr.value)
                        continue
                    p = ctypes.c_void_p()
                    _lib.check(lib.mp_ipc_open(handles[r], ctypes.byref(p)), "mp_ipc_open")
                    self._mapped.append(p)
                    ptrs.append(p.value)
                local = torch.as_tensor(_DeviceArray(ptr.value, (self.R, self.R, self.R)), device=self.device)
                self.sets.append(((ctypes.c_void_p * world_size)(*ptrs), local))
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)

    def next_set(self):
        s = self.sets[self._frame % len(self.sets)]
        self._frame += 1
        return s

    def barrier(self):
        """Stream-ordered barrier over the ranks: a one-element all-reduce enqueued behind this rank's kernel.  When it
        completes here, every rank's slab kernel has completed, i.e. this rank's volume is fully assembled."""
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self._token, group=self.group)

    def close(self):
        for p in self._mapped:
            self._lib.mp_ipc_close(p)
        self._mapped = []
        self.sets = []
        for p in self._own:
            self._lib.mp_ipc_free(p)
        self._own = []


def query_grid_fused(net, feat, calibs, R, b_min, b_max, peers):
    """Dense R^3 occupancy volume, z-slab sharded over the ranks of `peers` (a PeerVolumes): ONE kernel per rank computes
    its slab and stores it into every rank's volume; returns this rank's full [R,R,R] volume (valid on the current
    stream after the trailing barrier).  Consumers must be ordered on the same stream (the two alternating volume sets
    then make one barrier per frame sufficient)."""
    import ctypes
    from . import _lib
    from .modeling.geometry import perspective
    assert R == peers.R
    z0, nz = slab_bounds(R, peers.world)[peers.rank]
    ptrs, local = peers.next_set()
    with _lib.device_guard(feat.device):
        fh = net.feature_handle(feat)
        proj = _lib.PROJ_PERSPECTIVE if net.projection is perspective else _lib.PROJ_ORTHOGONAL
        _lib.check(_lib.load().mp_query_grid_peers(
            net.surface_classifier.handle(), fh.ptr, int(R), int(z0), int(nz), _lib.f3(b_min), _lib.f3(b_max),
            _lib.calib12(calibs), proj, ctypes.c_float(net.normalizer.scale), ptrs, peers.world, net._mode(),
            _lib.stream_ptr(feat.device)), "mp_query_grid_peers")
    peers.barrier()
    return local
