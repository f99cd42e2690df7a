"""CPU, world_size 2, gloo: the N>1 host path of the z-slab sharding (partition, padded single all-gather, reassembly).
The per-rank kernel call is replaced by an analytic slab so that no GPU is needed; the collective logic is the real one."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _field(R):
    z, y, x = torch.meshgrid(torch.arange(R), torch.arange(R), torch.arange(R), indexing="ij")
    return (x * 1.0 + y * 1000.0 + z * 1000000.0).float()


def _worker(rank, world, port, R, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from monoport_b200.shard import slab_bounds, gather_slabs
        z0, nz = slab_bounds(R, world)[rank]
        slab = _field(R)[z0:z0 + nz].contiguous()          # what net.query_grid(z0=z0, nz=nz) would return
        full = gather_slabs(slab, R, rank, world)
        ok = torch.equal(full, _field(R))
        # balanced ranges of the linear node order + ONE in-place all-gather inside the persistent buffer (ShardedVolume)
        from monoport_b200.shard import ShardedVolume, range_bounds
        bounds, per = range_bounds(R, world)
        assert per % 128 == 0 and sum(n for _, n in bounds) == R ** 3 and all(b[0] == r * per or b[1] == 0 for r, b in enumerate(bounds))
        sv = ShardedVolume(R, rank, world, "cpu")
        for frame in range(2):                              # the buffers are reused from frame to frame
            lin0, n = bounds[rank]
            sv.segment[:n] = (_field(R) + frame).reshape(-1)[lin0:lin0 + n]     # what net.query_grid_range(...) would write
            vol = sv.gather()
            ok = ok and torch.equal(vol, _field(R) + frame) and vol.data_ptr() == sv.flat.data_ptr()
        q.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("R", [17, 33])
def test_slab_allgather_world2(R):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == (R, R, R) for r in res)
