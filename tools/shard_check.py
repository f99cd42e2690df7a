"""torchrun --nproc-per-node N tools/shard_check.py : sharded dense volume == single-GPU volume, bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.distributed as dist
from oracle import spec
from helpers import build_net
from monoport_b200.shard import query_grid_sharded, query_grid_fused, PeerVolumes
from monoport_b200.recon import marching_cubes
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
Ws, bs = spec.make_weights(spec.G_CHANNELS, 3)
feat = spec.make_feat(256, 128, 128, 4, 0.5)
Ws, bs, feat, _ = spec.heightfield_person(Ws, bs, feat)
net = build_net("G", Ws, bs, device="cuda:%d" % local)
cal = spec.scene_calib(20, 33).cuda()
R = 129
full = query_grid_sharded(net, feat.cuda(), cal, R, (-1, -1, -1), (1, 1, 1), rank, world)
single = net.query_grid(feat.cuda(), cal, R, (-1, -1, -1), (1, 1, 1))
v, f = marching_cubes(full)
ok = torch.equal(full, single)
print("rank %d/%d: sharded == single: %s ; mesh %d verts %d faces" % (rank, world, ok, v.shape[0], f.shape[0]), flush=True)
t = torch.tensor([v.shape[0], f.shape[0], int(ok)], device="cuda")
lst = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(lst, t)
if rank == 0:
    assert all(bool(x[2]) for x in lst) and all(torch.equal(x, lst[0]) for x in lst), lst
    print("shard_check OK: identical volumes and mesh topology on all %d ranks" % world)
# fused slab exchange (peer-memory stores from the kernel epilogue instead of the all-gather): same volume, bit for bit,
# over several frames (alternating volume sets, changing features)
if "--fused" in sys.argv:
    peers = PeerVolumes(R, rank, world, "cuda:%d" % local)
    okf = True
    for it in range(6):
        f_it = (feat * (1.0 + 0.05 * it)).cuda()
        vol = query_grid_fused(net, f_it, cal, R, (-1, -1, -1), (1, 1, 1), peers)
        ref = net.query_grid(f_it, cal, R, (-1, -1, -1), (1, 1, 1))
        okf &= bool(torch.equal(vol, ref))
    torch.cuda.synchronize()
    t = torch.tensor([int(okf)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print("rank %d/%d: fused == single over 6 frames: %s" % (rank, world, okf), flush=True)
    if rank == 0:
        assert int(t.item()) == 1
        print("shard_check OK (fused slab exchange)")
    dist.barrier()
    peers.close()
# list-sharded coarse-to-fine engines (mp_octree_shard_*): the volume every rank receives == the single-GPU engine's volume,
# bit for bit, in the `faster` mode of RTL/main.py:195, in lossless mode (conflict loop) and for top-k
if "--octree" in sys.argv:
    import numpy as np, time
    from monoport_b200.engine import Seg3dLossless, Seg3dTopk, make_query_func
    b = np.array([[-1.0, -1.0, -1.0]], dtype=np.float32)
    res = [17, 33, 65, 129]
    dev = "cuda:%d" % local
    oko = True
    for name, make in (("faster", lambda: Seg3dLossless(make_query_func(net), b, -b, res, balance_value=0.5, faster=True)),
                       ("lossless", lambda: Seg3dLossless(make_query_func(net), b, -b, res, balance_value=0.5, faster=False)),
                       ("topk", lambda: Seg3dTopk(make_query_func(net), b, -b, res, num_points=[None, 3000, 9000, 30000]))):
        single = make().to(dev)
        ref = single(im_feat_list=[[feat.cuda()]], calib_tensor=cal)
        sharded = make().to(dev).shard(rank, world)
        for it in range(3):                       # several frames through the same mappings (alternating value lists)
            f_it = (feat * (1.0 + 0.05 * it)).cuda()
            got = sharded(im_feat_list=[[f_it]], calib_tensor=cal)
            want = single(im_feat_list=[[f_it]], calib_tensor=cal)
            same = (got is None and want is None) or bool(torch.equal(got, want))
            oko &= same and list(sharded.last_stats) == list(single.last_stats)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for it in range(10):
            sharded(im_feat_list=[[feat.cuda()]], calib_tensor=cal)
        torch.cuda.synchronize()
        ts = (time.perf_counter() - t0) / 10
        t0 = time.perf_counter()
        for it in range(10):
            single(im_feat_list=[[feat.cuda()]], calib_tensor=cal)
        torch.cuda.synchronize()
        t1 = (time.perf_counter() - t0) / 10
        print("rank %d/%d: octree %s sharded == single: %s ; %.0f us sharded vs %.0f us single ; evaluated %s"
              % (rank, world, name, oko, ts * 1e6, t1 * 1e6, sharded.last_stats), flush=True)
        dist.barrier()
        sharded.unshard()
        assert ref is not None
    t = torch.tensor([int(oko)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        assert int(t.item()) == 1
        print("shard_check OK (list-sharded octree engines)")
dist.destroy_process_group()
