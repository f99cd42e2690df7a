"""configs[4] sanity on one GPU: 513^3 ("512^3") dense query, 6-level octree, surface + mesh at 513^3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import spec
from helpers import build_net
from monoport_b200.engine import Seg3dLossless, make_query_func
from monoport_b200.recon import forward_vertices, marching_cubes
Ws, bs = spec.make_weights(spec.G_CHANNELS, 3)
feat = spec.make_feat(256, 128, 128, 4, 0.5)
Ws, bs, feat, _ = spec.heightfield_person(Ws, bs, feat)
net = build_net("G", Ws, bs)
cal = spec.scene_calib(20, 33).cuda()
f = feat.cuda()
R = 513
vol = net.query_grid(f, cal, R, (-1, -1, -1), (1, 1, 1)); torch.cuda.synchronize()
t0 = time.perf_counter(); vol = net.query_grid(f, cal, R, (-1, -1, -1), (1, 1, 1)); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("dense 513^3: %.1f ms, %.1f Mpts/s, occupied %.4f" % (dt * 1e3, R**3 / dt / 1e6, float((vol > 0.5).float().mean())), flush=True)
b = np.array([[-1.0, -1.0, -1.0]], dtype=np.float32)
eng = Seg3dLossless(make_query_func(net), b, -b, [17, 33, 65, 129, 257, 513], balance_value=0.5, faster=False).to("cuda")
sdf = eng(im_feat_list=[[f]], calib_tensor=cal); torch.cuda.synchronize()
t0 = time.perf_counter(); sdf = eng(im_feat_list=[[f]], calib_tensor=cal); torch.cuda.synchronize(); dt = time.perf_counter() - t0
mism = int(((sdf[0, 0] > 0.5) != (vol > 0.5)).sum())
print("lossless octree 513^3: %.2f ms, evaluated %s (%.2f%% of nodes), mask mismatches vs dense: %d" % (dt * 1e3, eng.last_stats, 100 * sum(eng.last_stats) / R**3, mism), flush=True)
X, Y, Z, n = forward_vertices(sdf, "front")
v, fc = marching_cubes(sdf[0, 0])
print("surface: %d visible verts; mesh %d verts %d faces" % (X.numel(), v.shape[0], fc.shape[0]))
