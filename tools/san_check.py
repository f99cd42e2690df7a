"""Tiny invocation of every kernel family for compute-sanitizer (memcheck): both tensor-core programs, the fp32 kernel,
the octree engine, marching cubes and the visible-surface kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import spec
from helpers import build_net
from monoport_b200.engine import Seg3dLossless, Seg3dTopk, make_query_func
from monoport_b200.recon import forward_vertices, marching_cubes
Ws, bs = spec.make_weights(spec.G_CHANNELS, 3)
feat = spec.make_feat(256, 128, 128, 4, 0.5)
Ws, bs, feat, _ = spec.heightfield_person(Ws, bs, feat)
net = build_net("G", Ws, bs)
cal = spec.scene_calib(20, 33).cuda()
pts = spec.make_points(700, 1).cuda()
for mode in ("fp32", "tc"):
    net.precision = mode
    out = net.query([[feat.cuda()]], pts, calibs=cal)[0]
    torch.cuda.synchronize()
    print(mode, float(out.sum()))
net.precision = "tc"
b = np.array([[-1.0, -1.0, -1.0]], dtype=np.float32)
for faster in (True, False):
    eng = Seg3dLossless(make_query_func(net), b, -b, [9, 17, 33], balance_value=0.5, faster=faster).to("cuda")
    sdf = eng(im_feat_list=[[feat.cuda()]], calib_tensor=cal)
    print("engine faster=%s" % faster, eng.last_stats, None if sdf is None else float(sdf.sum()))
eng = Seg3dTopk(make_query_func(net), b, -b, [9, 17, 33], num_points=[0, 500, 2000]).to("cuda")
sdf = eng(im_feat_list=[[feat.cuda()]], calib_tensor=cal)
X, Y, Z, n = forward_vertices(sdf, "front")
v, f = marching_cubes(sdf[0, 0])
torch.cuda.synchronize()
# the colour head's tensor-core program incl. the fused direct rendering, and a frame outside the range guard's limit
from monoport_b200.recon import colorization
cW, cb = spec.make_weights(spec.C_CHANNELS, 5)
netC = build_net("C", cW, cb)
featC = spec.make_feat(512, 128, 128, 6, 0.5).cuda()
img = colorization(netC, [[featC]], X, Y, Z, cal, resolution=33)
big = net.query([[feat.cuda() * 60.0]], pts, calibs=cal)[0]
torch.cuda.synchronize()
print("ok", X.numel(), v.shape, f.shape, float(img.sum()), float(big.sum()))
