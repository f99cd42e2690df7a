import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import spec
from helpers import build_net
Ws, bs = spec.make_weights(spec.G_CHANNELS, 1234)
feat = spec.make_feat(256, 128, 128, 77).cuda()
cal = spec.scene_calib(20, -50).cuda()
net = build_net("G", Ws, bs)
net.precision = "tc"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 161
for _ in range(2):
    vol = net.query_grid(feat, cal, R, (-1, -1, -1), (1, 1, 1))
    torch.cuda.synchronize()
