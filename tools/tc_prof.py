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
import time
for _ in range(2):
    vol = net.query_grid(feat, cal, R, (-1, -1, -1), (1, 1, 1))
    torch.cuda.synchronize()
if not os.environ.get("MONOPORT_B200_TC_PROF") and not os.environ.get("MONOPORT_B200_TC_TRACE"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fh = net.feature_handle(feat)
    e0.record()
    for _ in range(10):
        net.query_grid(feat, cal, R, (-1, -1, -1), (1, 1, 1), out=vol, fh=fh)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("dense %d^3: %.3f ms per volume, %.1f Mpts/s" % (R, ms, R ** 3 / ms / 1e3))
