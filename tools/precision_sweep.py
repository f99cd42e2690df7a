"""Max |tc - oracle| of query() over several seeded heads / feature maps / calibs (parity bar: 1e-4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import spec
from helpers import build_net
torch.set_num_threads(32)
worst = {}
for seed in range(12):
    Ws, bs = spec.make_weights(spec.G_CHANNELS, 100 + seed)
    feat = spec.make_feat(256, 128, 128, 200 + seed)
    cal = spec.scene_calib(20, 30.0 * seed)
    pts = spec.make_points(20000, 300 + seed)
    want = spec.query_ref(feat, pts, cal, Ws, bs, spec.LAST_SIGMOID)[0]
    net = build_net("G", Ws, bs)
    row = []
    for mode in ("tc",):
        net.precision = mode
        got = net.query([[feat.cuda()]], pts.cuda(), calibs=cal.cuda())[0][0, 0].cpu()
        e = (got - want).abs()
        row.append((mode, e.max().item(), e.mean().item()))
        worst[mode] = max(worst.get(mode, 0), e.max().item())
    print("seed %2d: " % seed + "  ".join("%s max %.2e mean %.2e" % r for r in row), flush=True)
    net.surface_classifier.release()
print("worst:", worst)
