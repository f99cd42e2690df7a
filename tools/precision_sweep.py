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
# colour head (Tanh, 512-channel map): layer 3 multiplies by W3 as an fp16 pair -- bar 1e-4 on the Tanh output
worst_c = 0.0
for seed in range(6):
    Ws, bs = spec.make_weights(spec.C_CHANNELS, 400 + seed)
    feat = spec.make_feat(512, 128, 128, 500 + seed)
    cal = spec.scene_calib(20, 30.0 * seed)
    pts = spec.make_points(20000, 600 + seed)
    want = spec.query_ref(feat, pts, cal, Ws, bs, spec.LAST_TANH)
    net = build_net("C", Ws, bs)
    net.precision = "tc"
    got = net.query([[feat.cuda()]], pts.cuda(), calibs=cal.cuda())[0][0].cpu()
    e = (got - want).abs()
    worst_c = max(worst_c, e.max().item())
    print("colour seed %2d: max %.2e mean %.2e" % (seed, e.max().item(), e.mean().item()), flush=True)
    net.surface_classifier.release()
print("worst colour:", worst_c)
