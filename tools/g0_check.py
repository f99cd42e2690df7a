"""GPU check of the tcgen05 G0 GEMM (layer 0 hoisted to texels): program v3 with G0 from the tensor cores vs. the oracle,
vs. the fp32 CUDA-core G0 (MONOPORT_B200_G0=fp32, run as a child process because the switch is read once), and the cost
of one G0 refresh (query after a feature upload vs. query on an unchanged feature map)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import spec
from helpers import build_net

Ws, bs = spec.make_weights(spec.G_CHANNELS, 1234)
feat = spec.make_feat(256, 128, 128, 77)
cal = spec.scene_calib(20, -50)
net = build_net("G", Ws, bs)
net.precision = "tc_v3"
out = {}
for n in (100, 1000, 20000, 148 * 128 * 3 + 17):
    pts = spec.make_points(n, 5 + n)
    want = spec.query_ref(feat, pts, cal, Ws, bs, spec.LAST_SIGMOID)[0]
    got = net.query([[feat.cuda()]], pts.cuda(), calibs=cal.cuda())[0][0, 0]
    torch.cuda.synchronize()
    out[n] = got.cpu()
    err = (got.cpu() - want).abs()
    print("[%s] n=%7d  max|v3 - oracle| = %.3e  mean = %.3e" % (os.environ.get("MONOPORT_B200_G0", "tc"), n, err.max().item(), err.mean().item()), flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    torch.save(out, sys.argv[2])
    sys.exit(0)
env = dict(os.environ, MONOPORT_B200_G0="fp32")
tmp = "/tmp/g0_child.pt"
subprocess.run([sys.executable, os.path.abspath(__file__), "--child", tmp], env=env, check=True)
ref = torch.load(tmp)
for n in out:
    print("n=%7d  max|G0 tc - G0 fp32| on outputs = %.3e" % (n, (out[n] - ref[n]).abs().max().item()))
# cost of a G0 refresh: small query with / without a feature re-upload
f = feat.cuda(); c = cal.cuda()
pts = spec.make_points(4096, 3).cuda()
def run(reupload, iters=50):
    fs = [f.clone() for _ in range(2)]
    for i in range(5):
        net.query([[fs[i % 2] if reupload else f]], pts, calibs=c)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        net.query([[fs[i % 2] if reupload else f]], pts, calibs=c)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
a = run(False); b = run(True)
print("4096-point query: %.1f us cached G0, %.1f us with feature upload + G0 refresh (transpose + GEMM)" % (a, b))
