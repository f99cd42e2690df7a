"""Per-kernel timeline of one configs[1] frame (Seg3dLossless 257^3 fused + forward_vertices [+ colour]) via torch.profiler.
Usage: python tools/recon_trace.py [--color] [--frames N]"""
import sys, os, argparse, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--color", action="store_true")
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--no-profiler", action="store_true", help="just run frames inside a cudaProfilerStart/Stop range "
                                                          "(ncu --profile-from-start off)")
ap.add_argument("--mc", action="store_true", help="also run marching cubes on every frame")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
from monoport_b200.modeling import PIFuNetG
chans, Ws, bs, feats = bench.synthetic(n_feat=4)
net = PIFuNetG()
net.surface_classifier.load_state_dict(
    {**{"filters.%d.weight" % l: W[:, :, None] for l, W in enumerate(Ws)},
     **{"filters.%d.bias" % l: b for l, b in enumerate(bs)}})
net.surface_classifier.to(dev)
net.eval()
feats = [f.to(dev) for f in feats]
cal = bench.scene_calib().to(dev)
from monoport_b200.engine import Seg3dLossless, make_query_func
from monoport_b200.recon import forward_vertices, colorization, marching_cubes
b = np.array([bench.B_MIN], dtype=np.float32)
eng = Seg3dLossless(make_query_func(net), b, -b, [17, 33, 65, 129, 257], balance_value=0.5, faster=True).to(dev)
netC = featC = None
if a.color:
    from monoport_b200.modeling import PIFuNetC
    netC = PIFuNetC(); netC.surface_classifier.to(dev); netC.eval()
    gC = torch.Generator().manual_seed(11)
    featC = [[(torch.randn(1, 512, 128, 128, generator=gC) * 0.5).to(dev)]]

def frame(i):
    sdf = eng(im_feat_list=[[feats[i % 4]]], calib_tensor=cal)
    X, Y, Z, nrm = forward_vertices(sdf, "front")
    if a.mc:
        marching_cubes(sdf[0, 0])
    if a.color:
        return colorization(netC, featC, X, Y, Z, cal)
    return X

for i in range(3):
    frame(i)
torch.cuda.synchronize()
if a.no_profiler:
    torch.cuda.profiler.start()
    for i in range(a.frames):
        frame(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
t0 = time.perf_counter()
for i in range(20):
    frame(i)
torch.cuda.synchronize()
print("wall per frame: %.1f us" % ((time.perf_counter() - t0) / 20 * 1e6))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(a.frames):
        frame(i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = {}
for e in evs:
    k = e.name[:70]
    t = tot.setdefault(k, [0, 0.0]); t[0] += 1; t[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
print("%-72s %6s %10s" % ("kernel (per frame)", "calls", "us"))
busy = 0.0
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %6.1f %10.1f" % (k, n / a.frames, us / a.frames)); busy += us / a.frames
print("GPU busy per frame: %.1f us" % busy)
# timeline of the last frame
last = evs[-int(len(evs) / a.frames):]
t00 = last[0].time_range.start
for e in last:
    print("  +%8.1f us  %8.1f us  %s" % (e.time_range.start - t00, e.time_range.end - e.time_range.start, e.name[:60]))
