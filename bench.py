#!/usr/bin/env python
"""bench.py -- occupancy Mpoints/s (+ recon frames/s @256^3) of the fused hot path on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        (the reference's CPU path: the oracle port, host cores)

A "step" = one pass of the hot path over one synthetic frame: upload-free channel-last repack of the [1,256,128,128]
feature map + the fused sample+MLP kernel over the dense 257^3 node grid ("256^3", RTL/main.py:187), z-slab
sharded over the ranks, + ONE all-gather of the occupancy volume (N>1).  `value` = whole-job Mpoints/s with inputs
resident in HBM; `e2e` = the same through the C-ABI host-buffer entry point (feature map H2D + volume D2H inside the
timed region).  `recon` reports BASELINE configs[1]: coarse-to-fine (Seg3dLossless, faster=True) recon frames/s.
Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

R_GRID = 257
FLOP_PER_POINT = 2363906          # 2*(257*1024+1281*512+769*256+513*128+385*1)  
B_MIN, B_MAX = (-1.0, -1.0, -1.0), (1.0, 1.0, 1.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("MONOPORT_B200_MODE", "auto"), choices=["auto", "tc", "fp32"])
    ap.add_argument("--res", type=int, default=R_GRID)
    ap.add_argument("--fused-gather", action="store_true", default=os.environ.get("MONOPORT_B200_FUSED_GATHER", "0") == "1",
                    help="N>1: store the slab into every rank's volume from the kernel epilogue (peer memory) instead of "
                         "the NCCL all-gather (experimental, opt-in)")
    ap.add_argument("--no-recon", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(source="measured", tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"])
    return dict(source="fallback", tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def synthetic(seed_w=3, seed_f=4, n_feat=1):
    """Seeded random-init netG head + feature maps of the encoder's shape; the 'person' wiring keeps the cost per
    point unchanged (oracle/spec.py:heightfield_person is test infra -- the same construction is inlined here so
    that the product bench does not import the oracle)."""
    import math
    import torch
    chans = [257, 1024, 512, 256, 128, 1]
    g = torch.Generator().manual_seed(seed_w)
    Ws, bs = [], []
    for l in range(5):
        cin = chans[l] + (chans[0] if l else 0)
        k = 1.0 / math.sqrt(cin)
        Ws.append((torch.rand(chans[l + 1], cin, generator=g) * 2 - 1) * k)
        bs.append((torch.rand(chans[l + 1], generator=g) * 2 - 1) * k)
    feats = []
    H = W = 128
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    for i in range(n_feat):
        gf = torch.Generator().manual_seed(seed_f + i)
        f = torch.randn(1, 256, H, W, generator=gf) * 0.5
        body = 1 - ((xx / 0.28) ** 2 + ((yy - 0.05) / 0.62) ** 2)
        head = 1 - ((xx / 0.16) ** 2 + ((yy + 0.68) / 0.17) ** 2)
        arm = 1 - (((xx.abs() - 0.40) / 0.10) ** 2 + ((yy + 0.05 - 0.02 * i) / 0.40) ** 2)
        f[0, 0] = torch.maximum(torch.maximum(body, head), arm).clamp(min=-1.0) * 0.35
        feats.append(f.contiguous())
    C = 256
    Ws[3][0].zero_(); Ws[3][1].zero_()
    Ws[3][0, 256 + C] = 1.0; Ws[3][1, 256 + C] = -1.0
    bs[3][0] = 0.0; bs[3][1] = 0.0
    Ws[4].zero_(); bs[4].zero_()
    Ws[4][0, 128 + 0] = 40.0
    Ws[4][0, 0] = -40.0 / 0.99; Ws[4][0, 1] = -40.0 / 0.99
    return chans, Ws, bs, feats


def scene_calib():
    """pifu_calib of the demo camera at yaw 20 deg, pitch 33 deg (RTL/scene.py:108-135, RTL/recon.py:4-25)."""
    import math
    import numpy as np
    from monoport_b200.recon import pifu_calib

    def rot(rx, ry):
        cx, sx, cy, sy = math.cos(rx), math.sin(rx), math.cos(ry), math.sin(ry)
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        return Ry @ Rx
    E = np.eye(4)
    E[:3, :3] = rot(math.radians(20), 0) @ rot(0, math.radians(33))
    E[:3, 3] = [0, 0, -2.0]
    K = np.diag([1.0, 1.0, -0.2, 1.0]); K[2, 3] = -1.0
    return pifu_calib(E, K, device="cpu")


# ------------------------------------------------------------------------------------------------------------
def best_threads(fn):
    """The torch-CPU port does not scale to every hardware thread of a big host: time a small sample at a few thread
    counts and keep the fastest (reported as `cores`)."""
    import torch
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores // 2, cores) if 1 <= c <= cores})
    best, best_t = cores, None
    for c in cands:
        torch.set_num_threads(c)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def run_reference(args):
    """The reference's own CPU implementation of the path = the oracle port (the Python reference cannot travel
    to the GPU box), all host threads, on a bounded sample of the same workload per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import spec
    chans, Ws, bs, feats = synthetic()
    cal = scene_calib()
    S = 48                                                  # 48^3 = 110 592 node centres of the same [-1,1]^3 grid
    coords = spec._grid_coords(S, 1)
    pts = spec.level_points(coords, S, B_MIN, B_MAX).t().contiguous()
    cores = best_threads(lambda: spec.query_ref(feats[0], pts[:, :16384], cal, Ws, bs, spec.LAST_SIGMOID))
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        spec.query_ref(feats[0], pts, cal, Ws, bs, spec.LAST_SIGMOID)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    mpts = S ** 3 * len(times) / total / 1e6
    line = {
        "impl": "reference", "metric": "occupancy_mpoints_per_s", "value": mpts, "unit": "Mpoints/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "netG query() over dense grid node centres, [1,256,128,128] features, scene calib "
                               "yaw20/pitch33; bounded sample 48^3 = 110592 points per step of the 257^3 job"},
        "cpu_baseline": {"value": mpts, "unit": "Mpoints/s", "cores": cores, "kind": "port",
                         "sample": "48^3 = 110592 points per step, torch CPU fp32, %d threads" % cores},
        "e2e": {"value": mpts, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from monoport_b200 import _lib
    from monoport_b200.modeling import PIFuNetG
    from monoport_b200.shard import slab_bounds, gather_slabs, PeerVolumes, query_grid_fused

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 through torch.distributed.run (see docstring)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)

    R = args.res
    chans, Ws, bs, feats_cpu = synthetic(n_feat=4)
    net = PIFuNetG()
    net.surface_classifier.load_state_dict(
        {**{"filters.%d.weight" % l: W[:, :, None] for l, W in enumerate(Ws)},
         **{"filters.%d.bias" % l: b for l, b in enumerate(bs)}})
    net.surface_classifier.to(dev)
    net.eval()
    net.precision = args.mode
    mode_used = "tc" if (args.mode in ("auto", "tc") and net.surface_classifier.tc_supported()) else "fp32"
    if args.mode == "tc" and mode_used != "tc":
        raise SystemExit("tcgen05 kernel unavailable")
    cal_cpu = scene_calib()
    cal = cal_cpu.to(dev)
    feats = [f.to(dev) for f in feats_cpu]
    z0, nz = slab_bounds(R, world)[rank]
    n_pts_total = R ** 3
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2
    slab = torch.empty((nz, R, R), dtype=torch.float32, device=dev)
    full = torch.empty((R, R, R), dtype=torch.float32, device=dev)

    fused = bool(args.fused_gather and world > 1)
    peers = PeerVolumes(R, rank, world, dev) if fused else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        f = feats[i % len(feats)]
        if fused:
            query_grid_fused(net, f, cal_cpu, R, B_MIN, B_MAX, peers)
            return
        net.query_grid(f, cal_cpu, R, B_MIN, B_MAX, z0=z0, nz=nz, out=slab)
        if world > 1:
            gather_slabs(slab, R, rank, world, out=full)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush.fill_(float(i))                      # L2 flush between timed iterations (not timed)
        f = feats[i % len(feats)]
        ev[i][0].record()
        fh = net.feature_handle(f)                 # channel-last repack kernel (part of the step)
        kev[i][0].record()
        if fused:
            query_grid_fused(net, f, cal_cpu, R, B_MIN, B_MAX, peers)      # kernel with peer stores + barrier
            kev[i][1].record()
        else:
            net.query_grid(f, cal_cpu, R, B_MIN, B_MAX, z0=z0, nz=nz, out=slab, fh=fh)
            kev[i][1].record()
            if world > 1:
                gather_slabs(slab, R, rank, world, out=full)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop() if sampler else None
    t_ms = sum(a.elapsed_time(b) for a, b in ev)
    k_ms = sum(a.elapsed_time(b) for a, b in kev)
    tt = torch.tensor([t_ms, k_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, k_ms = tt.tolist()
    value = n_pts_total * args.steps / (t_ms * 1e-3) / 1e6

    # ---- e2e through the C-ABI with HOST buffers (pinned), rank-local slab ---------------------------------
    lib = _lib.load()
    feat_pinned = [f.pin_memory() for f in feats_cpu]
    out_host = torch.empty((nz, R, R), dtype=torch.float32).pin_memory()
    fh = net.feature_handle(feats[0])
    cal12 = _lib.calib12(cal_cpu)
    st = _lib.stream_ptr(dev)
    mode_code = _lib.MODE_TC if mode_used == "tc" else _lib.MODE_FP32

    def e2e_step(i):
        f = feat_pinned[i % len(feat_pinned)]
        _lib.check(lib.mp_query_grid_host(net.surface_classifier.handle(), fh.ptr, ctypes.c_void_p(f.data_ptr()), R, z0, nz,
                                          _lib.f3(B_MIN), _lib.f3(B_MAX), cal12, 0, ctypes.c_float(net.normalizer.scale),
                                          ctypes.c_void_p(out_host.data_ptr()), mode_code, st))
    e2e_steps = max(3, min(args.steps, 10))
    e2e_step(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_pts_total * e2e_steps / te.item() / 1e6
    fh.key = None

    # ---- configs[1]: coarse-to-fine recon frames/s (rank 0, N=1 only) ---------------------------------------
    recon = None
    if rank == 0 and world == 1 and not args.no_recon and R == R_GRID:
        from monoport_b200.engine import Seg3dLossless, make_query_func
        from monoport_b200.recon import forward_vertices, marching_cubes
        b = np.array([B_MIN], dtype=np.float32)
        eng = Seg3dLossless(make_query_func(net), b, -b, [17, 33, 65, 129, 257], balance_value=0.5, faster=True).to(dev)
        for i in range(3):                                  # warm-up: engine, surface kernel, marching cubes
            sdf = eng(im_feat_list=[[feats[i % 4]]], calib_tensor=cal)
            X, Y, Z, nrm = forward_vertices(sdf, "front")
            v, fcs = marching_cubes(sdf[0, 0])
        torch.cuda.synchronize()
        nfr = 20
        t0 = time.perf_counter()
        for i in range(nfr):
            sdf = eng(im_feat_list=[[feats[i % 4]]], calib_tensor=cal)
            X, Y, Z, nrm = forward_vertices(sdf, "front")
        torch.cuda.synchronize()
        fps_fv = nfr / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for i in range(nfr):
            sdf = eng(im_feat_list=[[feats[i % 4]]], calib_tensor=cal)
            v, fcs = marching_cubes(sdf[0, 0])
        torch.cuda.synchronize()
        fps_mc = nfr / (time.perf_counter() - t0)
        recon_stats = list(eng.last_stats)
        # configs[2]: geometry + colour -- netC (513-wide head, fp32 fused kernel) queried at the visible vertices
        from monoport_b200.modeling import PIFuNetC
        from monoport_b200.recon import colorization
        netC = PIFuNetC()
        netC.surface_classifier.to(dev)
        netC.eval()
        gC = torch.Generator().manual_seed(11)
        featC = [[(torch.randn(1, 512, 128, 128, generator=gC) * 0.5).to(dev)]]
        img = colorization(netC, featC, X, Y, Z, cal)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nfr):
            sdf = eng(im_feat_list=[[feats[i % 4]]], calib_tensor=cal)
            X, Y, Z, nrm = forward_vertices(sdf, "front")
            img = colorization(netC, featC, X, Y, Z, cal)
        torch.cuda.synchronize()
        fps_color = nfr / (time.perf_counter() - t0)
        # configs[3]-style image stream on one GPU: synthetic 512x512 frames -> HG encoder (PyTorch, fp32) -> coarse-to-fine
        # recon -> visible surface, frames overlapped by FramePipeline (1 lane = sequential, 2 lanes = overlapped)
        from monoport_b200.pipeline import FramePipeline
        net.image_filter.to(dev)
        gI = torch.Generator().manual_seed(5)
        frames = [(torch.rand(1, 3, 512, 512, generator=gI) * 2 - 1).to(dev) for _ in range(4)]

        def stage_encode(img):
            return net.filter(img)

        def stage_recon(fs):
            return eng(im_feat_list=fs, calib_tensor=cal)

        def stage_surface(sdf_):
            return forward_vertices(sdf_, "front")

        stream_fps = {}
        with torch.no_grad():
            for lanes in (1, 2):
                pipe = FramePipeline([stage_encode, stage_recon, stage_surface], dev, n_lanes=lanes)
                list(pipe.run(frames[i % 4] for i in range(4)))               # warm-up (cudnn autotune, handles)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                nst = 24
                outs = list(pipe.run(frames[i % 4] for i in range(nst)))
                torch.cuda.synchronize()
                stream_fps[lanes] = nst / (time.perf_counter() - t0)
                pipe.close()
        recon = {"workload": "configs[1]: netG 256^3 Seg3dLossless(faster=True) from resident features, per frame",
                 "frames_per_s_stream_with_pytorch_encoder": {
                     "1_lane": stream_fps[1], "2_lanes": stream_fps[2],
                     "note": "512x512 frame -> HGFilter (PyTorch fp32, random init => noise field, %d points/frame) -> recon -> "
                             "forward_vertices; lanes = overlapped frames (FramePipeline)" % int(sum(eng.last_stats))},
                 "frames_per_s_geometry_plus_netC_colour": fps_color,
                 "frames_per_s_with_forward_vertices": fps_fv, "frames_per_s_with_marching_cubes": fps_mc,
                 "points_evaluated_per_frame": int(sum(recon_stats)), "per_level": recon_stats,
                 "visible_vertices": int(X.numel()), "mesh_vertices": int(v.shape[0]), "mesh_faces": int(fcs.shape[0])}

    # ---- CPU baseline (reported, not the target): the oracle port on a bounded sample, rank 0, N=1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import spec
        S = 64                                                     # BASELINE configs[0]: dense 64^3 on CPU
        pts = spec.level_points(spec._grid_coords(S, 1), S, B_MIN, B_MAX).t().contiguous()
        cores = best_threads(lambda: spec.query_ref(feats_cpu[0], pts[:, :16384], cal_cpu, Ws, bs, spec.LAST_SIGMOID))
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 10.0 or reps < 1:
            spec.query_ref(feats_cpu[0], pts, cal_cpu, Ws, bs, spec.LAST_SIGMOID)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": S ** 3 * reps / dt / 1e6, "unit": "Mpoints/s", "cores": cores, "kind": "port",
               "sample": "dense 64^3 = 262144 points x %d reps (%.1f s), torch CPU fp32 oracle port of MonoPortNet.query" % (reps, dt)}

    my_pts = nz * R * R
    if rank == 0:
        pk = peaks()
        # nchw_to_nhwc repack + (tensor-core program v3, every query size: per-texel layer-0 GEMM g0_tc_kernel) + fused query kernel
        launches_per_step = 3 if mode_used == "tc" else 2
        k_avg_s = k_ms * 1e-3 / args.steps
        achieved_tf = FLOP_PER_POINT * my_pts / k_avg_s / 1e12
        peak_tf = pk["tf_sustained"]
        line = {
            "metric": "occupancy_mpoints_per_s", "value": value, "unit": "Mpoints/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16" if mode_used == "tc" else "f32", "data": "synthetic",
            "config": {"workload": "netG dense %d^3 grid query (%d points/step), [1,256,128,128] features, scene calib "
                                   "yaw20/pitch33, z-slab sharded over %d GPU(s) + %s" % (R, n_pts_total, world, "peer-memory stores from the kernel epilogue + 1 barrier" if fused else "1 all-gather"),
                       "kernel_mode": mode_used, "l2_flush_between_steps": True, "grid": R,
                       "accumulate": "fp32", "last_layer": "fp32"},
            "e2e": {"value": e2e_value, "unit": "Mpoints/s", "h2d_bytes_per_step": 256 * 128 * 128 * 4 + 48,
                    "d2h_bytes_per_step": nz * R * R * 4, "steps": e2e_steps,
                    "api": "mp_query_grid_host (pinned host feature map in, host occupancy slab out)"},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from profiles/r01_final_ncu_tc_summary.txt
                         # (ncu --set full, 257^3, N=1, final code of round 1): 45.55 MB + 38.89 MB per launch
                         "traffic": (84.44e6 if (launches_per_step == 3 and world == 1 and R == R_GRID) else None),
                         "traffic_unit": "bytes/launch (ncu capture r01, not re-measured in this run)",
                         "kernel": ("query_tc3_kernel (+ g0_tc_kernel, the per-frame per-texel layer-0 GEMM, 26 us)" if launches_per_step == 3
                                    else "query_%s_kernel" % mode_used),
                         "kernel_ms": 1e3 * k_avg_s, "peak_source": pk["source"] + " bf16 sustained (cuBLAS loop)",
                         "frac_of_burst": achieved_tf / pk["tf_burst"],
                         "algorithmic_flop_per_point": FLOP_PER_POINT},
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if recon:
            line["recon"] = recon
        print(json.dumps(line))
    if peers is not None:
        barrier()
        peers.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
