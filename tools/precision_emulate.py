"""CPU emulation of the roundings of the v3 tensor-core program (no GPU needed): which fp16 rounding contributes how much
to |tc - exact| on the sigmoid output.  Variants of the layer-0 sampling (gen) arithmetic can be compared:
  fp32acc : G0 taps (fp16) converted to fp32, lerp + bias in fp32, one rounding to fp16   (current kernel)
  fp16acc : lerp accumulated with fp16 FMAs (weights rounded to fp16, one rounding per FMA)
Usage: python tools/precision_emulate.py [n_points] [n_seeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import spec
torch.set_num_threads(32)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4
h = lambda t: t.half().float()
lrelu = lambda t: torch.maximum(t, t * spec.LEAKY_SLOPE)


def hfma(a, b, c):
    """fp16 fused multiply-add: exact product+sum (float64 holds it), one rounding to fp16."""
    return (a.double() * b.double() + c.double()).half().float()


def run(seed, variant):
    Ws, bs = spec.make_weights(spec.G_CHANNELS, 100 + seed)
    feat = spec.make_feat(256, 128, 128, 200 + seed)[0]              # [256,128,128]
    F = feat.permute(1, 2, 0).reshape(-1, 256)                       # texel-major
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, generator=g) * 126.9
    v = torch.rand(N, generator=g) * 126.9
    z = (torch.rand(N, generator=g) * 2 - 1) * spec.Z_SCALE
    x0, y0 = u.floor().long(), v.floor().long()
    wx, wy = u - x0, v - y0
    offs = [y0 * 128 + x0, y0 * 128 + x0 + 1, (y0 + 1) * 128 + x0, (y0 + 1) * 128 + x0 + 1]
    wg = [(1 - wx) * (1 - wy), wx * (1 - wy), (1 - wx) * wy, wx * wy]
    # exact reference (float64)
    xd = sum(w[:, None].double() * F[o].double() for w, o in zip(wg, offs))
    xin = torch.cat([xd, z[:, None].double()], 1)
    y = xin
    for l in range(5):
        inp = y if l == 0 else torch.cat([y, xin], 1)
        y = inp @ Ws[l].double().t() + bs[l].double()
        if l < 4:
            y = torch.maximum(y, y * spec.LEAKY_SLOPE)
    want = torch.sigmoid(y[:, 0])
    # emulated
    W0f, w0z = Ws[0][:, :256], Ws[0][:, 256]
    G0 = h(h(F) @ h(W0f).t())                                        # per-texel layer-0 product, fp16
    if variant.startswith("fp32acc"):
        pre = sum(w[:, None] * G0[o] for w, o in zip(wg, offs)) + (bs[0] + w0z * z[:, None])
        H0 = lrelu(h(pre))
    else:
        acc = hfma(h(w0z)[None, :].expand(N, -1), h(z)[:, None].expand(-1, 1024), h(bs[0])[None, :].expand(N, -1))
        for w, o in zip(wg, offs):
            acc = hfma(h(w)[:, None].expand(-1, 1024), G0[o], acc)
        H0 = lrelu(acc)
    H0 = h(H0)
    x32 = sum(w[:, None] * F[o] for w, o in zip(wg, offs))           # fp32 taps
    X = h(x32)
    if "x16" in variant:                                             # X operand from an fp16 copy of the map
        X = h(sum(w[:, None] * h(F)[o] for w, o in zip(wg, offs)))
    hid, act = [0, 1024, 512, 256], H0
    for l in (1, 2, 3):
        Wh, Wx, wz = Ws[l][:, :hid[l]], Ws[l][:, hid[l]:hid[l] + 256], Ws[l][:, hid[l] + 256]
        pre = act @ h(Wh).t() + X @ h(Wx).t() + (bs[l] + wz * z[:, None])
        act = lrelu(h(pre)) if l < 3 else lrelu(pre)
        if l < 3:
            act = h(act)
    W4 = Ws[4]
    if "s4tex" in variant:                                           # last-layer skip hoisted to texels (fp32), then lerp
        S4 = F @ W4[:, 128:384].t()
        s4 = sum(w[:, None] * S4[o] for w, o in zip(wg, offs))
    else:
        s4 = x32 @ W4[:, 128:384].t()
    logit = act @ W4[:, :128].t() + s4 + W4[:, 384] * z[:, None] + bs[4]
    got = torch.sigmoid(logit[:, 0])
    return (got.double() - want).abs()


for variant in ("fp32acc", "fp16acc", "fp16acc+s4tex", "fp16acc+s4tex+x16"):
    worst, mean = 0.0, 0.0
    for s in range(S):
        e = run(s, variant)
        worst, mean = max(worst, e.max().item()), mean + e.mean().item() / S
    print("%-18s  max |tc - exact| = %.3e   mean = %.3e   (%d points x %d seeds)" % (variant, worst, mean, N, S))
