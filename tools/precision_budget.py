"""Which fp16 rounding of the tensor-core programs costs how much on what query() returns (CPU, torch; no GPU needed).
Model of program v3 / the colour program: G0 (per-texel layer-0 product) fp16, sampled layer-0 chunk H0 fp16, skip operand
X fp16 (rounded once after fp32 interpolation), weights fp16, H1 / H2 fp16, fp32 accumulation, layer 4 + S4 in fp32.
Every rounding class can be switched off (exact operand) and X / H can be split hi+lo (two MMAs) per layer.
Usage: python tools/precision_budget.py G|C [feat_scale] [n_points] [n_seeds]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import spec
torch.set_num_threads(min(32, os.cpu_count() or 8))
KIND = sys.argv[1] if len(sys.argv) > 1 else "G"
SCALE = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
N = int(sys.argv[3]) if len(sys.argv) > 3 else 6000
S = int(sys.argv[4]) if len(sys.argv) > 4 else 3
CH = spec.G_CHANNELS if KIND == "G" else spec.C_CHANNELS
C = CH[0] - 1
D = torch.float64


def h(t):
    return t.to(torch.float16).to(D)


def hl(t):            # hi + lo split: exact to ~22 bits
    hi = h(t)
    return hi + h(t - hi)


lrelu = lambda t: torch.maximum(t, t * spec.LEAKY_SLOPE)


def run(seed, cfg):
    """cfg: dict of rounding functions per operand class: W1..W3, X1..X3, H0, H1, H2, G0 (h = fp16, hl = split, id = exact)."""
    Ws, bs = spec.make_weights(CH, 100 + seed)
    Ws = [w.to(D) for w in Ws]; bs = [b.to(D) for b in bs]
    feat = spec.make_feat(C, 128, 128, 200 + seed, SCALE)[0]
    F = feat.permute(1, 2, 0).reshape(-1, C).to(D)
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(N, generator=g) * 126.9
    v = torch.rand(N, generator=g) * 126.9
    z = ((torch.rand(N, generator=g) * 2 - 1) * spec.Z_SCALE).to(D)
    x0, y0 = u.floor().long(), v.floor().long()
    wx, wy = (u - x0).to(D), (v - y0).to(D)
    offs = [y0 * 128 + x0, y0 * 128 + x0 + 1, (y0 + 1) * 128 + x0, (y0 + 1) * 128 + x0 + 1]
    wg = [(1 - wx) * (1 - wy), wx * (1 - wy), (1 - wx) * wy, wx * wy]
    x = sum(w[:, None] * F[o] for w, o in zip(wg, offs))
    xin = torch.cat([x, z[:, None]], 1)
    y = xin
    for l in range(5):
        y = (y if l == 0 else torch.cat([y, xin], 1)) @ Ws[l].t() + bs[l]
        if l < 4:
            y = lrelu(y)
    last = torch.sigmoid if KIND == "G" else torch.tanh
    want = last(y)
    # ---- model of the tensor-core program
    G0 = cfg["G0"](cfg["F0"](F) @ cfg["W0"](Ws[0][:, :C]).t())
    pre0 = sum(w[:, None] * G0[o] for w, o in zip(wg, offs)) + bs[0] + Ws[0][:, C] * z[:, None]
    act = cfg["H0"](lrelu(cfg["H0"](pre0)))
    hid = [0, 1024, 512, 256]
    for l in (1, 2, 3):
        Wh, Wx, wz = Ws[l][:, :hid[l]], Ws[l][:, hid[l]:hid[l] + C], Ws[l][:, hid[l] + C]
        fh = cfg.get("W%dh" % l, cfg["W%d" % l]); fx = cfg.get("W%dx" % l, cfg["W%d" % l])
        pre = act @ fh(Wh).t() + cfg["X%d" % l](x) @ fx(Wx).t() + bs[l] + wz * z[:, None]
        act = lrelu(pre)
        if l < 3:
            act = cfg["H%d" % l](act)
    W4 = Ws[4]
    logit = act @ W4[:, :128].t() + x @ W4[:, 128:128 + C].t() + W4[:, 128 + C] * z[:, None] + bs[4]
    return (last(logit) - want).abs().max().item()


ident = lambda t: t
BASE = dict(G0=h, F0=h, W0=h, H0=h, H1=h, H2=h, W1=h, W2=h, W3=h, X1=h, X2=h, X3=h)


def report(name, cfg):
    worst = max(run(s, cfg) for s in range(S))
    print("%-34s max |tc - exact| = %.3e" % (name, worst), flush=True)
    return worst


print("head %s, features x%g, %d points x %d seeds (bar 1e-4 on the returned value)" % (KIND, SCALE, N, S))
report("all fp16 (current program)", BASE)
for k in ("G0", "H0", "H1", "H2", "W1", "W2", "W3", "X1", "X2", "X3"):
    report("  exact %s" % k, dict(BASE, **{k: ident}))
report("  exact X1+X2+X3", dict(BASE, X1=ident, X2=ident, X3=ident))
report("  exact W1+W2+W3", dict(BASE, W1=ident, W2=ident, W3=ident))
report("  exact H0+H1+H2+G0", dict(BASE, H0=ident, H1=ident, H2=ident, G0=ident))
report("  split X3", dict(BASE, X3=hl))
report("  split X2+X3", dict(BASE, X2=hl, X3=hl))
report("  split X3 + H2", dict(BASE, X3=hl, H2=hl))
report("  split X2+X3 + H2", dict(BASE, X2=hl, X3=hl, H2=hl))
report("  split X3 + W3", dict(BASE, X3=hl, W3=hl))
report("  split X3+H2+W3 (layer 3 ~fp32)", dict(BASE, X3=hl, H2=hl, W3=hl))
report("  split X2+X3+H1+H2+W2+W3", dict(BASE, X2=hl, X3=hl, H1=hl, H2=hl, W2=hl, W3=hl))
report("  layer-3 skip hoisted to texels (X3, W3x exact)", dict(BASE, X3=ident, W3x=ident))
report("  + W3h split", dict(BASE, X3=ident, W3x=ident, W3h=hl))
report("  layer-2+3 skip hoisted", dict(BASE, X3=ident, W3x=ident, X2=ident, W2x=ident))
