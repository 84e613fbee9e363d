"""Where do the tc3x MiniMLP gradients differ from the SIMT engine's on the block's real inputs?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import diffusion_net_b200 as dn
n, m, K, C = 84, 84, 128, 128
mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(n, m, K, seed=3, device="cuda")
V = n * m
params = {k: v.cuda() for k, v in dn.synthetic.block_weights(C, seed=3).items()}
g = torch.Generator().manual_seed(5)
x = torch.randn(V, C, generator=g).cuda(); R = torch.randn(V, C, generator=g).cuda()
ws = [params["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)] for i in range(3)]
bs = [params["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)] for i in range(3)]
t = params["diffusion.diffusion_time"]
dn.set_engine("simt")
xd = dn.ops.DiffusionFn.apply(x, t.clone(), mass, evals, evecs)
ft = torch.randn_like(xd)

def run(engine):
    dn.set_engine(engine)
    inputs = [i.clone().requires_grad_(True) for i in [x, xd, ft] + ws + bs]
    y = dn.ops.mlp_apply(inputs[:3], inputs[3:6], inputs[6:9], residual=inputs[0])
    sv = y.grad_fn.saved_tensors
    hidden = [s.detach().clone() for s in sv[6:8]]
    gs = torch.autograd.grad((y * R).sum(), inputs)
    return y.detach(), hidden, gs
y0, h0, g0 = run("simt")
y1, h1, g1 = run("tc3x")
rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))
print("out", rel(y1, y0), "hidden", [rel(a, b) for a, b in zip(h1, h0)])
for l in range(2):
    flips = ((h1[l] > 0) != (h0[l] > 0))
    print("layer", l, "mask flips:", int(flips.sum()), "of", flips.numel(), " frac zeros simt {:.3f} tc {:.3f}".format(float((h0[l] == 0).float().mean()), float((h1[l] == 0).float().mean())))
    if flips.any():
        idx = flips.nonzero()[:10]
        for r, c in idx.tolist():
            print("    ", r, c, float(h0[l][r, c]), float(h1[l][r, c]))
d = (g1[1] - g0[1]).abs()
print("g_xd err max", float(d.max()), "ref max", float(g0[1].abs().max()))
bad = (d > 1e-4 * g0[1].abs().max())
print("bad elements", int(bad.sum()), "rows", int(bad.any(1).sum()), "cols", int(bad.any(0).sum()))
rows = bad.any(1).nonzero().flatten()
print("bad rows (first 40):", rows[:40].tolist())
print("bad rows mod 128 histogram:", torch.bincount(rows % 128, minlength=128).tolist())
print("bad row tiles:", torch.unique(rows // 128).tolist()[:60])
