"""Isolate the tensor-core backward pieces: every output of each autograd Function's backward, tc3x vs the exact SIMT
engine (validated against the reference's autograd by the golden tests), at V = 7056 / C = 128."""
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
x, xd, ft, R = (torch.randn(V, C, generator=g).cuda() for _ in range(4))
ws = [params["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)] for i in range(3)]
bs = [params["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)] for i in range(3)]
A_re, A_im = params["gradient_features.A_re.weight"], params["gradient_features.A_im.weight"]
t = params["diffusion.diffusion_time"]
gops = dn.ops.GradOperators(gX, gY)
rel = lambda a, b: float((a - b).abs().max() / (b.abs().max() + 1e-30))

def grads(fn, inputs):
    inputs = [i.clone().requires_grad_(True) for i in inputs]
    y = fn(*inputs)
    gs = torch.autograd.grad((y * R).sum(), inputs)
    return [y.detach()] + [gg.detach() for gg in gs]

cases = {
    "mlp (x,xd,ft,W0,W1,W2,b0,b1,b2)": (lambda a, b, c, w0, w1, w2, b0, b1, b2: dn.ops.mlp_apply([a, b, c], [w0, w1, w2], [b0, b1, b2], residual=a),
                                          [x, xd, ft] + ws + bs),
    "single layer w/ relu chain (x,W1,W2)": (lambda a, w1, w2: dn.ops.mlp_apply([a], [w1, w2], [None, None]), [x, ws[1], ws[2]]),
    "gradfeat (xd,A_re,A_im)": (lambda a, r, i: dn.ops.GradFeaturesFn.apply(a, r, i, gops), [xd, A_re, A_im]),
    "diffusion (x,t)": (lambda a, tt: dn.ops.DiffusionFn.apply(a, tt, mass, evals, evecs), [x, t]),
}
for name, (fn, inp) in cases.items():
    dn.set_engine("simt"); g0 = grads(fn, inp)
    dn.set_engine("tc3x"); g1 = grads(fn, inp)
    print("{:40s} out {:.1e} | grads ".format(name, rel(g1[0], g0[0])) + " ".join("{:.1e}".format(rel(a, b)) for a, b in zip(g1[1:], g0[1:])), flush=True)
os.environ["X"] = "1"
