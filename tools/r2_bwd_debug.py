"""Backward bring-up at BASELINE config 2's shape: per-gradient error of each engine vs fp64 autograd of the torch
oracle, and device time of each autograd Function's forward / backward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import diffusion_net_b200 as dn
import dn_oracle_torch as T

n, m, K, C = 84, 84, 128, 128
mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(n, m, K, seed=3, device="cuda")
V = n * m
params = dn.synthetic.block_weights(C, seed=3)
g = torch.Generator().manual_seed(5)
x = torch.randn(V, C, generator=g); R = torch.randn(V, C, generator=g)
rel = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-300))

d = torch.float64
prm = {k: v.to(d).requires_grad_(True) for k, v in params.items()}
x64 = x.to(d).unsqueeze(0).requires_grad_(True)
gold = T.block_forward(x64, mass.cpu().to(d).unsqueeze(0), evals.cpu().to(d).unsqueeze(0), evecs.cpu().to(d).unsqueeze(0),
                       [gX.cpu().to(d)], [gY.cpu().to(d)], prm)
(gold[0] * R.to(d)).sum().backward()

def t_us(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n

for engine in ("simt", "tc3x"):
    dn.set_engine(engine)
    blk = dn.DiffusionNetBlock(C_width=C, mlp_hidden_dims=[C, C], dropout=False)
    blk.load_state_dict(params, strict=True)
    blk = blk.cuda().train()
    xg = x.cuda().unsqueeze(0).requires_grad_(True)
    out = blk(xg, mass.unsqueeze(0), None, evals.unsqueeze(0), evecs.unsqueeze(0), [gX], [gY])
    (out[0] * R.cuda()).sum().backward()
    print("== engine", engine, " out err {:.2e}  dx err {:.2e}".format(rel(out[0].detach(), gold[0].detach()), rel(xg.grad[0], x64.grad[0])))
    for name, p_ in blk.named_parameters():
        print("   {:40s} {:.2e}".format(name, rel(p_.grad, prm[name].grad)))
    # component timings
    gops = dn.ops.prepare_operators(gX, gY)
    xc = x.cuda()
    A_re, A_im = blk.gradient_features.weights()
    lins = blk.mlp.linears()
    ws, bs = [l.weight for l in lins], [l.bias for l in lins]
    def run(fn_fwd):
        xr = xc.clone().requires_grad_(True)
        y = fn_fwd(xr)
        gr = torch.randn_like(y)
        tf = t_us(lambda: fn_fwd(xr))
        tb = t_us(lambda: torch.autograd.grad(fn_fwd(xr), xr, gr)) - tf
        return tf, tb
    print("   diffusion   fwd {:.0f} us  bwd {:.0f} us".format(*run(lambda t: dn.ops.DiffusionFn.apply(t, blk.diffusion.diffusion_time, mass, evals, evecs))))
    print("   gradfeat    fwd {:.0f} us  bwd {:.0f} us".format(*run(lambda t: dn.ops.GradFeaturesFn.apply(t, A_re, A_im, gops))))
    print("   mlp         fwd {:.0f} us  bwd {:.0f} us".format(*run(lambda t: dn.ops.mlp_apply([t, xc, xc], ws, bs, residual=t))))
    l0 = dn._lib.load().dn_kernel_launch_count()
    xr = xc.clone().requires_grad_(True)
    y = dn.ops.mlp_apply([xr, xc, xc], ws, bs, residual=xr); torch.autograd.grad(y, xr, torch.ones_like(y))
    print("   mlp fwd+bwd launches:", dn._lib.load().dn_kernel_launch_count() - l0)
