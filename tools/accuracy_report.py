"""Per-stage and block-level error of each engine vs the fp64 oracle (checker only) at full size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import diffusion_net_b200 as dn
import dn_oracle as O

def run(n, m, K, C, seed=1):
    ops_t = dn.synthetic.structural_operators(n, m, K, seed=seed, device="cuda")
    mass, L, evals, evecs, gX, gY = ops_t
    params = dn.synthetic.block_weights(C, seed=seed)
    x = torch.randn(n * m, C, generator=torch.Generator().manual_seed(seed)).cuda()
    V = n * m; f = np.float64
    gxc, gyc = gX.coalesce().cpu(), gY.coalesce().cpu()
    cX = O.coo_to_csr(gxc.indices()[0].numpy(), gxc.indices()[1].numpy(), gxc.values().numpy().astype(f), (V, V))
    cY = O.coo_to_csr(gyc.indices()[0].numpy(), gyc.indices()[1].numpy(), gyc.values().numpy().astype(f), (V, V))
    p64 = {k: v.numpy().astype(f) for k, v in params.items()}
    gold, inter = O.diffusion_net_block(x.cpu().numpy().astype(f), mass.cpu().numpy().astype(f), evals.cpu().numpy().astype(f),
                                        evecs.cpu().numpy().astype(f), cX, cY, p64, return_intermediates=True)
    for eng in ("simt", "tc3x", "tc1x"):
        dn.set_engine(eng)
        blk = dn.DiffusionNetBlock(C_width=C, mlp_hidden_dims=[C, C], dropout=False)
        blk.load_state_dict(params); blk = blk.cuda().eval()
        with torch.no_grad():
            xd = blk.diffusion(x[None], None, mass[None], evals[None], evecs[None])[0]
            gops = dn.prepare_operators(gX, gY)
            A_re, A_im = blk.gradient_features.weights()
            feat = dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, gops)
            out = blk(x[None], mass[None], None, evals[None], evecs[None], [gX], [gY])[0]
        print("V={} K={} C={} {:5s} x_diffuse {:.2e}  features {:.2e}  block_out {:.2e}".format(
            V, K, C, eng, O.rel_err(xd.cpu().numpy(), inter["x_diffuse"]),
            O.rel_err(feat.cpu().numpy(), inter["x_grad_features"]), O.rel_err(out.cpu().numpy(), gold)), flush=True)

if __name__ == "__main__":
    run(70, 100, 128, 128)
    run(400, 500, 128, 128)
