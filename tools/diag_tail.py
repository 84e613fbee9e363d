"""Diagnostics for the TMA tail path of rows_chain_tma_kernel: where do results differ from the SIMT engine?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
V, C = int(os.environ.get("DIAG_V", 200000)), 128
g = torch.Generator().manual_seed(0)
x, xd, ft = (torch.randn(V, C, generator=g).cuda() for _ in range(3))
p = dn.synthetic.block_weights(C, seed=0)
ws = [p["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)].cuda() for i in range(3)]
bs = [p["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)].cuda() for i in range(3)]
with torch.no_grad():
    dn.set_engine("simt")
    y0 = dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
    m0 = dn.ops.mlp_apply([x, xd, ft], ws, bs)
    dn.set_engine("tc3x")
    for it in range(8):
        y = dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
        torch.cuda.synchronize()
        e = (y - y0).abs()
        bad = e > 1e-4
        rows = bad.any(1).nonzero().flatten()
        print("call {}: rel err {:.3e}, bad elements {}, bad rows {}".format(it, float(e.max() / y0.abs().max()), int(bad.sum()), rows.numel()), flush=True)
        if rows.numel():
            r = rows.cpu()
            print("   first rows", r[:12].tolist(), "last", r[-4:].tolist())
            print("   tile index (row//128) distinct:", torch.unique(r // 128).numel(), "first", torch.unique(r // 128)[:12].tolist())
            print("   tile%148 distinct:", torch.unique((r // 128) % 148)[:20].tolist())
            print("   (tile//148) histogram:", torch.bincount((r // 128) // 148).tolist())
            print("   row%128 quarter histogram:", torch.bincount((r % 128) // 32, minlength=4).tolist())
            cols = bad.any(0).nonzero().flatten().cpu()
            print("   bad col blocks of 32:", torch.bincount(cols // 32, minlength=4).tolist())
            rr = int(r[0])
            cb = int(bad[rr].nonzero()[0])
            print("   sample row", rr, "col", cb, "got", float(y[rr, cb]), "want", float(y0[rr, cb]), "mlp-only", float(m0[rr, cb]), "x", float(x[rr, cb]))
            # is it the right mlp value plus a residual from another row?
            d = (y[rr] - m0[rr])[cb // 32 * 32:cb // 32 * 32 + 32]
            cand = (x[:, cb // 32 * 32:cb // 32 * 32 + 32] - d[None, :]).abs().max(1).values
            print("   residual used matches x row", int(cand.argmin()), "err", float(cand.min()))
