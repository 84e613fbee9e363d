"""Timing ablations of the fused MiniMLP chain kernel (DN_TC_VARIANT bits; outputs are invalid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
V, C = 200000, 128
dn.set_engine("tc3x")
g = torch.Generator().manual_seed(0)
x, xd, ft = (torch.randn(V, C, generator=g).cuda() for _ in range(3))
p = dn.synthetic.block_weights(C, seed=0)
ws = [p["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)].cuda() for i in range(3)]
bs = [p["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)].cuda() for i in range(3)]
evecs = torch.randn(V, 128, generator=g).cuda(); S = torch.randn(128, C, generator=g).cuda()
def t_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
names = {0: "baseline", 512: "only: WITH L2 bulk prefetch", 30: "no fence/loads/mma/tmemld", 62: "  + no weight TMA", 94: "  + no STS (w/ weight TMA)",
         158: "  + no global stores", 286: "  + no bias/residual loads", 510: "all off", 32: "only: no weight TMA",
         64: "only: no STS", 128: "only: no global stores", 256: "only: no bias/residual"}
with torch.no_grad():
    for v, nm in names.items():
        os.environ["DN_TC_VARIANT"] = str(v)
        t1 = t_ms(lambda: dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x))
        t2 = t_ms(lambda: dn.from_basis(S, evecs))
        print("variant {:2d} {:28s} mlp {:.3f} ms   from_basis {:.3f} ms".format(v, nm, t1, t2), flush=True)
