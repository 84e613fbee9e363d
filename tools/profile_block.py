"""A few DiffusionNetBlock forwards at the BASELINE metric size, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn

n, m, K, C = 400, 500, 128, int(os.environ.get("DN_PROFILE_C", "128"))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dn.set_engine(os.environ.get("DN_B200_ENGINE", "tc3x"))
mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(n, m, K, seed=0, device="cuda")
params = dn.synthetic.block_weights(C, seed=0)
x = torch.randn(n * m, C, generator=torch.Generator().manual_seed(0)).cuda()
blk = dn.DiffusionNetBlock(C_width=C, mlp_hidden_dims=[C, C], dropout=False)
blk.load_state_dict(params)
blk = blk.cuda().eval()
with torch.no_grad():
    for _ in range(steps):
        out = blk(x[None], mass[None], None, evals[None], evecs[None], [gX], [gY])
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
