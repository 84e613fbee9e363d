"""A/B of the gradient-features gather variants inside the fused block forward (V=200k, C=128): per-stage device times
from dn_block_fwd_profile and parity vs the exact SIMT engine; each setting in its own process (env read once).
DN_SPMM_BLK: 0 = round-1 warp-per-row kernel, 1 = block gather (64 rows per CTA, metadata staged in smem, FFMA2; default at C=128)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import diffusion_net_b200 as dn
C = 128
for permute in (False, True):
    mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(400, 500, 128, seed=0, device="cuda", permute=permute)
    V = mass.shape[0]
    params = dn.synthetic.block_weights(C, seed=0)
    x = torch.randn(V, C, generator=torch.Generator().manual_seed(1)).cuda()
    blk = dn.DiffusionNetBlock(C_width=C, mlp_hidden_dims=[C, C], dropout=False)
    blk.load_state_dict(params, strict=True)
    blk = blk.cuda().eval()
    gops = dn.ops.GradOperators(gX, gY)
    A_re, A_im = blk.gradient_features.weights()
    lins = blk.mlp.linears()
    run = lambda prof=None: dn.ops.block_forward_raw(x, mass, evals, evecs, gops, blk.diffusion.diffusion_time, A_re, A_im,
                                                     [l.weight for l in lins], [l.bias for l in lins], True, profile=prof)
    with torch.no_grad():
        dn.set_engine("simt"); y0 = run(); dn.set_engine("tc3x")
        y = run(); torch.cuda.synchronize()
        err = float((y - y0).abs().max() / y0.abs().max())
        acc = [0.0] * 6
        for it in range(12):
            prof = []; run(prof)
            if it >= 2: acc = [a + b for a, b in zip(acc, prof)]
    print("permute={} err {:.2e} stages_us {}".format(permute, err, {n: round(100 * a, 1) for n, a in zip(dn.ops.PROFILE_STAGES, acc)}), flush=True)
''' % ROOT
for v in sys.argv[1:] or ["0", "1"]:
    env = dict(os.environ, DN_SPMM_BLK=v, DN_SPMM_PATCH="0")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=400)
    out = "\n".join(l for l in r.stdout.splitlines() if l.startswith("permute"))
    print("DN_SPMM_BLK={}:\n{}".format(v, out or r.stderr.strip()[-800:]), flush=True)
