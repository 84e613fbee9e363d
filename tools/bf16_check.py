"""Bring-up of DN_ENGINE_BF16 (rows_chain16_kernel, dn_chain16.cu) and of C_width = 256 (BASELINE config 3):
parity of single layers / chains against the exact SIMT engine, then the whole fused block forward at
V = 200k, K = 128 for C in (128, 256) with per-stage device times (dn_block_fwd_profile), per engine.

    python tools/bf16_check.py [quick]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn

torch.manual_seed(0)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())


def lin(k, n, g):
    return ((torch.rand(n, k, generator=g) * 2 - 1) / k ** 0.5).cuda(), ((torch.rand(n, generator=g) * 2 - 1) / k ** 0.5).cuda()


def chains(V):
    g = torch.Generator().manual_seed(V)
    for dims, nsrc in (([128, 128], 1), ([256, 256], 1), ([128, 256], 1), ([64, 32], 1), ([192, 64, 96], 1),
                       ([384, 128, 128, 128], 3), ([768, 256, 256, 256], 3), ([256, 256, 256, 256, 256], 1),
                       ([128, 64, 128], 1)):
        w = dims[0] // nsrc
        srcs = [torch.randn(V, w, generator=g).cuda() for _ in range(nsrc)]
        wb = [lin(dims[i], dims[i + 1], g) for i in range(len(dims) - 1)]
        ws, bs = [a for a, _ in wb], [b for _, b in wb]
        res = srcs[0] if dims[-1] == w else None
        with torch.no_grad():
            dn.set_engine("simt")
            y0 = dn.ops.mlp_apply(srcs, ws, bs, residual=res)
            dn.set_engine("bf16")
            worst = 0.0
            first, nondet = None, 0
            for _ in range(6):
                y = dn.ops.mlp_apply(srcs, ws, bs, residual=res)
                torch.cuda.synchronize()
                worst = max(worst, rel(y, y0))
                if first is None:
                    first = y.clone()
                elif not torch.equal(y, first):
                    nondet += 1
                    bad = (y != first)
                    rows = bad.any(1).nonzero().flatten()
                    print("      NON-DETERMINISTIC call: {} elements, {} rows (first {}), cols {}..{}, max diff {:.2e}".format(
                        int(bad.sum()), rows.numel(), rows[:6].tolist(), int(bad.any(0).nonzero().min()),
                        int(bad.any(0).nonzero().max()), float((y - first).abs().max())), flush=True)
        print("   V={:6d} dims {}: bf16 vs exact {:.2e}{}".format(V, dims, worst, "  NONDET x{}".format(nondet) if nondet else ""), flush=True)


def block(C, engines, n=400, m=500):
    ops_t = dn.synthetic.structural_operators(n, m, 128, seed=0, device="cuda")
    mass, L, evals, evecs, gradX, gradY = ops_t
    V = n * m
    params = dn.synthetic.block_weights(C, seed=0)
    x = torch.randn(V, C, generator=torch.Generator().manual_seed(1)).cuda()
    blk = dn.DiffusionNetBlock(C_width=C, mlp_hidden_dims=[C, C], dropout=False)
    blk.load_state_dict(params, strict=True)
    blk = blk.cuda().eval()
    gops = dn.ops.GradOperators(gradX, gradY)
    A_re, A_im = blk.gradient_features.weights()
    lins = blk.mlp.linears()
    run = lambda prof=None: dn.ops.block_forward_raw(x, mass, evals, evecs, gops, blk.diffusion.diffusion_time, A_re, A_im,
                                                     [l.weight for l in lins], [l.bias for l in lins], True, profile=prof)
    with torch.no_grad():
        dn.set_engine("simt")
        y0 = run()
        for eng in engines:
            dn.set_engine(eng)
            y = run(); torch.cuda.synchronize()
            for rep in range(4):
                y2 = run(); torch.cuda.synchronize()
                if not torch.equal(y, y2):
                    bad = (y != y2)
                    rows = bad.any(1).nonzero().flatten()
                    print("   NON-DETERMINISTIC block ({}): {} elements in {} rows, first rows {}".format(
                        eng, int(bad.sum()), rows.numel(), rows[:8].tolist()), flush=True)
            # which stage?  repeat each op of the unfused route
            xd1 = dn.ops.DiffusionFn.apply(x, blk.diffusion.diffusion_time, mass, evals, evecs)
            xd2 = dn.ops.DiffusionFn.apply(x, blk.diffusion.diffusion_time, mass, evals, evecs)
            ft1 = dn.ops.GradFeaturesFn.apply(xd1, A_re, A_im, gops)
            ft2 = dn.ops.GradFeaturesFn.apply(xd1, A_re, A_im, gops)
            print("   stage repeatability ({}): diffusion {} features {}".format(eng, torch.equal(xd1, xd2), torch.equal(ft1, ft2)), flush=True)
            acc = [0.0] * 6
            for it in range(8):
                prof = []; run(prof)
                if it >= 2: acc = [a + b for a, b in zip(acc, prof)]
            st = {k: round(1000 * a / 6, 1) for k, a in zip(dn.ops.PROFILE_STAGES, acc)}
            print("block V={} C={} engine {}: err vs exact {:.2e}  total {:.1f} us  stages_us {}".format(
                V, C, eng, rel(y, y0), sum(st.values()), st), flush=True)


if __name__ == "__main__":
    quick = len(sys.argv) > 1
    for V in ((100, 5000) if quick else (100, 128, 129, 1000, 18945, 40000)):
        chains(V)
    block(128, ("tc3x", "bf16"), 30, 40)
    block(256, ("tc3x", "bf16"), 30, 40)
    if not quick:
        block(128, ("tc3x", "bf16"))
        block(256, ("tc3x", "bf16"))
