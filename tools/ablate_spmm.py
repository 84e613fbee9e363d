"""A/B timing of the spmm_features kernel variants (DN_SPMM_VARIANT is read once per process)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import torch
    import diffusion_net_b200 as dn
    dn.set_engine("tc3x")
    mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(400, 500, 128, seed=0, device="cuda", permute=(sys.argv[2] == "perm"))
    gops = dn.prepare_operators(gX, gY)
    C = 128
    xd = torch.randn(200000, C).cuda(); pw = dn.synthetic.block_weights(C)
    A_re, A_im = pw["gradient_features.A_re.weight"].cuda(), pw["gradient_features.A_im.weight"].cuda()
    def t_ms(fn, n=10):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    with torch.no_grad():
        t = t_ms(lambda: dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, gops))
        f = dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, gops)
    print("variant {} order {} grad_features(PQ gemm + spmm) {:.3f} ms checksum {:.6f}".format(sys.argv[1], sys.argv[2], t, float(f.double().sum())), flush=True)
else:
    for order in ("grid", "perm"):
        for v in ("0", "1", "2", "3"):
            env = dict(os.environ, DN_SPMM_VARIANT=v)
            subprocess.run([sys.executable, __file__, v, order], env=env)
