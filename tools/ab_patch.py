"""Fused gradient-features forward (PQ GEMM chain + gather) at V=200k, C=128: plain gather kernel vs the
shared-memory staged patch kernel (dn_patches) for a few patch shapes; also the permuted (worst-locality) order."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
dn.set_engine("tc3x")
C = 128
def t_us(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n
for permute in (False, True):
    mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(400, 500, 128, seed=0, device="cuda", permute=permute)
    V = mass.shape[0]
    g = torch.Generator().manual_seed(0)
    xd = torch.randn(V, C, generator=g).cuda()
    p = dn.synthetic.block_weights(C, seed=0)
    A_re, A_im = p["gradient_features.A_re.weight"].cuda(), p["gradient_features.A_im.weight"].cuda()
    with torch.no_grad():
        plain = dn.ops.GradOperators(gX, gY)
        ref = dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, plain)
        t0 = t_us(lambda: dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, plain))
        print("permute={}: plain gather: {:.1f} us (PQ chain + gather)".format(permute, t0), flush=True)
        for T, R in ((32, 72), (64, 144), (24, 48)):
            o = dn.ops.GradOperators(gX, gY)
            torch.cuda.synchronize(); w0 = time.perf_counter()
            o.build_patches(T, R)
            torch.cuda.synchronize(); w1 = time.perf_counter()
            out = dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, o)
            t1 = t_us(lambda: dn.ops.GradFeaturesFn.apply(xd, A_re, A_im, o))
            print("   patches T={:3d} R={:3d}: {:.1f} us  ({:+.1f})  identical={}  build {:.0f} ms  {}".format(
                T, R, t1, t1 - t0, bool(torch.equal(out, ref)), 1e3 * (w1 - w0), o.patch_stats), flush=True)
