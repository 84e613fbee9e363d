"""Bring-up: repeatability of rows_chain16_kernel per chain shape (and per DN_C16_DBG switch)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
import diffusion_net_b200 as dn
dn.set_engine("bf16")
def lin(k, n, g):
    return ((torch.rand(n, k, generator=g) * 2 - 1) / k ** 0.5).cuda(), ((torch.rand(n, generator=g) * 2 - 1) / k ** 0.5).cuda()
for V in (100, 5000, 60000):
    g = torch.Generator().manual_seed(V)
    for dims, nsrc in (([384, 128], 1), ([384, 128], 3), ([384, 128, 128], 3), ([384, 128, 128, 128], 3), ([384, 128, 128, 128], 1),
                       ([768, 256, 256, 256], 3), ([768, 256], 1), ([192, 64], 3), ([256, 128, 128, 128], 1), ([128, 128, 128, 128], 1)):
        w = dims[0] // nsrc
        srcs = [torch.randn(V, w, generator=g).cuda() for _ in range(nsrc)]
        wb = [lin(dims[i], dims[i + 1], g) for i in range(len(dims) - 1)]
        ws, bs = [a for a, _ in wb], [b for _, b in wb]
        with torch.no_grad():
            first, bad = None, 0
            for _ in range(10):
                y = dn.ops.mlp_apply(srcs, ws, bs)
                torch.cuda.synchronize()
                if first is None: first = y.clone()
                elif not torch.equal(y, first): bad += 1
        print("   V={:6d} dims {} nsrc {}: {}".format(V, dims, nsrc, "NONDET x{}".format(bad) if bad else "ok"), flush=True)
''' % ROOT
for v in sys.argv[1:] or ["0"]:
    env = dict(os.environ, DN_C16_DBG=v)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    print("DN_C16_DBG={}:\n{}".format(v, "\n".join(l for l in r.stdout.splitlines() if l.startswith("   ")) or r.stderr[-600:]), flush=True)
