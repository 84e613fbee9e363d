"""Bring-up probe for the tcgen05 engine: small GEMMs through the C-ABI vs fp64 torch, with
one-hot probes that print the operand-layout mapping when a result is wrong.
Usage: python tools/tc_debug.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import diffusion_net_b200 as dn  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def from_basis_case(V, K, C, engine, seed=0):
    g = torch.Generator().manual_seed(seed)
    basis = torch.randn(V, K, generator=g).cuda()
    vals = torch.randn(K, C, generator=g).cuda()
    dn.set_engine(engine)
    out = dn.from_basis(vals, basis)
    torch.cuda.synchronize()
    ref = basis.double() @ vals.double()
    return rel(out, ref), out, ref


def to_basis_case(V, K, C, engine, seed=0):
    g = torch.Generator().manual_seed(seed)
    basis = torch.randn(V, K, generator=g).cuda()
    x = torch.randn(V, C, generator=g).cuda()
    mass = (torch.rand(V, generator=g) + 0.5).cuda()
    dn.set_engine(engine)
    out = dn.to_basis(x, basis, mass)
    torch.cuda.synchronize()
    ref = basis.double().t() @ (x.double() * mass.double()[:, None])
    return rel(out, ref), out, ref


def onehot_probe():
    """A = e_(m0,k0), B = e_(n0,k0): D should be 1 at (m0,n0) only."""
    dn.set_engine("tc1x")
    V, K, C = 128, 32, 16
    for (m0, k0, n0) in [(0, 0, 0), (1, 0, 0), (8, 0, 0), (0, 1, 0), (0, 4, 0), (0, 8, 0), (0, 16, 0), (0, 0, 1),
                         (0, 0, 8), (37, 21, 11)]:
        basis = torch.zeros(V, K).cuda()
        vals = torch.zeros(K, C).cuda()
        basis[m0, k0] = 1.0
        vals[k0, n0] = 1.0
        out = dn.from_basis(vals, basis)
        torch.cuda.synchronize()
        nz = out.nonzero().tolist()
        print("  onehot A(m={},k={}) B(k={},n={}) -> nonzeros {}".format(m0, k0, k0, n0, nz[:6]), flush=True)


if __name__ == "__main__":
    print("device", torch.cuda.get_device_name(0), flush=True)
    for eng in ("simt", "tc1x", "tc3x"):
        for (V, K, C) in [(128, 32, 16), (128, 128, 128), (1000, 128, 128), (4096, 64, 256)]:
            e, out, ref = from_basis_case(V, K, C, eng)
            print("from_basis {:5s} V={} K={} C={} rel_err={:.3e}".format(eng, V, K, C, e), flush=True)
    e, _, _ = from_basis_case(128, 32, 16, "tc1x")
    if e > 1e-2:
        onehot_probe()
    for eng in ("simt", "tc1x", "tc3x"):
        for (V, K, C) in [(64, 128, 128), (1000, 128, 128), (5000, 64, 32), (200000, 128, 128)]:
            e, out, ref = to_basis_case(V, K, C, eng)
            print("to_basis   {:5s} V={} K={} C={} rel_err={:.3e}".format(eng, V, K, C, e), flush=True)
    # accumulation-chain probe: long K through the chain kernel (accumulator rounding behaviour)
    for eng in ("simt", "tc3x"):
        for K in (128, 384, 2048):
            e, _, _ = from_basis_case(2048, K, 128, eng, seed=3)
            print("chain K={} {:5s} rel_err={:.3e}".format(K, eng, e), flush=True)
