"""Device timings of the data-side neighbours of the block at BASELINE size (V=200k, K=128):
HKS features (dn_compute_hks) and the operator-cache CSC -> device CSR conversion (dn_csr_transpose)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
mass, L, evals, evecs, gX, gY = dn.synthetic.structural_operators(400, 500, 128, seed=0, device="cuda")
V = mass.shape[0]
def t_us(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n
scales = torch.logspace(-2, 0, 16, device="cuda")
t = t_us(lambda: dn.geometry.compute_hks(evals, evecs, scales))
print("compute_hks V={} K=128 S=16: {:.1f} us  ({:.0f} GB/s of evecs read)".format(V, t, V * 128 * 4 / t / 1e3))
def ref_hks():
    pc = torch.exp(-evals.unsqueeze(0) * scales.unsqueeze(-1))
    return ((evecs * evecs) @ pc.t())
t2 = t_us(ref_hks)
print("  torch (square + matmul, not the reference's (V,S,K) expansion): {:.1f} us".format(t2))
g = dn.ops.prepare_operators(gX, gY)
rp, ci, va = (a.cpu().numpy() for a in g.csr[1:])
torch.cuda.synchronize()
t0 = time.perf_counter()
o = dn.ops.GradOperators.from_csc(V, rp, ci, va[0::2], va[1::2], "cuda")
torch.cuda.synchronize()
print("from_csc (H2D of CSC arrays + dn_csr_transpose), nnz={}: {:.2f} ms wall".format(g.nnz, 1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter()
o2 = dn.ops.GradOperators(gX, gY)
_ = o2.csr_t
torch.cuda.synchronize()
print("COO path (dn_csr_from_coo + argsort transpose, operands already on device): {:.2f} ms wall".format(1e3 * (time.perf_counter() - t0)))
st = dn._lib.dn_csr(g.csr[1].data_ptr(), g.csr[2].data_ptr(), g.csr[3].data_ptr(), g.nnz)
import ctypes as C
rpo = torch.empty(V + 1, dtype=torch.int32, device="cuda"); cio = torch.empty(g.nnz, dtype=torch.int32, device="cuda")
vo = torch.empty(2 * g.nnz, dtype=torch.float32, device="cuda"); scr = torch.empty(V, dtype=torch.int32, device="cuda")
lib = dn._lib.load()
t3 = t_us(lambda: lib.dn_csr_transpose(C.byref(st), V, rpo.data_ptr(), cio.data_ptr(), vo.data_ptr(), scr.data_ptr(), 4 * V,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
print("dn_csr_transpose kernels alone: {:.1f} us".format(t3))
