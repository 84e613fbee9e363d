"""Per-warp clock64 timeline of CTA 0 of the fused MiniMLP chain kernel (DN_TRACE events in dn_tc.cu)."""
import os, sys, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
V, C, NW, NE = 200000, 128, 10, 4096
dn.set_engine("tc3x")
g = torch.Generator().manual_seed(0)
x, xd, ft = (torch.randn(V, C, generator=g).cuda() for _ in range(3))
p = dn.synthetic.block_weights(C, seed=0)
ws = [p["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)].cuda() for i in range(3)]
bs = [p["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)].cuda() for i in range(3)]
lib = dn._lib.load()
raw = ctypes.CDLL(dn._lib.LIB_PATH)
with torch.no_grad():
    for _ in range(3): dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
    buf = torch.zeros(NW * NE * 2, dtype=torch.int64, device="cuda")
    raw.dn_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
    torch.cuda.synchronize()
    raw.dn_debug_set_trace(ctypes.c_void_p(0))
t = buf.cpu().view(NW, NE, 2)
names = {1: "L0 step begin", 2: "L0 empty-wait done", 3: "L0 split+STS done", 4: "L0 fence+syncwarp done", 5: "L0 arrive done",
         10: "epi wait d_full", 11: "epi d_full ready", 12: "epi tmem_ld done", 13: "epi bias/relu/res done", 14: "epi store done",
         15: "epi empty-wait done", 16: "epi STS+fence+arrive done", 20: "mma wait full", 21: "mma full ready", 22: "mma issued+committed",
         30: "tma wait empty", 31: "tma empty ready"}
for w in range(NW):
    ev = [(int(e), int(c)) for e, c in t[w].tolist() if e != 0]
    if not ev: continue
    t0, t1 = ev[0][1], ev[-1][1]
    dur = collections.defaultdict(list)
    for (e0, c0), (e1, c1) in zip(ev[:-1], ev[1:]):
        dur[(e0, e1)].append(c1 - c0)
    print("warp {} : {} events, span {} cycles".format(w, len(ev), t1 - t0))
    for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) * 50 < (t1 - t0): continue
        print("    {:>26s} -> {:<28s} n={:4d} mean={:7.0f} total={:8d} ({:4.1f}%)".format(
            names.get(k[0], str(k[0])), names.get(k[1], str(k[1])), len(v), sum(v) / len(v), sum(v), 100.0 * sum(v) / (t1 - t0)))
