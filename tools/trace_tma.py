"""Per-warp clock64 timeline of CTA 0 of the two-tiles-in-flight MiniMLP chain kernel (DN_TRACE events)."""
import os, sys, ctypes, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
V, C, NW, NE = 200000, 128, 20, 4096
ENG = os.environ.get("TRACE_ENGINE", "tc3x")
dn.set_engine(ENG)
print("engine", ENG, "variant", os.environ.get("DN_TC_VARIANT", "0"))
g = torch.Generator().manual_seed(0)
x, xd, ft = (torch.randn(V, C, generator=g).cuda() for _ in range(3))
p = dn.synthetic.block_weights(C, seed=0)
ws = [p["mlp.miniMLP_mlp_layer_{:03d}.weight".format(i)].cuda() for i in range(3)]
bs = [p["mlp.miniMLP_mlp_layer_{:03d}.bias".format(i)].cuda() for i in range(3)]
dn._lib.load()
raw = ctypes.CDLL(dn._lib.LIB_PATH)
with torch.no_grad():
    for _ in range(3): dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
    buf = torch.zeros(NW * NE * 2, dtype=torch.int64, device="cuda")
    raw.dn_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    dn.ops.mlp_apply([x, xd, ft], ws, bs, residual=x)
    torch.cuda.synchronize()
    raw.dn_debug_set_trace(ctypes.c_void_p(0))
t = buf.cpu().view(NW, NE, 2)
names = {1: "put begin", 2: "put a_empty ok", 5: "put arrived", 30: "cvt wait box", 31: "cvt box ok",
         10: "epi wait d_full", 11: "epi d_full ok", 12: "epi acc in regs", 14: "epi done",
         20: "mma chunk begin", 25: "mma b ok", 21: "mma a ok", 22: "mma issued", 23: "mma job begin", 24: "mma d_empty ok"}
for w in ((2, 4) if os.environ.get("TRACE_BRIEF") else (2, 4, 8, 12, 16)):
    ev = [(int(e), int(c)) for e, c in t[w].tolist() if e != 0]
    if not ev: continue
    t0, t1 = ev[0][1], ev[-1][1]
    dur = collections.defaultdict(list)
    for (e0, c0), (e1, c1) in zip(ev[:-1], ev[1:]):
        dur[(e0, e1)].append(c1 - c0)
    print("warp {} : {} events, span {} cycles".format(w, len(ev), t1 - t0))
    for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
        if sum(v) * 100 < (t1 - t0): continue
        print("    {:>18s} -> {:<18s} n={:4d} mean={:7.0f} max={:7d} total={:8d} ({:4.1f}%)".format(
            names.get(k[0], str(k[0])), names.get(k[1], str(k[1])), len(v), sum(v) / len(v), max(v), sum(v),
            100.0 * sum(v) / (t1 - t0)))
# timeline of the MMA warp's jobs for the second pair of tiles
ev = [(int(e), int(c)) for e, c in t[2].tolist() if e != 0]
jobs = [c for e, c in ev if e == 23]
print("mma job starts (cycles since first):", [c - jobs[0] for c in jobs[:20]])
for w in (4,):
    ev = [(int(e), int(c)) for e, c in t[w].tolist() if e != 0]
    print("warp 4 first 120 events (ev, dt):", [(e, c - ev[0][1]) for e, c in ev[:120]])
