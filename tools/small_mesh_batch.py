"""BASELINE config 4: a batch of 32 small meshes (V~2k, K=128, C=128), 4-block DiffusionNet forward, one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffusion_net_b200 as dn
dn.set_engine(os.environ.get("DN_B200_ENGINE", "tc3x"))
K, C = 128, 128
meshes = []
for i in range(32):
    n, m = 36 + i % 9, 50
    ops_t = dn.synthetic.structural_operators(n, m, K, seed=i, device="cuda")
    x = torch.randn(n * m, 3, generator=torch.Generator().manual_seed(i)).cuda()
    meshes.append((x, ops_t))
net = dn.DiffusionNet(C_in=3, C_out=8, C_width=C, N_block=4, dropout=False).cuda().eval()
with torch.no_grad():
    for p in net.parameters():
        if p.dim() == 1 and p.shape[0] == C and "diffusion_time" in [n for n, q in net.named_parameters() if q is p][0]:
            p.uniform_(1e-3, 0.3)
def fwd_all():
    outs = []
    with torch.no_grad():
        for x, (mass, L, evals, evecs, gX, gY) in meshes:
            outs.append(net(x, mass, L=None, evals=evals, evecs=evecs, gradX=gX, gradY=gY))
    return outs
for _ in range(3): fwd_all()
torch.cuda.synchronize()
t0 = time.perf_counter(); n_rep = 5
for _ in range(n_rep): fwd_all()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n_rep
V = sum(m[0].shape[0] for m in meshes)
l0 = dn._lib.load().dn_kernel_launch_count(); fwd_all(); torch.cuda.synchronize()
print("32 meshes sumV={} 4 blocks: {:.2f} ms per batch = {:.2f} Mverts/s (net forward), {} kernel launches".format(
    V, dt * 1e3, V / dt / 1e6, dn._lib.load().dn_kernel_launch_count() - l0))

# ---- CUDA-graph replay, meshes round-robin on 4 streams
gn = dn.graphs.GraphedNet(net, n_streams=4)
items = [dict(x_in=x, mass=o[0], evals=o[2], evecs=o[3], gradX=o[4], gradY=o[5]) for x, o in meshes]
ref = fwd_all()
outs = gn.forward_batch(items)
torch.cuda.synchronize()
err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, ref))
for _ in range(3): gn.forward_batch(items)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n_rep): gn.forward_batch(items)
torch.cuda.synchronize()
dt2 = (time.perf_counter() - t0) / n_rep
print("graphs: {:.2f} ms per batch = {:.2f} Mverts/s (net forward); max rel diff vs eager {:.2e}".format(dt2 * 1e3, V / dt2 / 1e6, err))
