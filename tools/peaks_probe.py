"""Round-2 probe (GPU box): cuBLAS TF32 / bf16 / fp32 GEMM peaks and the torch-eager block forward on the B200."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import diffusion_net_b200 as dn
import dn_oracle_torch as T

dev = torch.device("cuda", 0)


def gemm_rate(dtype, tf32, n=8192, secs=2.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype)
    b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t0 = time.time(); cnt = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            a @ b
        cnt += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / cnt
    f = 2.0 * n ** 3 / 1e12
    return f / (best * 1e-3), f / (sus * 1e-3)


out = {}
for name, dt, tf in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False), ("fp32", torch.float32, False)):
    b, s = gemm_rate(dt, tf, secs=2.0 if name != "fp32" else 1.0)
    out[name] = {"burst_tflops": b, "sustained_tflops": s}
    print("cuBLAS {} 8192^3: burst {:.1f} TFLOP/s, sustained {:.1f}".format(name, b, s), flush=True)
torch.backends.cuda.matmul.allow_tf32 = False
ops_t = dn.synthetic.structural_operators(400, 500, 128, seed=0, device="cuda")
mass, L, evals, evecs, gradX, gradY = ops_t
params = {k: v.cuda() for k, v in dn.synthetic.block_weights(128, seed=0).items()}
x = torch.randn(200000, 128, generator=torch.Generator().manual_seed(100)).cuda()


def step():
    with torch.no_grad():
        return T.block_forward(x.unsqueeze(0), mass.unsqueeze(0), evals.unsqueeze(0), evecs.unsqueeze(0),
                               [gradX], [gradY], params)


for _ in range(3):
    y = step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
out["torch_eager_block_ms"] = ms
print("torch-eager (fp32, TF32 off) block forward on this GPU: {:.3f} ms = {:.1f} Mverts/s".format(ms, 0.2 / ms * 1e3), flush=True)
print(json.dumps(out))
