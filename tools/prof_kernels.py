"""Run the three heaviest kernels at the bench shape (B=1024, G=60530, H=1024, 5%) a few times: target for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mmvae_b200 import ops
from mmvae_b200.synth import synth_csr
B, G, H = int(os.environ.get("PB", 1024)), 60530, 1024
crow, col, val = synth_csr(B, G, 0.05, 1)
crow, col, val = (torch.from_numpy(a).cuda() for a in (crow, col, val))
nnz = int(col.numel())
Wt16 = (torch.randn(G, H, device="cuda") * 0.03).bfloat16()
Wout16 = (torch.randn(G, H, device="cuda") * 0.03).bfloat16()
bias = torch.zeros(H, device="cuda"); bout = torch.zeros(G, device="cuda")
h16 = torch.relu(torch.randn(B, H, device="cuda")).bfloat16()
dY16 = torch.randn(B, H, device="cuda").bfloat16()
tp, packed = ops.csr_tile_ptr(crow, col, val, G, nnz)
Y = torch.empty(B, H, device="cuda"); dWt = torch.empty(G, H, device="cuda")
ldd = (G + 63) // 64 * 64
dl = torch.zeros(B, ldd, device="cuda", dtype=torch.bfloat16); ls = torch.zeros(1, dtype=torch.float64, device="cuda")
dW = torch.empty(G, H, device="cuda"); dh = torch.empty(B, H, device="cuda")
n_adam = 125_080_192 if B <= 1024 else 1024
pa, ga, ma, va = (torch.zeros(n_adam, device="cuda") for _ in range(4))
p16a = torch.zeros(n_adam, device="cuda", dtype=torch.bfloat16); nsq = torch.ones(1, dtype=torch.float64, device="cuda")
Wg = torch.empty(G, H, device="cuda")
for it in range(3):
    ops.csr_linear_fwd_tc(packed, tp, B, G, Wt16, bias, out=Y)
    ops.csr_linear_bwd_w_tc(packed, tp, B, G, dY16, dWt)
    ops.decoder_mse_fused(h16, Wout16, bout, G, crow, col, val, dl, ls)
    ops.gemm(dl, 1, h16, 1, G, H, B, C32=dW)
    ops.gemm(dl, 0, Wout16, 1, B, H, G, C32=dh)
    ops.csr_linear_fwd(crow, col, val, G, Wt16, bias, out=Y)
    ops.clip_adam(pa, ga, ma, va, p16a, nsq, 10.0, 1.0, 5e-3, 0.9, 0.999, 1e-8, 1e-6, it + 1)
torch.cuda.synchronize()
print("ok")
