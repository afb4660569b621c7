"""time the three tile kernels alone at the bench shape (debug knobs via env)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mmvae_b200._lib as _L
if os.environ.get('CMMVAE_LIB_PATH'): _L.LIB_PATH = os.environ['CMMVAE_LIB_PATH']; _L.needs_build = lambda: False
from mmvae_b200 import ops
from oracle.cmmvae_oracle import synth_csr
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=20):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
r = {"fwd": timeit(lambda: ops.csr_linear_fwd_tc(packed, tp, B, G, Wt16, bias, out=Y)),
     "bwd": timeit(lambda: ops.csr_linear_bwd_w_tc(packed, tp, B, G, dY16, dWt)),
     "dec": timeit(lambda: ops.decoder_mse_fused(h16, Wout16, bout, G, crow, col, val, dl, ls, tile_ptr=tp))}
print(os.environ.get("CMMVAE_LIB_PATH", "default")[-12:], os.environ.get("CMMVAE_SPMM_DBG", "0"), os.environ.get("CMMVAE_DEC_DBG", "0"), {k: round(v, 4) for k, v in r.items()})
