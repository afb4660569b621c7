import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmvae_b200 import ops
from oracle.cmmvae_oracle import synth_csr
B, G, H = 1024, 60530, 1024
crow, col, val = synth_csr(B, G, 0.05, 1)
crow, col, val = (torch.from_numpy(a).cuda() for a in (crow, col, val))
nnz = int(col.numel())
Wout16 = (torch.randn(G, H, device="cuda") * 0.03).bfloat16(); bout = torch.zeros(G, device="cuda")
h16 = torch.relu(torch.randn(B, H, device="cuda")).bfloat16()
tp, packed = ops.csr_tile_ptr(crow, col, val, G, nnz)
ldd = (G + 63) // 64 * 64
dl = torch.zeros(B, ldd, device="cuda", dtype=torch.bfloat16); ls = torch.zeros(1, dtype=torch.float64, device="cuda")
flush = torch.empty(200 << 20, dtype=torch.uint8, device="cuda")
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
print("dbg", os.environ.get("CMMVAE_DEC_DBG", "0"), "decoder ms %.4f" % t(lambda: ops.decoder_mse_fused(h16, Wout16, bout, G, crow, col, val, dl, ls, tile_ptr=tp)),
      "tile_ptr+pack ms %.4f" % t(lambda: ops.csr_tile_ptr(crow, col, val, G, nnz, tp, packed)))
