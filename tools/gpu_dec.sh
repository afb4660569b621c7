mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -k "decoder" 2>&1 | grep -E "passed|failed|^E  " | cut -c1-200 | head -5
python - <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from mmvae_b200 import ops
from mmvae_b200.synth import synth_csr
G, H = 60530, 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B in (1024, 4096):
    crow, col, val = (torch.from_numpy(a).cuda() for a in synth_csr(B, G, 0.05, 1))
    nnz = int(col.numel())
    h = torch.relu(torch.randn(B, H, device="cuda")).bfloat16(); W = (torch.randn(G, H, device="cuda") * 0.03).bfloat16()
    bout = torch.zeros(G, device="cuda"); ldd = (G + 63) // 64 * 64
    dl = torch.zeros(B, ldd, dtype=torch.bfloat16, device="cuda"); ls = torch.zeros(1, dtype=torch.float64, device="cuda")
    tp, packed = ops.csr_tile_ptr(crow, col, val, G, nnz)
    ts = []
    for i in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.decoder_mse_fused(h, W, bout, G, crow, col, val, dl, ls, tile_ptr=tp); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); t = ts[len(ts) // 2]
    print(f"decoder B={B}: {t:.4f} ms  {2.0 * B * G * H / t / 1e9:.0f} TFLOP/s")
PY
