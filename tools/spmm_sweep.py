"""BASELINE config 5: density x batch sweep of the expert-encoder SpMM (forward + weight gradient) on one B200.
Writes a markdown table (ms, achieved GB/s on ALGORITHMIC bytes, fraction of the measured HBM peak, FMA TFLOP/s)
for the gather kernels and the tensor-pipe kernels, and says which one the engine picks."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmvae_b200 import ops  # noqa: E402

G, H = 60530, 1024
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    HBM = 6650.0


def synth(B, d, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    crow = [torch.zeros(1, dtype=torch.int32, device="cuda")]
    cols, n = [], 0
    for r0 in range(0, B, 1024):
        m = torch.rand(min(1024, B - r0), G, device="cuda", generator=g) < d
        cnt = m.sum(1)
        crow.append((n + cnt.cumsum(0)).to(torch.int32))
        n += int(cnt.sum())
        cols.append(m.nonzero()[:, 1].to(torch.int32))
    col = torch.cat(cols)
    val = (torch.rand(col.numel(), device="cuda", generator=g) * 6 + 0.5)
    return torch.cat(crow), col, val


def timed(fn, flush, n=5):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    Wt16 = (torch.randn(G, H, device="cuda") * 0.03).bfloat16()
    bias = torch.zeros(H, device="cuda")
    dWt = torch.empty(G, H, device="cuda")
    rows = []
    for B in (512, 1024, 2048, 4096, 8192, 16384):
        Y = torch.empty(B, H, device="cuda")
        dY = torch.randn(B, H, device="cuda")
        dY16 = dY.bfloat16()
        for d in (0.01, 0.02, 0.05, 0.10, 0.20):
            crow, col, val = synth(B, d, 7)
            nnz = int(col.numel())
            fwd_bytes = nnz * 8 + (B + 1) * 4 + G * H * 2 + B * H * 4 + H * 4
            bwd_bytes = nnz * 8 + B * H * 2 + G * H * 4
            t_g = timed(lambda: ops.csr_linear_fwd(crow, col, val, G, Wt16, bias, out=Y), flush)
            tp, packed = ops.csr_tile_ptr(crow, col, val, G, nnz)
            t_prep = timed(lambda: ops.csr_tile_ptr(crow, col, val, G, nnz, tp, packed), flush)
            t_t = timed(lambda: ops.csr_linear_fwd_tc(packed, tp, B, G, Wt16, bias, out=Y), flush)
            t_bt = timed(lambda: ops.csr_linear_bwd_w_tc(packed, tp, B, G, dY16, dWt), flush)
            cptr, ridx, cval = ops.csr_transpose(crow, col, val, G, nnz)
            t_bg = timed(lambda: (ops.csr_transpose(crow, col, val, G, nnz, cptr, ridx, cval),
                                  ops.csr_linear_bwd_w(cptr, ridx, cval, B, G, dY, dWt)), flush, n=3)
            pick = "tensor" if nnz >= 0.015 * B * G else "gather"
            best_f = t_t + t_prep if pick == "tensor" else t_g
            rows.append((B, d, nnz, t_g, t_t, t_prep, t_bg, t_bt, pick,
                         fwd_bytes / best_f / 1e6, fwd_bytes / best_f / 1e6 / HBM, 2.0 * nnz * H / best_f / 1e9,
                         bwd_bytes / (t_bt if pick == "tensor" else t_bg) / 1e6 / HBM))
            print(rows[-1], flush=True)
            del crow, col, val, tp, packed, cptr, ridx, cval
    out = os.path.join(ROOT, "gpurun_out", "spmm_sweep.md")
    with open(out, "w") as f:
        f.write("# Expert-encoder SpMM sweep (BASELINE config 5), 1x B200, G=60530, H=1024, bf16 weight\n\n"
                "Bernoulli(d) sparsity per (cell, gene); CUDA-event times, L2 flushed between launches; GB/s and "
                f"roofline fraction are ALGORITHMIC bytes / time against the measured {HBM:.0f} GB/s copy peak "
                "(fwd bytes = 8 nnz + 4(B+1) + 2 G H + 4 B H; bwd bytes = 8 nnz + 2 B H + 4 G H). "
                "`pick` = kernel family the engine selects (tensor above 1.5 % density); tensor fwd time includes "
                "the pointer-table/packing pre-pass.\n\n"
                "| B | d | nnz | fwd gather ms | fwd tensor ms | prep ms | bwd gather(+CSC) ms | bwd tensor ms | pick | "
                "fwd GB/s | fwd frac of HBM | fwd FMA TFLOP/s | bwd frac of HBM |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write(f"| {r[0]} | {r[1]:.0%} | {r[2]} | {r[3]:.3f} | {r[4]:.3f} | {r[5]:.3f} | {r[6]:.3f} | {r[7]:.3f} | "
                    f"{r[8]} | {r[9]:.0f} | {r[10]:.3f} | {r[11]:.1f} | {r[12]:.3f} |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
