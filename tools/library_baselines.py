"""Same-box library timings next to every kernel claim: cuBLASLt bf16 GEMMs at the decoder /
dWout / dh shapes (+ the separate ReLU/MSE passes the fused decoder kernel absorbs), torch.sparse CSR addmm
(cuSPARSE) at the SpMM shape, and torch's fused Adam -- each beside this repo's kernel, CUDA events, L2 flushed
between launches.  Writes gpurun_out/library_baselines.json (summarised under profiles/)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmvae_b200 import ops  # noqa: E402
from mmvae_b200.synth import synth_csr  # noqa: E402

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    G, H = 60530, 1024
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": []}
    for B in (1024, 4096, 8192):
        crow, col, val = (torch.from_numpy(a).to(dev) for a in synth_csr(B, G, 0.05, 1))
        nnz = int(col.numel())
        h = torch.relu(torch.randn(B, H, device=dev)).bfloat16()
        W = (torch.randn(G, H, device=dev) * 0.03).bfloat16()
        bout = torch.zeros(G, device=dev)
        ldd = (G + 63) // 64 * 64
        dl = torch.zeros(B, ldd, dtype=torch.bfloat16, device=dev)
        ls = torch.zeros(1, dtype=torch.float64, device=dev)
        tp, packed = ops.csr_tile_ptr(crow, col, val, G, nnz)
        x_csr = torch.sparse_csr_tensor(crow, col, val, size=(B, G))
        flops = 2.0 * B * G * H
        row = {"B": B, "G": G, "H": H, "nnz": nnz, "gflop_dense": flops / 1e9}
        # decoder: ours (GEMM + bias + ReLU + MSE-vs-CSR + dlogits in one kernel)
        row["decoder_mse_fused_ms"] = timed(lambda: ops.decoder_mse_fused(h, W, bout, G, crow, col, val, dl, ls, tile_ptr=tp))
        # library: cuBLASLt GEMM alone, then with the passes the reference runs after it
        logits = torch.empty(B, G, dtype=torch.bfloat16, device=dev)
        row["cublaslt_gemm_bf16_ms"] = timed(lambda: torch.matmul(h, W.t(), out=logits))
        row["cublaslt_linear_bias_bf16_ms"] = timed(lambda: torch.nn.functional.linear(h, W, bout.bfloat16()))
        if B <= 4096:
            xd = x_csr.to_dense()

            def ref_chain():
                y = torch.nn.functional.linear(h, W, bout.bfloat16())
                xh = torch.relu(y).float()
                return torch.nn.functional.mse_loss(xh, xd, reduction="sum")
            row["torch_linear_relu_mse_ms"] = timed(ref_chain, reps=8)
            row["torch_to_dense_ms"] = timed(lambda: x_csr.to_dense(), reps=8)
            del xd
        # dWout = dlogits^T h  and  dh = dlogits Wout
        gW = torch.empty(G, H, device=dev)
        dh = torch.empty(B, H, device=dev)
        dlv = dl[:, :G]
        row["dWout_ours_ms"] = timed(lambda: ops.gemm(dl, 1, h, 1, G, H, B, C32=gW))
        row["dWout_cublaslt_bf16_ms"] = timed(lambda: torch.matmul(dlv.t(), h))
        row["dh_ours_ms"] = timed(lambda: ops.gemm(dl, 0, W, 1, B, H, G, C32=dh))
        row["dh_cublaslt_bf16_ms"] = timed(lambda: torch.matmul(dlv, W))
        # SpMM forward / weight gradient: ours (tensor pipe) vs cuSPARSE through torch.sparse
        Wt16 = (torch.randn(G, H, device=dev) * 0.03).bfloat16()
        b1 = torch.zeros(H, device=dev)
        Y = torch.empty(B, H, device=dev)
        row["spmm_fwd_tc_ms"] = timed(lambda: ops.csr_linear_fwd_tc(packed, tp, B, G, Wt16, b1, out=Y))
        row["spmm_fwd_gather_ms"] = timed(lambda: ops.csr_linear_fwd(crow, col, val, G, Wt16, b1, out=Y))
        Wt32 = Wt16.float()
        try:
            row["cusparse_addmm_f32_ms"] = timed(lambda: torch.addmm(b1, x_csr, Wt32), reps=8)
        except Exception as e:  # noqa: BLE001
            row["cusparse_addmm_f32_ms"] = f"error: {str(e)[:100]}"
        dY16 = torch.randn(B, H, device=dev).bfloat16()
        dWt = torch.empty(G, H, device=dev)
        row["spmm_bwd_tc_ms"] = timed(lambda: ops.csr_linear_bwd_w_tc(packed, tp, B, G, dY16, dWt))
        try:
            xt = x_csr.t().to_sparse_csr()
            dY32 = dY16.float()
            row["cusparse_XtdY_f32_ms"] = timed(lambda: torch.mm(xt, dY32), reps=8)
        except Exception as e:  # noqa: BLE001
            row["cusparse_XtdY_f32_ms"] = f"error: {str(e)[:100]}"
        for k in ("decoder_mse_fused_ms", "cublaslt_gemm_bf16_ms", "dWout_ours_ms", "dWout_cublaslt_bf16_ms",
                  "dh_ours_ms", "dh_cublaslt_bf16_ms"):
            row[k.replace("_ms", "_tflops")] = flops / (row[k] * 1e-3) / 1e12
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
        del W, dl, logits, gW, x_csr
        torch.cuda.empty_cache()
    # Adam over the human expert (125 M params): ours vs torch fused
    n = 125_080_192
    p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    p16 = torch.empty(n, dtype=torch.bfloat16, device=dev)
    ns = torch.ones(1, dtype=torch.float64, device=dev)
    adam = {"n": n}
    adam["clip_adam_ours_ms"] = timed(lambda: ops.clip_adam(p, g, m, v, p16, ns, 10.0, 1.0, 5e-3, 0.9, 0.999, 1e-8, 1e-6, 3), reps=8)
    pp = torch.nn.Parameter(p.clone())
    pp.grad = g.clone()
    opt = torch.optim.Adam([pp], lr=5e-3, weight_decay=1e-6, fused=True)

    def torch_adam():
        torch.nn.utils.clip_grad_norm_([pp], 10.0)
        opt.step()
    adam["torch_clip+fused_adam_ms"] = timed(torch_adam, reps=8)
    out["adam"] = adam
    print(json.dumps(adam))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "library_baselines.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
