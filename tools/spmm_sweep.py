"""BASELINE config 5: density x batch sweep of the expert-encoder SpMM (forward + weight gradient), with the roof
that binds every point named next to it.

  python tools/spmm_sweep.py                 one GPU, uniform columns + a Zipf-column variant
  python tools/spmm_sweep.py --shard N       one GPU, the kernel shapes a rank runs under the gene-sharded
                                             data-parallel route at N GPUs (N*B cells x G/N genes), local routes
  torchrun --nproc-per-node N tools/spmm_sweep.py --dp     the routed forward kernel on N GPUs, partial sums stored
                                             into the owners' buffers over NVLink (csrc/peer.cu), max over ranks

Roofs per point (MEASURED_PEAKS.json: HBM copy GB/s, sustained bf16 TFLOP/s):
  t_hbm    = algorithmic bytes / HBM            (fwd bytes = 8 nnz + 4(B+1) + 2 G H + 4 B H; bwd = 8 nnz + 2 B H + 4 G H)
  t_tensor = 2 B G H / tensor peak              what the densified tile product ISSUES (zeros are multiplied)
  t_fma    = 2 nnz H / 75 TFLOP/s               CUDA-core FMA roof of a gather kernel
The tensor family is bound by max(t_hbm, t_tensor), the gather family by max(t_hbm, t_fma) and in practice by L2
(every non-zero pulls a 2 KB weight row through L2).  `frac` = t_roof / t_measured for the family's own roof;
`hbm frac` = t_hbm / t_measured (the BASELINE metric).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmvae_b200 import ops  # noqa: E402
from mmvae_b200.synth import synth_csr  # noqa: E402

G, H = 60530, 1024
try:
    PK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:  # noqa: BLE001
    PK = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
HBM, TENSOR, FMA = PK["hbm_gbs"] * 1e9, PK["bf16_tflops_sustained"] * 1e12, 75e12


def timed(fn, flush, n=6):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def gpu_csr(B, d, seed, g_lo=0, g_hi=G):
    """Bernoulli(d) sparsity per (cell, gene), generated on the GPU (fast at 100 K cells); optionally only the
    columns [g_lo, g_hi), rebased"""
    gen = torch.Generator(device="cuda").manual_seed(seed)
    crow = [torch.zeros(1, dtype=torch.int32, device="cuda")]
    cols, n = [], 0
    for r0 in range(0, B, 1024):
        m = torch.rand(min(1024, B - r0), G, device="cuda", generator=gen)[:, g_lo:g_hi] < d
        cnt = m.sum(1)
        crow.append((n + cnt.cumsum(0)).to(torch.int32))
        n += int(cnt.sum())
        cols.append(m.nonzero()[:, 1].to(torch.int32))
    col = torch.cat(cols)
    val = torch.rand(col.numel(), device="cuda", generator=gen) * 6 + 0.5
    return torch.cat(crow), col, val


def point(B, d, zipf, flush, Wt16, shard=1):
    """one sweep point; shard > 1: the per-rank shape of the gene-sharded route (shard * B cells x G / shard genes)"""
    per = -(-G // shard)
    per = (per + 127) // 128 * 128 if shard > 1 else G
    Bc = B * shard
    if zipf:
        crow, col, val = (torch.from_numpy(a).cuda() for a in synth_csr(Bc, G, d, seed=7, zipf=zipf))
    else:      # shard > 1: only the entries of gene shard 0 (what a rank receives from the all-to-all)
        crow, col, val = gpu_csr(Bc, d, 7, 0, min(per, G))
    nnz = int(col.numel())
    Gs = per
    W = Wt16[:Gs]
    bias = torch.zeros(H, device="cuda")
    Y = torch.empty(Bc, H, device="cuda")
    dY16 = torch.randn(Bc, H, device="cuda").bfloat16()
    dWt = torch.empty(Gs, H, device="cuda")
    fwd_bytes = nnz * 8 + (Bc + 1) * 4 + Gs * H * 2 + Bc * H * 4
    bwd_bytes = nnz * 8 + Bc * H * 2 + Gs * H * 4
    t_hbm_f, t_hbm_b = fwd_bytes / HBM, bwd_bytes / HBM
    t_tensor = 2.0 * Bc * Gs * H / TENSOR
    t_fma = 2.0 * nnz * H / FMA
    tp, packed = ops.csr_tile_ptr(crow, col, val, Gs, nnz)
    r = dict(B=B, d=d, zipf=zipf, shard=shard, nnz=nnz)
    r["prep"] = timed(lambda: ops.csr_tile_ptr(crow, col, val, Gs, nnz, tp, packed), flush)
    r["fwd_tc"] = timed(lambda: ops.csr_linear_fwd_tc(packed, tp, Bc, Gs, W, bias, out=Y), flush)
    r["bwd_tc"] = timed(lambda: ops.csr_linear_bwd_w_tc(packed, tp, Bc, Gs, dY16, dWt), flush)
    if shard == 1 and nnz <= 30_000_000:
        r["fwd_gather"] = timed(lambda: ops.csr_linear_fwd(crow, col, val, Gs, W, bias, out=Y), flush, n=4)
    from mmvae_b200.engine import StepEngine
    pick = "tensor" if (shard > 1 or StepEngine.spmm_picks_tensor(nnz, Bc, Gs, H)) else "gather"
    t_f = r["fwd_tc"] if pick == "tensor" else r.get("fwd_gather", r["fwd_tc"])
    roof_f = max(t_hbm_f, t_tensor) if pick == "tensor" else max(t_hbm_f, t_fma)
    r.update(pick=pick, t_hbm_f=t_hbm_f, t_tensor=t_tensor, t_fma=t_fma,
             bound_f=("tensor issue" if t_tensor >= t_hbm_f else "HBM") if pick == "tensor" else
                     ("FMA / L2 gather" if t_fma >= t_hbm_f else "HBM (L2 gather in practice)"),
             frac_f=roof_f / t_f, hbm_frac_f=t_hbm_f / t_f, gbs_f=fwd_bytes / t_f / 1e9,
             fma_tflops=2.0 * nnz * H / t_f / 1e12, tensor_frac_f=t_tensor / r["fwd_tc"],
             bound_b="tensor issue" if t_tensor >= t_hbm_b else "HBM", frac_b=max(t_hbm_b, t_tensor) / r["bwd_tc"],
             hbm_frac_b=t_hbm_b / r["bwd_tc"],
             other_family_faster=("fwd_gather" in r and ((pick == "tensor") != (r["fwd_tc"] + r["prep"] <= r["fwd_gather"]))))
    print(json.dumps(r), flush=True)
    return r


def write_table(rows, path, title):
    with open(path, "w") as f:
        f.write(f"# {title}\n\nG={G}, H={H}, bf16 weight; CUDA events, L2 flushed between launches, median of 6. "
                f"Peaks: HBM {HBM / 1e9:.0f} GB/s, tensor {TENSOR / 1e12:.0f} TFLOP/s sustained (MEASURED_PEAKS.json), "
                "CUDA-core FMA 75 TFLOP/s.  `bound` = the roof that binds the picked family at this point; `frac` = "
                "that roof / measured; `HBM frac` = algorithmic bytes at HBM speed / measured.\n\n"
                "| B | d | cols | shard | nnz | pick | fwd ms | prep ms | other fwd ms | bound (fwd) | frac | HBM frac | GB/s | "
                "useful FMA TFLOP/s | issued / tensor peak | bwd ms | bound (bwd) | frac | HBM frac |\n" + "|---" * 19 + "|\n")
        for r in rows:
            other = r.get("fwd_gather") if r["pick"] == "tensor" else r["fwd_tc"]
            t_f = r["fwd_tc"] if r["pick"] == "tensor" else r.get("fwd_gather", r["fwd_tc"])
            f.write(f"| {r['B']} | {r['d']:.0%} | {'zipf ' + str(r['zipf']) if r['zipf'] else 'uniform'} | {r['shard']} | "
                    f"{r['nnz']} | {r['pick']} | {t_f * 1e3:.3f} | {r['prep'] * 1e3:.3f} | "
                    f"{(other * 1e3 if other else float('nan')):.3f} | {r['bound_f']} | {r['frac_f']:.2f} | {r['hbm_frac_f']:.3f} | "
                    f"{r['gbs_f']:.0f} | {r['fma_tflops']:.1f} | {r['tensor_frac_f']:.2f} | {r['bwd_tc'] * 1e3:.3f} | "
                    f"{r['bound_b']} | {r['frac_b']:.2f} | {r['hbm_frac_b']:.3f} |\n")
        bad = [r for r in rows if r.get("other_family_faster")]
        f.write(f"\nPoints where the other kernel family would have been faster than the pick: {len(bad)}"
                + ("".join(f"\n- B={r['B']} d={r['d']:.0%}: tensor {(r['fwd_tc'] + r['prep']) * 1e3:.3f} ms (incl. prep) "
                           f"vs gather {r['fwd_gather'] * 1e3:.3f} ms" for r in bad)) + "\n")


def dp_sweep(args):
    """routed forward kernel on N GPUs: every rank multiplies (N * B cells) x (its G / N genes) and stores the partial
    sums into the owners' buffers over NVLink; time = max over ranks of the kernel + flag round trip"""
    import torch.distributed as dist
    from mmvae_b200.peer import PeerComm
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    comm = PeerComm(torch.device("cuda"))
    per = (-(-G // world) + 127) // 128 * 128
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    W = (torch.randn(per, H, device="cuda") * 0.03).bfloat16()
    rows = []
    step = 0
    for B in args.batches:
        # as in the step engine (_dp_setup): with few cell tiles the product is also cut along the genes, one slab each
        units = ((world * B + 127) // 128) * ((H + 255) // 256)
        S = max(1, min(4, 148 // units))
        Yin = comm.alloc(f"Y{B}", world * S * B * H * 4)
        for d in args.densities:
            crow, col, val = gpu_csr(world * B, d, 7, rank * per, min(G, (rank + 1) * per))
            nnz = int(col.numel())
            tp, packed = ops.csr_tile_ptr(crow, col, val, per, nnz)
            route = [p + rank * S * B * H * 4 for p in Yin.ptr]

            def run():
                nonlocal step
                step += 1
                ops.csr_linear_fwd_tc_routed(packed, tp, world * B, per, W, route, B, S, B * H)
                ops.peer_signal(comm.flag_ptrs("Y"), step)
                ops.peer_wait(comm.local_flags("Y"), world, step)
            dist.barrier()
            t = timed(run, flush)
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt)
            fwd_bytes = nnz * 8 + (world * B + 1) * 4 + per * H * 2 + world * B * H * 4
            t_tensor = 2.0 * world * B * per * H / TENSOR
            rows.append(dict(B=B, d=d, world=world, S=S, nnz=nnz, ms=t * 1e3, tensor_frac=t_tensor / t,
                             hbm_frac=fwd_bytes / HBM / t, nvlink_gbs=(world - 1) * S * B * H * 4 / t / 1e9))
            if rank == 0:
                print(json.dumps(rows[-1]), flush=True)
    if rank == 0:
        out = os.path.join(ROOT, "gpurun_out", f"spmm_sweep_dp{world}.md")
        with open(out, "w") as f:
            f.write(f"# Routed first-layer product on {world} GPUs (gene-sharded data-parallel route)\n\nPer rank: {world} x B cells "
                    f"x {per} genes, H={H}; partial sums stored into the owners' buffers over NVLink by the epilogue, then "
                    "flag round trip; max over ranks, median of 6, L2 flushed (cold weights: inside a step the kernel runs "
                    "~0.1 ms faster, `kernels_ms.csr_linear_fwd` of bench.py).  `S` = gene pieces per rank (own slab each) "
                    "when there are fewer cell tiles than SMs.\n\n| B per rank | d | S | nnz in shard | ms | "
                    "issued / tensor peak | HBM frac | NVLink out GB/s per rank |\n|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r['B']} | {r['d']:.0%} | {r['S']} | {r['nnz']} | {r['ms']:.3f} | {r['tensor_frac']:.2f} | "
                        f"{r['hbm_frac']:.3f} | {r['nvlink_gbs']:.0f} |\n")
        print("wrote", out)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard", type=int, default=1)
    ap.add_argument("--dp", action="store_true")
    ap.add_argument("--batches", type=int, nargs="+", default=[512, 1024, 4096, 16384])
    ap.add_argument("--densities", type=float, nargs="+", default=[0.01, 0.02, 0.05, 0.10, 0.20])
    ap.add_argument("--no-zipf", action="store_true")
    args = ap.parse_args()
    if args.dp:
        return dp_sweep(args)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    Wt16 = (torch.randn(G, H, device="cuda") * 0.03).bfloat16()
    rows = []
    for B in args.batches:
        for d in args.densities:
            if B * args.shard * d * G > 2.2e8:
                continue
            rows.append(point(B, d, 0.0, flush, Wt16, args.shard))
    if not args.no_zipf and args.shard == 1:
        for B in (512, 4096):
            for d in (0.01, 0.05):
                rows.append(point(B, d, 1.0, flush, Wt16, 1))
    name = "spmm_sweep.md" if args.shard == 1 else f"spmm_sweep_shard{args.shard}.md"
    title = ("Expert-encoder SpMM sweep (BASELINE config 5), 1x B200" if args.shard == 1 else
             f"Expert-encoder SpMM sweep at the per-rank shape of the gene-sharded route, {args.shard} GPUs "
             f"({args.shard} x B cells x G/{args.shard} genes), measured on one GPU with local routes")
    out = os.path.join(ROOT, "gpurun_out", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    write_table(rows, out, title)
    json.dump(rows, open(out.replace(".md", ".json"), "w"))
    print("wrote", out)


if __name__ == "__main__":
    main()
