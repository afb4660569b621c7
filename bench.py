"""bench.py -- train cells/sec of the CMMVAE training step on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 2|3|4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the single-GPU case the metric is quoted on): single-species core VAE
(human expert only, 60 530 genes, 1024-512 | 512-256-Z128 | 128-256-512 | 512-1024-G), 1024 cells per
GPU per step, synthetic CSR at 5 % density, bf16 tensor-core GEMMs / fp32 elsewhere.  Weak scaling:
every rank trains its own 1024-cell batch, gradients are all-reduced (mean) over NCCL.

One JSON line on rank 0:
  value     whole-job cells/s with the batch already resident in HBM (CUDA-event timed, max over ranks)
  e2e       the same metric through the public API (CMMVAEModel.training_step on a torch.sparse_csr batch
            built from PINNED HOST arrays each step; H2D copy and the D2H read of the loss scalars inside
            the timed region)
  roofline  the dominant kernel (fused decoder GEMM + ReLU + sum-MSE epilogue): algorithmic FLOPs / its
            mean launch duration (CUDA events on the launch stream) vs MEASURED_PEAKS.json bf16 peak
  cpu_baseline  the oracle port of the reference's CPU training step on this box's host cores
``--impl reference`` times that CPU path alone (rank 0) and prints the same line with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DENSITY = 0.05


class Dims:
    """model dimensions of a BASELINE config: gene panels, expert hidden sizes, VAE hidden size, latent size"""

    def __init__(self, config: int):
        if config == 4:   # configV2.yaml:16-84: wider VAE, latent 256, its own gene panels, 8192 cells per step
            self.G_HUMAN, self.G_MOUSE, self.H1, self.H2, self.HV, self.Z = 60664, 52417, 1024, 768, 512, 256
        else:             # config.yaml:34-105 / human_only.yaml:28-102
            self.G_HUMAN, self.G_MOUSE, self.H1, self.H2, self.HV, self.Z = 60530, 52437, 1024, 512, 256, 128
        self.species = {"human": self.G_HUMAN} if config == 2 else {"human": self.G_HUMAN, "mouse": self.G_MOUSE}
METRIC = "train cells/sec (fwd+bwd+ELBO)"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synth_batches(n, B, G, density, seed):
    from oracle.cmmvae_oracle import synth_csr
    return [synth_csr(B, G, density, seed + i) for i in range(n)]


def build_model(config: int):
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Adversarial, Expert, Experts, FCBlockConfig, KLAnnealingFn
    import pandas as pd
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    d = Dims(config)
    species, H1, H2, HV, Z = d.species, d.H1, d.H2, d.HV, d.Z
    experts = Experts([Expert(s, FCBlockConfig([g, H1, H2], dropout_rate=0.1, use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([H2, H1, g], activation_fn=relu)) for s, g in species.items()])
    vae = CLVAE(FCBlockConfig([H2, HV], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([Z, HV, H2], activation_fn=relu), latent_dim=Z, hidden_z=(config == 3))
    advs, conds = [], {}
    if config == 3:
        conds = {"assay": 8, "dataset_id": 272}
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "human"))
        for c, n in conds.items():
            pd.DataFrame([f"{c}_{i}" for i in range(n)]).to_csv(
                os.path.join(tmp, "human", f"unique_expression_{c}.csv"), header=False, index=False)
        Adversarial.labels.clear()
        advs = [Adversarial(FCBlockConfig([HV, 128, 64], activation_fn=relu), FCBlockConfig([64]), list(conds), tmp),
                Adversarial(FCBlockConfig([Z, 64], activation_fn=relu), FCBlockConfig([64]), list(conds), tmp)]
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    model = CMMVAEModel(CMMVAE(vae, experts, advs), adv_weight=1.0,
                        autograd_config=AutogradConfig(clip(), clip(), clip()), kl_annealing_fn=KLAnnealingFn(1.0))
    return model, species, conds


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        clocks, reasons, mx = [], set(), None
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clocks.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if clocks:
            busy = [c for c in clocks if c >= 0.5 * max(clocks)]
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": mx, "reasons": sorted(reasons),
                   "samples": len(clocks)}
        return out


def cpu_reference_leg(config, B, steps, warmup, threads=None):
    """The reference's CPU training step (oracle port: same ATen calls -- sparse-CSR addmm, dense GEMMs,
    batch-norm, to_dense + mse, autograd backward, clip, Adam) on this box's host cores."""
    from oracle import cmmvae_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    O.FAST_CSR = True
    d = Dims(config)
    species, H1, H2, HV, Z = d.species, d.H1, d.H2, d.HV, d.Z
    spec = O.ModelSpec(
        experts={s: {"encoder": O.BlockSpec.make([g, H1, H2], bn=True), "decoder": O.BlockSpec.make([H2, H1, g])}
                 for s, g in species.items()},
        vae_encoder=O.BlockSpec.make([H2, HV], bn=True, return_hidden=True),
        vae_decoder=O.BlockSpec.make([Z, HV, H2]), latent_dim=Z)
    gen = torch.Generator().manual_seed(0)
    P = {}

    def lin(prefix, n_in, n_out):
        P[f"{prefix}.weight"] = torch.randn(n_out, n_in, generator=gen) * (2.0 / n_out) ** 0.5
        P[f"{prefix}.bias"] = torch.zeros(n_out)

    def bn(prefix, n):
        P[f"{prefix}.weight"], P[f"{prefix}.bias"] = torch.ones(n), torch.zeros(n)
        P[f"{prefix}.running_mean"], P[f"{prefix}.running_var"] = torch.zeros(n), torch.ones(n)
        P[f"{prefix}.num_batches_tracked"] = torch.tensor(0)

    for s, g in species.items():
        for i, (a, b) in enumerate(((g, H1), (H1, H2))):
            lin(f"experts.{s}.encoder.fc_layers.{i}.lin", a, b)
            bn(f"experts.{s}.encoder.fc_layers.{i}.bn", b)
        for i, (a, b) in enumerate(((H2, H1), (H1, g))):
            lin(f"experts.{s}.decoder.fc_layers.{i}.lin", a, b)
    lin("vae.encoder.fc.fc_layers.0.lin", H2, HV)
    bn("vae.encoder.fc.fc_layers.0.bn", HV)
    lin("vae.encoder.mean_encoder", HV, Z)
    lin("vae.encoder.var_encoder", HV, Z)
    lin("vae.decoder.fc_layers.0.lin", Z, HV)
    lin("vae.decoder.fc_layers.1.lin", HV, H2)
    opt = {}
    names = list(species)
    batches = {s: synth_batches(2, B, g, DENSITY, 7000) for s, g in species.items()}
    times = []
    for t in range(warmup + steps):
        s = names[t % len(names)]
        crow, col, val = batches[s][t % 2]
        eps = torch.randn(B, Z, generator=gen)
        t0 = time.perf_counter()
        O.train_step(spec, P, opt, s, crow, col, val, eps, 1.0, return_grads=False)
        dt = time.perf_counter() - t0
        if t >= warmup:
            times.append(dt)
    O.FAST_CSR = False
    sec = statistics.median(times)
    return {"value": B / sec, "unit": "cells/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps (median) after {warmup} warm-up of the same workload (B={B}/step), "
                      f"oracle port of the reference CPU step, torch {torch.__version__} fp32, {threads} threads",
            "ms_per_step": sec * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-diag", action="store_true", help="extra legs that split the end-to-end time")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = Dims(args.config)
    G_HUMAN, G_MOUSE, H1, H2, HV, Z = d.G_HUMAN, d.G_MOUSE, d.H1, d.H2, d.HV, d.Z
    B = args.batch or {2: 1024, 3: 4096, 4: 8192}[args.config]
    workload = (f"config{args.config}: " + {2: "single-species core VAE (human expert only)",
                                            3: "two-species CMMVAE + 2 GRL adversaries",
                                            4: f"two-species CMMVAE, latent {Z} ({H1}-{H2}|{H2}-{HV}-Z{Z}), "
                                               "no output discriminator"}[args.config] +
                f", {B} cells/GPU/step, G={G_HUMAN}" + (f"/{G_MOUSE}" if args.config != 2 else "") +
                f", CSR {DENSITY:.0%} nnz, {args.precision}")
    config = {"workload": workload, "batch_per_gpu": B, "global_batch": B * world, "genes": G_HUMAN,
              "density": DENSITY, "parallelism": f"dp{world}",
              "l2": "inputs larger than L2: 1.0 GB of fp32 weights+grads and 2.5 GB of optimizer state are "
                    "streamed every step; 4 rotating batches"}

    if args.impl == "reference":
        if rank != 0:
            return
        cpu_B = min(B, 1024)
        leg = cpu_reference_leg(args.config, cpu_B, max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "cells/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": leg,
                "e2e": {"value": leg["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))   # one launch thread per rank matters
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")   # keep stdout clean: the only stdout line is the JSON result
        torch.distributed.init_process_group("nccl", device_id=dev)
    from mmvae_b200 import layers as L, ops
    import pandas as pd
    L.set_precision(args.precision)
    model, species, conds = build_model(args.config)
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    eng.pipeline_optimizer = True   # single GPU: output-layer clip+Adam runs underneath the next forward pass
    names = list(species)
    NB = 4
    host = {s: synth_batches(NB, B, g, DENSITY, 1000 * (rank + 1)) for s, g in species.items()}
    resident = {s: [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host[s]] for s in names}
    rng = np.random.default_rng(rank)
    metas = [pd.DataFrame({c: [f"{c}_{i}" for i in rng.integers(0, n, size=B)] for c, n in conds.items()})
             for _ in range(NB)]
    labels = [{c: torch.tensor([int(v.split("_")[-1]) for v in m[c]], dtype=torch.int64, device=dev) for c in conds}
              for m in metas] if conds else [None] * NB

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_step(t):
        s = names[t % len(names)]
        crow, col, val = resident[s][t % NB]
        eng.train_step(s, crow, col, val, int(col.numel()), 1.0, labels=labels[t % NB])

    # ---------------- device-resident throughput ----------------
    for t in range(args.warmup):
        resident_step(t)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # inside the timed region only the two kernels the roofline lines are computed from carry CUDA events (an
    # event record between two kernels ends a programmatic-dependent-launch chain); the other sections are
    # timed in a short separate pass afterwards
    eng.timers, eng.timer_filter = {}, {"decoder_mse_fused", "csr_linear_fwd"}
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(args.steps):
        resident_step(args.warmup + t)
    eng.finish()      # the last step's background optimizer work belongs to the timed region
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    ms = e0.elapsed_time(e1)
    t_dec, t_spmm = eng.timer_ms("decoder_mse_fused"), eng.timer_ms("csr_linear_fwd")
    eng.timers, eng.timer_filter = {}, None
    for t in range(min(args.steps, 20)):
        resident_step(args.warmup + args.steps + t)
    barrier()
    t_dw, t_dh, t_adam = eng.timer_ms("dWout_gemm"), eng.timer_ms("dh_gemm"), eng.timer_ms("norm+clip_adam")
    t_spbw = eng.timer_ms("csr_linear_bwd_w+bn")
    t_dp = {k: eng.timer_ms(k) for k in ("csr_prep", "mid_fwd", "mid_bwd", "dp_wait_shadow_first",
                                         "dp_wait_shadow_rest", "dp_wait_grads")}
    eng.timers = None
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API, host buffers ----------------
    # host side of a step: the batch's three CSR arrays (numpy, as the reference's batcher emits them) go through
    # mmvae_b200.feed.CSRStager -- packed into one pinned block, one async H2D copy on the copy stream
    from mmvae_b200.feed import CSRStager
    max_nnz = max(int(b[1].size) for s in names for b in host[s])
    stager = CSRStager(max_cells=B, max_nnz=max_nnz, device=dev, depth=NB * len(names))
    # the (synthetic) batcher has written each rotating batch into its pinned block once; a step ships its block
    blocks = {}
    for s in names:
        for i, (crow, col, val) in enumerate(host[s]):
            blk = stager.reserve(B, int(col.size))
            blk.crow[:], blk.col[:], blk.val[:] = crow, col, val
            blocks[(s, i)] = blk

    def stage(t):
        s = names[t % len(names)]
        return s, stager.commit(blocks[(s, t % NB)], species[s])

    def api_step(t, staged):
        s, ticket = staged
        x = stager.get(ticket)
        # every step ends with a D2H copy of its scalar block (loss, KL, norms) into pinned memory; with
        # sync_logging off the host reads it one step later, so the copy never stalls the launch queue
        model.training_step((x, metas[t % NB], s), t)
        stager.release(ticket)
        v = model.logged_metrics.get(f"loss/training/{s}")
        return float(v) if v is not None else None

    model.sync_logging = False
    nxt = stage(0)
    for t in range(max(args.warmup, 3)):
        cur, nxt = nxt, stage(t + 1)
        api_step(t, cur)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    loss = None
    for t in range(args.steps):
        cur, nxt = nxt, stage(t + 1 + max(args.warmup, 3))
        loss = api_step(t, cur)
    model.flush_logs()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None
    diag = None
    if args.e2e_diag:
        # where does the end-to-end leg lose time?  (a) the same API calls on batches already in HBM (no H2D);
        # (b) the H2D copies alone (no compute)
        def api_resident(t):
            s = names[t % len(names)]
            crow, col, val = resident[s][t % NB]
            model.training_step((torch.sparse_csr_tensor(crow, col, val, size=(B, species[s])), metas[t % NB], s), t)
        for t in range(3):
            api_resident(t)
        barrier()
        e0.record()
        for t in range(args.steps):
            api_resident(t)
        model.flush_logs()
        e1.record()
        barrier()
        a_ms = e0.elapsed_time(e1) / args.steps
        barrier()
        e0.record()
        for t in range(args.steps):
            _, tk = stage(t)
            stager.arrays(tk)
            stager.release(tk)
        e1.record()
        barrier()
        c_ms = e0.elapsed_time(e1) / args.steps
        tt = torch.tensor([a_ms, c_ms], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        diag = {"api_resident_ms_per_step": float(tt[0]), "h2d_only_ms_per_step": float(tt[1]),
                "h2d_only_gbs_per_gpu": stager.bytes_staged / (float(tt[1]) * 1e-3) / 1e9}
    nnz = int(host[names[0]][0][1].size)
    h2d = int(stager.bytes_staged)
    d2h = int(eng.last["sc"].numel()) * 8

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    pk, pk_kind = peaks()
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape (ncu --set full)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_topk_dram_bytes.json")))
        traffic = tj.get("decoder_mse_fused_kernel") if (B == 1024 and args.config == 2) else None
    except Exception:  # noqa: BLE001
        pass
    flops = 2.0 * B * G_HUMAN * H1
    ach = flops / (t_dec * 1e-3) / 1e12 if t_dec > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    spmm_bytes = nnz * 8 + (B + 1) * 4 + G_HUMAN * H1 * 2 + B * H1 * 4 + H1 * 4
    line = {
        "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": B * world / (e2e_ms / args.steps * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "last_loss": loss},
        "gpu_launches": int(launches),
        "roofline": {"kernel": "decoder_mse_fused_kernel (K5-K7: tcgen05 GEMM + ReLU + sum-MSE-vs-CSR epilogue)",
                     "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                     "traffic": traffic, "peak_source": f"{pk_kind} bf16_tflops_sustained (kernel timed inside a step)",
                     "ms_per_launch": t_dec, "flops_per_launch": flops},
        "kernels_ms": {"decoder_mse_fused": t_dec, "dWout_gemm": t_dw, "dh_gemm": t_dh, "csr_linear_fwd": t_spmm,
                       "csr_linear_bwd_w+bn_bwd": t_spbw, "norm+clip_adam": t_adam, **t_dp},
        "spmm": {"algorithmic_bytes": spmm_bytes, "ms": t_spmm,
                 "achieved_gbs": spmm_bytes / (t_spmm * 1e-3) / 1e9 if t_spmm > 0 else 0.0,
                 "peak_gbs": pk["hbm_gbs"], "frac": (spmm_bytes / (t_spmm * 1e-3) / 1e9 / pk["hbm_gbs"]) if t_spmm > 0 else 0.0,
                 "fma_tflops": 2.0 * nnz * H1 / (t_spmm * 1e-3) / 1e12 if t_spmm > 0 else 0.0},
        "clocks": clocks,
    }
    if diag:
        line["e2e_diag"] = diag
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_leg(args.config, min(B, 1024), 2, 1)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
