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
  e2e       the same metric through the public API: the reference batcher's iteration over PAGEABLE scipy CSR
            chunks (mmvae_b200.feed.StagedCSRBatches: packing on worker threads, uint16 gene ids, one pinned
            block + one async H2D copy per batch) -> CMMVAEModel.training_step -> D2H read of the loss scalars,
            all inside the timed region
  roofline  the dominant kernel (fused decoder GEMM + ReLU + sum-MSE epilogue): algorithmic FLOPs / its
            mean launch duration (CUDA events on the launch stream) vs MEASURED_PEAKS.json bf16 peak
  parity_check  one extra step from the trained weights against one oracle step (outside the timed regions);
            the bench prints no line if they disagree
  torch_cuda_baseline  the UNMODIFIED reference (baseline/_ref) as stock PyTorch eager on this GPU, fp32 and
            bf16 autocast: the on-box library path the kernels have to beat
  cpu_baseline  the unmodified reference's CPU training step on this box's host cores (bounded sample)
  also.config3  (N > 1) BASELINE configs[2] -- two species + GRL adversaries, 4096 cells/GPU -- at the same N
``--impl reference`` times the unmodified reference on the host cores alone (rank 0) and prints the same line
with impl=reference (oracle port only if baseline/_ref is absent).
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


def synth_csr(B, G, density, seed, zipf=0.0):
    from mmvae_b200.synth import synth_csr as gen
    return gen(B, G, density, seed, zipf)


def synth_batches(n, B, G, density, seed):
    return [synth_csr(B, G, density, seed + i) for i in range(n)]


def build_model(config: int, only=None, discriminators: bool = True):
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Adversarial, Expert, Experts, FCBlockConfig, KLAnnealingFn
    import pandas as pd
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    d = Dims(config)
    species = {s: g for s, g in d.species.items() if only is None or s in only}
    H1, H2, HV, Z = d.H1, d.H2, d.HV, d.Z
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
    extra = {}
    if config == 4 and discriminators:      # BASELINE configs[3]: "+ output discriminator" (one per species, trained inside the step)
        from mmvae_b200.modules import create_discriminators
        extra["output_discriminators"] = create_discriminators(species)
    model = CMMVAEModel(CMMVAE(vae, experts, advs), adv_weight=1.0,
                        autograd_config=AutogradConfig(clip(), clip(), clip()), kl_annealing_fn=KLAnnealingFn(1.0),
                        **extra)
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


def oracle_spec(config, species, dropout=0.0, conds=None):
    from oracle import cmmvae_oracle as O
    d = Dims(config)
    H1, H2, HV, Z = d.H1, d.H2, d.HV, d.Z
    advs = []
    if conds:
        advs = [O.AdversarySpec(O.BlockSpec.make([HV, 128, 64]), dict(conds)),
                O.AdversarySpec(O.BlockSpec.make([Z, 64]), dict(conds))]
    return O.ModelSpec(
        experts={s: {"encoder": O.BlockSpec.make([g, H1, H2], bn=True, dropout=dropout),
                     "decoder": O.BlockSpec.make([H2, H1, g])} for s, g in species.items()},
        vae_encoder=O.BlockSpec.make([H2, HV], bn=True, return_hidden=True),
        vae_decoder=O.BlockSpec.make([Z, HV, H2]), latent_dim=Z, hidden_z=bool(conds), adversarials=advs)


def cpu_port_leg(config, B, steps, warmup, threads=None, budget_s=170.0):
    """The reference's CPU training step restated by the oracle (same ATen calls -- sparse-CSR addmm, dense
    GEMMs, batch-norm, to_dense + mse, autograd backward, clip, Adam) on this box's host cores.  Only used when
    the unmodified reference (baseline/_ref) is not installed."""
    from oracle import cmmvae_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    O.FAST_CSR = True
    d = Dims(config)
    species, H1, H2, HV, Z = d.species, d.H1, d.H2, d.HV, d.Z
    spec = oracle_spec(config, species)
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
    times, spent = [], 0.0
    for t in range(warmup + steps):
        s = names[t % len(names)]
        crow, col, val = batches[s][t % 2]
        eps = torch.randn(B, Z, generator=gen)
        t0 = time.perf_counter()
        O.train_step(spec, P, opt, s, crow, col, val, eps, 1.0, return_grads=False)
        dt = time.perf_counter() - t0
        if t >= warmup:
            times.append(dt)
            spent += dt
            if spent > budget_s:
                break
    O.FAST_CSR = False
    sec = statistics.median(times)
    return {"value": B / sec, "unit": "cells/s", "cores": threads, "kind": "port", "steps_run": len(times),
            "sample": f"{len(times)} steps (median) after {warmup} warm-up of the same workload (B={B}/step), "
                      f"oracle port of the reference CPU step, torch {torch.__version__} fp32, {threads} threads",
            "ms_per_step": sec * 1e3}


def cpu_reference_leg(config, B, steps, warmup, threads=None, budget_s=170.0):
    """The reference's own CPU training step on this box's host cores: the UNMODIFIED reference package
    (baseline/_ref, installed by tools/install_reference.sh) driven through CMMVAEModel.training_step
    (kind "reference"); the oracle port (kind "port") only if that install is absent."""
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    from baseline import ref_arm
    if not ref_arm.available():
        return cpu_port_leg(config, B, steps, warmup, threads, budget_s)
    r = ref_arm.time_reference_steps(Dims(config), config, B, steps, warmup, device="cpu", budget_s=budget_s,
                                     density=DENSITY)
    return {"value": r["value"], "unit": "cells/s", "cores": threads, "kind": "reference",
            "steps_run": r["steps_run"], "ms_per_step": r["ms_per_step"], "last_loss": r["last_loss"],
            "sample": f"{r['steps_run']} steps (median) after {warmup} warm-up of the same workload (B={B}/step): "
                      f"unmodified reference CMMVAEModel.training_step (baseline/_ref) on torch.sparse_csr batches, "
                      f"torch {torch.__version__} fp32 CPU, {threads} threads"}


def torch_cuda_baseline(config, B, dev):
    """SURVEY.md 8d / BASELINE.md 4.4: the reference's own GPU path on this box -- the unmodified reference
    modules as stock PyTorch eager on ``cuda`` (cuSPARSE addmm, cuBLASLt GEMMs, to_dense + mse_loss, autograd,
    clip_grad_norm_, torch.optim.Adam), fp32 and bf16 autocast.  The number every "beats" claim stands next to."""
    from baseline import ref_arm
    if not ref_arm.available():
        return {"unavailable": "baseline/_ref not installed (tools/install_reference.sh)"}
    out = {"how": "unmodified reference CMMVAEModel.training_step, stock PyTorch eager on cuda, "
                  f"torch {torch.__version__}, B={B}, median of 8 steps after 3 warm-up, host-synchronised per step"}
    for name, kw in (("fp32", {}), ("bf16_autocast", {"autocast": True})):
        for dense in (False, True):
            try:
                r = ref_arm.time_reference_steps(Dims(config), config, B, 8, 3, device=str(dev), density=DENSITY,
                                                 dense_input=dense, **kw)
                out[name] = {k: r[k] for k in ("value", "ms_per_step", "last_loss", "input")}
                break
            except Exception as e:  # noqa: BLE001   (e.g. sparse-CSR addmm backward unsupported on cuda)
                out[name] = {"error": f"{type(e).__name__}: {str(e)[:160]}", "input": "dense" if dense else "sparse_csr"}
            finally:
                torch.cuda.empty_cache()
    return out


def parity_check(model, eng, config, species, conds, host, metas_src, B, rank, world, dev):
    """One extra training step outside every timed region, from the weights the benchmark has trained, on rank 0's
    batch 0 with injected noise and dropout masks -- against ONE oracle step on the same state and inputs (rank 0).
    The bench refuses to print a line whose arithmetic disagrees with the oracle."""
    from mmvae_b200 import layers as L
    model.flush_logs()
    eng.finish()
    s = list(species)[0]
    G = species[s]
    d = Dims(config)
    P = {k[len("module."):]: v.detach().cpu().clone() for k, v in model.state_dict().items()}   # collective when world > 1
    crow, col, val = host[s][0]
    gen = torch.Generator().manual_seed(99 + rank)
    eps = torch.randn(B, d.Z, generator=gen)
    keep = [(torch.rand(B, n, generator=gen) >= 0.1) for n in (d.H1, d.H2)]
    L.inject_noise(eps.to(dev))
    L.inject_dropout_masks({f"experts.{s}.encoder.fc_layers.{j}.dr": k.to(torch.uint8).to(dev)
                            for j, k in enumerate(keep)})
    meta = metas_src[0]
    x = torch.sparse_csr_tensor(*(torch.from_numpy(a).to(dev) for a in (crow, col, val)), size=(B, G))
    was = model.sync_logging
    model.sync_logging = True
    model.logged_metrics.clear()
    model.training_step((x, meta.copy(), s), 0)
    model.sync_logging = was
    got = {k.split("/")[0]: float(v) for k, v in model.logged_metrics.items()
           if k.split("/")[0] in ("loss", "recon_loss", "kl_loss")}
    if rank != 0:
        return None
    from oracle import cmmvae_oracle as O
    O.FAST_CSR = True
    labels = {c: torch.tensor([int(v.split("_")[-1]) for v in meta[c]], dtype=torch.int64) for c in conds} or None
    ref = O.train_step(oracle_spec(config, species, dropout=0.1, conds=conds), P, {}, s, crow, col, val, eps, 1.0,
                       labels=labels, return_grads=False,
                       dropout_masks={f"experts.{s}.encoder.fc_layers.{j}.dr": k.float() for j, k in enumerate(keep)})
    O.FAST_CSR = False
    want = ref["logs"]
    err = {k: abs(got[k] - want[k]) / max(abs(want[k]), 1e-30) for k in got}
    tol = {"loss": 1e-3, "recon_loss": 1e-3, "kl_loss": 1e-2}
    ok = all(err[k] <= tol[k] for k in err)
    return {"ok": ok, "what": "one fused bf16 step vs one oracle step, same trained weights / batch / eps / dropout "
                              "masks, outside the timed region", "loss_gpu": got["loss"], "loss_oracle": want["loss"],
            "rel_err": err, "tol": tol}


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
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--also-config3", type=int, default=None,
                    help="extra leg: BASELINE configs[2] (two species + GRL adversaries, 4096 cells/GPU) at the same N; "
                         "default on when N > 1")
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
                                               "+ per-species output discriminator G-128-64-1 on the "
                                               "reconstruction"}[args.config] +
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
        leg = cpu_reference_leg(args.config, cpu_B, max(1, args.steps), max(0, args.warmup))
        config = dict(config, reference_sample=f"each step = the same workload at {cpu_B} cells/step on the host cores")
        line = {"impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "cells/s",
                "n_gpus": args.gpus, "steps": leg["steps_run"], "warmup": args.warmup,
                "steps_requested": args.steps, "ms_per_step": leg["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": leg,
                "e2e": {"value": leg["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    # a stalled run (e.g. a rank that died while its peers spin on its flags) ends itself instead of holding the GPUs:
    # every thread's stack is dumped and the process exits (the normal run takes 1-3 minutes)
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG", 1500)), exit=True)
    same_gpu = os.environ.get("BENCH_SAME_GPU") == "1"     # debugging aid: all ranks share GPU 0 (CUDA IPC, gloo)
    if same_gpu:
        local_rank = 0
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and same_gpu:
        torch.distributed.init_process_group("gloo")
    elif world > 1:
        torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))   # one launch thread per rank matters
        # NCCL's communicator lines (stderr) stay visible: the driver counts ranks from them
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        torch.distributed.init_process_group("nccl", device_id=dev)
        if rank == 0:
            print(f"[bench] NCCL communicator up: nranks={torch.distributed.get_world_size()}", file=sys.stderr)
    from mmvae_b200 import layers as L, ops
    import pandas as pd
    L.set_precision(args.precision)
    model, species, conds = build_model(args.config)
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    eng.pipeline_optimizer = True   # output-layer clip+Adam runs underneath the next forward pass
    # `value` leg: stream launches with programmatic dependent launch (fastest device-resident mode: 1.50 vs 1.55 ms
    # per step replayed from graphs); the `e2e` leg goes through CMMVAEModel.training_step, whose pipelined mode
    # replays the step from CUDA graphs (the host then never limits the step).  BENCH_GRAPH=1: graphs in both legs.
    eng.use_graph = os.environ.get("BENCH_GRAPH", "0") == "1"
    names = list(species)
    NB = 4
    host = {s: synth_batches(NB, B, g, DENSITY, 1000 * (rank + 1)) for s, g in species.items()}
    resident = {s: [tuple(torch.from_numpy(a).to(dev) for a in b) for b in host[s]] for s in names}
    rng = np.random.default_rng(rank)
    metas = [pd.DataFrame({c: [f"{c}_{i}" for i in rng.integers(0, n, size=B)] for c, n in conds.items()})
             if conds else pd.DataFrame({"cell": np.arange(B)}) for _ in range(NB)]
    labels = [{c: torch.tensor([int(v.split("_")[-1]) for v in m[c]], dtype=torch.int64, device=dev) for c in conds}
              for m in metas] if conds else [None] * NB

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_batch(t):
        s = names[t % len(names)]
        return (s,) + resident[s][t % NB]

    def resident_step(t):
        s, crow, col, val = resident_batch(t)
        eng.train_step(s, crow, col, val, int(col.numel()), 1.0, labels=labels[t % NB], nnz_cap=int(col.numel()))
        if world > 1:      # the exchange of the NEXT batch's CSR records runs underneath this step
            eng.prefetch(*resident_batch(t + 1), ready=True)

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
    graph_mode = eng.use_graph and world == 1
    if graph_mode:     # the two roofline kernels are timed by event nodes INSIDE the replayed graphs
        eng.graph_timers = {"decoder_mse_fused", "csr_linear_fwd"}
        for t in range(2 * NB * len(names)):          # visit every rotating batch twice: eager warm-up, then capture
            resident_step(args.warmup + t)
        barrier()
    else:
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
    if graph_mode:
        t_dec, t_spmm = eng.graph_timer_ms("decoder_mse_fused"), eng.graph_timer_ms("csr_linear_fwd")
        launches = None     # replayed launches do not pass through the C ABI: counted from the eager pass below
    else:
        t_dec, t_spmm = eng.timer_ms("decoder_mse_fused"), eng.timer_ms("csr_linear_fwd")
    eng.timers, eng.timer_filter = {}, None
    l1 = ops.launch_count()
    for t in range(min(args.steps, 20)):
        resident_step(args.warmup + args.steps + t)
    barrier()
    if launches is None:   # kernels per step (eager pass, same launch sequence the graphs replay) x timed steps
        launches = (ops.launch_count() - l1) // min(args.steps, 20) * args.steps
    t_dw, t_dh, t_adam = eng.timer_ms("dWout_gemm"), eng.timer_ms("dh_gemm"), eng.timer_ms("norm+clip_adam")
    t_spbw = eng.timer_ms("csr_linear_bwd_w+bn")
    t_dp = {k: eng.timer_ms(k) for k in ("csr_prep", "mid_fwd", "mid_bwd", "dp_wait_shadow_first",
                                         "dp_wait_shadow_rest", "dp_wait_grads", "dp_wait_Y", "dp_wait_h",
                                         "dp_wait_dh") if eng.timers.get(k)}
    eng.timers = None
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        ms = float(tt.item())
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---------------- end to end through the public API, host buffers ----------------
    # Host side of a step, all INSIDE the timed region: the reference's batcher semantics
    # (SparseCSRMatrixBatcherDataPipe, cellxgene_datapipe.py:169-193) over pageable scipy CSR chunks ->
    # mmvae_b200.feed.StagedCSRBatches packs each batch (native packer on worker threads, uint16 gene ids) into a
    # pinned block, one async H2D copy on the copy stream -> CMMVAEModel.training_step on the torch.sparse_csr
    # batch -> D2H read of the step's scalar block (one step late, so the copy never stalls the launch queue).
    import scipy.sparse as sp
    from mmvae_b200.feed import StagedCSRBatches
    workers = int(os.environ.get("BENCH_WORKERS", 0)) or max(1, min(6, (os.cpu_count() or 8) // world - 1))

    def chunk_source(s):
        chunk = sp.vstack([sp.csr_matrix((v, c, r), shape=(B, species[s])) for r, c, v in host[s]], format="csr")
        assert chunk.indices.dtype == np.int32 and chunk.data.dtype == np.float32 and chunk.has_sorted_indices
        frame = pd.concat(metas, ignore_index=True)
        while True:            # an endless epoch over the (pageable) chunk
            yield chunk, frame

    # (BENCH_MAIN_PINNED=1: diagnostic -- the headline leg itself on page-locked chunks)
    main_kw = dict(workers=0, ahead=int(os.environ.get("BENCH_PINNED_AHEAD", 4)), pin_chunks=True) \
        if os.environ.get("BENCH_MAIN_PINNED") == "1" else dict(workers=workers)
    feeds = {s: StagedCSRBatches(chunk_source(s), B, device=dev, **main_kw) for s in names}
    iters = {s: iter(f) for s, f in feeds.items()}

    pending = {}

    def fetch(t):
        s = names[t % len(names)]
        x, meta = next(iters[s])
        return x, meta, s

    host_t = {"fetch": 0.0, "training_step": 0.0}     # host wall time per phase (diagnostic: is a leg host-bound?)

    def api_step(t):
        t0 = time.perf_counter()
        batch = pending.pop(t, None) or fetch(t)
        t1 = time.perf_counter()
        model.training_step(batch, t)
        t2 = time.perf_counter()
        if world > 1:      # data parallel: hand the next batch over so its CSR exchange runs under this step
            pending[t + 1] = fetch(t + 1)
            model.prefetch_batch(pending[t + 1])
        host_t["fetch"] += t1 - t0 + (time.perf_counter() - t2)
        host_t["training_step"] += t2 - t1
        v = model.logged_metrics.get(f"loss/training/{batch[2]}")
        return float(v) if v is not None else None

    model.sync_logging = False
    # (every graph the pipelined mode replays is visited twice before the clock starts: eager warm-up, capture)
    w_api = max(args.warmup, 4 * len(names))
    for t in range(w_api):
        api_step(t)
    barrier()
    host_t.update(fetch=0.0, training_step=0.0)
    prof = None
    if os.environ.get("BENCH_PROFILE") == "1" and rank == 0:      # debugging aid: where does the host loop spend its time?
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    e0.record()
    loss = None
    for t in range(args.steps):
        loss = api_step(t + w_api)
    host_main = {k: v * 1e3 / args.steps for k, v in host_t.items()}
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(35)
    model.flush_logs()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    # the same leg with the chunks page-locked in place once (StagedCSRBatches(pin_chunks=True)): a batch is its
    # rebased crow plus two DMA transfers out of the chunk's own arrays -- no host pass, no packing threads.  Reported
    # next to the headline (which keeps the reference loader's pageable chunks), to show what the step sustains when
    # the host's DRAM bandwidth is not spent on packing (it is what limits `e2e` at N=8)
    pinned_leg = None
    if os.environ.get("BENCH_PINNED_LEG", "0") == "1":     # (opt-in: a third leg lengthens the run into the power cap)
        pfeeds = {s: StagedCSRBatches(chunk_source(s), B, device=dev, workers=0,
                                      ahead=int(os.environ.get("BENCH_PINNED_AHEAD", 4)), pin_chunks=True)
                  for s in names}
        main_iters, iters = iters, {s: iter(f) for s, f in pfeeds.items()}
        pending.clear()
        for t in range(w_api):
            api_step(t)
        barrier()
        host_t.update(fetch=0.0, training_step=0.0)
        e0.record()
        for t in range(args.steps):
            api_step(t + w_api)
        host_pinned = {k: v * 1e3 / args.steps for k, v in host_t.items()}
        model.flush_logs()
        e1.record()
        barrier()
        p_ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([p_ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            p_ms = float(tt.item())
        pinned_leg = {"value": B * world / (p_ms / args.steps * 1e-3), "unit": "cells/s",
                      "ms_per_step": p_ms / args.steps,
                      "h2d_bytes_per_step": int(pfeeds[names[0]].stager.bytes_staged),
                      "host_ms_per_step": host_pinned,
                      "host_side": "chunks page-locked in place once (cudaHostRegister), int32 gene ids, no packing"}
        pending.clear()
        iters = main_iters
        for f in pfeeds.values():
            f.close()
    clocks = sampler.stop() if rank == 0 else None
    h2d = int(feeds[names[0]].stager.bytes_staged)
    narrow = bool(feeds[names[0]].stager.narrow)
    diag = None
    if args.e2e_diag:
        # where does the end-to-end leg lose time?  (a) the same API calls on batches already in HBM (no H2D, no
        # packing); (b) packing + H2D copies alone (no compute)
        def api_resident(t):
            s, crow, col, val = resident_batch(t)
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
        t0 = time.perf_counter()
        for t in range(args.steps):
            next(iters[names[t % len(names)]])
        torch.cuda.synchronize()
        c_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        tt = torch.tensor([a_ms, c_ms], device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        diag = {"api_resident_ms_per_step": float(tt[0]), "pack+h2d_only_ms_per_step": float(tt[1]),
                "pack+h2d_only_gbs_per_gpu": h2d / (float(tt[1]) * 1e-3) / 1e9, "pack_workers": workers}
    for f in feeds.values():
        f.close()
    nnz = int(host[names[0]][0][1].size)
    d2h = int(eng.last["sc"].numel()) * 8

    parity = None
    if not args.no_parity_check:
        parity = parity_check(model, eng, args.config, species, conds, host, metas, B, rank, world, dev)

    also = None
    want_c3 = args.also_config3 if args.also_config3 is not None else int(world > 1 and args.config == 2)
    if want_c3:
        also = also_config3(rank, world, dev, min(args.steps, 40), max(3, min(args.warmup, 5)))

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    if parity is not None and not parity["ok"]:
        raise SystemExit(f"bench.py: the fused step disagrees with the oracle, no line printed: {json.dumps(parity)}")
    pk, pk_kind = peaks()
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape (ncu --set full)
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_dram_bytes.json")))
        traffic = tj.get(f"decoder_mse_fused_kernel@B{B}") if args.config == 2 else None
    except Exception:  # noqa: BLE001
        pass
    flops = 2.0 * B * G_HUMAN * H1      # per rank and launch, also when the layer is gene-sharded (N*B cells x G/N genes)
    ach = flops / (t_dec * 1e-3) / 1e12 if t_dec > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    spmm_bytes = nnz * 8 + (B + 1) * 4 + G_HUMAN * H1 * 2 + B * H1 * 4 + H1 * 4
    line = {
        "metric": METRIC, "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": config,
        "e2e": {"value": B * world / (e2e_ms / args.steps * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps, "last_loss": loss,
                "host_ms_per_step": host_main,
                "host_side": f"StagedCSRBatches over pageable scipy CSR chunks, {workers} packing threads inside the "
                             f"timed region, gene ids on the wire: {'uint16' if narrow else 'int32'}"},
        "gpu_launches": int(launches),
        "launch_mode": ("2 CUDA graphs per step, captured from this library's own launch sequence; gpu_launches = "
                        "kernels inside the replayed graphs") if graph_mode else "stream launches through the C ABI",
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
    if pinned_leg is not None:
        line["e2e"]["pinned_chunks"] = pinned_leg
    if parity is not None:
        line["parity_check"] = parity
    if also:
        line["also"] = also
    if diag:
        line["e2e_diag"] = diag
    if world == 1 and not args.no_torch_baseline:
        line["torch_cuda_baseline"] = torch_cuda_baseline(args.config, B, dev)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_leg(args.config, min(B, 1024), 8, 2, budget_s=25.0)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def also_config3(rank, world, dev, steps, warmup):
    """BASELINE configs[2] at the same N: two species + two GRL adversaries, 4096 cells per GPU per step,
    device-resident cells/s (CUDA events, max over ranks).  Returned as a sub-record of the main line."""
    import pandas as pd
    model, species, conds = build_model(3)
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    eng.pipeline_optimizer = True
    B, NB = 4096, 2
    names = list(species)
    resident = {s: [tuple(torch.from_numpy(a).to(dev) for a in synth_csr(B, g, DENSITY, 5000 * (rank + 1) + i))
                    for i in range(NB)] for s, g in species.items()}
    rng = np.random.default_rng(rank)
    labels = [{c: torch.from_numpy(rng.integers(0, n, size=B)).to(dev) for c, n in conds.items()} for _ in range(NB)]

    def batch(t):
        s = names[t % len(names)]
        return (s,) + resident[s][(t // len(names)) % NB]

    def step(t):
        s, crow, col, val = batch(t)
        eng.train_step(s, crow, col, val, int(col.numel()), 1.0, labels=labels[t % NB])
        if world > 1:
            eng.prefetch(*batch(t + 1), ready=True)

    for t in range(warmup):
        step(t)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        step(warmup + t)
    eng.finish()
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
    ms = float(tt.item()) / steps
    sc = eng.scalars()
    del model, eng
    torch.cuda.empty_cache()
    return {"config3": {"workload": f"BASELINE configs[2]: two-species CMMVAE + 2 GRL adversaries, {B} cells/GPU/step, "
                                    f"dp{world}, alternating human/mouse, device-resident",
                        "value": B * world / (ms * 1e-3), "unit": "cells/s", "ms_per_step": ms, "steps": steps,
                        "warmup": warmup, "n_gpus": world, "last_loss": sc["loss"]}}


if __name__ == "__main__":
    main()
