"""Data-parallel fused step on 2 GPUs (NCCL) against the oracle: per-rank batches of the same species,
averaged gradients, clip on the averaged gradient, identical post-step weights on both ranks.
Skipped on boxes with fewer than 2 GPUs."""
import os

import numpy as np
import pandas as pd
import pytest
import torch
import torch.multiprocessing as mp

from helpers import GoldenCase, build_b200_model, csr_batch, rel_l2
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, tmp, out, precision="fp32", by_inputs="1"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CMMVAE_DP_BY_INPUTS=by_inputs)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from mmvae_b200 import layers as L
        L.set_precision(precision)
        gc = GoldenCase("core_human")
        from mmvae_b200.modules.base import KLAnnealingFn
        model = build_b200_model(gc, os.path.join(tmp, str(rank)), kl_fn=KLAnnealingFn(0.5))
        model.load_state_dict({f"module.{k}": v for k, v in gc.state("init").items()})
        model.cuda().train()
        model.configure_optimizers()
        s = gc.step(rank)          # rank r trains the reference's step-r batch, both from the init weights
        L.inject_noise(s["eps"].cuda())
        model.training_step((csr_batch(s["crow"], s["col"], s["val"], gc.genes["human"], f"cuda:{rank}"),
                             pd.DataFrame({"a": np.arange(gc.dims["B"])}), "human"), 0)
        torch.cuda.synchronize()
        sd = {k[len("module."):]: v.detach().cpu() for k, v in model.state_dict().items()}
        logs = {k: float(v) for k, v in model.logged_metrics.items()}
        out.put((rank, sd, logs))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_oracle(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path), out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    gc = GoldenCase("core_human")
    spec = gc.spec()
    # oracle: per-rank gradients from the same weights, averaged, clipped, one Adam step
    grads, new_buffers = [], []
    for r in range(world):
        P = gc.state("init")
        s = gc.step(r)
        o = O.train_step(spec, P, {}, "human", s["crow"], s["col"], s["val"], s["eps"], 0.5)
        grads.append(o["grads"])
        new_buffers.append({k: v for k, v in P.items() if "running" in k})
    P = gc.state("init")
    mean = {k: (grads[0][k] + grads[1][k]) / 2 for k in grads[0]}
    norms = {}
    for group in ("vae", "experts/human"):
        gsel = {k: v for k, v in mean.items() if O.group_of(k) == group}
        norms[group] = O.apply_group_step(P, gsel, O.OptState(), spec, 10.0)
    for r in range(world):
        _, sd, logs = res[r]
        assert logs["grad_norms/vae"] == pytest.approx(norms["vae"], rel=2e-4)
        assert logs["grad_norms/expert_human"] == pytest.approx(norms["experts/human"], rel=2e-4)
        for k, v in P.items():
            if "running" in k or k.endswith("num_batches_tracked"):
                continue  # BatchNorm statistics stay per rank (sync_batchnorm: false)
            if k.endswith(".lin.bias") and k.replace(".lin.bias", ".bn.weight") in P:
                continue
            assert rel_l2(sd[k].numpy(), v.numpy()) < 5e-5, (r, k)
        for k, v in new_buffers[r].items():
            if k.endswith("running_mean"):
                assert np.abs(sd[k].numpy() - v.numpy()).max() < 1e-5
            else:
                assert rel_l2(sd[k].numpy(), v.numpy()) < 1e-5, (r, k)
    # replicas stay identical
    for k in res[0][1]:
        if "running" in k or k.endswith("num_batches_tracked"):
            continue
        assert torch.equal(res[0][1][k], res[1][1][k]), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_first_layer_gradient_by_gathered_inputs_equals_reduce_scatter(tmp_path):
    """bf16 route, 2 ranks: computing each rank's gene shard of the summed first-layer weight gradient from
    all-gathered inputs (packed CSR, window pointers, dY) must give the step the reduce-scatter of the full
    per-rank gradients gives -- same bf16 operands, fp32 accumulation, only the summation order differs.
    G = 264 on 2 ranks also exercises a shard that is mostly padding rows (256 rows per rank)."""
    world = 2
    ctx = mp.get_context("spawn")
    runs = {}
    for mode in ("1", "0"):
        out = ctx.Queue()
        port = 29600 + (os.getpid() + 7 + int(mode)) % 2000
        procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path / mode), out, "bf16", mode))
                 for r in range(world)]
        for p in procs:
            p.start()
        runs[mode] = sorted([out.get(timeout=300) for _ in procs], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
    for r in range(world):
        _, sd1, logs1 = runs["1"][r]
        _, sd0, logs0 = runs["0"][r]
        for k in ("grad_norms/vae", "grad_norms/expert_human", "loss/training/human"):
            assert logs1[k] == pytest.approx(logs0[k], rel=1e-5), k
        for k in sd0:
            if k.endswith("num_batches_tracked"):
                continue
            assert rel_l2(sd1[k].float().numpy(), sd0[k].float().numpy()) < 2e-5, (r, k)
    # replicas stay identical in the by-inputs mode too
    for k in runs["1"][0][1]:
        if "running" in k or k.endswith("num_batches_tracked"):
            continue
        assert torch.equal(runs["1"][0][1][k], runs["1"][1][1][k]), k
