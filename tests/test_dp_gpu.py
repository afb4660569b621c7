"""Data-parallel fused step over peer memory (gene-sharded first / last layer, csrc/peer.cu), checked three ways:

  * one process, CMMVAE_FORCE_DP=1: the data-parallel ROUTE with itself as only peer must reproduce the ordinary
    single-GPU step (same kernels, different decomposition) -- runs on any GPU box;
  * two processes sharing ONE GPU (CUDA IPC between processes, gloo for the handshake): real two-rank exchange --
    flags, slabs, gene shards incl. a shard that is mostly padding -- against (a) the sum of the two single-process
    gradients computed with the same kernels and (b) the oracle's averaged-gradient step (DDP semantics: mean of
    the per-rank gradients, clip on the mean, per-rank BatchNorm statistics and logged losses);
  * two processes on two GPUs over NVLink (NCCL handshake); skipped on boxes with one GPU.
"""
import os

import numpy as np
import pandas as pd
import pytest
import torch
import torch.multiprocessing as mp

from helpers import csr_batch, rel_l2
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu

DIMS = dict(G=3000, H1=256, H2=128, Hv=64, Z=32, B=160)
TINY = dict(G=264, H1=64, H2=32, Hv=32, Z=16, B=24)      # 2 ranks: rank 1 owns 8 real gene rows + 248 rows of padding


def _build(d, dropout=0.0, od=False):
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Expert, Experts, FCBlockConfig, KLAnnealingFn
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    experts = Experts([Expert("human", FCBlockConfig([d["G"], d["H1"], d["H2"]], use_batch_norm=True, activation_fn=relu,
                                                     dropout_rate=dropout),
                              FCBlockConfig([d["H2"], d["H1"], d["G"]], activation_fn=relu))])
    vae = CLVAE(FCBlockConfig([d["H2"], d["Hv"]], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([d["Z"], d["Hv"], d["H2"]], activation_fn=relu), latent_dim=d["Z"])
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    extra = {}
    if od:
        from mmvae_b200.modules import create_discriminators
        torch.manual_seed(1)
        extra["output_discriminators"] = create_discriminators({"human": d["G"]})
    return CMMVAEModel(CMMVAE(vae, experts, []), autograd_config=AutogradConfig(clip(), clip(), clip()),
                       kl_annealing_fn=KLAnnealingFn(1.0), **extra)


def _spec(d):
    return O.ModelSpec(experts={"human": {"encoder": O.BlockSpec.make([d["G"], d["H1"], d["H2"]], bn=True),
                                          "decoder": O.BlockSpec.make([d["H2"], d["H1"], d["G"]])}},
                       vae_encoder=O.BlockSpec.make([d["H2"], d["Hv"]], bn=True, return_hidden=True),
                       vae_decoder=O.BlockSpec.make([d["Z"], d["Hv"], d["H2"]]), latent_dim=d["Z"])


def _inputs(d, rank, t=0):
    crow, col, val = O.synth_csr(d["B"], d["G"], 0.06, seed=400 + 10 * t + rank)
    eps = torch.randn(d["B"], d["Z"], generator=torch.Generator().manual_seed(50 + 10 * t + rank))
    return crow, col, val, eps


def _batch(d, rank, t, device):
    crow, col, val, eps = _inputs(d, rank, t)
    rng = np.random.default_rng(7000 + 10 * t + rank)
    meta = pd.DataFrame({"cell": np.arange(d["B"]),
                         "assay": [f"assay_{i}" for i in rng.integers(0, 5, d["B"])],
                         "dataset_id": [f"dataset_id_{i}" for i in rng.integers(0, 11, d["B"])]})
    return (csr_batch(crow, col, val, d["G"], device), meta, "human"), eps


def _step(model, d, rank, t=0, device="cuda", batch=None, after=None):
    from mmvae_b200 import layers as L
    batch, eps = batch or _batch(d, rank, t, device)
    L.inject_noise(eps.to(device))
    model.logged_metrics.clear()
    model.training_step(batch, t)
    if after is not None:
        after()
    if not model.sync_logging:
        model.flush_logs()
    torch.cuda.synchronize()
    return {k.split("/")[0] if not k.startswith("grad_norms") else k: float(v) for k, v in model.logged_metrics.items()}


def _grads(model):
    """every parameter's gradient as left by the step (row-sharded ones: this rank's rows are valid, others 0)"""
    return {(n[len("module."):] if n.startswith("module.") else n): p.grad.detach().float().cpu().clone()
            for n, p in model.named_parameters()}


def _single_process_reference(d, world, steps=1):
    """the ordinary single-GPU engine on each rank's batch from the same weights: losses and gradients"""
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    os.environ.pop("CMMVAE_FORCE_DP", None)
    model = _build(d, od=bool(d.get("od")))
    init = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.cuda().train()
    model.configure_optimizers()
    assert model.engine().comm is None
    out = []
    for r in range(world):
        model.load_state_dict(init)
        for g in model.engine().groups.values():
            g.m.zero_(); g.v.zero_(); g.step_count = 0
        for name, buf in model.named_buffers():
            buf.copy_(init[name])
        logs = _step(model, d, r)
        out.append((logs, _grads(model)))
    return init, out


def test_forced_dp_route_single_process_equals_plain_step():
    """the gene-sharded route with one rank (gathered CSR -> shard CSR -> routed SpMM / decoder blocks / routed dh
    -> slab sums) gives the step the plain route gives"""
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    for d in (DIMS, TINY):
        init, ref = _single_process_reference(d, 1)
        os.environ["CMMVAE_FORCE_DP"] = "1"
        try:
            model = _build(d)
            model.load_state_dict(init)
            model.cuda().train()
            model.configure_optimizers()
            eng = model.engine()
            assert eng.comm is not None and eng.world == 1
            logs = _step(model, d, 0)
            grads = _grads(model)
        finally:
            os.environ.pop("CMMVAE_FORCE_DP", None)
        rlogs, rgrads = ref[0]
        for k, tol in (("loss", 2e-5), ("recon_loss", 2e-5), ("kl_loss", 2e-4), ("grad_norms/vae", 1e-3),
                       ("grad_norms/expert_human", 1e-3)):    # (summation order differs; KL sits behind two BatchNorms)
            assert logs[k] == pytest.approx(rlogs[k], rel=tol), (k, logs[k], rlogs[k])
        for k, g in rgrads.items():
            if k.endswith(".lin.bias") and k.replace(".lin.bias", ".bn.weight") in rgrads:
                continue
            # same bf16 kernels, different summation order: a last-bit difference in a pre-activation flips a bf16
            # rounding or a ReLU mask now and then, which moves encoder-side gradients by ~1e-2 (see
            # tests/test_parity_fullshape_gpu.py); a routing / sharding bug moves them by O(1)
            assert rel_l2(grads[k].numpy(), g.numpy()) < 3e-2, (d["G"], k, rel_l2(grads[k].numpy(), g.numpy()))


def _worker(rank, world, port, out, d, same_gpu, steps, pipelined=None):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    os.environ.pop("CMMVAE_FORCE_DP", None)
    dev = 0 if same_gpu else rank
    torch.cuda.set_device(dev)
    if same_gpu:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    try:
        from mmvae_b200 import layers as L
        L.set_precision("bf16")
        if d.get("adv"):      # two GRL adversaries (their small groups are summed over ranks inside the step)
            import tempfile
            from test_graph_gpu import _model
            model = _model(d["G"], True, tempfile.mkdtemp())
        else:
            model = _build(d, od=bool(d.get("od")))
        model.cuda().train()
        model.configure_optimizers()
        eng = model.engine()
        assert eng.comm is not None and eng.world == world
        if d.get("adv"):
            for g in eng.groups.values():
                g.lr = 2e-4       # (free-running comparison of two launch modes: keep the trajectories close)
        res = []
        if pipelined is not None:
            # the training loop's mode: results trail by a step, the next batch's records are exchanged underneath
            # the running step, and (pipelined == "graph") the step is replayed from CUDA graphs
            model.sync_logging = False
            model.use_cuda_graphs = pipelined == "graph"
        nxt = None
        for t in range(steps):
            cur = nxt or _batch(d, rank, t, f"cuda:{dev}")
            nxt = _batch(d, rank, t + 1, f"cuda:{dev}") if pipelined is not None and t + 1 < steps else None
            # (step 2's successor is NOT announced: its records are exchanged at the start of step 3 instead)
            pf = (lambda: model.prefetch_batch(nxt[0])) if nxt is not None and t != 2 else None
            logs = _step(model, d, rank, t, f"cuda:{dev}", batch=cur, after=pf)
            res.append(logs)
        if pipelined == "graph":
            replays = sum(e.get("replays", 0) for e in eng._graphs.values())
            assert replays >= steps - 4, (replays, list(eng._graphs))
        grads = _grads(model)                      # of the last step
        sd = {(k[len("module."):] if k.startswith("module.") else k): v.detach().cpu()
              for k, v in model.state_dict().items()}   # gathers the rows
        # numpy arrays travel by value (torch tensors would travel as file descriptors of a process about to exit)
        out.put((rank, res, {k: v.numpy() for k, v in grads.items()}, {k: v.numpy() for k, v in sd.items()}))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run_ranks(world, d, same_gpu, steps=1, pipelined=None):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    _run_ranks.n = getattr(_run_ranks, "n", 0) + 1
    port = 29600 + (os.getpid() * 7 + world + int(same_gpu) + 13 * _run_ranks.n) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, d, same_gpu, steps, pipelined)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


def _check_two_ranks(d, res, world=2):
    init, ref = _single_process_reference(d, world)
    names = list(ref[0][1])
    want = {k: sum(ref[r][1][k] for r in range(world)) for k in names}        # SUM of the per-rank gradients
    # (1) per-rank losses are the rank's own batch's; gradient = sum over ranks (rows owned by the rank / replicated)
    sharded = ("experts.human.encoder.fc_layers.0.lin.weight", "experts.human.decoder.fc_layers.1.lin.weight",
               "experts.human.decoder.fc_layers.1.lin.bias")
    per = -(-d["G"] // world)
    per = (per + 127) // 128 * 128
    for r in range(world):
        _, logs, grads, sd = res[r]
        for k, tol in (("loss", 2e-5), ("recon_loss", 2e-5), ("kl_loss", 2e-4)):
            assert logs[0][k] == pytest.approx(ref[r][0][k], rel=tol), (r, k)
        for k in names:
            if k.endswith(".lin.bias") and k.replace(".lin.bias", ".bn.weight") in want:
                continue
            a, b = torch.from_numpy(grads[k]), want[k]
            if k in sharded:      # gene rows [r * per, (r + 1) * per) are this rank's; gene axis is dim 1 of W1
                lo, hi = r * per, min(d["G"], (r + 1) * per)
                if k.endswith("encoder.fc_layers.0.lin.weight"):
                    a, b = a[:, lo:hi], b[:, lo:hi]
                else:
                    a, b = a[lo:hi], b[lo:hi]
            assert rel_l2(a.numpy(), b.numpy()) < 3e-2, (r, k, rel_l2(a.numpy(), b.numpy()))
    # (2) replicas agree after the step (row-sharded parameters gathered by state_dict)
    for k in res[0][3]:
        if "running" in k or k.endswith("num_batches_tracked"):
            continue      # BatchNorm statistics stay per rank (sync_batchnorm: false)
        assert np.array_equal(res[0][3][k], res[1][3][k]), k
    # (3) DDP semantics against the oracle: mean gradient, clip on the mean, one Adam step
    spec = _spec(d)
    ograds = []
    for r in range(world):
        P = {k[len("module."):]: v.clone() for k, v in init.items()}
        crow, col, val, eps = _inputs(d, r)
        ograds.append(O.train_step(spec, P, {}, "human", crow, col, val, eps, 1.0)["grads"])
    mean = {k: sum(g[k] for g in ograds) / world for k in ograds[0]}
    P = {k[len("module."):]: v.clone() for k, v in init.items()}
    norms = {}
    for group in ("vae", "experts/human"):
        norms[group] = O.apply_group_step(P, {k: v for k, v in mean.items() if O.group_of(k) == group}, O.OptState(),
                                          spec, 10.0)
    _, logs, _, sd = res[0]
    assert logs[0]["grad_norms/vae"] == pytest.approx(norms["vae"], rel=3e-2)
    assert logs[0]["grad_norms/expert_human"] == pytest.approx(norms["experts/human"], rel=3e-2)
    for k in ("experts.human.encoder.fc_layers.0.lin.weight", "experts.human.decoder.fc_layers.1.lin.weight",
              "experts.human.encoder.fc_layers.1.lin.weight", "vae.encoder.mean_encoder.weight"):
        p0 = init["module." + k].double().flatten()
        ua, ub = torch.from_numpy(sd[k]).double().flatten() - p0, P[k].double().flatten() - p0
        cos = float((ua * ub).sum() / (ua.norm() * ub.norm()).clamp_min(1e-30))
        assert cos > 0.95, (k, cos)


def test_two_ranks_output_discriminator_on_the_gene_sharded_route():
    """BASELINE config 4 data parallel: the discriminator's first layer is sharded by genes like the expert's; the
    two halves of xhat W1^T are routed to the cells' owners, d(a1) is gathered for the rows of dW1.  Per-rank loss =
    the single-process loss on that rank's batch; gradients = the SUM over ranks of the single-process gradients
    (own gene rows of W1 / replicated small layers); replicas agree bit for bit after the step."""
    d = dict(DIMS, od=True)
    world = 2
    res = _run_ranks(world, d, same_gpu=True)
    init, ref = _single_process_reference(d, world)
    per = (-(-d["G"] // world) + 127) // 128 * 128
    pre = "output_discriminators.human."
    names = [k for k in ref[0][1] if k.startswith(pre)]
    assert len(names) == 6
    for r in range(world):
        _, logs, grads, sd = res[r]
        assert logs[0]["meta_disc"] == pytest.approx(ref[r][0]["meta_disc"], rel=2e-3), r
        assert logs[0]["loss"] == pytest.approx(ref[r][0]["loss"], rel=2e-5), r
        for k in names:
            a, b = torch.from_numpy(grads[k]), sum(ref[q][1][k] for q in range(world))
            if k == pre + "0.weight":      # [128, G]: gene axis is dim 1; this rank's genes
                lo, hi = r * per, min(d["G"], (r + 1) * per)
                a, b = a[:, lo:hi], b[:, lo:hi]
            assert rel_l2(a.numpy(), b.numpy()) < 3e-2, (r, k, rel_l2(a.numpy(), b.numpy()))
    for k in res[0][3]:
        if "running" in k or k.endswith("num_batches_tracked"):
            continue
        assert np.array_equal(res[0][3][k], res[1][3][k]), k
    # the discriminator moved (Adam at lr 1e-3 on the mean gradient)
    k = pre + "2.weight"
    assert not np.array_equal(res[0][3][k], init[k].numpy())


@pytest.mark.parametrize("d", [DIMS, TINY], ids=["mid", "tiny-padding-shard"])
def test_two_ranks_sharing_one_gpu_over_cuda_ipc(d):
    res = _run_ranks(2, d, same_gpu=True)
    _check_two_ranks(d, res)


def _compare_runs(a, b, steps, extra=()):
    for r in range(2):
        for t in range(steps):
            for k, tol in (("loss", 2e-3), ("recon_loss", 2e-3), ("kl_loss", 2e-2), ("grad_norms/expert_human", 5e-2)) \
                    + tuple(extra):
                assert a[r][1][t][k] == pytest.approx(b[r][1][t][k], rel=tol), (r, t, k)
        for k in ("experts.human.encoder.fc_layers.0.lin.weight", "experts.human.decoder.fc_layers.1.lin.weight",
                  "vae.encoder.mean_encoder.weight"):
            assert rel_l2(a[r][3][k], b[r][3][k]) < 2e-3, (r, k, rel_l2(a[r][3][k], b[r][3][k]))


@pytest.mark.parametrize("adv", [False, True], ids=["core", "adversaries"])
def test_two_ranks_graph_replay_equals_stream_launches(adv):
    """data parallel, pipelined mode: the step replayed from CUDA graphs (flag values read from the device block,
    one graph pair per buffer parity, records prefetched or exchanged inline; with adversaries: their replicated
    groups summed over ranks inside the graphs) walks the trajectory the stream-launched step walks; after the run
    the replicas still agree bit for bit"""
    steps = 8
    d = dict(DIMS, adv=True, G=2000, Z=32, B=128) if adv else DIMS
    eager = _run_ranks(2, d, same_gpu=True, steps=steps, pipelined="stream")
    graph = _run_ranks(2, d, same_gpu=True, steps=steps, pipelined="graph")
    extra = (("discriminator_1", 2e-2), ("generator_2", 2e-2)) if adv else ()
    _compare_runs(eager, graph, steps, extra)
    for k in graph[0][3]:
        if "running" in k or k.endswith("num_batches_tracked"):
            continue
        assert np.array_equal(graph[0][3][k], graph[1][3][k]), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_ranks_on_two_gpus_over_nvlink():
    res = _run_ranks(2, DIMS, same_gpu=False, steps=3)
    assert all(np.isfinite(list(r.values())).all() for _, logs, _, _ in res for r in logs)
    res1 = _run_ranks(2, DIMS, same_gpu=False)
    _check_two_ranks(DIMS, res1)
