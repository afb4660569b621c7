"""N>1 host logic on CPU: two gloo ranks.  DDP semantics of the fused step = SUM all-reduce of the flat
gradient buffers + 1/world folded into clip/Adam: must equal averaging the per-rank gradients, clipping
the average and stepping (what Lightning DDP gives the reference)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmvae_b200 import dp
from oracle import cmmvae_oracle as O


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        n = 1000
        p0 = torch.randn(n, generator=torch.Generator().manual_seed(7))
        grad = torch.randn(n, generator=g) * 5
        flat = grad.clone()
        # chunked reduction (what overlapping the exchange with backward does) == one-shot reduction
        for lo, hi in dp.chunk_bounds(n, [300, 300, 640]):
            dp.allreduce_sum_(flat[lo:hi])
        gathered = [torch.empty(n) for _ in range(world)]
        dist.all_gather(gathered, grad)
        mean = torch.stack(gathered).mean(0)
        assert torch.allclose(flat / world, mean, atol=1e-6)
        # norm of the averaged gradient from the summed buffer
        assert dp.reduced_norm(float(flat.double().pow(2).sum())) == pytest.approx(float(mean.norm()), rel=1e-6)
        # fused semantics: coef from the averaged norm, gradient scale 1/world -> same update as clip(mean)+Adam
        total = float(flat.norm()) / world
        coef = min(1.0, 10.0 / (total + 1e-6))
        mine, _, _ = O.adam_update(p0, flat * (coef / world), torch.zeros(n), torch.zeros(n), 1, 5e-3, 1e-6,
                                   (0.9, 0.999), 1e-8)
        clipped, _ = O.clip_by_norm([mean], 10.0)
        ref, _, _ = O.adam_update(p0, clipped[0], torch.zeros(n), torch.zeros(n), 1, 5e-3, 1e-6, (0.9, 0.999), 1e-8)
        assert torch.allclose(mine, ref, atol=1e-7)
        # every rank picks the same species for every step
        sched = [dp.species_for_step(t, ["human", "mouse"], seed=3) for t in range(50)]
        objs = [None] * world
        dist.all_gather_object(objs, sched)
        assert all(o == sched for o in objs) and {"human", "mouse"} <= set(sched)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_by_inputs(rank, world, port, out):
    """the identity the data-parallel first-layer gradient relies on: rank r's gene shard of the SUMMED
    gradient  sum_s X_s^T dY_s  equals  X_all^T dY_all  restricted to the shard, so exchanging the inputs
    (CSR + dY, all-gather) replaces exchanging the outputs (reduce-scatter of a [G, H] matrix)"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, G, H = 24, 300, 16
        crow, col, val = O.synth_csr(B, G, 0.1, seed=40 + rank)
        X = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(col), torch.from_numpy(val),
                                    size=(B, G)).to_dense()
        dY = torch.randn(B, H, generator=torch.Generator().manual_seed(90 + rank))
        # outputs route: full per-rank gradient, summed over ranks (what DDP all-reduces), then sliced
        full = X.t() @ dY
        dist.all_reduce(full)
        per = dp.shard_rows(G, world)
        assert per % 128 == 0 and per * world >= G and per * (world - 1) < G + per
        lo, hi = rank * per, min((rank + 1) * per, G)
        # inputs route: gather every rank's cells and dY, multiply only this rank's gene rows
        Xs, dYs = [torch.empty_like(X) for _ in range(world)], [torch.empty_like(dY) for _ in range(world)]
        dist.all_gather(Xs, X)
        dist.all_gather(dYs, dY)
        shard = torch.cat(Xs)[:, lo:hi].t() @ torch.cat(dYs)
        assert torch.allclose(shard, full[lo:hi], rtol=1e-5, atol=1e-5)
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_first_layer_gradient_by_inputs_identity_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + 501) % 2000
    procs = [ctx.Process(target=_worker_by_inputs, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res


def _worker_gene_sharded(rank, world, port, out):
    """Executable statement of the round-2 plan ("Next"): with every rank holding all ranks'
    cells and only ITS gene rows of the two gene-sized matrices, the forward and backward of both layers need
    no weight or weight-gradient exchange:
      Y       = reduce_scatter_r( X_all[:, genes_r] @ W1t[genes_r] )            (first layer, all cells)
      recon   = sum_r || relu(h_all @ Wout[genes_r].T + b[genes_r]) - X_all[:, genes_r] ||^2
      dWout_r = dlogits_r.T @ h_all,   dW1t_r = X_all[:, genes_r].T @ dY_all       (rank-local, already summed)
      dh      = reduce_scatter_r( dlogits_r @ Wout[genes_r] )
    checked against the replicated-weights computation on the concatenated batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, G, H = 12, 300, 8

        def close(a, b):                                   # fp32 sums in different orders: compare at the tensor's scale
            return (a - b).abs().max() <= 2e-5 * b.abs().max()

        gen = torch.Generator().manual_seed(5)           # same weights on every rank
        W1t, Wout, bout = torch.randn(G, H, generator=gen), torch.randn(G, H, generator=gen), torch.randn(G, generator=gen)
        crow, col, val = O.synth_csr(B, G, 0.1, seed=60 + rank)
        X = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(col), torch.from_numpy(val),
                                    size=(B, G)).to_dense()
        Xs = [torch.empty_like(X) for _ in range(world)]
        dist.all_gather(Xs, X)
        X_all = torch.cat(Xs)                              # [world*B, G]
        per = dp.shard_rows(G, world, align=4)
        lo, hi = rank * per, min((rank + 1) * per, G)
        mine = slice(rank * B, (rank + 1) * B)            # this rank's cells inside the global batch

        # ---- reference: replicated weights, global batch
        Y_ref = X_all @ W1t
        h_ref = torch.tanh(Y_ref)                          # stand-in for the middle of the network
        logits = h_ref @ Wout.t() + bout
        xhat = torch.relu(logits)
        recon_ref = ((xhat - X_all) ** 2).sum()
        dlog_ref = 2 * (xhat - X_all) * (logits > 0)
        dWout_ref, dh_ref = dlog_ref.t() @ h_ref, dlog_ref @ Wout
        dY_ref = dh_ref * (1 - h_ref ** 2)
        dW1t_ref = X_all.t() @ dY_ref

        # ---- gene-sharded: only rows [lo, hi) of W1t / Wout / bout are touched on this rank
        part = X_all[:, lo:hi] @ W1t[lo:hi]
        dist.all_reduce(part)                              # gloo has no reduce_scatter: all-reduce, keep my cells
        Y = part[mine]
        assert close(Y, Y_ref[mine])
        h = torch.tanh(Y)
        hs = [torch.empty_like(h) for _ in range(world)]
        dist.all_gather(hs, h)
        h_all = torch.cat(hs)
        logits_r = h_all @ Wout[lo:hi].t() + bout[lo:hi]
        xhat_r = torch.relu(logits_r)
        recon = ((xhat_r - X_all[:, lo:hi]) ** 2).sum()
        dist.all_reduce(recon)
        assert close(recon, recon_ref)
        dlog_r = 2 * (xhat_r - X_all[:, lo:hi]) * (logits_r > 0)
        assert close(dlog_r.t() @ h_all, dWout_ref[lo:hi])
        dh_part = dlog_r @ Wout[lo:hi]
        dist.all_reduce(dh_part)
        dh = dh_part[mine]
        assert close(dh, dh_ref[mine])
        dY = dh * (1 - h ** 2)
        dYs = [torch.empty_like(dY) for _ in range(world)]
        dist.all_gather(dYs, dY)
        assert close(X_all[:, lo:hi].t() @ torch.cat(dYs), dW1t_ref[lo:hi])
        out.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gene_sharded_layers_identity_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + 977) % 2000
    procs = [ctx.Process(target=_worker_gene_sharded, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res


def test_shard_rows():
    assert dp.shard_rows(60530, 8) == 7680 and dp.shard_rows(60530, 2) == 30336 and dp.shard_rows(264, 2) == 256
    assert dp.shard_rows(128, 1) == 128 and dp.shard_rows(129, 1) == 256


def test_dp_semantics_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_helpers():
    assert dp.world_size() == 1 and dp.rank() == 0
    assert dp.allreduce_sum_(torch.ones(3)) is None
    assert dp.chunk_bounds(10, [0, 4, 4, 10, 12]) == [(0, 4), (4, 10)]


def _worker_module_route_guard(rank, world, port, out):
    """a topology that trains through the module route has no gradient exchange: with more than one process its
    training_step must refuse to run instead of letting the replicas drift apart"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pandas as pd
        from mmvae_b200.config import AutogradConfig, GradientClipConfig
        from mmvae_b200.models import CMMVAEModel
        from mmvae_b200.modules import CLVAE, CMMVAE
        from mmvae_b200.modules.base import Expert, Experts, FCBlockConfig, KLAnnealingFn
        relu = torch.nn.ReLU
        experts = Experts([Expert("human", FCBlockConfig([40, 16, 8], use_batch_norm=True, activation_fn=relu),
                                  FCBlockConfig([8, 16, 40], activation_fn=relu))])
        vae = CLVAE(FCBlockConfig([8, 8], use_batch_norm=True, activation_fn=relu), FCBlockConfig([4, 8, 8], activation_fn=relu),
                    latent_dim=4)
        clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
        model = CMMVAEModel(CMMVAE(vae, experts, []), autograd_config=AutogradConfig(clip(), clip(), clip()),
                            kl_annealing_fn=KLAnnealingFn(1.0))
        model.use_fused_engine = False
        assert model.engine() is None
        x = torch.eye(4, 40).to_sparse_csr()
        try:
            model.training_step((x, pd.DataFrame({"cell": range(4)}), "human"), 0)
            out.put((rank, "no error"))
        except RuntimeError as e:
            out.put((rank, "ok" if "single-process" in str(e) else repr(e)))
    except Exception as e:  # noqa: BLE001
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_module_route_refuses_to_run_data_parallel():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() + 1511) % 2000
    procs = [ctx.Process(target=_worker_module_route_guard, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
