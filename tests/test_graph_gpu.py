"""CUDA-graph mode of the fused step: replayed steps must train exactly like stream-launched ones -- same logged
scalars step by step (dropout seeds, KL weight, Adam bias corrections and the batch's nnz are read from device
memory by the replayed launches), same final weights -- also when consecutive batches of different density reuse
the same staging addresses."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


def _model(G, with_adv, tmp):
    import os
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Adversarial, Expert, Experts, FCBlockConfig, LinearKLAnnealingFn
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    experts = Experts([Expert("human", FCBlockConfig([G, 256, 128], use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([128, 256, G], activation_fn=relu))])
    vae = CLVAE(FCBlockConfig([128, 64], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([32, 64, 128], activation_fn=relu), latent_dim=32, hidden_z=with_adv)
    advs = []
    if with_adv:
        os.makedirs(os.path.join(tmp, "human"), exist_ok=True)
        for c, n in (("assay", 5), ("dataset_id", 11)):
            pd.DataFrame([f"{c}_{i}" for i in range(n)]).to_csv(os.path.join(tmp, "human", f"unique_expression_{c}.csv"),
                                                                header=False, index=False)
        Adversarial.labels.clear()
        advs = [Adversarial(FCBlockConfig([64, 64, 32], activation_fn=relu), FCBlockConfig([32]),
                            ["assay", "dataset_id"], tmp),
                Adversarial(FCBlockConfig([32, 32], activation_fn=relu), FCBlockConfig([32]), ["assay", "dataset_id"], tmp)]
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    return CMMVAEModel(CMMVAE(vae, experts, advs), autograd_config=AutogradConfig(clip(), clip(), clip()),
                       kl_annealing_fn=LinearKLAnnealingFn(min_kl_weight=0.1, max_kl_weight=1.0, warmup_steps=2,
                                                           climax_steps=9))


@pytest.mark.parametrize("with_adv", [False, True])
def test_graph_replay_trains_like_stream_launches(with_adv, tmp_path):
    from mmvae_b200 import layers as L
    from mmvae_b200.feed import CSRStager
    L.set_precision("bf16")
    G, B, steps = 2000, 128, 12
    rng = np.random.default_rng(0)
    batches = [O.synth_csr(B, G, d, seed=50 + i) for i, d in enumerate((0.05, 0.03, 0.08, 0.05, 0.06, 0.04))]
    metas = [pd.DataFrame({"assay": [f"assay_{i}" for i in rng.integers(0, 5, B)],
                           "dataset_id": [f"dataset_id_{i}" for i in rng.integers(0, 11, B)]}) for _ in batches]
    runs = []
    for use_graph in (False, False, True):       # two stream-launched runs measure the run-to-run noise floor
        model = _model(G, with_adv, str(tmp_path / str(use_graph)))
        model.cuda().train()
        model.configure_optimizers()
        model.sync_logging = False
        model.use_cuda_graphs = use_graph
        for g in model.engine().groups.values():
            g.lr = 2e-4        # the reference's hard-coded 5e-3 makes 12 free-running steps chaotic (run-to-run noise of
                               # several per cent, occasionally far more); the comparison needs trajectories that stay close
        stager = CSRStager(max_cells=B, max_nnz=int(0.09 * B * G), device="cuda", depth=2, narrow_col=True)
        logs = []
        for t in range(steps):
            crow, col, val = batches[t % len(batches)]
            tk = stager.put(crow, col, val, G)
            L.inject_noise(torch.randn(B, 32, generator=torch.Generator().manual_seed(900 + t)).cuda())
            model.logged_metrics.clear()
            model.training_step((stager.get(tk), metas[t % len(batches)].copy(), "human"), t)
            stager.release(tk)
            logs.append({k: float(v) for k, v in model.logged_metrics.items()})
        model.flush_logs()
        logs.append({k: float(v) for k, v in model.logged_metrics.items()})
        eng = model.engine()
        n_graphs = sum(1 for e in eng._graphs.values() if "gA" in e)
        assert n_graphs == (1 if use_graph else 0)          # one graph pair serves every batch of the expert
        runs.append((logs, {k: v.detach().cpu() for k, v in model.state_dict().items()}))
    def deviation(x, y):
        worst = {}
        for t, (a, b) in enumerate(zip(x, y)):
            assert a.keys() == b.keys()
            for k in a:
                kind = "grad" if k.startswith("grad_norms") else ("mean" if k.startswith("Mean") else "other")
                dev = abs(a[k] - b[k]) / max(abs(a[k]), 1e-2 if kind == "mean" else 1e-12)
                if dev > worst.get(kind, (0.0,))[0]:
                    worst[kind] = (dev, t, k, a[k], b[k])
        return worst

    (la, sa), (lb, sb), (lg, sg) = runs
    # free-running trajectories (lr 5e-3, no re-synchronisation): two stream-launched runs already differ -- atomics'
    # summation order, amplified step by step -- so the replayed run is held to that noise floor, not to zero
    floor, got = deviation(la, lb), deviation(la, lg)
    for kind, dev in got.items():
        assert dev[0] <= max(4 * floor.get(kind, (0.0,))[0], 1e-2), (kind, dev, floor)
    for k, v in sa.items():
        if v.dtype.is_floating_point and not k.endswith("lin.bias"):
            ref = float((sb[k].double() - v.double()).norm() / v.double().norm().clamp_min(1e-30))
            err = float((sg[k].double() - v.double()).norm() / v.double().norm().clamp_min(1e-30))
            assert err <= max(4 * ref, 1e-3), (k, err, ref)
        elif not v.dtype.is_floating_point:
            assert torch.equal(sg[k], v), k
