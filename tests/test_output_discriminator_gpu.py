"""Output discriminator inside the fused step (BASELINE config 4) against the oracle's restatement of
the reference network (runners/meta_discriminators.py:33-49,112-148), MLP level: the discriminator of the step's
species trains on the step's (detached) reconstruction.  The reconstruction never exists in HBM on the B200 side
(it is recovered from dlogits + the CSR batch); the oracle materialises it."""
import numpy as np
import pandas as pd
import pytest
import torch

from helpers import csr_batch, rel_l2
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


def _build(G, dims):
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE, create_discriminators
    from mmvae_b200.modules.base import Expert, Experts, FCBlockConfig, KLAnnealingFn
    H1, H2, Hv, Z = dims
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    experts = Experts([Expert(s, FCBlockConfig([G[s], H1, H2], use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([H2, H1, G[s]], activation_fn=relu)) for s in G])
    vae = CLVAE(FCBlockConfig([H2, Hv], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([Z, Hv, H2], activation_fn=relu), latent_dim=Z)
    return CMMVAEModel(CMMVAE(vae, experts, []), kl_annealing_fn=KLAnnealingFn(1.0),
                       output_discriminators=create_discriminators(G))


@pytest.mark.parametrize("graph", [False, True], ids=["stream", "graph"])
def test_output_discriminator_step_matches_oracle(graph):
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    G = {"human": 3000, "mouse": 2777}
    dims = (256, 128, 64, 32)
    H1, H2, Hv, Z = dims
    B = 256
    model = _build(G, dims)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    P = {k[len("module."):]: v.clone() for k, v in sd.items() if k.startswith("module.")}
    Pd = {s: {k[len(f"output_discriminators.{s}."):]: v.clone() for k, v in sd.items()
              if k.startswith(f"output_discriminators.{s}.")} for s in G}
    assert set(Pd["human"]) == {"0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"}
    spec = O.ModelSpec(experts={s: {"encoder": O.BlockSpec.make([g, H1, H2], bn=True),
                                    "decoder": O.BlockSpec.make([H2, H1, g])} for s, g in G.items()},
                       vae_encoder=O.BlockSpec.make([H2, Hv], bn=True, return_hidden=True),
                       vae_decoder=O.BlockSpec.make([Z, Hv, H2]), latent_dim=Z)
    model.cuda().train()
    opts = model.configure_optimizers()
    assert model.optimizer_map["output_discriminators"] == {"human": len(opts) - 2, "mouse": len(opts) - 1}
    if graph:
        model.sync_logging = False
    opt, dopt = {}, {s: O.OptState() for s in G}
    steps = [("human", 0), ("mouse", 1), ("human", 2), ("mouse", 3), ("human", 4), ("human", 5)]
    for t, (sp, seed) in enumerate(steps):
        crow, col, val = O.synth_csr(B, G[sp], 0.06, seed=70 + seed)
        eps = torch.randn(B, Z, generator=torch.Generator().manual_seed(seed))
        # oracle: x-hat of THIS step's forward (before any update), then the discriminator step, then the main step
        _, _, _, xhat, _ = O.forward(spec, P, sp, (crow, col, val), G[sp], eps, True, {})
        want = O.output_discriminator_step(Pd[sp], dopt[sp], xhat.detach(), {"human": 0.0, "mouse": 1.0}[sp])
        ref = O.train_step(spec, P, opt, sp, crow, col, val, eps, 1.0)
        L.inject_noise(eps.cuda())
        model.logged_metrics.clear()
        model.training_step((csr_batch(crow, col, val, G[sp]), pd.DataFrame({"cell": np.arange(B)}), sp), t)
        model.flush_logs()
        torch.cuda.synchronize()
        got = {k: float(v) for k, v in model.logged_metrics.items()}
        assert got[f"meta_disc/md_{sp}"] == pytest.approx(want["loss"], rel=2e-3), (t, sp)
        assert got[f"loss/training/{sp}"] == pytest.approx(ref["logs"]["loss"], rel=1e-3)
        disc = model.output_discriminators[sp]
        for k, g in want["grads"].items():
            mine = dict(disc.named_parameters())[k].grad.detach().cpu().numpy()
            assert rel_l2(mine, g.numpy()) < 3e-2, (t, sp, k, rel_l2(mine, g.numpy()))
        # same-state comparison next step: load the oracle's weights (main model and discriminators)
        new = {f"module.{k}": v for k, v in P.items()}
        new.update({f"output_discriminators.{s}.{k}": v.detach() for s in G for k, v in Pd[s].items()})
        mine_sd = model.state_dict()
        for k in ("0.weight", "4.bias"):
            a = mine_sd[f"output_discriminators.{sp}.{k}"].cpu().double().flatten()
            b = Pd[sp][k].detach().double().flatten()
            p0 = sd[f"output_discriminators.{sp}.{k}"].double().flatten() if t < 2 else None
            if p0 is not None:      # first Adam step of this discriminator: every element moves by +-lr: directions
                ua, ub = a - p0, b - p0
                cos = float((ua * ub).sum() / (ua.norm() * ub.norm()).clamp_min(1e-30))
                assert cos > 0.9, (t, sp, k, cos)
        model.load_state_dict(new)
        # the discriminators' Adam moments follow the oracle's too (same-state comparison of later steps)
        eng = model.engine()
        for s in G:
            g = eng.groups[f"output_discriminators/{s}"]
            for k, p in model.output_discriminators[s].named_parameters():
                if k in dopt[s].m:
                    g.logical(p, g.m).copy_(dopt[s].m[k])
                    g.logical(p, g.v).copy_(dopt[s].v[k])
            g.step_count = max(dopt[s].step.values()) if dopt[s].step else 0
