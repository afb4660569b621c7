"""Engine state handling on the GPU: bounded workspaces under varying nnz, Adam moments in
optimizer.state_dict(), bf16 shadows refreshed on load_state_dict, Lightning progress advanced by the fused step,
the pre-dropout hidden-representation rule."""
import numpy as np
import pandas as pd
import pytest
import torch

from helpers import csr_batch
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


def _model(G=1500, H1=256, H2=128, Hv=64, Z=32, venc_dropout=0.0):
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Expert, Experts, FCBlockConfig, KLAnnealingFn
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    experts = Experts([Expert("human", FCBlockConfig([G, H1, H2], use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([H2, H1, G], activation_fn=relu))])
    vae = CLVAE(FCBlockConfig([H2, Hv], use_batch_norm=True, activation_fn=relu, return_hidden=True,
                              dropout_rate=venc_dropout),
                FCBlockConfig([Z, Hv, H2], activation_fn=relu), latent_dim=Z)
    return CMMVAEModel(CMMVAE(vae, experts, []), kl_annealing_fn=KLAnnealingFn(1.0))


def _step(model, B, G, density, seed, t=0, Z=32):
    from mmvae_b200 import layers as L
    crow, col, val = O.synth_csr(B, G, density, seed=seed)
    L.inject_noise(torch.randn(B, Z, generator=torch.Generator().manual_seed(seed)).cuda())
    model.training_step((csr_batch(crow, col, val, G), pd.DataFrame({"cell": np.arange(B)}), "human"), t)


def test_workspaces_stay_bounded_when_nnz_varies():
    """real CSR batches never repeat an nnz: the nnz-sized workspaces are high-water buffers, not one per shape"""
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    model = _model()
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    for i, dens in enumerate((0.05, 0.051, 0.052)):      # warm-up: every workspace exists afterwards
        _step(model, 96, 1500, dens, i)
    torch.cuda.synchronize()
    n_ws, mem = len(eng._ws), torch.cuda.memory_allocated()
    for i in range(40):
        _step(model, 96, 1500, 0.03 + 0.0005 * i, 10 + i)   # 40 distinct nnz, all below the high-water mark
        eng.spmm_tc = (i % 2 == 0)                          # both SpMM routes (packed records / CSC copy)
    torch.cuda.synchronize()
    assert len(eng._ws) <= n_ws + 6
    assert torch.cuda.memory_allocated() <= mem + (8 << 20)


def test_optimizer_state_dict_round_trip_resumes_adam():
    """FlatAdam speaks torch.optim.Adam's checkpoint format: a model + optimizers restored from state_dicts
    continues exactly like the original (moments and step counts included)"""
    from mmvae_b200 import layers as L
    L.set_precision("fp32")
    try:
        a = _model()
        a.cuda().train()
        opts_a = a.configure_optimizers()
        for t in range(3):
            _step(a, 64, 1500, 0.05, 100 + t, t)
        sd_model = {k: v.clone() for k, v in a.state_dict().items()}
        sd_opts = [o.state_dict() for o in opts_a]
        st = sd_opts[0]["state"]
        assert len(st) == len(opts_a[0].flat.params) and float(st[0]["step"]) == 3.0
        w = a.module.experts["human"].encoder.fc_layers[0].lin.weight
        idx = [id(p) for p in opts_a[0].flat.params].index(id(w))
        assert tuple(st[idx]["exp_avg"].shape) == tuple(w.shape)          # logical [H1, G] shape, like torch's
        b = _model()
        b.cuda().train()
        opts_b = b.configure_optimizers()
        b.load_state_dict(sd_model)
        for o, sd in zip(opts_b, sd_opts):
            o.load_state_dict(sd)
        _step(a, 64, 1500, 0.05, 777, 3)
        _step(b, 64, 1500, 0.05, 777, 3)
        sa, sb = a.state_dict(), b.state_dict()
        for k, va in sa.items():      # (atomics make summation order vary run to run: equal to rounding, not bitwise)
            if va.dtype.is_floating_point:
                err = float((va.double() - sb[k].double()).norm() / va.double().norm().clamp_min(1e-30))
                assert err < 2e-6, (k, err)
            else:
                assert torch.equal(va, sb[k]), k
        # a resumed run WITHOUT the optimizer state restarts Adam's bias correction: it must differ
        c = _model()
        c.cuda().train()
        c.configure_optimizers()
        c.load_state_dict(sd_model)
        _step(c, 64, 1500, 0.05, 777, 3)
        k = "module.vae.encoder.mean_encoder.weight"
        assert float((c.state_dict()[k] - sa[k]).abs().max()) > 1e-4      # first-step Adam moves every weight by lr
    finally:
        L.set_precision("bf16")


def test_load_state_dict_refreshes_bf16_shadows():
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    model = _model()
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    sd = {k: (v * 1.5 if v.dtype.is_floating_point and k.endswith("lin.weight") else v.clone())
          for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    for g in eng.groups.values():
        assert torch.equal(g.p16, g.p.bfloat16()), g.name


def test_fused_step_advances_optimizer_progress():
    """the reference steps vae + expert (+ each adversary) optimizers per batch; Lightning's global_step counts
    those calls.  The fused step applies the update itself and then calls step() on the same optimizers, which
    only consume the 'applied' mark"""
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    model = _model()
    model.cuda().train()
    opts = model.configure_optimizers()
    calls = []
    for o in opts:
        orig = o.step
        o.step = (lambda orig=orig, o=o: (calls.append(o), orig())[1])
    w_before = model.module.vae.encoder.mean_encoder.weight.detach().clone()
    _step(model, 64, 1500, 0.05, 1)
    assert len(calls) == 2 and model.trainer.global_step == 2
    g = model.engine().groups["vae"]
    assert g.step_count == 1 and not g.applied
    w1 = model.module.vae.encoder.mean_encoder.weight.detach().clone()
    assert not torch.equal(w1, w_before)
    opts[-1].step()      # an explicit step() by a caller still works (single process): a second Adam update
    assert g.step_count == 2


def test_hidden_representation_with_dropout_leaves_the_fused_step():
    """the reference hands the adversary the activation BEFORE dropout (components.py:309-313); the fused step
    cannot, so such a topology is routed to the module path instead of silently differing"""
    model = _model(venc_dropout=0.2)
    model.cuda().train()
    assert model.engine() is None and "dropout" in model._module_route_reason
