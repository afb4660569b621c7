"""Parity of the bf16 tcgen05 training step against the oracle AT THE SHAPES bench.py TIMES, with dropout on
(masks injected on both sides), and over a 100-step trajectory (VERDICT r1 "next" #1).

  config 2  human expert, G=60 530, 1024 cells, 1024-512 | 512-256-Z128, dropout 0.1          (full check)
  config 3  human + mouse experts, two GRL adversaries (hidden-256 and z heads), 4096 cells     (all logged scalars)
  config 4  G=60 664, 8192 cells, 1024-768 | 768-512-Z256                                      (all logged scalars)
  curve     100 optimisation steps, lr 5e-3, dropout 0.1: the bf16 ELBO curve stays within 1 % of the oracle's

Stated tolerances (bf16 operands, fp32 accumulation, fp32 everything else): loss / recon 5e-4, KL 5e-3,
adversary CE 1e-2, gradient norms 3e-2, per-tensor gradients <= 8e-2 rel-L2.

Why gradients sit at several 1e-2 while the loss agrees to 1e-5: a bf16 operand carries 2^-9 relative rounding,
so pre-activations differ by ~0.3 % and ~0.3 % of the ReLU masks of a layer flip; a flipped element contributes
its whole gradient as error, i.e. sqrt(0.003) ~ 5 % rel-L2 per ReLU layer on everything upstream of it.  Measured
(tools/diag_precision.py, config 2): with every GEMM on bf16 operands the encoder-side tensors sit at 9-10 %
(torch's own bf16 autocast: 9e-2 on W1.grad); the small GEMMs between the two gene-sized layers
therefore run on TF32 operands (tcgen05 kind::tf32, rounded by the TMA unit), which brings them to 3-6 %; what
remains comes from the bf16 operands of the two gene-sized layers themselves.
"""
import os

import numpy as np
import pandas as pd
import psutil
import pytest
import torch

from helpers import csr_batch, rel_l2, untag
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _sparse_addmm_in_the_oracle():
    """at these sizes the oracle's first layer goes through torch's sparse-CSR addmm -- the very ATen call the
    reference makes (components.py:276) -- instead of the explicit per-non-zero restatement, which materialises
    [nnz, 1024] floats (12.7 GB at 1024 cells); tests/test_oracle_golden.py pins that both forms agree"""
    O.FAST_CSR = True
    yield
    O.FAST_CSR = False


TOL = dict(loss=5e-4, kl=5e-3, adv=1e-2, norm=3e-2, grad=8e-2)


def bias_feeds_batchnorm(name, state):
    return name.endswith(".lin.bias") and name.replace(".lin.bias", ".bn.weight") in state


def oracle_spec(species, H1, H2, Hv, Z, conds=None, hidden_z=False, dropout=0.1, adv_dims=None):
    advs = []
    if conds:
        a1, a2 = adv_dims or ([Hv, 128, 64], [Z, 64])
        advs = [O.AdversarySpec(O.BlockSpec.make(a1), dict(conds)), O.AdversarySpec(O.BlockSpec.make(a2), dict(conds))]
    return O.ModelSpec(
        experts={s: {"encoder": O.BlockSpec.make([g, H1, H2], bn=True, dropout=dropout),
                     "decoder": O.BlockSpec.make([H2, H1, g])} for s, g in species.items()},
        vae_encoder=O.BlockSpec.make([H2, Hv], bn=True, return_hidden=True),
        vae_decoder=O.BlockSpec.make([Z, Hv, H2]), latent_dim=Z, hidden_z=hidden_z, adversarials=advs, adv_weight=1.0)


def dropout_masks(species_id, widths, B, p, seed):
    """keep-masks for the expert encoder's dropout layers: float {0,1} for the oracle, uint8 for the kernels"""
    g = torch.Generator().manual_seed(seed)
    cpu, gpu = {}, {}
    for j, n in enumerate(widths):
        keep = (torch.rand(B, n, generator=g) >= p)
        key = f"experts.{species_id}.encoder.fc_layers.{j}.dr"
        cpu[key] = keep.float()
        gpu[key] = keep.to(torch.uint8).cuda()
    return cpu, gpu


def state_of(model):
    return {k[len("module."):]: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def check_scalars(got, ref, tag=""):
    assert set(got) == set(ref), sorted(set(got) ^ set(ref))
    for k, v in ref.items():
        if k in ("loss", "recon_loss"):
            tol = TOL["loss"]
        elif k in ("kl_loss", "Mean", "Variance"):
            tol = TOL["kl"]
        elif "adversarial_loss" in k:
            tol = TOL["adv"]
        elif k.startswith("grad_norms"):
            tol = TOL["norm"]
        else:
            tol = 1e-6
        assert got[k] == pytest.approx(v, rel=tol, abs=1e-6), (tag, k, got[k], v)


def run_pair(model, spec, P, opt, sp, G, crow, col, val, eps, labels, conds, masks, t):
    """one step on both sides from the same state; returns (oracle record, logged scalars of the B200 step)"""
    from mmvae_b200 import layers as L
    cpu_masks, gpu_masks = masks
    ref = O.train_step(spec, P, opt, sp, crow, col, val, eps, 1.0, labels=labels, dropout_masks=cpu_masks)
    L.inject_noise(eps.cuda())
    L.inject_dropout_masks(gpu_masks)
    B = len(crow) - 1
    meta = pd.DataFrame({c: [f"{c}_{int(i)}" for i in labels[c]] for c in conds}) if conds else \
        pd.DataFrame({"cell": np.arange(B)})
    model.logged_metrics.clear()
    model.training_step((csr_batch(crow, col, val, G), meta, sp), t)
    torch.cuda.synchronize()
    return ref, untag({k: float(v) for k, v in model.logged_metrics.items()}, sp)


def test_config2_full_shape_step_with_dropout():
    """BASELINE config 2 exactly as bench.py builds it (bench.build_model(2): dropout 0.1 in the expert encoder):
    every logged scalar, every gradient tensor and the direction of the Adam update against the oracle."""
    import bench
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    model, species, _ = bench.build_model(2)
    d = bench.Dims(2)
    G, B = d.G_HUMAN, 1024
    spec = oracle_spec(species, d.H1, d.H2, d.HV, d.Z)
    P = state_of(model)
    model.cuda().train()
    model.configure_optimizers()
    opt = {}
    for t in range(2):
        crow, col, val = bench.synth_csr(B, G, 0.05, seed=900 + t)
        eps = torch.randn(B, d.Z, generator=torch.Generator().manual_seed(t))
        before = {k: v.clone() for k, v in P.items()} if t == 0 else None
        ref, got = run_pair(model, spec, P, opt, "human", G, crow, col, val, eps, None, None,
                            dropout_masks("human", (d.H1, d.H2), B, 0.1, 40 + t), t)
        check_scalars(got, ref["logs"], f"step{t}")
        if t == 0:
            params = dict(model.named_parameters())
            worst = {}
            for k, g in ref["grads"].items():
                if bias_feeds_batchnorm(k, P):
                    continue      # analytically zero gradient: rounding noise on both sides
                worst[k] = rel_l2(params[f"module.{k}"].grad.detach().cpu().numpy(), g.numpy())
            assert max(worst.values()) < TOL["grad"], sorted(worst.items(), key=lambda kv: -kv[1])[:4]
            mine = state_of(model)
            for k in ("experts.human.encoder.fc_layers.0.lin.weight", "experts.human.decoder.fc_layers.1.lin.weight",
                      "experts.human.decoder.fc_layers.1.lin.bias", "vae.encoder.mean_encoder.weight"):
                ua = (mine[k].double() - before[k].double()).flatten()
                ub = (P[k].double() - before[k].double()).flatten()
                cos = float((ua * ub).sum() / (ua.norm() * ub.norm()).clamp_min(1e-30))
                assert cos > 0.95, (k, cos)     # Adam normalises every element to +-lr: compare directions
        model.load_state_dict({f"module.{k}": v for k, v in P.items()})     # same-state comparison next step


def _write_label_csvs(tmp_path, conds):
    os.makedirs(tmp_path / "human", exist_ok=True)
    for c, n in conds.items():
        pd.DataFrame([f"{c}_{i}" for i in range(n)]).to_csv(tmp_path / "human" / f"unique_expression_{c}.csv",
                                                            header=False, index=False)


def test_config3_step_scalars_both_species_with_adversaries():
    """BASELINE config 3 (bench.build_model(3)): human then mouse step, 4096 cells, two GRL adversaries --
    every scalar the reference logs (ELBO terms, per-condition discriminator / generator CE, all grad norms)."""
    if psutil.virtual_memory().available < 40 << 30:
        pytest.skip("the oracle needs ~20 GB of host memory at 4096 x 60 530")
    import bench
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    model, species, conds = bench.build_model(3)
    d = bench.Dims(3)
    B = 4096
    spec = oracle_spec(species, d.H1, d.H2, d.HV, d.Z, conds=conds, hidden_z=True)
    P = state_of(model)
    model.cuda().train()
    model.configure_optimizers()
    opt = {}
    rng = np.random.default_rng(3)
    for t, sp in enumerate(("human", "mouse")):
        G = species[sp]
        crow, col, val = bench.synth_csr(B, G, 0.05, seed=700 + t)
        eps = torch.randn(B, d.Z, generator=torch.Generator().manual_seed(10 + t))
        labels = {c: torch.from_numpy(rng.integers(0, n, size=B)) for c, n in conds.items()}
        ref, got = run_pair(model, spec, P, opt, sp, G, crow, col, val, eps, labels, conds,
                            dropout_masks(sp, (d.H1, d.H2), B, 0.1, 60 + t), t)
        check_scalars(got, ref["logs"], sp)
        model.load_state_dict({f"module.{k}": v for k, v in P.items()})


def test_config4_shape_step_scalars():
    """BASELINE config 4's shape: G=60 664, 8192 cells, 1024-768 | 768-512-Z256 (decoder GEMM + loss bound)."""
    if psutil.virtual_memory().available < 64 << 30:
        pytest.skip("the oracle needs ~35 GB of host memory at 8192 x 60 664")
    import bench
    from mmvae_b200 import layers as L
    L.set_precision("bf16")
    # (the core step at config 4's shape; its output discriminator has its own parity test,
    # tests/test_output_discriminator_gpu.py)
    model, species, _ = bench.build_model(4, only=("human",), discriminators=False)
    d = bench.Dims(4)
    G, B = d.G_HUMAN, 8192
    spec = oracle_spec({"human": G}, d.H1, d.H2, d.HV, d.Z)
    P = state_of(model)
    model.cuda().train()
    model.configure_optimizers()
    crow, col, val = bench.synth_csr(B, G, 0.05, seed=800)
    eps = torch.randn(B, d.Z, generator=torch.Generator().manual_seed(4))
    ref, got = run_pair(model, spec, P, {}, "human", G, crow, col, val, eps, None, None,
                        dropout_masks("human", (d.H1, d.H2), B, 0.1, 80), 0)
    check_scalars(got, ref["logs"], "config4")


def test_elbo_curve_100_steps_bf16_tracks_oracle(tmp_path):
    """ELBO curve parity (vae.py:136-152; SURVEY.md 7.6: within 1 % after 100 steps): 100 free-running steps
    (no re-synchronisation) of the bf16 fused step and of the oracle from the same initial state, on the same
    rotating batches, noise and dropout masks, lr 5e-3 as hard-coded in the reference.  Gene panel = human / 10
    so that the oracle finishes in seconds; every tile / split-K path of the kernels is active at this size."""
    from mmvae_b200 import layers as L
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Expert, Experts, FCBlockConfig, KLAnnealingFn
    L.set_precision("bf16")
    G, H1, H2, Hv, Z, B, steps = 6053, 512, 256, 128, 64, 512, 100
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    experts = Experts([Expert("human", FCBlockConfig([G, H1, H2], dropout_rate=0.1, use_batch_norm=True,
                                                     activation_fn=relu), FCBlockConfig([H2, H1, G], activation_fn=relu))])
    vae = CLVAE(FCBlockConfig([H2, Hv], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([Z, Hv, H2], activation_fn=relu), latent_dim=Z)
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    model = CMMVAEModel(CMMVAE(vae, experts, []), autograd_config=AutogradConfig(clip(), clip(), clip()),
                        kl_annealing_fn=KLAnnealingFn(1.0))
    spec = oracle_spec({"human": G}, H1, H2, Hv, Z)
    P = state_of(model)
    model.cuda().train()
    model.configure_optimizers()
    batches = [O.synth_csr(B, G, 0.05, seed=5000 + i) for i in range(8)]
    opt, curve_ref, curve_got = {}, [], []
    for t in range(steps):
        crow, col, val = batches[t % len(batches)]
        eps = torch.randn(B, Z, generator=torch.Generator().manual_seed(100 + t))
        ref, got = run_pair(model, spec, P, opt, "human", G, crow, col, val, eps, None, None,
                            dropout_masks("human", (H1, H2), B, 0.1, 200 + t), t)
        curve_ref.append([ref["logs"][k] for k in ("loss", "recon_loss", "kl_loss")])
        curve_got.append([got[k] for k in ("loss", "recon_loss", "kl_loss")])
    a, b = np.asarray(curve_got), np.asarray(curve_ref)
    dev = np.abs(a - b) / np.abs(b)
    np.save(tmp_path / "elbo_curve.npy", np.stack([a, b]))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        np.save(os.path.join(out_dir, "elbo_curve_100.npy"), np.stack([a, b]))
    print("ELBO curve: first/last oracle loss", b[0, 0], b[-1, 0], "max rel dev loss/recon/kl", dev.max(0),
          "mean dev last 10", dev[-10:].mean(0))
    assert b[-1, 0] < 0.97 * b[0, 0]                        # the run actually trains
    assert dev[:, 0].max() < 1e-2, (int(dev[:, 0].argmax()), dev[:, 0].max())          # ELBO, every step
    assert dev[:, 1].max() < 1e-2                                                     # reconstruction term
    assert dev[-10:, 2].mean() < 5e-2, dev[-10:, 2]         # KL term (O(1e-5) of the ELBO here), last 10 steps
