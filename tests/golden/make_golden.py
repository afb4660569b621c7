"""Generate golden vectors by executing the UNMODIFIED reference (read-only /root/reference)
in the build container.  Run:  python tests/golden/make_golden.py

Writes tests/golden/<case>.npz holding, per case: the initial state_dict, every step's inputs
(CSR arrays, injected reparameterisation noise, labels, species) and the reference's outputs
(every logged scalar, z, pre-clip gradients of the first and last step, the state_dict after
the last step, and one eval-mode validation step).  The fixtures are committed; this script is
what made them.  /root/reference is NOT available on the GPU box -- tests read only the npz.
"""
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _lightning_standin  # noqa: E402

_lightning_standin.install()
sys.path.insert(0, "/root/reference/src")

import torch.distributions.normal as _tdn  # noqa: E402
from cmmvae.models import CMMVAEModel  # noqa: E402  (the reference)
from cmmvae.modules import CMMVAE, CLVAE  # noqa: E402
from cmmvae.modules.base import (  # noqa: E402
    Adversarial, Expert, Experts, FCBlockConfig, KLAnnealingFn, LinearKLAnnealingFn)
from cmmvae.config import AutogradConfig, GradientClipConfig  # noqa: E402
from oracle.cmmvae_oracle import synth_csr  # noqa: E402

DIMS = dict(G={"human": 264, "mouse": 200}, H1=64, H2=32, Hv=32, Z=16, B=24)
CONDITIONS = {"assay": 5, "dataset_id": 11}


class _Eps:
    """Inject the reparameterisation noise: Normal.rsample draws eps from
    torch.distributions.normal._standard_normal (torch/distributions/normal.py)."""

    def __init__(self):
        self.next = None

    def __call__(self, shape, dtype, device):
        assert self.next is not None and tuple(self.next.shape) == tuple(shape)
        return self.next.to(dtype=dtype, device=device)


def build(species, with_adv, labels_dir, kl_fn, adv_weight, conditional=False):
    d = DIMS
    torch.manual_seed(0)
    experts = Experts([
        Expert(
            id=s,
            encoder_config=FCBlockConfig(layers=[d["G"][s], d["H1"], d["H2"]], dropout_rate=0.0,
                                         use_batch_norm=True, activation_fn=torch.nn.ReLU),
            decoder_config=FCBlockConfig(layers=[d["H2"], d["H1"], d["G"][s]], dropout_rate=0.0,
                                         activation_fn=torch.nn.ReLU),
        ) for s in species])
    cond_kwargs = {}
    if conditional:   # the topology of configs/model/human_only.yaml:53-79 (parallel conditionals + concat layer)
        from cmmvae.modules.base import ConcatBlockConfig
        cond_kwargs = dict(
            conditional_config=FCBlockConfig(layers=[d["Z"]], use_layer_norm=True, activation_fn=None),
            concat_config=ConcatBlockConfig(activation_fn=torch.nn.ReLU),
            conditionals_directory=labels_dir, conditionals=["assay", "dataset_id", "species"],
            selection_order=["parallel"])
    vae = CLVAE(
        encoder_config=FCBlockConfig(layers=[d["H2"], d["Hv"]], use_batch_norm=True,
                                     activation_fn=torch.nn.ReLU, return_hidden=True),
        decoder_config=FCBlockConfig(layers=[d["Z"], d["Hv"], d["H2"]], activation_fn=torch.nn.ReLU),
        latent_dim=d["Z"], hidden_z=with_adv, **cond_kwargs)
    advs = []
    if with_adv:
        Adversarial.labels.clear()
        for enc_layers in ([d["Hv"], 24, 16], [d["Z"], 16]):
            advs.append(Adversarial(
                encoder=FCBlockConfig(layers=enc_layers, activation_fn=torch.nn.ReLU),
                heads=FCBlockConfig(layers=[16], activation_fn=None),
                conditions=list(CONDITIONS), labels_dir=labels_dir))
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    model = CMMVAEModel(
        module=CMMVAE(vae=vae, experts=experts, adversarials=advs),
        adv_weight=adv_weight,
        autograd_config=AutogradConfig(adversarial_gradient_clip=clip(), vae_gradient_clip=clip(),
                                       expert_gradient_clip=clip()),
        kl_annealing_fn=kl_fn)
    # give BatchNorm affine params and biases non-trivial values so parity exercises them
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("bn.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    return model


def run_case(name, species_schedule, with_adv, kl_fn, adv_weight, density=0.10, conditional=False):
    import random
    d = DIMS
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "human"))
        os.makedirs(os.path.join(tmp, "shared"))
        for cond, n in CONDITIONS.items():
            pd.DataFrame([f"{cond}_{i}" for i in range(n)]).to_csv(
                os.path.join(tmp, "human", f"unique_expression_{cond}.csv"), header=False, index=False)
        pd.DataFrame([f"assay_{i}" for i in range(CONDITIONS["assay"])]).to_csv(
            os.path.join(tmp, "shared", "unique_expression_assay.csv"), header=False, index=False)
        model = build(sorted(set(species_schedule)), with_adv, tmp, kl_fn, adv_weight, conditional)
    model.train()
    eps_src = _Eps()
    _tdn._standard_normal = eps_src
    for k, v in model.state_dict().items():
        out[f"init/{k}"] = v.detach().numpy().copy()
    captured = {}
    model.module.register_forward_hook(lambda m, i, o: captured.__setitem__("z", o[2].detach().clone()))
    rng = np.random.default_rng(123)
    n_steps = len(species_schedule)
    for t, sp in enumerate(species_schedule):
        G = d["G"][sp]
        crow, col, val = synth_csr(d["B"], G, density, seed=1000 + t)
        eps = torch.randn(d["B"], d["Z"], generator=torch.Generator().manual_seed(50 + t))
        eps_src.next = eps
        meta = pd.DataFrame({c: [f"{c}_{i}" for i in rng.integers(0, n, size=d["B"])]
                             for c, n in CONDITIONS.items()})
        x = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(col), torch.from_numpy(val),
                                    size=(d["B"], G))
        out[f"step{t}/kl_weight_in"] = np.float64(model.kl_annealing_fn.kl_weight)
        model.logged.clear()
        model.pre_clip_grads.clear()
        random.seed(4242 + t)   # ConditionalLayers shuffles its (parallel) order with python's random
        model.training_step((x, meta, sp), t)
        out[f"step{t}/species"] = np.array(sp)
        out[f"step{t}/crow"], out[f"step{t}/col"], out[f"step{t}/val"] = crow, col, val
        out[f"step{t}/eps"] = eps.numpy()
        for c in CONDITIONS:
            out[f"step{t}/labels/{c}"] = np.array([int(v.split("_")[-1]) for v in meta[c]], dtype=np.int64)
        out[f"step{t}/z"] = captured["z"].numpy()
        out[f"step{t}/log_keys"] = np.array(list(model.logged.keys()))
        out[f"step{t}/log_vals"] = np.array(list(model.logged.values()), dtype=np.float64)
        if t in (0, n_steps - 1):
            for k, g in model.pre_clip_grads.items():
                out[f"step{t}/grad/{k}"] = g.numpy()
    for k, v in model.state_dict().items():
        out[f"final/{k}"] = v.detach().numpy().copy()
    # one validation step (eval mode, running statistics)
    model.eval()
    model.trainer.training, model.trainer.validating = False, True
    sp = species_schedule[0]
    G = d["G"][sp]
    crow, col, val = synth_csr(d["B"], G, density, seed=2000)
    eps = torch.randn(d["B"], d["Z"], generator=torch.Generator().manual_seed(99))
    eps_src.next = eps
    x = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(col), torch.from_numpy(val), size=(d["B"], G))
    meta = pd.DataFrame({c: [f"{c}_0"] * d["B"] for c in CONDITIONS})
    model.logged.clear()
    with torch.no_grad():
        random.seed(777)
        model.validation_step((x, meta, sp))
        eps_src.next = eps
        random.seed(777)
        qz, pz, z, xhats, hid = model.module(x, meta, sp)
    out["val/species"] = np.array(sp)
    out["val/crow"], out["val/col"], out["val/val"], out["val/eps"] = crow, col, val, eps.numpy()
    out["val/log_keys"] = np.array(list(model.logged.keys()))
    out["val/log_vals"] = np.array(list(model.logged.values()), dtype=np.float64)
    out["val/z"] = z.numpy()
    out["val/xhat"] = xhats[sp].numpy()
    out["meta/dims"] = np.array([d["H1"], d["H2"], d["Hv"], d["Z"], d["B"]])
    out["meta/genes_keys"] = np.array(list(d["G"].keys()))
    out["meta/genes_vals"] = np.array(list(d["G"].values()))
    out["meta/with_adv"] = np.array(with_adv)
    out["meta/conditional"] = np.array(conditional)
    out["meta/adv_weight"] = np.float64(adv_weight if adv_weight else 1.0)
    out["meta/n_steps"] = np.array(n_steps)
    out["meta/torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
    print(name, "steps", n_steps, "keys", len(out))


if __name__ == "__main__":
    torch.set_num_threads(1)
    run_case("core_human", ["human", "human", "human"], False, KLAnnealingFn(0.5), None)
    run_case("two_species_adv", ["human", "mouse", "human", "mouse"], True,
             LinearKLAnnealingFn(min_kl_weight=0.1, max_kl_weight=1.0, warmup_steps=1, climax_steps=4), 2.0)
    run_case("human_conditional", ["human", "human", "human"], True, KLAnnealingFn(0.5), 1.0, conditional=True)
