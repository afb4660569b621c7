"""Shared test helpers: golden-case loader and comparison utilities."""
import os

import numpy as np
import torch

from oracle import cmmvae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONDITIONS = {"assay": 5, "dataset_id": 11}
COND_NAMES = ("assay", "dataset_id", "species")      # conditional layers of the human_conditional case


class GoldenCase:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
        H1, H2, Hv, Z, B = [int(v) for v in self.z["meta/dims"]]
        self.dims = dict(H1=H1, H2=H2, Hv=Hv, Z=Z, B=B)
        self.genes = {str(k): int(v) for k, v in zip(self.z["meta/genes_keys"], self.z["meta/genes_vals"])}
        self.with_adv = bool(self.z["meta/with_adv"])
        self.conditional = bool(self.z["meta/conditional"]) if "meta/conditional" in self.z.files else False
        self.adv_weight = float(self.z["meta/adv_weight"])
        self.n_steps = int(self.z["meta/n_steps"])

    def state(self, which="init"):
        """state_dict of the reference CMMVAEModel without the leading 'module.'"""
        out = {}
        pre = f"{which}/module."
        for k in self.z.files:
            if k.startswith(pre):
                out[k[len(pre):]] = torch.from_numpy(self.z[k].copy())
        return out

    def species_present(self):
        return sorted({k.split(".")[1] for k in self.state() if k.startswith("experts.")})

    def spec(self):
        d = self.dims
        experts = {}
        for s in self.species_present():
            G = self.genes[s]
            experts[s] = {
                "encoder": O.BlockSpec.make([G, d["H1"], d["H2"]], bn=True),
                "decoder": O.BlockSpec.make([d["H2"], d["H1"], G]),
            }
        advs = []
        if self.with_adv:
            for enc in ([d["Hv"], 24, 16], [d["Z"], 16]):
                advs.append(O.AdversarySpec(O.BlockSpec.make(enc), dict(CONDITIONS)))
        # make_golden.py: conditionals = [assay (shared), dataset_id (per species), species], parallel selection
        n_cond = len(COND_NAMES) if self.conditional else 1
        dec = [d["Z"], d["Hv"], d["H2"]]
        return O.ModelSpec(
            experts=experts,
            vae_encoder=O.BlockSpec.make([d["H2"], d["Hv"]], bn=True, return_hidden=True),
            vae_decoder=O.BlockSpec.make(([n_cond * d["Z"]] if n_cond > 1 else []) + dec),
            latent_dim=d["Z"], hidden_z=self.with_adv, adversarials=advs, adv_weight=self.adv_weight,
            conditionals=O.CondSpec(names=list(COND_NAMES), species_specific=["dataset_id"])
            if self.conditional else None)

    def cond(self, s, seed):
        """what the reference's ConditionalLayers.forward saw for step record ``s``: the formatted condition key of
        every row and the order it drew after ``random.seed(seed)`` (make_golden.py)"""
        import random
        if not self.conditional:
            return None
        random.seed(seed)
        order = random.sample(list(COND_NAMES), len(COND_NAMES))
        if "labels" in s:
            keys = {c: [f"{c}_{int(i)}" for i in s["labels"][c]] for c in CONDITIONS}
        else:       # the validation record: every row carries value 0 (make_golden.py)
            keys = {c: [f"{c}_0"] * self.dims["B"] for c in CONDITIONS}
        return dict(keys=keys, order=order)

    def step(self, t):
        p = f"step{t}/" if t != "val" else "val/"
        z = self.z
        rec = dict(
            species=str(z[p + "species"]), crow=z[p + "crow"], col=z[p + "col"], val=z[p + "val"],
            eps=torch.from_numpy(z[p + "eps"].copy()),
            logs={str(k): float(v) for k, v in zip(z[p + "log_keys"], z[p + "log_vals"])},
            z=z[p + "z"],
        )
        if t != "val":
            rec["kl_weight"] = float(z[p + "kl_weight_in"])
            rec["labels"] = {c: torch.from_numpy(z[p + f"labels/{c}"].copy()) for c in CONDITIONS}
            gp = p + "grad/module."
            rec["grads"] = {k[len(gp):]: z[k] for k in z.files if k.startswith(gp)}
        else:
            rec["xhat"] = z["val/xhat"]
        return rec


def untag(logs, species, stage="training"):
    """Strip the reference's tag decoration so keys compare against the oracle's plain keys:
    'loss/training/human' -> 'loss';  'discriminator_1/training/human/adversarial_loss/assay'
    -> 'discriminator_1/adversarial_loss/assay'."""
    out = {}
    for k, v in logs.items():
        parts = [p for p in k.split("/") if p not in (stage, species)]
        out["/".join(parts)] = v
    return out


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def build_b200_model(gc: "GoldenCase", tmpdir, kl_fn=None, dropout=0.0):
    """mmvae_b200 CMMVAEModel with the golden case's topology (constructor calls read like the
    reference's: same class names and keyword arguments)."""
    import pandas as pd
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import Adversarial, Expert, Experts, FCBlockConfig, KLAnnealingFn

    d = gc.dims
    os.makedirs(os.path.join(tmpdir, "human"), exist_ok=True)
    os.makedirs(os.path.join(tmpdir, "shared"), exist_ok=True)
    for cond, n in CONDITIONS.items():
        pd.DataFrame([f"{cond}_{i}" for i in range(n)]).to_csv(
            os.path.join(tmpdir, "human", f"unique_expression_{cond}.csv"), header=False, index=False)
    pd.DataFrame([f"assay_{i}" for i in range(CONDITIONS["assay"])]).to_csv(
        os.path.join(tmpdir, "shared", "unique_expression_assay.csv"), header=False, index=False)
    experts = Experts([
        Expert(id=s,
               encoder_config=FCBlockConfig(layers=[gc.genes[s], d["H1"], d["H2"]], dropout_rate=dropout,
                                            use_batch_norm=True, activation_fn=torch.nn.ReLU),
               decoder_config=FCBlockConfig(layers=[d["H2"], d["H1"], gc.genes[s]], dropout_rate=0.0,
                                            activation_fn=torch.nn.ReLU))
        for s in gc.species_present()])
    cond_kwargs = {}
    if gc.conditional:   # topology of configs/model/human_only.yaml:53-79
        from mmvae_b200.modules.base import ConcatBlockConfig
        cond_kwargs = dict(
            conditional_config=FCBlockConfig(layers=[d["Z"]], use_layer_norm=True, activation_fn=None),
            concat_config=ConcatBlockConfig(activation_fn=torch.nn.ReLU),
            conditionals_directory=str(tmpdir), conditionals=["assay", "dataset_id", "species"],
            selection_order=["parallel"])
    vae = CLVAE(encoder_config=FCBlockConfig(layers=[d["H2"], d["Hv"]], use_batch_norm=True,
                                             activation_fn=torch.nn.ReLU, return_hidden=True),
                decoder_config=FCBlockConfig(layers=[d["Z"], d["Hv"], d["H2"]], activation_fn=torch.nn.ReLU),
                latent_dim=d["Z"], hidden_z=gc.with_adv, **cond_kwargs)
    advs = []
    if gc.with_adv:
        Adversarial.labels.clear()
        for enc_layers in ([d["Hv"], 24, 16], [d["Z"], 16]):
            advs.append(Adversarial(encoder=FCBlockConfig(layers=enc_layers, activation_fn=torch.nn.ReLU),
                                    heads=FCBlockConfig(layers=[16], activation_fn=None),
                                    conditions=list(CONDITIONS), labels_dir=str(tmpdir)))
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    model = CMMVAEModel(module=CMMVAE(vae=vae, experts=experts, adversarials=advs), adv_weight=gc.adv_weight,
                        autograd_config=AutogradConfig(adversarial_gradient_clip=clip(), vae_gradient_clip=clip(),
                                                       expert_gradient_clip=clip()),
                        kl_annealing_fn=kl_fn or KLAnnealingFn(1.0))
    return model


def csr_batch(crow, col, val, n_genes, device="cuda"):
    """the reference batch format: torch.sparse_csr (cellxgene_datapipe.py:178-183)"""
    return torch.sparse_csr_tensor(torch.from_numpy(np.asarray(crow)), torch.from_numpy(np.asarray(col)),
                                   torch.from_numpy(np.asarray(val)), size=(len(crow) - 1, n_genes)).to(device)
