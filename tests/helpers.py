"""Shared test helpers: golden-case loader and comparison utilities."""
import os

import numpy as np
import torch

from oracle import cmmvae_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONDITIONS = {"assay": 5, "dataset_id": 11}


class GoldenCase:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
        H1, H2, Hv, Z, B = [int(v) for v in self.z["meta/dims"]]
        self.dims = dict(H1=H1, H2=H2, Hv=Hv, Z=Z, B=B)
        self.genes = {str(k): int(v) for k, v in zip(self.z["meta/genes_keys"], self.z["meta/genes_vals"])}
        self.with_adv = bool(self.z["meta/with_adv"])
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
        return O.ModelSpec(
            experts=experts,
            vae_encoder=O.BlockSpec.make([d["H2"], d["Hv"]], bn=True, return_hidden=True),
            vae_decoder=O.BlockSpec.make([d["Z"], d["Hv"], d["H2"]]),
            latent_dim=d["Z"], hidden_z=self.with_adv, adversarials=advs, adv_weight=self.adv_weight)

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
