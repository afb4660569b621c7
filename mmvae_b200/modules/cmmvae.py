"""CMMVAE composition: species expert encoder -> shared VAE -> species expert decoder(s) (mirror of
``cmmvae.modules.cmmvae``; reference: src/cmmvae/modules/cmmvae.py -- ``__init__`` 36-49, ``forward``
51-113, ``get_latent_embeddings`` 115-142)."""
from __future__ import annotations

import warnings
from typing import Optional

import pandas as pd
import torch
from torch import nn

from mmvae_b200.constants import REGISTRY_KEYS as RK
from mmvae_b200.modules.base import Adversarial, Experts
from mmvae_b200.modules.clvae import CLVAE


class CMMVAE(nn.Module):
    """``forward(x, metadata, expert_id, cross_generate=False) -> (qz, pz, z, xhats, hidden)`` where
    ``xhats`` maps species -> reconstruction (only ``expert_id`` unless cross-generating).

    Unlike the reference, ``adversarials`` always exists (an empty ``ModuleList`` when none are
    given), so ``training_step`` works for ``adversarials: null`` YAMLs too."""

    def __init__(self, vae: CLVAE, experts: Experts, adversarials: Optional[list] = None):
        super().__init__()
        self.vae = vae
        self.experts = experts
        self.adversarials = nn.ModuleList([a for a in (adversarials or []) if a])

    def forward(self, x: torch.Tensor, metadata: pd.DataFrame, expert_id: str, cross_generate: bool = False):
        shared_x = self.experts[expert_id].encode(x)
        qz, pz, z, shared_xhat, hidden = self.vae(shared_x, metadata, species=expert_id)
        if cross_generate:
            if self.training:
                warnings.warn("CMMVAE is cross-generating during training, which could cause gradients to be "
                              "accumulated for cross-generation passes")
            targets = list(self.experts)
        else:
            targets = [expert_id]
        xhats = {e: self.experts[e].decode(shared_xhat) for e in targets}
        return qz, pz, z, xhats, hidden

    @torch.no_grad()
    def get_latent_embeddings(self, x: torch.Tensor, metadata: pd.DataFrame, expert_id: str) -> dict:
        _, z, _ = self.vae.encode(self.experts[expert_id].encode(x))
        metadata["species"] = expert_id
        return {RK.Z: (z, metadata)}
