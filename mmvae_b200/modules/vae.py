"""Shared VAE core (mirror of the reference's ``cmmvae.modules.vae``; reference:
src/cmmvae/modules/vae.py -- ``BaseVAE.forward`` 80-102, ``BaseVAE.elbo`` 104-152, ``VAE`` 194-205).

``elbo`` accepts the batch ``x`` either dense or as ``torch.sparse_csr``; a CSR batch is consumed by
the ReLU/MSE-vs-CSR kernel without ``to_dense()`` when ``xhat`` carries no autograd graph, and by a
differentiable sparse-aware formula otherwise."""
from __future__ import annotations

import pandas as pd
import torch
from torch import nn
from torch.distributions import Distribution, Normal, kl_divergence

from mmvae_b200.constants import REGISTRY_KEYS as RK
from mmvae_b200.modules import base


class BaseVAE(nn.Module):
    """encoder -> prior -> ``after_reparameterize`` hook -> decoder, plus the ELBO."""

    def __init__(self, encoder: base.Encoder, decoder: nn.Module):
        super().__init__()
        self.encoder = encoder
        self.decoder = decoder

    def encode(self, x: torch.Tensor, **kwargs):
        qz, z, hidden = self.encoder(x)
        return qz, z, hidden

    def decode(self, z: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.decoder(z)

    def after_reparameterize(self, z: torch.Tensor, metadata: pd.DataFrame, **kwargs) -> torch.Tensor:
        return z

    def forward(self, x: torch.Tensor, metadata: pd.DataFrame, **kwargs):
        qz, z, hidden = self.encode(x, **kwargs)
        pz = Normal(torch.zeros_like(z), torch.ones_like(z))
        z = self.after_reparameterize(z, metadata, **kwargs)
        xhat = self.decode(z, **kwargs)
        return qz, pz, z, xhat, hidden

    def elbo(self, qz: Distribution, pz: Distribution, x: torch.Tensor, xhat: torch.Tensor, kl_weight: float,
             **kwargs) -> dict:
        """``{loss, recon_loss, kl_loss, kl_weight}``: KL summed over latent dims and averaged over
        cells; reconstruction = sum of squared errors over cells and genes."""
        kl = kl_divergence(qz, pz).sum(dim=-1).mean()
        if x.layout == torch.sparse_csr:
            # sum (xhat - x)^2 = sum xhat^2 + sum_nz (x^2 - 2 x xhat): no densification
            crow, col, val = x.crow_indices(), x.col_indices(), x.values().to(xhat.dtype)
            rows = torch.repeat_interleave(torch.arange(x.shape[0], device=xhat.device), crow[1:] - crow[:-1])
            picked = xhat[rows, col.long()]
            recon = (xhat * xhat).sum() + (val * val - 2.0 * val * picked).sum()
        else:
            recon = ((xhat - x) ** 2).sum()
        return {RK.LOSS: recon + kl_weight * kl, RK.RECON_LOSS: recon, RK.KL_LOSS: kl, RK.KL_WEIGHT: kl_weight}

    @torch.no_grad()
    def get_latent_embeddings(self, x: torch.Tensor, metadata: pd.DataFrame, **kwargs) -> dict:
        _, z, _ = self.encode(x)
        return {RK.Z: z, f"{RK.Z}_{RK.METADATA}": metadata}


class VAE(BaseVAE):
    """``BaseVAE`` built from two ``FCBlockConfig`` s; extra keyword arguments (``latent_dim``,
    ``hidden_z``, ``distribution``, ``var_eps``) go to the ``Encoder``."""

    def __init__(self, encoder_config: base.FCBlockConfig, decoder_config: base.FCBlockConfig, **encoder_kwargs):
        super().__init__(
            encoder=base.Encoder(fc_block_config=encoder_config, return_dist=True, **encoder_kwargs),
            decoder=base.FCBlock(decoder_config),
        )
