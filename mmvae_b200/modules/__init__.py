"""nn.Modules of the CMMVAE network (mirror of the reference's ``cmmvae.modules``)."""
from mmvae_b200.modules import base
from mmvae_b200.modules.vae import VAE
from mmvae_b200.modules.clvae import CLVAE
from mmvae_b200.modules.cmmvae import CMMVAE

__all__ = ["base", "CLVAE", "CMMVAE", "VAE"]
