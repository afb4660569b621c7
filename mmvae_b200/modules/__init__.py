"""The CMMVAE network as nn.Modules -- what ``cmmvae.modules`` exports: the VAE family and the
mixture-of-experts wrapper; building blocks live in ``.base``."""
from . import base
from .clvae import CLVAE
from .cmmvae import CMMVAE
from .output_discriminator import OutputDiscriminator, create_discriminators
from .vae import VAE

__all__ = ["VAE", "CLVAE", "CMMVAE", "OutputDiscriminator", "create_discriminators", "base"]
