"""VAE with optional conditional layers on z (mirror of ``cmmvae.modules.clvae``; reference:
src/cmmvae/modules/clvae.py -- ``CLVAE.__init__`` 31-87, ``after_reparameterize`` 89-111)."""
from __future__ import annotations

import warnings
from typing import Optional

import pandas as pd
import torch

from mmvae_b200.modules.base import ConcatBlockConfig, ConditionalLayers, FCBlockConfig
from mmvae_b200.modules.vae import VAE


class CLVAE(VAE):
    """``conditional_config`` + ``conditionals`` + ``conditionals_directory`` switch the conditional
    layers on; with ``selection_order == ["parallel"]`` their outputs are concatenated and
    ``concat_config`` describes the extra decoder layer that consumes the concatenation."""

    def __init__(self, encoder_config: FCBlockConfig, decoder_config: FCBlockConfig,
                 conditional_config: Optional[FCBlockConfig] = None, conditionals_directory: Optional[str] = None,
                 conditionals: Optional[list] = None, selection_order: Optional[list] = None,
                 concat_config: Optional[ConcatBlockConfig] = None, **encoder_kwargs):
        conditional_module = None
        if conditional_config and conditionals and conditionals_directory:
            conditional_module = ConditionalLayers(directory=conditionals_directory, conditionals=conditionals,
                                                   fc_block_config=conditional_config,
                                                   selection_order=selection_order)
        else:
            warnings.warn("No conditionals found for vae")

        if selection_order and selection_order[0] == "parallel":
            if not concat_config:
                raise RuntimeError("Please define concat_config when selection_order = parallel")
            width = len(conditional_module.selection_order) * conditional_config.layers[-1]
            decoder_config.layers = [width] + decoder_config.layers
            for option in ("activation_fn", "dropout_rate", "return_hidden", "use_layer_norm", "use_batch_norm"):
                setattr(decoder_config, option, [getattr(concat_config, option)] + getattr(decoder_config, option))

        super().__init__(encoder_config=encoder_config, decoder_config=decoder_config, **encoder_kwargs)
        self.conditionals = conditional_module

    def after_reparameterize(self, z: torch.Tensor, metadata: pd.DataFrame, **kwargs) -> torch.Tensor:
        if self.conditionals:
            return self.conditionals(z, metadata, **kwargs)
        return z
