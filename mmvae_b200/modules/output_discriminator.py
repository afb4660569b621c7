"""Output discriminator on the reconstruction (BASELINE config 4).

Mirror of the per-species networks of the reference's ``create_discriminators``
(src/cmmvae/runners/meta_discriminators.py:33-49): ``Linear(G, 128) Sigmoid Linear(128, 64) Sigmoid Linear(64, 1)
Sigmoid`` as an ``nn.Sequential`` (same sub-module indices, hence the same ``state_dict`` keys), trained with
``binary_cross_entropy(reduction='mean')`` against the species label (human 0, mouse 1; lines 125-148) by
``Adam(lr=1e-3)`` (line 103), on the DETACHED reconstruction -- the CMMVAE receives no gradient from it.

The reference trains these networks after the fact on a frozen model.  Here they train inside the fused step
(``CMMVAEModel(output_discriminators=...)``), chained off the fused decoder kernel: the first layer is another
[cells x genes] contraction over x-hat, which never exists in HBM -- it is recovered from the decoder's
``dlogits`` and the CSR batch (csrc/dense_basic.cu, ``mask_vals_by_dl``)."""
from __future__ import annotations

import torch
from torch import nn

SPECIES_LABEL = {"human": 0.0, "mouse": 1.0}


class OutputDiscriminator(nn.Sequential):
    def __init__(self, n_genes: int, hidden=(128, 64)):
        h1, h2 = hidden
        super().__init__(nn.Linear(n_genes, h1), nn.Sigmoid(), nn.Linear(h1, h2), nn.Sigmoid(), nn.Linear(h2, 1),
                         nn.Sigmoid())
        self.n_genes = int(n_genes)

    @property
    def linears(self):
        return self[0], self[2], self[4]


def create_discriminators(gene_panels: dict) -> nn.ModuleDict:
    """one discriminator per species gene panel (the reference hard-codes 52417 / 60664 genes)"""
    return nn.ModuleDict({s: OutputDiscriminator(g) for s, g in gene_panels.items()})
