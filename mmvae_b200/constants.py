"""Dictionary-key registry shared by the loss dicts, loggers and prediction writers.

Same field names and string values as the reference's ``cmmvae.constants.REGISTRY_KEYS``
(reference: src/cmmvae/constants.py:7-67) -- the strings are part of the drop-in contract
(``loss``, ``recon_loss``, ``kl_loss``, ``kl_weight``, ``adversarial_loss`` ...).
"""
from collections import namedtuple

_FIELDS = dict(
    LOSS="loss", RECON_LOSS="recon_loss", KL_LOSS="kl_loss", KL_WEIGHT="kl_weight", LABELS="labels",
    PX="px", QZ="qz", PZ="pz", QZM="qzm", QZV="qzv", Z="z", Z_STAR="z_star", X="x", xhat="xhat", Y="Y",
    METADATA="metadata", EXPERT="expert", HUMAN="human", MOUSE="mouse", ELBO="elbo", REGISTRY="registry",
    EXPERT_ID="expert_id", ADV_LOSS="adversarial_loss", ADV_WEIGHT="adverserial_weight",
    UMAP_EMBEDDINGS="umap_embeddings", PREDICT_SAMPLES="data",
    FILTER_CATEGORIES=["sex", "dev_stage", "tissue", "cell_type", "assay"],
)
REGISTRY_KEYS_NT = namedtuple("REGISTRY_KEYS_NT", list(_FIELDS), defaults=list(_FIELDS.values()))
REGISTRY_KEYS = REGISTRY_KEYS_NT()
