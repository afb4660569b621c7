"""LightningModule side of the drop-in (what ``cmmvae.models`` exports): the step orchestration
(``CMMVAEModel``) on top of the logging / optimizer plumbing of ``BaseModel``."""
from .cmmvae_model import CMMVAEModel
from .base_model import BaseModel, tag_log_dict

__all__ = ["CMMVAEModel", "BaseModel", "tag_log_dict"]
