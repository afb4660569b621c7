"""Training orchestration (mirror of the reference's ``cmmvae.models``)."""
from mmvae_b200.models.base_model import BaseModel, tag_log_dict
from mmvae_b200.models.cmmvae_model import CMMVAEModel

__all__ = ["BaseModel", "CMMVAEModel", "tag_log_dict"]
