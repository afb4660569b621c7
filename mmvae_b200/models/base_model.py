"""LightningModule base of the B200 CMMVAE models (mirror of ``cmmvae.models.base_model``; reference:
src/cmmvae/models/base_model.py -- ``tag_log_dict`` 14-48, ``BaseModel.__init__`` 60-104,
``init_weights`` 106-109, ``log_gradient_norms`` 111-123, ``stage_name`` 157-176, ``auto_log`` 265-297).

Only the training-path surface is rebuilt; prediction-file writing and TensorBoard histogram dumps of
the reference (``save_latent_predictions``, ``save_gradients`` ...) are out of scope.
"""
from __future__ import annotations

from typing import Iterable, Literal, Optional, Union

import torch

from mmvae_b200._lightning import LightningModule
from mmvae_b200.modules.base import KLAnnealingFn
import mmvae_b200.modules.base.init as init


def tag_log_dict(log_dict: dict, tags: Iterable[str] = [], sep: str = "/",
                 key_pos: Union[Literal["first"], Literal["last"]] = "first") -> dict:
    """Decorate every key with ``tags`` joined by ``sep``: ``key/tag1/tag2`` (``key_pos='first'``) or
    ``tag1/tag2/key`` (``'last'``); keys are untouched when there are no tags."""
    if key_pos not in ("first", "last"):
        raise ValueError(f"Key position {key_pos} is not supported!")
    joined = sep.join(tags)
    if not joined:
        return dict(log_dict)
    if key_pos == "first":
        return {f"{k}{sep}{joined}": v for k, v in log_dict.items()}
    return {f"{joined}{sep}{k}": v for k, v in log_dict.items()}


class BaseModel(LightningModule):
    def __init__(self, record_gradients: bool = False, save_gradients_interval: int = 25,
                 gradient_record_cap: int = 20, kl_annealing_fn: Optional[KLAnnealingFn] = None,
                 predict_dir: str = "", predict_save_interval: int = 600, initial_save_index: int = -1,
                 use_he_init_weights: bool = True):
        super().__init__()
        self.save_hyperparameters(ignore=["module"], logger=False)
        self._record_gradients = record_gradients
        self.save_gradients_interval = save_gradients_interval
        self.gradient_record_cap = gradient_record_cap
        self.predict_dir = predict_dir
        self.predict_save_interval = predict_save_interval
        self._running_predictions = []
        self._curr_save_idx = initial_save_index
        self.kl_annealing_fn = kl_annealing_fn or KLAnnealingFn(1.0)
        self._use_he_init_weights = use_he_init_weights

    def init_weights(self):
        if self._use_he_init_weights:
            init.he_init_weights(self)

    @property
    def stage_name(self) -> str:
        t = self.trainer
        for flag, name in (("training", "training"), ("validating", "validation"),
                           ("sanity_checking", "sanity_checking"), ("predicting", "prediction"),
                           ("testing", "test")):
            if getattr(t, flag, False):
                return name
        return ""

    def log_gradient_norms(self, optimizer_dict, tag_prefix="grad_norms"):
        """Log sqrt(sum_p ||grad_p||^2) per optimizer.  For the flat-buffer optimizers of this package
        that is ONE sum-of-squares launch and one host read per optimizer (the reference pays one
        ``.item()`` per parameter)."""
        for name, optimizer in optimizer_dict.items():
            if isinstance(optimizer, dict):
                self.log_gradient_norms(optimizer, f"{tag_prefix}/{name}")
                continue
            flat = getattr(optimizer, "flat", None)
            if flat is not None:
                ns = torch.zeros(1, dtype=torch.float64, device=flat.g.device)
                flat.grad_norm_sq(ns)
                total = float(ns.sqrt())
            else:
                sq = [p.grad.detach().float().pow(2).sum() for g in optimizer.param_groups for p in g["params"]
                      if p.grad is not None]
                total = float(torch.stack(sq).sum().sqrt()) if sq else 0.0
            self.log(f"{tag_prefix}/{name}", total)

    def auto_log(self, log_dict: dict, tags: Iterable[str] = [], sep: str = "/",
                 key_pos: Literal["first", "last"] = "first", log_sanity_checking: bool = False):
        if self.trainer and self.trainer.sanity_checking and not log_sanity_checking:
            return
        self.log_dict(tag_log_dict(log_dict, tags, sep, key_pos), on_step=self.trainer.training, on_epoch=True,
                      logger=True)
