"""Minimal stand-in for ``lightning.pytorch`` so the UNMODIFIED reference ``cmmvae.models`` imports
and ``CMMVAEModel.training_step`` runs in the build container (lightning is not installed and
there is no network).  Used only by ``make_golden.py``; never shipped, never on the GPU box.

It provides exactly the LightningModule surface the reference touches
(src/cmmvae/models/base_model.py, cmmvae_model.py): save_hyperparameters, log, log_dict,
optimizers, manual_backward, clip_gradients and ``trainer`` stage flags.  ``clip_gradients``
restates Lightning's precision-plugin behaviour for ``gradient_clip_algorithm='norm'``:
``torch.nn.utils.clip_grad_norm_(optimizer params, clip_val)``.
"""
import sys
import types

import torch


class _Trainer:
    def __init__(self):
        self.training = True
        self.validating = False
        self.sanity_checking = False
        self.predicting = False
        self.testing = False
        self.evaluating = False
        self.global_step = 0


class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.trainer = _Trainer()
        self.automatic_optimization = True
        self.logged = {}
        self._optimizers = None
        self.pre_clip_grads = {}
        self.logger = None

    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, name, value, **k):
        self.logged[name] = float(value)

    def log_dict(self, d, **k):
        for key, v in d.items():
            self.logged[key] = float(v)

    def optimizers(self):
        if self._optimizers is None:
            self._optimizers = self.configure_optimizers()
        return self._optimizers

    def manual_backward(self, loss):
        loss.backward()

    def clip_gradients(self, optimizer, gradient_clip_val=None, gradient_clip_algorithm=None):
        params = [p for g in optimizer.param_groups for p in g["params"]]
        ids = {id(p): n for n, p in self.named_parameters()}
        for p in params:
            if p.grad is not None:
                self.pre_clip_grads[ids[id(p)]] = p.grad.detach().clone()
        if gradient_clip_val is None:
            return
        assert gradient_clip_algorithm == "norm"
        torch.nn.utils.clip_grad_norm_(params, gradient_clip_val)


def install():
    lightning = types.ModuleType("lightning")
    pl = types.ModuleType("lightning.pytorch")
    pl.LightningModule = LightningModule
    lightning.pytorch = pl
    sys.modules["lightning"] = lightning
    sys.modules["lightning.pytorch"] = pl
