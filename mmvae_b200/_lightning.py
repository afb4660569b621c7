"""LightningModule base for ``mmvae_b200.models``.

When ``lightning`` (or ``pytorch_lightning``) is importable the real ``LightningModule`` is used and the
models plug into ``Trainer.fit`` exactly like the reference's.  Otherwise (this image ships neither and
has no network) a small stand-in with the same method names keeps ``training_step`` /
``validation_step`` / ``configure_optimizers`` usable from a plain Python loop (bench.py, tests).
"""
from __future__ import annotations

import torch

try:  # pragma: no cover - depends on the environment
    import lightning.pytorch as pl  # type: ignore
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        import pytorch_lightning as pl  # type: ignore
        LightningModule = pl.LightningModule
        HAVE_LIGHTNING = True
    except Exception:  # noqa: BLE001
        HAVE_LIGHTNING = False

        class _StageFlags:
            """the ``trainer`` attributes the models read"""

            def __init__(self):
                self.training, self.validating, self.sanity_checking = True, False, False
                self.predicting, self.testing, self.evaluating = False, False, False
                self.global_step = 0

            def set_stage(self, stage: str):
                for name in ("training", "validating", "sanity_checking", "predicting", "testing"):
                    setattr(self, name, name == stage)
                self.evaluating = stage in ("validating", "testing")

        class LightningModule(torch.nn.Module):  # type: ignore[no-redef]
            def __init__(self):
                super().__init__()
                self.trainer = _StageFlags()
                self.automatic_optimization = True
                self.logger = None
                self.logged_metrics = {}
                self._opt_cache = None

            def save_hyperparameters(self, *args, **kwargs):
                return None

            def log(self, name, value, **kwargs):
                self.logged_metrics[name] = value

            def log_dict(self, dictionary, **kwargs):
                self.logged_metrics.update(dictionary)

            def optimizers(self):
                if self._opt_cache is None:
                    self._opt_cache = self.configure_optimizers()
                return self._opt_cache

            def manual_backward(self, loss, *args, **kwargs):
                loss.backward(*args, **kwargs)

            def clip_gradients(self, optimizer, gradient_clip_val=None, gradient_clip_algorithm=None):
                if gradient_clip_val is None:
                    return
                params = [p for g in optimizer.param_groups for p in g["params"]]
                if gradient_clip_algorithm == "value":
                    torch.nn.utils.clip_grad_value_(params, gradient_clip_val)
                else:
                    torch.nn.utils.clip_grad_norm_(params, gradient_clip_val)
