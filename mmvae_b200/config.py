"""Gradient-clipping configuration (boundary objects of ``cmmvae.config``, reference src/cmmvae/config.py:4-26).

Both objects are plain records.  ``GradientClipConfig`` unpacks as ``(val, algorithm)`` so that it can be
splatted into ``clip_gradients(optimizer, *cfg)`` the way ``CMMVAEModel.training_step`` does
(cmmvae_model.py:126-129, 203-209); the fused step reads ``val`` when ``algorithm == "norm"`` and applies the
clip inside the Adam launch.
"""
from dataclasses import astuple, dataclass
from typing import Iterator, Optional, Union

_ALGORITHMS = (None, "norm", "value")


@dataclass
class GradientClipConfig:
    val: Optional[Union[int, float]] = None
    algorithm: Optional[str] = None

    def __post_init__(self):
        if self.algorithm not in _ALGORITHMS:
            raise ValueError(f"unknown clip algorithm {self.algorithm!r} (expected one of {_ALGORITHMS[1:]})")

    def __iter__(self) -> Iterator:
        return iter(astuple(self))

    def __bool__(self) -> bool:     # a config object is always "present", even with val=None
        return True


@dataclass
class AutogradConfig:
    """One optional clip per optimizer family; ``None`` leaves that family unclipped."""
    adversarial_gradient_clip: Optional[GradientClipConfig] = None
    vae_gradient_clip: Optional[GradientClipConfig] = None
    expert_gradient_clip: Optional[GradientClipConfig] = None
