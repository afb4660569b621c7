"""Gradient-clipping configuration objects (reference: src/cmmvae/config.py:4-26).

``GradientClipConfig`` unpacks as ``(val, algorithm)`` so it can be splatted into
``clip_gradients(optimizer, *cfg)`` exactly like the reference's object.
"""
from typing import Optional, Union


class GradientClipConfig:
    def __init__(self, val: Optional[Union[int, float]] = None, algorithm: Optional[str] = None):
        if algorithm not in (None, "norm", "value"):
            raise ValueError(f"unknown clip algorithm {algorithm!r}")
        self.val, self.algorithm = val, algorithm

    def __iter__(self):
        yield self.val
        yield self.algorithm

    def __bool__(self):
        return True

    def __repr__(self):
        return f"GradientClipConfig(val={self.val}, algorithm={self.algorithm!r})"


class AutogradConfig:
    """Which optimizer groups get clipped (None = that group is not clipped)."""

    def __init__(self, adversarial_gradient_clip: Optional[GradientClipConfig] = None,
                 vae_gradient_clip: Optional[GradientClipConfig] = None,
                 expert_gradient_clip: Optional[GradientClipConfig] = None):
        self.adversarial_gradient_clip = adversarial_gradient_clip
        self.vae_gradient_clip = vae_gradient_clip
        self.expert_gradient_clip = expert_gradient_clip
