"""Building-block nn.Modules and schedules (mirror of the reference's ``cmmvae.modules.base``)."""
from mmvae_b200.modules.base.components import (
    Adversarial,
    ConcatBlockConfig,
    ConditionalLayer,
    ConditionalLayers,
    Encoder,
    Expert,
    Experts,
    FCBlock,
    FCBlockConfig,
    GradientReversalFunction,
)
from mmvae_b200.modules.base.annealing_fn import KLAnnealingFn, LinearKLAnnealingFn

__all__ = [
    "Adversarial", "ConditionalLayer", "ConditionalLayers", "ConcatBlockConfig", "Encoder", "Expert", "Experts",
    "FCBlock", "FCBlockConfig", "GradientReversalFunction", "KLAnnealingFn", "LinearKLAnnealingFn",
]
