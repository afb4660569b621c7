"""Building blocks (what ``cmmvae.modules.base`` exports): fully-connected blocks and their configs, the
latent encoder, species experts, conditional layers, GRL adversaries, KL schedules."""
from .annealing_fn import KLAnnealingFn, LinearKLAnnealingFn
from .components import (FCBlock, FCBlockConfig, ConcatBlockConfig, Encoder, Expert, Experts, ConditionalLayer,
                         ConditionalLayers, Adversarial, GradientReversalFunction)

__all__ = ["FCBlock", "FCBlockConfig", "ConcatBlockConfig", "Encoder", "Expert", "Experts", "ConditionalLayer",
           "ConditionalLayers", "Adversarial", "GradientReversalFunction", "KLAnnealingFn", "LinearKLAnnealingFn"]
