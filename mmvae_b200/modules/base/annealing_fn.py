"""KL-weight schedules (host-side scalars; reference: src/cmmvae/modules/base/annealing_fn.py:1-42)."""


class KLAnnealingFn:
    """Constant KL weight; ``step()`` is a hook for schedules."""

    def __init__(self, kl_weight: float):
        self._kl_weight = kl_weight

    @property
    def kl_weight(self) -> float:
        return self._kl_weight

    @kl_weight.setter
    def kl_weight(self, value: float) -> None:
        self._kl_weight = value

    def step(self) -> None:
        return None


class LinearKLAnnealingFn(KLAnnealingFn):
    """Hold ``min_kl_weight`` for ``warmup_steps`` calls, then ramp linearly with slope
    (max-min)/climax_steps, clamped to [min, max]."""

    def __init__(self, min_kl_weight: float = 1e-7, max_kl_weight: float = 1e-5, warmup_steps: float = 1e3,
                 climax_steps: float = 1e4):
        super().__init__(min_kl_weight)
        self._min, self._max = min_kl_weight, max_kl_weight
        self._warmup_steps, self._climax_steps = warmup_steps, climax_steps
        self.m = (max_kl_weight - min_kl_weight) / climax_steps
        self.b = min_kl_weight
        self.x = -warmup_steps

    def step(self) -> None:
        self.x += 1
        if self.x >= 0:
            self.kl_weight = min(self._max, max(self._min, self.m * self.x + self.b))
