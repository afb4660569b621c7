"""KL-weight schedules: host-side scalars handed to the step as ``kl_weight`` (boundary objects of
``cmmvae.modules.base``; behaviour of reference src/cmmvae/modules/base/annealing_fn.py:1-42)."""


class KLAnnealingFn:
    """A KL weight that stays where it was put.  ``step()`` is called once per training step
    (cmmvae_model.py:213); schedules override it."""

    def __init__(self, kl_weight: float):
        self._kl_weight = kl_weight

    def _get(self) -> float:
        return self._kl_weight

    def _set(self, value: float) -> None:
        self._kl_weight = value

    kl_weight = property(_get, _set, doc="current weight of the KL term (readable and assignable)")

    def step(self) -> None:
        pass


class LinearKLAnnealingFn(KLAnnealingFn):
    """Linear warm-up of the KL weight.

    The weight is left alone (``min_kl_weight`` unless someone assigned another value) for the first
    ``warmup_steps`` calls of ``step()``; from then on call number ``warmup_steps + k`` sets it to
    ``min + k * (max - min) / climax_steps`` clamped into ``[min, max]`` (k = 0 gives ``min``).
    ``x`` (calls so far minus the warm-up), ``m`` (slope) and ``b`` (intercept) keep the reference's names.
    """

    def __init__(self, min_kl_weight: float = 1e-7, max_kl_weight: float = 1e-5, warmup_steps: float = 1e3,
                 climax_steps: float = 1e4):
        KLAnnealingFn.__init__(self, min_kl_weight)
        self._bounds = (min_kl_weight, max_kl_weight)
        self._warmup_steps, self._climax_steps = warmup_steps, climax_steps
        self.b = min_kl_weight
        self.m = (max_kl_weight - min_kl_weight) / climax_steps
        self.x = -warmup_steps

    @property
    def _min(self) -> float:
        return self._bounds[0]

    @property
    def _max(self) -> float:
        return self._bounds[1]

    def ramp(self, k: float) -> float:
        """weight k steps after the warm-up ended"""
        lo, hi = self._bounds
        return min(hi, max(lo, self.b + self.m * k))

    def step(self) -> None:
        self.x = self.x + 1
        if self.x >= 0:
            self.kl_weight = self.ramp(self.x)
