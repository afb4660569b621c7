"""Run the UNMODIFIED reference (zdebruine/MMVAE ``cmmvae``, installed into ``baseline/_ref`` by
``tools/install_reference.sh``) through its own public API -- ``CMMVAEModel.training_step`` on a
``torch.sparse_csr`` batch -- for ``bench.py --impl reference`` (CPU, the box's host cores) and for the
``torch_cuda_baseline`` leg (the same stock PyTorch modules on ``cuda``: cuSPARSE addmm, cuBLASLt GEMMs,
``to_dense`` + ``mse_loss``, autograd, ``clip_grad_norm_``, ``torch.optim.Adam``).

None of this repo's kernels, modules or engine is on this path.  ``lightning`` is not installed in the image, so
the reference's ``LightningModule`` base is a 25-line stand-in that provides exactly the surface the reference
touches (log, log_dict, optimizers, manual_backward, clip_gradients = ``clip_grad_norm_``, trainer flags) --
the same stand-in role as ``tests/golden/_lightning_standin.py``, minus the gradient capture.
"""
from __future__ import annotations

import os
import statistics
import sys
import tempfile
import time
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF, "cmmvae", "models", "cmmvae_model.py"))


def _install_lightning_standin():
    try:
        import lightning.pytorch  # noqa: F401
        return
    except Exception:  # noqa: BLE001
        pass

    class _Trainer:
        def __init__(self):
            self.training, self.validating, self.sanity_checking = True, False, False
            self.predicting, self.testing, self.evaluating, self.global_step = False, False, False, 0

    class LightningModule(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.trainer, self.automatic_optimization, self.logger = _Trainer(), True, None
            self.logged, self._optimizers = {}, None

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, name, value, **k):
            self.logged[name] = float(value)       # the host read Lightning's logger connector performs

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
            if gradient_clip_val is None:
                return
            assert gradient_clip_algorithm in (None, "norm")
            torch.nn.utils.clip_grad_norm_([p for g in optimizer.param_groups for p in g["params"]], gradient_clip_val)

    lightning = types.ModuleType("lightning")
    pl = types.ModuleType("lightning.pytorch")
    pl.LightningModule = LightningModule
    lightning.pytorch = pl
    sys.modules["lightning"], sys.modules["lightning.pytorch"] = lightning, pl


def import_reference():
    """``cmmvae`` from baseline/_ref (never from this repo's compat alias)"""
    if not available():
        raise RuntimeError("baseline/_ref is empty: run tools/install_reference.sh (needs /root/reference)")
    for name in [m for m in sys.modules if m == "cmmvae" or m.startswith("cmmvae.")]:
        del sys.modules[name]
    _install_lightning_standin()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import cmmvae.models  # noqa: F401
    import cmmvae
    assert os.path.realpath(cmmvae.__file__).startswith(os.path.realpath(REF)), cmmvae.__file__
    return cmmvae


def build_reference_model(dims, config: int, only=None):
    """the reference's CMMVAEModel with the topology bench.build_model(config) gives the B200 model"""
    import pandas as pd
    import_reference()
    from cmmvae.config import AutogradConfig, GradientClipConfig
    from cmmvae.models import CMMVAEModel
    from cmmvae.modules import CLVAE, CMMVAE
    from cmmvae.modules.base import Adversarial, Expert, Experts, FCBlockConfig, KLAnnealingFn
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    species = {s: g for s, g in dims.species.items() if only is None or s in only}
    H1, H2, HV, Z = dims.H1, dims.H2, dims.HV, dims.Z
    experts = Experts([Expert(s, FCBlockConfig([g, H1, H2], dropout_rate=0.1, use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([H2, H1, g], activation_fn=relu)) for s, g in species.items()])
    vae = CLVAE(FCBlockConfig([H2, HV], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([Z, HV, H2], activation_fn=relu), latent_dim=Z, hidden_z=(config == 3))
    advs, conds = [], {}
    if config == 3:
        conds = {"assay": 8, "dataset_id": 272}
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "human"))
        for c, n in conds.items():
            pd.DataFrame([f"{c}_{i}" for i in range(n)]).to_csv(
                os.path.join(tmp, "human", f"unique_expression_{c}.csv"), header=False, index=False)
        Adversarial.labels.clear()
        advs = [Adversarial(FCBlockConfig([HV, 128, 64], activation_fn=relu), FCBlockConfig([64]), list(conds), tmp),
                Adversarial(FCBlockConfig([Z, 64], activation_fn=relu), FCBlockConfig([64]), list(conds), tmp)]
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    model = CMMVAEModel(CMMVAE(vae, experts, advs), adv_weight=1.0,
                        autograd_config=AutogradConfig(clip(), clip(), clip()), kl_annealing_fn=KLAnnealingFn(1.0))
    return model, species, conds


def time_reference_steps(dims, config: int, B: int, steps: int, warmup: int, device: str = "cpu",
                         autocast: bool = False, dense_input: bool = False, budget_s: float = 170.0, density=0.05,
                         only=None):
    """median seconds per ``training_step`` of the reference on ``device``.  Returns a dict (value = cells/s).
    The run stops early once ``budget_s`` of timed work has been spent (the line reports the steps it ran)."""
    import numpy as np
    import pandas as pd
    from mmvae_b200.synth import synth_csr
    model, species, conds = build_reference_model(dims, config, only)
    model.to(device).train()
    names = list(species)
    rng = np.random.default_rng(0)
    batches, metas = {}, []
    for s, g in species.items():
        batches[s] = []
        for i in range(2):
            crow, col, val = synth_csr(B, g, density, 7000 + i)
            x = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(col), torch.from_numpy(val),
                                        size=(B, g)).to(device)
            batches[s].append(x.to_dense() if dense_input else x)
    for i in range(2):
        metas.append(pd.DataFrame({c: [f"{c}_{k}" for k in rng.integers(0, n, size=B)] for c, n in conds.items()})
                     if conds else pd.DataFrame({"cell": np.arange(B)}))
    cuda = torch.device(device).type == "cuda"
    times, spent, loss = [], 0.0, None
    for t in range(warmup + steps):
        s = names[t % len(names)]
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and cuda):
            model.training_step((batches[s][t % 2], metas[t % 2].copy(), s), t)
        if cuda:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        loss = model.logged.get(f"loss/training/{s}")
        if t >= warmup:
            times.append(dt)
            spent += dt
            if spent > budget_s:
                break
    sec = statistics.median(times)
    return {"value": B / sec, "unit": "cells/s", "ms_per_step": sec * 1e3, "steps_run": len(times),
            "warmup_run": warmup, "last_loss": loss, "device": device, "autocast_bf16": bool(autocast and cuda),
            "input": "dense" if dense_input else "sparse_csr"}
