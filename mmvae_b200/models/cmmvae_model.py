"""``CMMVAEModel``: the LightningModule whose ``training_step`` is the accelerated hot path (mirror of
``cmmvae.models.cmmvae_model``; reference: src/cmmvae/models/cmmvae_model.py -- ``__init__`` 40-57,
``training_step`` 138-217, ``validation_step`` 219-248, ``predict_step`` 250-264, ``get_optimizers``
267-297, ``configure_optimizers`` 299-324, ``convert_to_flat_list_and_map`` 327-351).

Same constructor, same optimizer list / ``optimizer_map``, same logged keys.  ``training_step`` hands
the CSR batch to ``mmvae_b200.engine.StepEngine`` (one fused sequence of sm_100a kernel launches, no
autograd graph, no per-parameter host syncs) and logs the reference's keys from one small host read.
"""
from __future__ import annotations

from typing import Optional

import collections
import numpy as np
import pandas as pd
import torch

from mmvae_b200 import layers as L
from mmvae_b200.config import AutogradConfig
from mmvae_b200.constants import REGISTRY_KEYS as RK
from mmvae_b200.engine import FlatAdam, StepEngine, UnsupportedTopology
from mmvae_b200.models.base_model import BaseModel
from mmvae_b200.modules import CMMVAE
from mmvae_b200.modules.base.components import Adversarial


def convert_to_flat_list_and_map(d: dict, flat_list: Optional[list] = None) -> dict:
    """Flatten a nested dict of values into ``flat_list`` (depth first, insertion order) and return the
    same nesting with list indices in place of the values."""
    if flat_list is None:
        flat_list = []
    mapping = {}
    for key, value in d.items():
        if isinstance(value, dict):
            mapping[key] = convert_to_flat_list_and_map(value, flat_list)
        else:
            mapping[key] = len(flat_list)
            flat_list.append(value)
    return mapping


class CMMVAEModel(BaseModel):
    def __init__(self, module: CMMVAE, adv_weight: Optional[float] = None,
                 autograd_config: Optional[AutogradConfig] = None, *args, output_discriminators=None,
                 output_discriminator_lr: float = 1e-3, **kwargs):
        super().__init__(*args, **kwargs)
        self.module = module
        self.output_discriminator_lr = float(output_discriminator_lr)
        self.automatic_optimization = False
        self.adversarial_criterion = torch.nn.CrossEntropyLoss(reduction="sum")
        self.init_weights()
        # extension (BASELINE config 4): per-species discriminators on the reconstruction, trained inside
        # the step on the detached x-hat (mmvae_b200.modules.OutputDiscriminator); the reference only has the post-hoc
        # runner (runners/meta_discriminators.py), whose networks live outside the LightningModule and therefore
        # keep torch's default initialisation -- attached after init_weights() for the same reason
        self.output_discriminators = torch.nn.ModuleDict(dict(output_discriminators or {}))
        self.adv_weight = adv_weight if adv_weight else 1.0   # 0/None -> 1.0, as in the reference
        self.autograd_config = autograd_config or AutogradConfig()
        self._engine: Optional[StepEngine] = None
        # True: read the step's scalars back and log them inside the same training_step (one host sync per
        # step, the reference's timing).  False: the scalars are copied to pinned host memory asynchronously
        # and logged at the NEXT training_step (or flush_logs()), so the host never waits for the GPU.
        self.sync_logging = True
        # pipelined mode + batches from mmvae_b200.feed (fixed addresses per ring slot): the step is replayed from
        # captured CUDA graphs
        self.use_cuda_graphs = True
        # False: train through the module route (the same kernels under autograd) even when the fused engine
        # covers the topology
        self.use_fused_engine = True
        self._pending_log = collections.deque()
        # pipelined mode: how many steps the host may run ahead of the step whose scalars it logs.  With 1 the host
        # waits for step t-1 right after enqueueing step t, so any host hiccup longer than the slack of one step
        # stalls the GPU -- and, data parallel, every GPU (the ranks meet at the flags each step); 2 absorbs it
        self.log_lag = 2
        self._label_ring = {}    # (n conditions, B) -> [pinned blocks, events of their last copy, next slot]

    # ------------------------------------------------------------------------------------ engine
    @staticmethod
    def _clip_val(cfg):
        if not cfg:
            return None
        val, algorithm = tuple(cfg)
        if val is None:
            return None
        if algorithm not in (None, "norm"):
            raise NotImplementedError("only clip-by-norm is implemented in the fused step")
        return float(val)

    def engine(self) -> Optional[StepEngine]:
        """The fused step engine, or None when the topology is outside it (LayerNorm / non-ReLU activations in the
        main blocks, conditional blocks other than Linear [+ LayerNorm] [+ ReLU]): those models train through the
        module route (same kernels under autograd)."""
        if self._engine is None and not self.use_fused_engine:
            self._engine = False
            self._module_route_reason = "use_fused_engine is off"
        if self._engine is None:
            ac = self.autograd_config
            clip = {"vae": self._clip_val(ac.vae_gradient_clip), "expert": self._clip_val(ac.expert_gradient_clip),
                    "adversarial": self._clip_val(ac.adversarial_gradient_clip)}
            try:
                self._engine = StepEngine(self.module, adv_weight=self.adv_weight, clip=clip,
                                          output_discriminators=dict(self.output_discriminators.items()),
                                          output_discriminator_lr=self.output_discriminator_lr)
            except UnsupportedTopology as why:
                self._engine = False
                self._module_route_reason = str(why)
        return self._engine or None

    def configure_optimizers(self, optim_cls="Adam"):
        """``[Adam(expert_0), ..., Adam(vae), Adam(adv_1), ...]`` + ``self.optimizer_map`` with the
        reference's nesting; each optimizer is the flat-buffer fused Adam of its group."""
        if optim_cls != "Adam":
            raise NotImplementedError("the fused optimizer implements torch.optim.Adam semantics only")
        eng = self.engine()
        if eng is None:
            return self._configure_optimizers_module_route()
        optim_dict = {"experts": {eid: FlatAdam(eng.groups[f"experts/{eid}"]) for eid in self.module.experts.keys()},
                      "vae": FlatAdam(eng.groups["vae"], bank=eng.cond)}
        if len(self.module.adversarials):
            optim_dict["adversarials"] = {i: FlatAdam(eng.groups[f"adversarials/{i}"])
                                          for i in range(1, len(self.module.adversarials) + 1)}
        if len(self.output_discriminators):      # after the reference's entries: existing indices stay what they are
            optim_dict["output_discriminators"] = {s: FlatAdam(eng.groups[f"output_discriminators/{s}"])
                                                   for s in self.output_discriminators.keys()}
        optimizers = []
        self.optimizer_map = convert_to_flat_list_and_map(optim_dict, optimizers)
        if hasattr(self, "_opt_cache"):      # stand-in LightningModule: what ``self.optimizers()`` hands out
            self._opt_cache = optimizers
        return optimizers

    def _configure_optimizers_module_route(self):
        """same optimizer list / map, stock Adam objects (lr 5e-3, wd 1e-6) over ordinary parameters"""
        adam = lambda params: torch.optim.Adam(params, lr=5e-3, weight_decay=1e-6)  # noqa: E731
        optim_dict = {"experts": {eid: adam(m.parameters()) for eid, m in self.module.experts.items()},
                      "vae": adam(self.module.vae.parameters())}
        if len(self.module.adversarials):
            optim_dict["adversarials"] = {i: adam(m.parameters())
                                          for i, m in enumerate(self.module.adversarials, start=1)}
        optimizers = []
        self.optimizer_map = convert_to_flat_list_and_map(optim_dict, optimizers)
        return optimizers

    def get_optimizers(self, zero_all: bool = False):
        optimizers = self.optimizers()
        if zero_all:
            for opt in optimizers:
                opt.zero_grad()

        def resolve(node):
            return {k: resolve(v) for k, v in node.items()} if isinstance(node, dict) else optimizers[node]

        return resolve(self.optimizer_map)

    def state_dict(self, *args, **kwargs):
        if self._engine:     # ZeRO-1 sharded fp32 master copies are gathered before export
            for g in self._engine.groups.values():
                g.sync_master()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, state_dict, *args, **kwargs):
        """values land in the flat fp32 buffers (the parameters are views of them); the bf16 shadows the
        kernels read are re-derived right away, so a reload before evaluation / resume never runs stale weights"""
        if self._engine:
            self._engine.finish()
        out = super().load_state_dict(state_dict, *args, **kwargs)
        if self._engine:
            for g in self._engine.groups.values():
                g.refresh_shadow()
        return out

    def _mark_stepped(self, expert_id: str, n_adv: int):
        """The fused step has already applied clip+Adam.  Call ``step()`` on the optimizers the reference steps
        in this batch (cmmvae_model.py:130,211-212: each adversary, then vae, then the expert) so Lightning's
        manual-optimization progress (``trainer.global_step``) advances as it does for the reference;
        ``FlatAdam.step`` sees the group's ``applied`` mark and does not update twice."""
        opts = self.get_optimizers()
        stepped = [o for _, o in zip(range(n_adv), (opts.get("adversarials") or {}).values())]
        stepped += [opts["vae"], opts["experts"][expert_id]]
        if expert_id in self.output_discriminators:
            stepped.append(opts["output_discriminators"][expert_id])
        for o in stepped:
            o.step()
        if hasattr(self.trainer, "set_stage"):     # the stand-in trainer counts optimizer steps like Lightning
            self.trainer.global_step += len(stepped)

    # ------------------------------------------------------------------------------------- steps
    def _labels(self, metadata: pd.DataFrame, device):
        """int64 class ids per condition: row index of each value in the human csv (class-level
        ``Adversarial.labels``; an unknown value raises KeyError like the reference's dict lookup,
        cmmvae_model.py:111-115).  The reference walks the cells in Python (one dict lookup per cell and
        condition, ~10 ms for 4096 cells) and ships one tensor per condition; here the column is factorised
        once, only its distinct values go through the dict, and all conditions travel in ONE pinned block."""
        tables = Adversarial.labels
        if not tables:
            return {}
        B = len(metadata)
        block = np.empty((len(tables), B), dtype=np.int64)
        for i, (c, table) in enumerate(tables.items()):
            codes, uniques = pd.factorize(metadata[c].values)
            if (codes < 0).any():
                raise KeyError(f"missing value in metadata column {c!r}")
            lut = np.fromiter((table[u] for u in uniques), dtype=np.int64, count=len(uniques))
            block[i] = lut[codes]
        dev = torch.device(device)
        if dev.type != "cuda":
            t = torch.from_numpy(block).to(dev)
        else:
            # a small ring of pinned blocks: the copy is asynchronous and a block is not refilled while in flight
            # (one ring per block shape; a block is refilled only after the copy that last read it has finished)
            ring = self._label_ring.get((len(tables), B))
            if ring is None:
                ring = self._label_ring[(len(tables), B)] = [
                    [torch.empty((len(tables), B), dtype=torch.int64, pin_memory=True) for _ in range(4)],
                    [None] * 4, 0]
            pins, copied, slot = ring
            ring[2] = (slot + 1) % len(pins)
            if copied[slot] is not None:
                copied[slot].synchronize()
            pins[slot].numpy()[...] = block
            t = pins[slot].to(dev, non_blocking=True)
            copied[slot] = torch.cuda.Event()
            copied[slot].record()
        return {c: t[i] for i, c in enumerate(tables)}

    @staticmethod
    def _csr(x: torch.Tensor):
        if x.layout != torch.sparse_csr:
            raise NotImplementedError(
                "the fused step consumes torch.sparse_csr batches (reference datapipe default, "
                "cellxgene_manager.py:44); densified input is outside the hot path")
        return L.csr_parts(x)

    def _adversary_losses(self, hidden, labels, expert_id, through_grl: bool):
        """sum over conditions of CE(sum) per (hidden representation, adversary) pair; logs every term"""
        from mmvae_b200.modules.base.components import GradientReversalFunction
        tag = "generator" if through_grl else "discriminator"
        losses = []
        for i, (h, adversary) in enumerate(zip(hidden, self.module.adversarials), start=1):
            h = GradientReversalFunction.apply(h, 1) if through_grl else h.detach()
            code = adversary.encoder(h)
            terms = {c: self.adversarial_criterion(adversary.heads[c](code), y) for c, y in labels.items()}
            for c, v in terms.items():
                self.auto_log({c: v}, tags=[f"{tag}_{i}", self.stage_name, expert_id, RK.ADV_LOSS], key_pos="last")
            total = torch.stack(list(terms.values())).sum()
            self.auto_log({"summed": total}, tags=[f"{tag}_{i}", self.stage_name, expert_id, RK.ADV_LOSS],
                          key_pos="last")
            losses.append(total)
        return losses

    def _training_step_module_route(self, batch) -> None:
        """training_step for topologies outside the fused engine: the nn.Modules run on the same CUDA kernels
        through autograd Functions, optimisers are stock Adam; order of operations as in the reference
        (cmmvae_model.py:138-217): discriminator update first, then the generator loss through the GRL."""
        x, metadata, expert_id = batch
        from mmvae_b200 import dp
        if dp.world_size() > 1:      # no gradient exchange on this route: replicas would silently drift apart
            raise RuntimeError("this topology trains through the module route (" + str(getattr(self, "_module_route_reason", "")) +
                               "), which is single-process; the data-parallel step needs the fused engine")
        opts = self.get_optimizers()
        vae_opt, expert_opt = opts["vae"], opts["experts"][expert_id]
        adv_opts = opts.get("adversarials") or {}
        for o in [vae_opt, expert_opt, *adv_opts.values()]:
            o.zero_grad()
        qz, pz, z, xhats, hidden = self.module(x=x, metadata=metadata, expert_id=expert_id)
        ld = self.module.vae.elbo(qz, pz, x, xhats[expert_id], self.kl_annealing_fn.kl_weight)
        ld["Mean"], ld["Variance"] = qz.mean.mean(), qz.variance.mean()
        total = ld[RK.LOSS]
        ac = self.autograd_config
        if len(self.module.adversarials):
            labels = self._labels(metadata, z.device)
            for i, (loss_i, opt_i) in enumerate(zip(self._adversary_losses(hidden, labels, expert_id, False),
                                                    adv_opts.values()), start=1):
                self.manual_backward(loss_i)
                self.log_gradient_norms({f"discriminator_{i}": opt_i}, tag_prefix="grad_norms")
                if ac.adversarial_gradient_clip:
                    self.clip_gradients(opt_i, *ac.adversarial_gradient_clip)
                opt_i.step()
                opt_i.zero_grad()
            for loss_i in self._adversary_losses(hidden, labels, expert_id, True):
                total = total + loss_i * self.adv_weight
        self.manual_backward(total)
        ld[RK.LOSS] = total
        self.log_gradient_norms({"vae": vae_opt, f"expert_{expert_id}": expert_opt}, tag_prefix="grad_norms")
        for key, o in adv_opts.items():
            self.log_gradient_norms({f"generator_{key}": o}, tag_prefix="grad_norms")
        if ac.vae_gradient_clip:
            self.clip_gradients(vae_opt, *ac.vae_gradient_clip)
        if ac.expert_gradient_clip:
            self.clip_gradients(expert_opt, *ac.expert_gradient_clip)
        vae_opt.step()
        expert_opt.step()
        self.kl_annealing_fn.step()
        self.auto_log(ld, tags=[self.stage_name, expert_id])

    def training_step(self, batch, batch_idx: int) -> None:
        x, metadata, expert_id = batch
        metadata["species"] = expert_id
        eng = self.engine()
        if eng is None:
            return self._training_step_module_route(batch)
        crow, col, val, nnz = self._csr(x)
        labels = self._labels(metadata, x.device) if len(self.module.adversarials) else None
        eng.pipeline_optimizer = not self.sync_logging    # pipelined mode: results (logs, output-layer update) trail
        eng.use_graph = eng.pipeline_optimizer and self.use_cuda_graphs
        rec = eng.train_step(expert_id, crow, col, val, nnz, self.kl_annealing_fn.kl_weight, labels=labels,
                             nnz_cap=getattr(x, "_cmmvae_cap", None),
                             metadata=metadata if eng.cond is not None else None)
        self._mark_stepped(expert_id, rec["n_adv"])
        self.kl_annealing_fn.step()
        if self.sync_logging:
            self._log_step(eng.scalars(rec), expert_id)
        else:
            self._pending_log.append((eng.scalars_async(rec), rec, expert_id))
            while len(self._pending_log) > max(1, int(self.log_lag)):
                host, prev, eid = self._pending_log.popleft()
                self._log_step(eng.scalars(prev, host=host), eid)

    def prefetch_batch(self, batch) -> None:
        """Data parallel only (no-op otherwise): call right after ``training_step`` with the NEXT batch, so that
        the exchange of its CSR records between the ranks runs underneath the step that was just enqueued.
        Without it the exchange happens at the start of the next ``training_step`` (same results, ~0.1 ms later)."""
        eng = self._engine
        if eng and eng.comm is not None:
            x, _, expert_id = batch
            crow, col, val, _ = self._csr(x)
            # batches from mmvae_b200.feed carry the event of their H2D copy: the exchange then waits for the copy
            # only, not for the step that is running
            eng.prefetch(expert_id, crow, col, val, ready=getattr(x, "_cmmvae_ready", None))

    def flush_logs(self):
        """log the scalars of the last step when ``sync_logging`` is off"""
        while self._pending_log:
            host, rec, expert_id = self._pending_log.popleft()
            self._log_step(self.engine().scalars(rec, host=host), expert_id)
        if self._engine:
            self._engine.finish()

    def _log_step(self, s: dict, expert_id: str):
        stage = self.stage_name
        main = {k: s[k] for k in (RK.LOSS, RK.RECON_LOSS, RK.KL_LOSS, RK.KL_WEIGHT, "Mean", "Variance")}
        for key, v in s.items():
            if key.startswith("grad_norms/") or key.startswith("meta_disc/"):
                self.log(key, v)
            elif "/adversarial_loss/" in key:
                tag, _, cond = key.split("/")
                self.auto_log({cond: v}, tags=[tag, stage, expert_id, RK.ADV_LOSS], key_pos="last")
        self.auto_log(main, tags=[stage, expert_id])

    def validation_step(self, batch):
        x, metadata, expert_id = batch
        if self._engine:
            self._engine.finish()
        if self.engine() is None:
            with torch.no_grad():
                qz, pz, z, xhats, _ = self.module(x, metadata, expert_id)
                loss_dict = self.module.vae.elbo(qz, pz, x, xhats[expert_id], self.kl_annealing_fn.kl_weight)
        else:
            crow, col, val, _ = self._csr(x)
            out = self.engine().eval_step(expert_id, crow, col, val, kl_weight=self.kl_annealing_fn.kl_weight,
                                          metadata=metadata)
            loss_dict = {k: out[k] for k in (RK.LOSS, RK.RECON_LOSS, RK.KL_LOSS, RK.KL_WEIGHT)}
        self.auto_log(loss_dict, tags=[self.stage_name, expert_id])
        if self.trainer.validating:
            self.log("val_loss", loss_dict[RK.LOSS], logger=False, on_epoch=True)

    test_step = validation_step

    def predict_step(self, batch, batch_idx: int):
        x, metadata, species = batch
        if self._engine:
            self._engine.finish()
        return self.module.get_latent_embeddings(x, metadata, species)
