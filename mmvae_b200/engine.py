"""Fused CMMVAE training step on the sm_100a kernels (no autograd, no per-parameter host syncs).

``StepEngine`` restates ``CMMVAEModel.training_step`` (reference: src/cmmvae/models/cmmvae_model.py:138-217;
exact semantics in SURVEY.md Appendix A) as an explicit sequence of C-ABI kernel launches:

  forward   CSR SpMM (K1) -> BN/ReLU/dropout (K2/K3) -> tcgen05 GEMMs (K4) -> fused latent kernel (K8/K9)
            -> decoder GEMMs -> fused decoder GEMM + ReLU + sum-MSE-vs-CSR epilogue (K5-K7)
  adversary discriminator pass + clip + Adam, then generator pass through the (folded) GRL (K10/K11)
  backward  hand-written: dW/dX GEMMs (K12, MN-major operands, no transposed copies), BN/ReLU/dropout
            backward, CSC gather for the sparse weight gradient (K1b)
  optimiser one sum-of-squares + one clip+Adam launch per optimizer group over flat buffers (K13-K15),
            clip coefficient read on the device; bf16 shadows refreshed by the same launch
  DP        gene-sharded first / last layer over NVLink peer memory (no NCCL on the data path): see
            ``_train_step`` (data-parallel route) and csrc/peer.cu

Parameters of each optimizer group live in one flat fp32 buffer (``FlatGroup``); the nn.Parameters of
the modules are views into it, so ``state_dict()`` keeps the reference's names and shapes.  The first
expert-encoder weight is stored physically transposed ([genes, hidden]) for coalesced SpMM row reads
and exposed as a ``[hidden, genes]`` view.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import dp
from . import layers as L
from . import ops


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class FlatGroup:
    """One optimizer group (``experts/<id>``, ``vae`` or ``adversarials/<i>``) as flat buffers.

    Layout: ``[row-sharded parameters ... | tail]``.  With one process the distinction is moot.  With ``world``
    ranks the *row-sharded* parameters -- the two gene-sized matrices of an expert (first-layer weight, stored
    ``[genes, hidden]``, and output-layer weight ``[genes, hidden]``) and the output bias ``[genes]`` -- are
    owned by GENE RANGE: rank r holds the current value, the gradient and the Adam state of rows
    ``[r * per, (r + 1) * per)`` only (``per`` = rows per rank, 128-aligned, the matrix padded to ``world * per``
    rows).  Nothing of them is ever exchanged: the step computes each rank's rows of the SUMMED gradient locally
    (engine, data-parallel route).  The full-size fp32 buffer is kept so that ``state_dict()`` has its usual shapes;
    rows of other ranks are refreshed by ``sync_master()`` only.  The *tail* (biases, BatchNorm affine, small
    matrices) is replicated: gradients are summed over ranks, identical Adam on every rank."""

    ALIGN = 64

    def __init__(self, name: str, chains: List[List[tuple]], device, lr=5e-3, weight_decay=1e-6,
                 betas=(0.9, 0.999), eps=1e-8, sharded: Optional[List[tuple]] = None, world: int = 1, rank: int = 0,
                 background_last: bool = True):
        """``chains`` (tail): lists of ``(param, transposed)`` stored back to back (e.g. mean/var head weights =
        one matrix).  ``sharded``: ``(param, transposed)`` whose physical rows are gene rows."""
        self.name, self.lr, self.wd, self.betas, self.eps = name, lr, weight_decay, betas, eps
        self.world, self.rank = int(world), int(rank)
        self.params: List[nn.Parameter] = []
        self.offset: Dict[int, int] = {}
        self.transposed: Dict[int, bool] = {}
        self.rows_pad: Dict[int, int] = {}      # id(param) -> padded rows of a row-sharded parameter
        sharded = sharded or []
        self.sharded = self.world > 1 and len(sharded) > 0
        self.per = 0
        total = 0
        self.shard_blocks: List[tuple] = []     # (flat lo of the padded block, rows_pad, width)
        for p, tr in sharded:
            rows, width = self._rows_width(p, tr)
            total = _ceil(total, 256)
            self.params.append(p)
            self.offset[id(p)] = total
            self.transposed[id(p)] = tr
            if self.world > 1:
                per = dp.shard_rows(rows, self.world)
                assert self.per in (0, per), "row-sharded parameters of a group share the gene axis"
                self.per = per
                rows_p = self.world * per
            else:
                rows_p = rows
            self.rows_pad[id(p)] = rows_p
            self.shard_blocks.append((total, rows_p, width))
            total += rows_p * width
        total = _ceil(total, 256)
        self.tail_lo = total
        for chain in chains:
            total = _ceil(total, self.ALIGN)
            for p, tr in chain:
                self.params.append(p)
                self.offset[id(p)] = total
                self.transposed[id(p)] = tr
                total += p.numel()
        self.n = _ceil(max(total, 4), 4)
        self.p = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.g = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.p16 = torch.zeros(self.n, device=device, dtype=torch.bfloat16)
        # ranges the optimizer walks on this rank: (flat lo, flat hi, offset into m/v)
        self.ranges: List[tuple] = []
        mv = 0
        if self.sharded:
            for lo, rows_p, width in self.shard_blocks:
                own = lo + self.rank * self.per * width
                n_own = self.per * width
                self.ranges.append((own, own + n_own, mv))
                mv += _ceil(n_own, 4)
            self.ranges.append((self.tail_lo, self.n, mv))
            mv += self.n - self.tail_lo
        else:
            self.ranges.append((0, self.n, 0))
            mv = self.n
        self.m = torch.zeros(mv, device=device, dtype=torch.float32)
        self.v = torch.zeros(mv, device=device, dtype=torch.float32)
        # single process: the LAST row-sharded block (the output layer, which the next forward reads last) may be
        # updated on a background stream underneath the next step; data parallel: same, for this rank's rows
        self.bg_range = (len(self.shard_blocks) - 2) if (background_last and len(self.shard_blocks) >= 2) else None
        self.step_count = 0
        self.applied = False    # set by the fused step after its own clip+Adam launch; consumed by FlatAdam.step()
        self.bc_dev: Optional[torch.Tensor] = None   # device float[2] with this step's Adam bias corrections (graph mode)
        self.dyn_active = False                      # the engine is in graph mode: read bc_dev instead of step_count
        self._bg_snap = None
        self.master_dirty = False
        self._vec_range: Optional[tuple] = None
        self._deferred: Optional[torch.cuda.Event] = None   # output-layer update in flight on the background stream
        for p in self.params:
            phys = self.phys(p)
            src = p.data.to(device)
            phys.copy_(src.t() if self.transposed[id(p)] else src)
            p.data = phys.t() if self.transposed[id(p)] else phys
            gphys = self.phys(p, self.g)
            p.grad = gphys.t() if self.transposed[id(p)] else gphys
            L.SHADOWS[id(p)] = self.phys(p, self.p16)
        self.refresh_shadow()

    @staticmethod
    def _rows_width(p, tr):
        if p.dim() == 1:
            return p.shape[0], 1
        return (p.shape[1], p.shape[0]) if tr else (p.shape[0], p.shape[1])

    def zero_vector_grads(self):
        """one fill per step over the gradients of the 1-D parameters (biases, BatchNorm affine): the kernels
        that produce them accumulate, instead of each issuing its own memsets in the middle of the backward
        pass.  Matrix gradients are overwritten by their GEMMs and need no zeroing."""
        if self._vec_range is None:
            vec = [(self.offset[id(p)], self.offset[id(p)] + self.rows_pad.get(id(p), p.numel()))
                   for p in self.params if p.dim() == 1]
            if not vec:
                self._vec_range = (0, 0)
            elif self.shard_blocks:     # vectors are contiguous: [sharded bias | tail vectors first]
                self._vec_range = (min(a for a, _ in vec), max(b for _, b in vec))
            else:                       # small group with interleaved layout: clear everything
                self._vec_range = (0, self.n)
        lo, hi = self._vec_range
        if hi > lo:
            self.g[lo:hi].zero_()

    def phys(self, p: nn.Parameter, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """physical (row-major, contiguous) view of parameter ``p`` inside ``buf`` (default: values)"""
        buf = self.p if buf is None else buf
        o = self.offset[id(p)]
        shape = tuple(p.shape)
        if self.transposed[id(p)]:
            shape = shape[::-1]
        return buf[o:o + p.numel()].view(shape)

    def phys_padded(self, p: nn.Parameter, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """row-sharded parameter incl. its padding rows: ``[rows_pad, width]`` (``[rows_pad]`` for a vector)"""
        buf = self.p if buf is None else buf
        o = self.offset[id(p)]
        rows_p = self.rows_pad[id(p)]
        rows, width = self._rows_width(p, self.transposed[id(p)])
        t = buf[o:o + rows_p * width]
        return t if p.dim() == 1 else t.view(rows_p, width)

    def own_rows(self, p: nn.Parameter, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """this rank's rows ``[rank * per, (rank + 1) * per)`` of a row-sharded parameter"""
        t = self.phys_padded(p, buf)
        if not self.sharded:
            return t
        return t[self.rank * self.per:(self.rank + 1) * self.per]

    def refresh_shadow(self):
        ops.cast_bf16(self.p, self.p16)

    def grad_norm_sq(self, out: torch.Tensor, skip: Optional[List[nn.Parameter]] = None, shard_out=None):
        """out (double[1], pre-zeroed) += || sum_r g_r ||^2 over the replicated part (single process: the whole
        group; padding is zero).  ``skip``: row-sharded parameters whose contribution a producer kernel already
        added.  Data parallel: the sum of squares of THIS rank's rows goes to ``shard_out`` (the engine adds the
        other ranks' shares after the scalar exchange)."""
        skip_ids = {id(p) for p in (skip or [])}
        if self.sharded:
            for (lo, hi, _), p in zip(self.ranges, [q for q in self.params if id(q) in self.rows_pad]):
                if id(p) not in skip_ids:
                    ops.sumsq(self.g[lo:hi], shard_out)
            ops.sumsq(self.g[self.tail_lo:self.n], out)
            return
        if skip_ids:
            cuts = sorted((self.offset[id(p)], self.offset[id(p)] + p.numel()) for p in skip)
            lo = 0
            for a, b in cuts + [(self.n, self.n)]:
                a4, lo4 = a // 4 * 4, _ceil(lo, 4)   # ranges start/stop on parameter boundaries (64-aligned)
                if a4 > lo4:
                    ops.sumsq(self.g[lo4:a4], out)
                lo = b
            return
        ops.sumsq(self.g, out)

    def advance(self):
        """host-side bookkeeping of one optimizer step (what ``clip_adam`` does besides launching)"""
        self.step_count += 1
        self.applied = True
        self.master_dirty = True

    def bias_corrections(self):
        t = self.step_count
        return 1.0 - self.betas[0] ** t, 1.0 - self.betas[1] ** t

    def clip_adam(self, norm_sq: torch.Tensor, max_norm: Optional[float], grad_scale: float = 1.0,
                  background: Optional[torch.cuda.Stream] = None, defer_background: bool = False, advance: bool = True):
        """fused clip + Adam over the ranges this rank owns.  ``background``: the output layer's range (which the
        next forward pass reads last) is updated on that low-priority stream, after everything else, so the update
        runs underneath the next step's forward; ``join_background()`` joins it."""
        if advance:
            self.advance()               # (rows owned by other ranks are now stale in this rank's fp32 buffer)
        hyper = (self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count)
        bc = self.bc_dev if self.dyn_active else None
        if self.sharded:
            todo = list(self.ranges)
            bg = todo.pop(len(self.shard_blocks) - 2) if (background is not None and self.bg_range is not None) else None
        elif background is not None and self.bg_range is not None:
            lo1, rows_p, width = self.shard_blocks[self.bg_range]
            hi1 = _ceil(lo1 + rows_p * width, 4)
            todo = [(0, lo1, 0), (hi1, self.n, hi1)]
            bg = (lo1, hi1, lo1)
        else:
            todo, bg = list(self.ranges), None
        for lo, hi, mv in todo:
            n = hi - lo
            if n > 0:
                ops.clip_adam(self.p[lo:hi], self.g[lo:hi], self.m[mv:mv + n], self.v[mv:mv + n], self.p16[lo:hi],
                              norm_sq, max_norm or 0.0, grad_scale, *hyper, bc_dev=bc)
        if bg is None:
            return None
        lo, hi, mv = bg
        n = hi - lo
        if bc is not None:
            # graph mode: the scalar block and the bias corrections sit at fixed addresses that the NEXT step's first
            # graph rewrites while this background update may still be running -- it reads private copies instead
            # (rewritten only after join_background() of the next step)
            if self._bg_snap is None:
                self._bg_snap = (torch.zeros(1, dtype=torch.float64, device=self.p.device),
                                 torch.zeros(2, dtype=torch.float32, device=self.p.device))
            self._bg_snap[0].copy_(norm_sq)
            self._bg_snap[1].copy_(bc)
            norm_sq, bc = self._bg_snap

        def launch_background():
            first_done = torch.cuda.Event()
            first_done.record()
            # the clip norm lives in the step's scalar block: keep the allocator from recycling that block for a
            # later step while the background launch may still read it
            norm_sq.record_stream(background)
            with torch.cuda.stream(background), ops.stream_scope(background):
                background.wait_event(first_done)
                ops.clip_adam(self.p[lo:hi], self.g[lo:hi], self.m[mv:mv + n], self.v[mv:mv + n], self.p16[lo:hi],
                              norm_sq, max_norm or 0.0, grad_scale, *hyper, background=True, bc_dev=bc)
                self._deferred = torch.cuda.Event()
                self._deferred.record(background)

        if defer_background:      # graph capture: the caller launches it after the capture has ended
            return launch_background
        launch_background()
        return None

    def join_background(self):
        """make the current stream wait for an output-layer update still running on the background stream"""
        if self._deferred is not None:
            torch.cuda.current_stream().wait_event(self._deferred)
            self._deferred = None

    def logical(self, p: nn.Parameter, buf: torch.Tensor) -> torch.Tensor:
        """view of ``buf`` (laid out like the value buffer) with parameter ``p``'s logical shape"""
        t = self.phys(p, buf)
        return t.t() if self.transposed[id(p)] else t

    def _gather_rows(self, own: torch.Tensor, full: torch.Tensor):
        """full[world * n] <- all ranks' own[n] (rare: state_dict / checkpoint time; torch.distributed plumbing)"""
        if torch.distributed.get_backend() == "nccl":
            torch.distributed.all_gather_into_tensor(full, own.contiguous())
        else:
            parts = [torch.empty(own.numel(), dtype=own.dtype) for _ in range(self.world)]
            torch.distributed.all_gather(parts, own.detach().cpu().contiguous())
            full.copy_(torch.cat(parts))

    def full_moments(self):
        """Adam moments laid out like the value buffer (``[n]`` each).  Single process: the live buffers.
        Gene-sharded: fresh full-size copies, the row-sharded blocks all-gathered from their owners."""
        if not self.sharded:
            return self.m, self.v
        out = []
        for src in (self.m, self.v):
            full = torch.zeros(self.n, device=src.device, dtype=torch.float32)
            for (lo, rows_p, width), (own_lo, own_hi, mv) in zip(self.shard_blocks, self.ranges):
                self._gather_rows(src[mv:mv + own_hi - own_lo], full[lo:lo + rows_p * width])
            _, _, mv = self.ranges[-1]
            full[self.tail_lo:self.n].copy_(src[mv:mv + self.n - self.tail_lo])
            out.append(full)
        return out

    def store_moments(self, m_full: torch.Tensor, v_full: torch.Tensor):
        """inverse of ``full_moments`` (no-op single process: the views were written in place)"""
        if not self.sharded:
            return
        for src, dst in ((m_full, self.m), (v_full, self.v)):
            for (own_lo, own_hi, mv) in self.ranges:
                dst[mv:mv + own_hi - own_lo].copy_(src[own_lo:own_hi])

    def sync_master(self):
        """refresh the rows of the row-sharded fp32 parameters that other ranks own (before state_dict / fp32
        evaluation); also joins a background update"""
        self.join_background()
        if not self.sharded or not self.master_dirty:
            return
        self.master_dirty = False
        for (lo, rows_p, width), (own_lo, own_hi, _) in zip(self.shard_blocks, self.ranges):
            own = self.p[own_lo:own_hi].clone()
            self._gather_rows(own, self.p[lo:lo + rows_p * width])
        self.refresh_shadow()


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Optimizer`` face of a ``FlatGroup`` (what ``configure_optimizers`` returns, one per
    reference optimizer: Adam(lr=5e-3, weight_decay=1e-6), cmmvae_model.py:306-319).  ``step`` runs the
    fused clip+Adam launch; a clip value set through ``set_clip`` is applied inside that launch.

    The fused ``training_step`` launches clip+Adam itself and then calls ``step()`` on the (Lightning-wrapped)
    optimizers it updated: the group is marked ``applied`` and ``step`` only consumes the mark, so Lightning's
    progress tracking (``trainer.global_step``, ``max_steps``, step-based checkpoints) advances exactly as
    with the reference's ``optimizer.step()`` calls, without a second update.

    ``state_dict`` / ``load_state_dict`` speak torch.optim.Adam's format (``step``, ``exp_avg``,
    ``exp_avg_sq`` per parameter, in the parameter's logical shape), so checkpoints resume with their moments."""

    def __init__(self, group: FlatGroup, bank=None):
        """``bank``: a ``mmvae_b200.conditional.CondBank`` stepped with this group (the conditional layers belong to
        the VAE's optimizer, cmmvae_model.py:311); its parameters follow the group's, each slot with its own step"""
        self.flat = group
        self.bank = bank
        self._max_norm = None
        super().__init__(group.params + (bank.params if bank is not None else []),
                         dict(lr=group.lr, weight_decay=group.wd, betas=group.betas, eps=group.eps))

    def set_clip(self, max_norm: Optional[float]):
        self._max_norm = max_norm

    @torch.no_grad()
    def step(self, closure=None):
        if self.flat.applied:
            self.flat.applied = False
            return
        if self.flat.world > 1:
            raise NotImplementedError("with torch.distributed the exchange + step run inside training_step")
        if self.bank is not None:
            raise NotImplementedError("conditional layers are stepped inside training_step (values present in the "
                                      "batch only)")
        ns = torch.zeros(1, dtype=torch.float64, device=self.flat.p.device)
        self.flat.grad_norm_sq(ns)
        self.flat.clip_adam(ns, self._max_norm)
        self.flat.applied = False
        self._max_norm = None

    def zero_grad(self, set_to_none: bool = True):
        self.flat.g.zero_()

    # ---- checkpoint format of torch.optim.Adam ----
    def state_dict(self):
        g = self.flat
        m_full, v_full = g.full_moments()
        state = {}
        if g.step_count > 0:
            for i, p in enumerate(g.params):
                state[i] = {"step": torch.tensor(float(g.step_count)),
                            "exp_avg": g.logical(p, m_full).clone(), "exp_avg_sq": g.logical(p, v_full).clone()}
        n_all = len(g.params)
        if self.bank is not None:
            state.update(self.bank.state_entries(len(g.params)))
            n_all += len(self.bank.params)
        pg = dict(self.param_groups[0])
        pg["params"] = list(range(n_all))
        return {"state": state, "param_groups": [pg]}

    def load_state_dict(self, sd):
        g = self.flat
        state = sd.get("state", {})
        m_full, v_full = g.full_moments()
        steps = set()
        for i, p in enumerate(g.params):
            st = state.get(i, state.get(str(i)))
            if st is None:
                g.logical(p, m_full).zero_()
                g.logical(p, v_full).zero_()
                continue
            g.logical(p, m_full).copy_(st["exp_avg"])
            g.logical(p, v_full).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FlatAdam steps a whole group together; checkpoint holds steps {sorted(steps)}")
        g.step_count = steps.pop() if steps else 0
        g.store_moments(m_full, v_full)
        if self.bank is not None:
            self.bank.load_state_entries(state, len(g.params))
        for k, v in (sd.get("param_groups") or [{}])[0].items():
            if k in ("lr", "weight_decay", "betas", "eps"):
                self.param_groups[0][k] = v
        pg = self.param_groups[0]
        g.lr, g.wd, g.betas, g.eps = pg["lr"], pg["weight_decay"], tuple(pg["betas"]), pg["eps"]


@dataclass
class LayerPlan:
    """one lin[/bn][/relu][/dropout] layer bound to its flat-buffer views"""
    lin: nn.Linear
    bn: Optional[nn.BatchNorm1d]
    relu: bool
    p_drop: float
    group: FlatGroup
    sparse: bool = False
    K: int = 0
    N: int = 0
    W32: torch.Tensor = None   # physical: [N,K] dense, [G,H] sparse
    W16: torch.Tensor = None
    gW: torch.Tensor = None
    b: torch.Tensor = None
    gb: torch.Tensor = None
    gamma: torch.Tensor = None
    beta: torch.Tensor = None
    ggamma: torch.Tensor = None
    gbeta: torch.Tensor = None
    return_hidden: bool = False


class UnsupportedTopology(NotImplementedError):
    pass


def _plan_block(block, group: FlatGroup, sparse_first=False) -> List[LayerPlan]:
    plans = []
    for i, layer in enumerate(block.fc_layers):
        parts = dict(layer.named_children())
        if "ln" in parts:
            raise UnsupportedTopology("LayerNorm layers are outside the fused step")
        af = parts.get("af")
        if af is not None and type(af) is not nn.ReLU:
            raise UnsupportedTopology(f"activation {type(af).__name__} is outside the fused step")
        lin, bn, dr = parts["lin"], parts.get("bn"), parts.get("dr")
        lp = LayerPlan(lin=lin, bn=bn, relu=af is not None, p_drop=float(dr.p) if dr is not None else 0.0,
                       group=group, sparse=(sparse_first and i == 0), K=lin.in_features, N=lin.out_features,
                       return_hidden=bool(block.config.return_hidden[i]))
        lp.W32, lp.W16, lp.gW = group.phys(lin.weight), group.phys(lin.weight, group.p16), group.phys(lin.weight, group.g)
        lp.b, lp.gb = group.phys(lin.bias), group.phys(lin.bias, group.g)
        if bn is not None:
            lp.gamma, lp.beta = group.phys(bn.weight), group.phys(bn.bias)
            lp.ggamma, lp.gbeta = group.phys(bn.weight, group.g), group.phys(bn.bias, group.g)
        plans.append(lp)
    return plans


def _block_chains(block, sparse_first=False, kind="all"):
    """chains of a block; kind: "all", "matrix" (>= 2-D params) or "vector" (1-D params)"""
    chains = []
    for i, layer in enumerate(block.fc_layers):
        for name, p in layer.named_parameters():
            if kind == "matrix" and p.dim() < 2 or kind == "vector" and p.dim() >= 2:
                continue
            chains.append([(p, sparse_first and i == 0 and name == "lin.weight")])
    return chains


@dataclass
class AdvPlan:
    enc: List[LayerPlan]
    conditions: List[str]
    classes: List[int]
    Wh32: torch.Tensor = None   # [sumC, K]
    Wh16: torch.Tensor = None
    bh: torch.Tensor = None
    gWh: torch.Tensor = None
    gbh: torch.Tensor = None
    group: FlatGroup = None


class StepEngine:
    def __init__(self, module, adv_weight: float = 1.0, clip: Optional[Dict[str, Optional[float]]] = None,
                 precision: Optional[str] = None, device=None, output_discriminators=None,
                 output_discriminator_lr: float = 1e-3):
        self.module = module
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("StepEngine needs a CUDA device (no CPU fallback)")
        self.precision = precision or L.get_precision()
        # data parallel (one process per GPU): peer-memory exchange, gene-sharded first / last layer.
        # CMMVAE_FORCE_DP=1 runs that route with a single process (its own "peer"): the single-GPU test of the route
        self.comm = None
        self.world, self.rank = 1, 0
        if dp.world_size() > 1 or os.environ.get("CMMVAE_FORCE_DP") == "1":
            if self.precision != "bf16":
                raise RuntimeError("the data-parallel route runs the bf16 policy only (tensor-pipe SpMM, fused decoder)")
            from .peer import PeerComm
            self.comm = PeerComm(self.device)
            self.world, self.rank = self.comm.world, self.comm.rank
        self._dp = None          # symmetric buffers, allocated on the first step (needs B and nnz)
        self._dp_step = 0
        self._csr_pushed = {}    # (pointers, nnz) of a prefetched batch -> step it was pushed for
        # single process: run the step on a high-priority stream and the output layer's clip+Adam on a
        # low-priority one, underneath the next step's forward pass (37 % of a B=1024 step is optimizer HBM
        # traffic, half of it the output layer whose new value is only needed by the decoder kernel).  Weights
        # read outside the engine need ``finish()`` first, hence opt-in (training_step turns it on in its
        # pipelined mode, sync_logging=False)
        self.pipeline_optimizer = False
        self._hp = self._bg = None
        # CUDA-graph mode (single process, pipelined mode): a step is ~70 launches of this library, most of them a
        # few microseconds of GPU work -- replaying them from two captured graphs removes the launch-bound gaps
        # between the small kernels and nearly all host work.  Per-step scalars (dropout seed, KL weight, Adam bias
        # corrections, the batch's nnz) are read from device memory by the replayed launches (``_dyn`` block).
        self.use_graph = False
        self.graph_timers: Optional[set] = None    # section names timed INSIDE the captured graphs (bench roofline)
        self._graphs: Dict[tuple, dict] = {}
        self._dyn = None
        self._gmode = None
        self.adv_weight = adv_weight
        self.clip = clip or {"vae": 10.0, "expert": 10.0, "adversarial": 10.0}
        vae = module.vae
        # conditional layers on z: one Linear [+ LayerNorm] block per metadata value, run by
        # mmvae_b200.conditional.CondBank (grouped by value on the host, one launch per direction, Adam on the
        # values present in the batch only).  Built below, once the module sits on the device.
        cond = getattr(vae, "conditionals", None) or None
        if cond is not None and self.comm is not None:
            raise UnsupportedTopology("conditional layers are not part of the data-parallel route (SURVEY.md 8f-1)")
        self.cond = None
        enc = vae.encoder
        if isinstance(enc.z_transformation, nn.Softmax):
            raise UnsupportedTopology("distribution='ln' is outside the fused step")
        module.to(self.device)   # buffers (BN running stats) and not-yet-flattened params
        dev = self.device
        if cond is not None:
            from .conditional import CondBank, CondUnsupported
            try:
                self.cond = CondBank(cond, dev)
            except CondUnsupported as why:
                raise UnsupportedTopology(str(why))
        # ---- optimizer groups as flat buffers (order mirrors configure_optimizers) ----
        self.groups: Dict[str, FlatGroup] = {}
        self.enc_plan: Dict[str, List[LayerPlan]] = {}
        self.dec_plan: Dict[str, List[LayerPlan]] = {}
        for eid, expert in module.experts.items():
            # row-sharded by genes when data parallel: first-layer weight (stored [genes, hidden]), output-layer
            # weight [genes, hidden] and output bias [genes]; replicated tail: vectors first, then the small matrices
            w1 = expert.encoder.fc_layers[0].lin.weight
            last_lin = expert.decoder.fc_layers[len(expert.decoder.fc_layers) - 1].lin
            big = {id(w1), id(last_lin.weight), id(last_lin.bias)}
            mats = [c for c in _block_chains(expert.encoder, kind="matrix") + _block_chains(expert.decoder, kind="matrix")
                    if id(c[0][0]) not in big]
            vecs = [c for c in _block_chains(expert.encoder, kind="vector") + _block_chains(expert.decoder, kind="vector")
                    if id(c[0][0]) not in big]
            g = FlatGroup(f"experts/{eid}", vecs + mats, dev,
                          sharded=[(w1, True), (last_lin.weight, False), (last_lin.bias, False)],
                          world=self.world, rank=self.rank)
            self.groups[f"experts/{eid}"] = g
            self.enc_plan[eid] = _plan_block(expert.encoder, g, sparse_first=True)
            self.dec_plan[eid] = _plan_block(expert.decoder, g)
            last = self.dec_plan[eid][-1]
            if not last.relu or last.bn is not None or last.p_drop > 0:
                raise UnsupportedTopology("fused decoder loss expects Linear+ReLU as the output layer")
        chains = _block_chains(enc.fc)
        chains.append([(enc.mean_encoder.weight, False), (enc.var_encoder.weight, False)])
        chains.append([(enc.mean_encoder.bias, False), (enc.var_encoder.bias, False)])
        chains += _block_chains(vae.decoder)
        listed = {id(p) for c in chains for p, _ in c}
        listed |= {id(p) for p in (self.cond.params if self.cond else [])}
        extra = [p for p in vae.parameters() if id(p) not in listed]
        if extra:
            raise UnsupportedTopology("VAE has parameters outside encoder/decoder")
        gv = self.groups["vae"] = FlatGroup("vae", chains, dev)
        self.vaeenc_plan = _plan_block(enc.fc, gv)
        for lp in self.vaeenc_plan:
            # the reference hands the adversary the activation right after ``af``, BEFORE ``dr``
            # (components.py:309-313); the fused step takes the layer output, so the two only agree without dropout
            if lp.return_hidden and lp.relu and lp.p_drop > 0:
                raise UnsupportedTopology("return_hidden on a VAE-encoder layer with dropout is outside the fused step")
        self.vaedec_plan = _plan_block(vae.decoder, gv)
        self.Z = enc.mean_encoder.out_features
        self.Hv = enc.mean_encoder.in_features
        o = gv.offset[id(enc.mean_encoder.weight)]
        n = 2 * self.Z * self.Hv
        self.Wmv32, self.Wmv16, self.gWmv = (b[o:o + n].view(2 * self.Z, self.Hv) for b in (gv.p, gv.p16, gv.g))
        o = gv.offset[id(enc.mean_encoder.bias)]
        self.bmv, self.gbmv = gv.p[o:o + 2 * self.Z], gv.g[o:o + 2 * self.Z]
        self.var_eps = float(enc.var_eps)
        self.hidden_z = bool(enc.hidden_z)
        self.n_hidden = sum(1 for lp in self.vaeenc_plan if lp.return_hidden and lp.relu) + int(self.hidden_z)
        # ---- adversaries ----
        self.adv: List[AdvPlan] = []
        for i, adv in enumerate(module.adversarials):
            conds = list(adv.heads.keys())
            heads = [adv.heads[c] for c in conds]
            for h in heads:
                if len(h.fc_layers) != 1 or len(list(h.fc_layers[0].children())) != 1:
                    raise UnsupportedTopology("adversary heads must be single Linear layers")
            chains = _block_chains(adv.encoder)
            chains.append([(h.fc_layers[0].lin.weight, False) for h in heads])
            chains.append([(h.fc_layers[0].lin.bias, False) for h in heads])
            g = self.groups[f"adversarials/{i + 1}"] = FlatGroup(f"adversarials/{i + 1}", chains, dev)
            ap = AdvPlan(enc=_plan_block(adv.encoder, g), conditions=conds,
                         classes=[h.fc_layers[0].lin.out_features for h in heads], group=g)
            K = heads[0].fc_layers[0].lin.in_features
            sumC = sum(ap.classes)
            o = g.offset[id(heads[0].fc_layers[0].lin.weight)]
            ap.Wh32, ap.gWh = g.p[o:o + sumC * K].view(sumC, K), g.g[o:o + sumC * K].view(sumC, K)
            ap.Wh16 = g.p16[o:o + sumC * K].view(sumC, K)
            o = g.offset[id(heads[0].fc_layers[0].lin.bias)]
            ap.bh, ap.gbh = g.p[o:o + sumC], g.g[o:o + sumC]
            self.adv.append(ap)
        # ---- output discriminators on the reconstruction (BASELINE config 4) ----
        self.odisc: Dict[str, dict] = {}
        for sid, disc in (output_discriminators or {}).items():
            l1, l2, l3 = disc.linears
            disc.to(dev)
            # first-layer weight stored [genes, hidden] (like the expert encoder's): the operand layout of both the
            # dense and the sparse half of xhat W^T; torch.optim.Adam defaults (no weight decay) at lr 1e-3
            g = self.groups[f"output_discriminators/{sid}"] = FlatGroup(
                f"output_discriminators/{sid}", [[(l1.bias, False)], [(l2.bias, False)], [(l3.bias, False)],
                                                 [(l2.weight, False)], [(l3.weight, False)]], dev,
                lr=output_discriminator_lr, weight_decay=0.0, sharded=[(l1.weight, True)], background_last=False,
                world=self.world, rank=self.rank)
            ph = lambda p, buf=None, g=g: g.phys(p, buf)   # noqa: E731
            self.odisc[sid] = dict(l1=l1, group=g, G=l1.in_features, H1=l1.out_features, H2=l2.out_features,
                                   W1t16=ph(l1.weight, g.p16), gW1t=ph(l1.weight, g.g), b1=ph(l1.bias), gb1=ph(l1.bias, g.g),
                                   W2=ph(l2.weight), gW2=ph(l2.weight, g.g), b2=ph(l2.bias), gb2=ph(l2.bias, g.g),
                                   W3=ph(l3.weight), gW3=ph(l3.weight, g.g), b3=ph(l3.bias), gb3=ph(l3.bias, g.g))
        self._ws: Dict[tuple, torch.Tensor] = {}
        # small weight-gradient GEMMs are off the critical path (only the optimizer needs them): they run on a
        # side stream, concurrently with the dX chain of the main stream
        self.side = torch.cuda.Stream(device=self.device)
        self._side_used = False
        self.timers: Optional[Dict[str, list]] = None   # name -> [(start_event, end_event)] when profiling
        self.timer_filter = None   # optional set of section names to time (an event record ends a PDL chain)
        self._seed = 0x5EED
        self.mid_tf32 = os.environ.get("CMMVAE_MID_TF32", "1") != "0"
        self.spmm_tc = True                 # bf16 policy: expert-encoder SpMM on the tensor pipe when spmm_picks_tensor() says so
        self.last = None

    # ------------------------------------------------------------------------------------ utilities
    def ws(self, name: str, shape, dtype=torch.float32, zero=False) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
        return t

    def ws_cap(self, name: str, n: int, dtype=torch.float32) -> torch.Tensor:
        """1-D workspace whose length follows the batch (``nnz``-sized buffers): ONE buffer per name, grown with
        25 % headroom when a batch needs more than any before it, handed out as a ``[:n]`` view -- real batches
        almost never repeat an nnz, so keying by exact shape would allocate (and keep) a new buffer every step"""
        key = (name, "cap", dtype)
        t = self._ws.get(key)
        if t is None or t.numel() < n:
            t = self._ws[key] = torch.empty(max(int(n * 1.25) + 1024, 1024), dtype=dtype, device=self.device)
        return t[:n]

    def _t0(self, name):
        gm = self._gmode
        if gm is not None and gm.get("graphs") is not None:
            # capturing: timing events become event-record nodes of the graph (external events); after a replay they
            # hold that replay's timestamps
            if self.graph_timers and name in self.graph_timers:
                ev = (torch.cuda.Event(enable_timing=True, external=True),
                      torch.cuda.Event(enable_timing=True, external=True))
                ev[0].record()
                gm["events"].setdefault(name, []).append(ev)
                return ev
            return None
        if self.timers is None or (self.timer_filter is not None and name not in self.timer_filter):
            return None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        self.timers.setdefault(name, []).append(ev)
        return ev

    @staticmethod
    def _t1(ev):
        if ev is not None:
            ev[1].record()

    def timer_ms(self, name) -> float:
        """mean duration (ms) of the events recorded under ``name`` (call after a synchronize)"""
        evs = (self.timers or {}).get(name, [])
        return sum(a.elapsed_time(b) for a, b in evs) / max(len(evs), 1)

    def graph_timer_ms(self, name) -> float:
        """graph mode: mean duration (ms) of section ``name`` over the LAST replay of every captured step graph
        (call after a synchronize; ``graph_timers`` must have held the name when the graphs were captured)"""
        evs = [ev for e in self._graphs.values() if "gA" in e and e.get("replays", 0) > 0
               for ev in e["events"].get(name, [])]
        return sum(a.elapsed_time(b) for a, b in evs) / max(len(evs), 1)

    def _on_side(self, fn):
        """run ``fn`` (kernel launches reading tensors the main stream has produced so far) on the side stream"""
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.side), ops.stream_scope(self.side):
            self.side.wait_event(ev)
            fn()
        self._side_used = True

    def _join_side(self):
        if self._side_used:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_used = False

    def _tc(self, *dims) -> bool:
        return self.precision == "bf16" and all(d % 8 == 0 for d in dims)

    @staticmethod
    def spmm_picks_tensor(nnz: int, B: int, G: int, H: int = 1024) -> bool:
        """which first-layer kernel family is faster for this batch (cost model fitted to the density x batch sweep,
        profiles/r2_spmm_sweep.md, times in ms at G = 60 530, H = 1024 and scaled from there): the gather kernel
        costs ~0.12 us per 1000 non-zeros with a 0.09 ms floor, the densified tensor-pipe product ~0.085 ms per 1024
        cells plus index preparation and a small per-non-zero scatter term"""
        scale = (G / 60530.0) * (H / 1024.0)
        t_gather = max(0.09, 1.2e-7 * nnz * H / 1024.0 * (1.0 + 256.0 / max(B, 1)))   # (one CTA per cell: small batches under-fill)
        t_tensor = 8.5e-5 * B * scale + 2e-8 * nnz + 0.035
        return t_tensor < t_gather

    def _tf(self, *dims) -> bool:
        """small GEMMs of the middle chain: TF32 operands (fp32 activations / master weights) instead of bf16.
        2 % of the step's FLOPs, but their operand rounding is what flips ReLU masks upstream of every encoder
        gradient (tools/diag_precision.py); TF32 rounds 8x finer at no measurable cost."""
        return self.precision == "bf16" and self.mid_tf32 and all(d % 4 == 0 for d in dims)

    def _next_seed(self) -> int:
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) & ((1 << 63) - 1)
        return self._seed

    # ---------------------------------------------------------------------------------- dense layer
    def _linear(self, tag, lp: LayerPlan, x32, x16, B, fuse_relu):
        """Y = x W^T + b (optionally ReLU fused in the GEMM epilogue) -> (y32, y16|None)"""
        y32 = self.ws(tag + ".y32", (B, lp.N))
        if self._tf(lp.K, lp.N):
            y16 = self.ws(tag + ".y16", (B, lp.N), torch.bfloat16) if (fuse_relu and lp.N % 8 == 0) else None
            ops.gemm(x32, 0, lp.W32, 0, B, lp.N, lp.K, bias=lp.b, relu=fuse_relu, C32=y32, C16=y16, tf32=True)
            return y32, y16
        if self._tc(lp.K, lp.N):
            y16 = self.ws(tag + ".y16", (B, lp.N), torch.bfloat16) if fuse_relu else None
            ops.gemm(x16, 0, lp.W16, 0, B, lp.N, lp.K, bias=lp.b, relu=fuse_relu, C32=y32, C16=y16)
            return y32, y16
        ops.gemm(x32, 0, lp.W32, 0, B, lp.N, lp.K, bias=lp.b, relu=fuse_relu, C32=y32, use_tc=False)
        return y32, None

    def _layer_fwd(self, tag, lp: LayerPlan, x32, x16, B, csr=None, training=True, masks=None, Y_pre=None):
        """full layer; returns (out32, out16, cache).  ``Y_pre``: the layer's linear part was computed elsewhere
        (data-parallel first layer: partial sums from all ranks)"""
        want16 = self.precision == "bf16"
        plain = lp.bn is None and lp.p_drop == 0.0
        if Y_pre is not None:
            Y, y16, fused_relu = Y_pre, None, False
        elif lp.sparse:
            crow, col, val, G, tp = csr
            Wt = lp.W16 if self.precision == "bf16" else lp.W32
            ev = self._t0("csr_linear_fwd")
            if tp is not None:
                Y = ops.csr_linear_fwd_tc(tp[1], tp[0], B, G, lp.W16, lp.b, out=self.ws(tag + ".y32", (B, lp.N)))
            else:
                Y = ops.csr_linear_fwd(crow, col, val, G, Wt, lp.b, out=self.ws(tag + ".y32", (B, lp.N)))
            self._t1(ev)
            y16 = None
            fused_relu = False
        else:
            fused_relu = plain and lp.relu
            Y, y16 = self._linear(tag, lp, x32, x16, B, fused_relu)
        cache = dict(x32=x32, x16=x16, Y=Y, mean=None, rstd=None, seed=0, mask=None, p=0.0, seed_base=None)
        if plain and (fused_relu or not lp.relu):
            out32, out16 = Y, y16
            if want16 and out16 is None:
                out16 = ops.cast_bf16(out32, self.ws(tag + ".o16", (B, lp.N), torch.bfloat16))
        else:
            mean = rstd = None
            if lp.bn is not None:
                mean, rstd = self.ws(tag + ".mean", (lp.N,)), self.ws(tag + ".rstd", (lp.N,))
                if training:
                    ops.bn_stats(Y, lp.bn.eps, lp.bn.momentum, mean, rstd, lp.bn.running_mean, lp.bn.running_var)
                    lp.bn.num_batches_tracked += 1
                else:
                    mean = lp.bn.running_mean
                    ops.rstd_from_var(lp.bn.running_var, lp.bn.eps, rstd)
            p = lp.p_drop if training else 0.0
            mask = masks.get(tag) if (masks and p > 0) else None
            sb = None
            if p > 0 and mask is None and self._gmode is not None:
                # graph mode: per-step seed in device memory + a per-layer salt (the launch itself is replayed)
                seed, sb = (hash(tag) & 0xFFFFFFF) * 0x9E3779B1, self._dyn["seed"]
            else:
                seed = self._next_seed() if (p > 0 and mask is None) else 0
            out32 = self.ws(tag + ".o32", (B, lp.N))
            out16 = self.ws(tag + ".o16", (B, lp.N), torch.bfloat16) if want16 else None
            ops.bn_act_drop_fwd(Y, mean, rstd, lp.gamma if lp.bn is not None else None,
                                lp.beta if lp.bn is not None else None, lp.relu, p, seed, mask, out32, out16,
                                seed_base=sb)
            cache.update(mean=mean, rstd=rstd, seed=seed, mask=mask, p=p, seed_base=sb)
        cache.update(out32=out32, out16=out16)
        return out32, out16, cache

    def _layer_bwd(self, tag, lp: LayerPlan, cache, dOut32, B, need_dx=True, csc=None):
        """backward of one layer: parameter grads go to the flat grad buffer; returns dX32 (or None)"""
        want16 = self.precision == "bf16"
        has_tail = lp.bn is not None or lp.relu or cache["p"] > 0
        if has_tail:
            dY = self.ws(tag + ".dY", (B, lp.N))
            dY16 = self.ws(tag + ".dY16", (B, lp.N), torch.bfloat16) if want16 else None
            ops.bn_act_drop_bwd(dOut32, cache["Y"], cache["out32"], cache["mean"], cache["rstd"],
                                lp.gamma if lp.bn is not None else None, lp.relu, cache["p"], cache["seed"],
                                cache["mask"], dY, dY16, lp.ggamma, lp.gbeta, lp.gb, accumulate=True,
                                seed_base=cache["seed_base"])
        else:
            dY = dOut32
            dY16 = ops.cast_bf16(dY, self.ws(tag + ".dY16", (B, lp.N), torch.bfloat16)) if want16 else None
            ops.colsum(dY, lp.gb, accumulate=True)
        if lp.sparse:
            if csc[0] == "dp":
                self._dp_first_layer_grad(csc[1], lp, dY16, B)
            elif csc[0] == "tc":
                _, tp, G, ssq = csc
                ops.csr_linear_bwd_w_tc(tp[1], tp[0], B, G, dY16, lp.gW, sumsq_out=ssq)
            else:
                _, cptr, ridx, cval, G = csc
                ops.csr_linear_bwd_w(cptr, ridx, cval, B, G, dY, lp.gW)
            return None
        dX = self.ws(tag + ".dX", (B, lp.K)) if need_dx else None
        if self._tf(lp.K, lp.N):
            x32 = cache["x32"]
            self._on_side(lambda: ops.gemm(dY, 1, x32, 1, lp.N, lp.K, B, C32=lp.gW, tf32=True))
            if need_dx:
                ops.gemm(dY, 0, lp.W32, 1, B, lp.K, lp.N, C32=dX, tf32=True)
        elif self._tc(lp.K, lp.N):
            x16 = cache["x16"]
            self._on_side(lambda: ops.gemm(dY16, 1, x16, 1, lp.N, lp.K, B, C32=lp.gW))
            if need_dx:
                ops.gemm(dY16, 0, lp.W16, 1, B, lp.K, lp.N, C32=dX)
        else:
            x32 = cache["x32"]
            self._on_side(lambda: ops.gemm(dY, 1, x32, 1, lp.N, lp.K, B, C32=lp.gW, use_tc=False))
            if need_dx:
                ops.gemm(dY, 0, lp.W32, 1, B, lp.K, lp.N, C32=dX, use_tc=False)
        return dX

    # ------------------------------------------------------------------------------------ adversary
    def _adv_tc(self, ap: AdvPlan) -> bool:
        """adversary GEMMs go to the tensor pipe when every dimension is TMA-addressable (else CUDA-core fp32)"""
        dims = [d for lp in ap.enc for d in (lp.K, lp.N)] + [sum(ap.classes), ap.Wh32.shape[1]]
        return self.precision == "bf16" and all(d % 8 == 0 for d in dims)

    def _adv_fwd(self, tag, ap: AdvPlan, hid32, hid16, B):
        x, x16, caches = hid32, hid16, []
        tc = self._adv_tc(ap)
        saved_precision, self.precision = self.precision, ("bf16" if tc else "fp32")
        try:
            if tc and x16 is None:
                x16 = ops.cast_bf16(x, self.ws(tag + ".in16", tuple(x.shape), torch.bfloat16))
            for j, lp in enumerate(ap.enc):
                x, x16, c = self._layer_fwd(f"{tag}.e{j}", lp, x, x16, B)
                caches.append(c)
            sumC, K = ap.Wh32.shape
            logits = self.ws(tag + ".logits", (B, sumC))
            if tc and self.mid_tf32:
                ops.gemm(x, 0, ap.Wh32, 0, B, sumC, K, bias=ap.bh, C32=logits, tf32=True)
            elif tc:
                ops.gemm(x16, 0, ap.Wh16, 0, B, sumC, K, bias=ap.bh, C32=logits)
            else:
                ops.gemm(x, 0, ap.Wh32, 0, B, sumC, K, bias=ap.bh, C32=logits, use_tc=False)
        finally:
            self.precision = saved_precision
        return (x, x16), caches, logits

    def _adv_loss(self, tag, ap: AdvPlan, logits, labels, scale, B, ce_slots):
        dl = self.ws(tag + ".dlogits", logits.shape)
        o = 0
        for c, C, slot in zip(ap.conditions, ap.classes, ce_slots):
            ops.softmax_ce_sum(logits[:, o:o + C], C, labels[c], scale, dl[:, o:o + C], slot)
            o += C
        return dl

    def _adv_bwd(self, tag, ap: AdvPlan, code, caches, dl, B, need_dx):
        code32, code16 = code
        tc = self._adv_tc(ap)
        saved_precision, self.precision = self.precision, ("bf16" if tc else "fp32")
        try:
            sumC, K = ap.Wh32.shape
            ap.group.zero_vector_grads()
            ops.colsum(dl, ap.gbh, accumulate=True)
            d = self.ws(tag + ".dcode", (B, K))
            if tc and self.mid_tf32:
                self._on_side(lambda: ops.gemm(dl, 1, code32, 1, sumC, K, B, C32=ap.gWh, tf32=True))
                ops.gemm(dl, 0, ap.Wh32, 1, B, K, sumC, C32=d, tf32=True)
            elif tc:
                dl16 = ops.cast_bf16(dl, self.ws(tag + ".dl16", tuple(dl.shape), torch.bfloat16))
                self._on_side(lambda: ops.gemm(dl16, 1, code16, 1, sumC, K, B, C32=ap.gWh))
                ops.gemm(dl16, 0, ap.Wh16, 1, B, K, sumC, C32=d)
            else:
                ops.gemm(dl, 1, code32, 1, sumC, K, B, C32=ap.gWh, use_tc=False)
                ops.gemm(dl, 0, ap.Wh32, 1, B, K, sumC, C32=d, use_tc=False)
            for j in reversed(range(len(ap.enc))):
                d = self._layer_bwd(f"{tag}.e{j}", ap.enc[j], caches[j], d, B, need_dx=(need_dx or j > 0))
        finally:
            self.precision = saved_precision
        self._join_side()
        return d

    # --------------------------------------------------------------------------- output discriminator
    def _output_disc_step(self, od, expert_id, crow, col, val, nnz, tp, dl, B, G, loss_slot, norm_slot):
        """One optimisation step of the species' output discriminator on this step's (detached) reconstruction:
        Linear(G,128) Sigmoid Linear(128,64) Sigmoid Linear(64,1) Sigmoid, BCE(mean) against the species label,
        Adam(lr 1e-3) (meta_discriminators.py:33-49,112-148).  xhat is not in HBM: xhat W1^T = 1/2 dlogits W1^T +
        Xm W1^T (Xm: entries of the batch with non-zero dlogits), and likewise dW1 = xhat^T da1."""
        from mmvae_b200.modules.output_discriminator import SPECIES_LABEL
        g, H1, H2 = od["group"], od["H1"], od["H2"]
        y = SPECIES_LABEL.get(expert_id, 0.0)
        cap = self._gmode["cap"] if self._gmode is not None else nnz
        val_m = ops.mask_vals_by_dl(crow, col, val, dl, self.ws_cap("od.valm", max(cap, 1)))
        tp_buf = self.ws("od.tp64", (B * ((G + 63) // 64 + 1),), torch.int32)
        pk_buf = self.ws_cap("od.packed", (cap + 3) // 4 * 4 + 4, torch.int32)
        if self._gmode is not None:
            tpm = ops.csr_tile_ptr_dyn(crow, col, val_m, G, cap, tp_buf, pk_buf)
        else:
            tpm = ops.csr_tile_ptr(crow, col, val_m, G, nnz, tp_buf, pk_buf)
        # ---- forward
        a1pre = ops.csr_linear_fwd_tc(tpm[1], tpm[0], B, G, od["W1t16"], od["b1"], out=self.ws("od.a1pre", (B, H1)))
        half = self.ws("od.half", (B, H1))
        ops.gemm(dl, 0, od["W1t16"], 1, B, H1, G, C32=half)             # dlogits W1^T   (K = genes)
        ops.axpy(a1pre, half, 0.5)
        a1 = self.ws("od.a1", (B, H1))
        ops.sigmoid_fwd(a1pre, a1)
        a2pre, a2 = self.ws("od.a2pre", (B, H2)), self.ws("od.a2", (B, H2))
        ops.gemm(a1, 0, od["W2"], 0, B, H2, H1, bias=od["b2"], C32=a2pre, tf32=True)
        ops.sigmoid_fwd(a2pre, a2)
        a3 = self.ws("od.a3", (B, 1))
        ops.gemm(a2, 0, od["W3"], 0, B, 1, H2, bias=od["b3"], C32=a3, use_tc=False)
        da3 = self.ws("od.da3", (B, 1))
        ops.bce_sigmoid(a3, y, None, da3, loss_slot)
        # ---- backward
        ops.colsum(da3, od["gb3"])
        ops.gemm(da3, 1, a2, 1, 1, H2, B, C32=od["gW3"], use_tc=False)
        da2 = self.ws("od.da2", (B, H2))
        ops.gemm(da3, 0, od["W3"], 1, B, H2, 1, C32=da2, use_tc=False)
        da2pre = self.ws("od.da2pre", (B, H2))
        ops.sigmoid_bwd(da2, a2, da2pre)
        ops.colsum(da2pre, od["gb2"])
        ops.gemm(da2pre, 1, a1, 1, H2, H1, B, C32=od["gW2"], tf32=True)
        da1 = self.ws("od.da1", (B, H1))
        ops.gemm(da2pre, 0, od["W2"], 1, B, H1, H2, C32=da1, tf32=True)
        da1pre, da1pre16 = self.ws("od.da1pre", (B, H1)), self.ws("od.da1pre16", (B, H1), torch.bfloat16)
        ops.sigmoid_bwd(da1, a1, da1pre, da1pre16)
        ops.colsum(da1pre, od["gb1"])
        ops.csr_linear_bwd_w_tc(tpm[1], tpm[0], B, G, da1pre16, od["gW1t"])          # Xm^T da1
        halfw = self.ws("od.halfw", (G, H1))
        ops.gemm(dl, 1, da1pre16, 1, G, H1, B, C32=halfw)                            # dlogits^T da1
        ops.axpy(od["gW1t"], halfw, 0.5)
        g.grad_norm_sq(norm_slot)
        g.clip_adam(norm_slot, None, 1.0, advance=self._gmode is None)

    def _od_tail(self, od, a3, y, loss_slot, a1, B):
        """layers 2-3 + BCE + their backward on this rank's cells; returns d(loss)/d(a1 pre-activation) fp32 / bf16"""
        H1, H2 = od["H1"], od["H2"]
        a2pre, a2 = self.ws("od.a2pre", (B, H2)), self.ws("od.a2", (B, H2))
        ops.gemm(a1, 0, od["W2"], 0, B, H2, H1, bias=od["b2"], C32=a2pre, tf32=True)
        ops.sigmoid_fwd(a2pre, a2)
        ops.gemm(a2, 0, od["W3"], 0, B, 1, H2, bias=od["b3"], C32=a3, use_tc=False)
        da3 = self.ws("od.da3", (B, 1))
        ops.bce_sigmoid(a3, y, None, da3, loss_slot)
        ops.colsum(da3, od["gb3"])
        ops.gemm(da3, 1, a2, 1, 1, H2, B, C32=od["gW3"], use_tc=False)
        da2 = self.ws("od.da2", (B, H2))
        ops.gemm(da3, 0, od["W3"], 1, B, H2, 1, C32=da2, use_tc=False)
        da2pre = self.ws("od.da2pre", (B, H2))
        ops.sigmoid_bwd(da2, a2, da2pre)
        ops.colsum(da2pre, od["gb2"])
        ops.gemm(da2pre, 1, a1, 1, H2, H1, B, C32=od["gW2"], tf32=True)
        da1 = self.ws("od.da1", (B, H1))
        ops.gemm(da2pre, 0, od["W2"], 1, B, H1, H2, C32=da1, tf32=True)
        da1pre, da1pre16 = self.ws("od.da1pre", (B, H1)), self.ws("od.da1pre16", (B, H1), torch.bfloat16)
        ops.sigmoid_bwd(da1, a1, da1pre, da1pre16)
        ops.colsum(da1pre, od["gb1"])
        return da1pre, da1pre16

    def _dp_output_disc_step(self, od, expert_id, dpm, dl, B, loss_slot, norm_slot):
        """the output discriminator on the data-parallel route: its first layer [genes, 128] is sharded by genes like
        the expert's.  Rank r holds dlogits and the batch for (all cells) x (its genes): the two halves of
        xhat W1^T -- 1/2 dlogits W1[genes_r] and Xm[:, genes_r] W1[genes_r] -- are routed from the kernels' epilogues
        to the cells' owners and summed there; layers 2-3 and the loss run on the owner's cells; d(a1) is gathered
        on every rank for dW1[genes_r] = xhat[:, genes_r]^T d(a1); the small layers' gradients are summed like the
        other replicated groups.  DDP semantics: the mean over ranks of the per-rank BCE(mean) gradients."""
        from mmvae_b200.modules.output_discriminator import SPECIES_LABEL
        d, N, r, step = self._dp, self.world, self.rank, dpm["step"]
        g, H1, l1 = od["group"], od["H1"], od["l1"]
        per, NB = dpm["per"], dpm["NB"]
        y = SPECIES_LABEL.get(expert_id, 0.0)
        if "od" not in d:       # collective (every rank steps the same species): symmetric buffers of the route
            d["od"] = dict(s=self.comm.alloc("od.s", N * B * H1 * 4), g=self.comm.alloc("od.g", N * B * H1 * 4),
                           da=self.comm.alloc("od.da", N * B * H1 * 2))
        o = d["od"]
        W16, gW = g.own_rows(l1.weight, g.p16), g.own_rows(l1.weight, g.g)
        if W16.shape[0] != per:
            raise RuntimeError("output discriminator on the data-parallel route needs more than one process")
        n_rec = dpm["n_rec"]
        val_m = ops.mask_vals_by_dl_rows(dpm["rbeg"], dpm["rend"], dpm["col"], dpm["val"], dl,
                                         self.ws("od.dp.valm", (n_rec,)))
        tpm = ops.csr_tile_ptr_rows(dpm["rbeg"], dpm["rend"], dpm["col"], val_m, NB, per, n_rec - 8,
                                    self.ws("od.dp.tp64", (NB * ((per + 63) // 64 + 1),), torch.int32),
                                    self.ws("od.dp.packed", (n_rec + 8,), torch.int32))
        # ---- forward: both halves routed to the owners
        ops.csr_linear_fwd_tc_routed(tpm[1], tpm[0], NB, per, W16, [p + r * B * H1 * 4 for p in o["s"].ptr], B)
        ops.gemm_routed(dl, 0, W16, 1, NB, H1, per, H1, [p + r * B * H1 * 4 for p in o["g"].ptr], B)
        ops.peer_signal(self.comm.flag_ptrs("od.fwd"), step, step_dev=self._sd())
        ops.peer_wait(self.comm.local_flags("od.fwd"), N, step, step_dev=self._sd())
        a1pre, half = self.ws("od.a1pre", (B, H1)), self.ws("od.half", (B, H1))
        ops.slab_sum(o["s"].local.view(torch.float32), N, B * H1, B * H1, out32=a1pre, bias=od["b1"], H=H1)
        ops.slab_sum(o["g"].local.view(torch.float32), N, B * H1, B * H1, out32=half)
        ops.axpy(a1pre, half, 0.5)
        a1 = self.ws("od.a1", (B, H1))
        ops.sigmoid_fwd(a1pre, a1)
        da1pre, da1pre16 = self._od_tail(od, self.ws("od.a3", (B, 1)), y, loss_slot, a1, B)
        # ---- backward: d(a1) of every rank's cells on every rank, then this rank's gene rows of dW1
        ops.peer_push(da1pre16, B * H1 * 2, [p + r * B * H1 * 2 for p in o["da"].ptr], self.comm.flag_ptrs("od.da"), step,
                      self.comm.ticket, step_dev=self._sd())
        self._dp_allreduce_start(dpm, g, g.tail_lo, g.n)
        ops.peer_wait(self.comm.local_flags("od.da"), N, step, step_dev=self._sd())
        da_all = o["da"].local.view(torch.bfloat16).view(NB, H1)
        ssq = self.ws("od.dp.ssq", (1,), torch.float64)
        ssq.zero_()
        ops.csr_linear_bwd_w_tc(tpm[1], tpm[0], NB, per, da_all, gW)                   # Xm^T da1
        halfw = self.ws("od.dp.halfw", (per, H1))
        ops.gemm(dl, 1, da_all, 1, per, H1, NB, C32=halfw)                              # dlogits^T da1
        ops.axpy(gW, halfw, 0.5)
        ops.sumsq(gW, ssq)
        self._dp_allreduce_finish(dpm, g, g.tail_lo, g.n)
        ops.sumsq(g.g[g.tail_lo:g.n], norm_slot)
        dpm["od_ssq"] = ssq          # (the other ranks' rows join the logged norm in the scalar exchange)
        g.clip_adam(norm_slot, None, 1.0 / N, advance=self._gmode is None)

    # ------------------------------------------------------ data parallel over peer memory (gene shards)
    # Cells shard across ranks; the two gene-sized layers shard by GENES.  Rank r owns rows
    # [r * per, (r + 1) * per) of W1 (stored [genes, hidden]), Wout and bout -- values, gradients, Adam state -- and
    # multiplies them against the cells of ALL ranks:
    #   forward   Y_part[N*B, H1] = X_all[:, genes_r] W1[genes_r]   epilogue stores each cell block into its owner's
    #             buffer (peer memory), the owner sums the N slabs                               ("reduce-scatter")
    #             h of all ranks is pushed to every rank ("all-gather", 2 MB per rank), the fused decoder runs on
    #             h_all x Wout[genes_r] against X_all[:, genes_r]: dlogits, dWout[genes_r], dbout[genes_r] stay local
    #   backward  dh_part[N*B, Hd] = dlogits Wout[genes_r]          routed to the owners like Y_part
    #             dW1[genes_r] = X_all[:, genes_r]^T dY_all          (dY pushed like h)
    # What crosses NVLink per step and rank: the CSR records of the batch (prefetched one step ahead), two
    # [N*B, hidden] f32 partial-sum slabs, two [B, hidden] bf16 all-gathers and ~6 MB of small gradients -- instead
    # of 500 MB of gradient all-reduce.  No gradient, weight or optimizer state of the 125 M gene-sized parameters is
    # ever exchanged; the sums over ranks that DDP takes on gradients are taken on the layer's inputs.
    def _dp_setup(self, B: int, nnz: int, H1: int, Hd: int):
        """collective, first step: agree on capacities and allocate the symmetric buffers"""
        comm, N = self.comm, self.world
        mx, bmax, bmin = dp.host_allreduce_max([nnz, B, -B])
        if bmax != -bmin:
            raise RuntimeError("data-parallel step needs the same number of cells on every rank")
        # every rank receives, from every rank, the piece of the batch that falls into its gene shard: about
        # nnz / N entries per piece (uniform panels); headroom for skew, checked on the device every step
        head = float(os.environ.get("CMMVAE_DP_CSR_HEADROOM", "1.5"))
        cap = _ceil(int(mx / N * head) + 4096, 1024)
        d = dict(B=B, cap=cap, H1=H1, Hd=Hd)
        d["col_off"] = _ceil(4 * (B + 1), 256)
        d["val_off"] = d["col_off"] + 4 * cap
        d["slab"] = d["val_off"] + 4 * cap
        d["csr"] = [comm.alloc(f"csr{i}", N * d["slab"]) for i in range(2)]
        d["scratch"] = [tuple(torch.zeros(N * B, dtype=torch.int32, device=self.device) for _ in range(3)) +
                        (torch.zeros(2, dtype=torch.int32, device=self.device),) for _ in range(2)]
        # few cell tiles (small N * B): the routed first-layer product is also cut along the genes so that every SM
        # has a unit; each piece lands in its own slab on the owner
        units = ((N * B + 127) // 128) * ((H1 + 255) // 256)
        d["S"] = max(1, min(4, 148 // units))
        d["Yin"] = comm.alloc("Yin", N * d["S"] * B * H1 * 4)
        d["hall"] = comm.alloc("hall", N * B * Hd * 2)
        d["dhin"] = comm.alloc("dhin", N * B * Hd * 4)
        d["dYall"] = comm.alloc("dYall", N * B * H1 * 2)
        d["SC"] = 32
        d["scal"] = comm.alloc("scal", N * d["SC"] * 8)
        d["tails"] = {}
        d["stream"] = torch.cuda.Stream(self.device)
        d["ticket_pf"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._dp = d
        return d

    def set_dp_capacity(self, max_nnz_per_rank: int):
        """(before the first step, same value on every rank) densest batch, in non-zeros per rank, the exchange buffers
        are sized for (each of the N pieces of a batch gets 1/N of it); default: 1.5 x the densest first batch"""
        os.environ["CMMVAE_DP_CSR_HEADROOM"] = "1.0"
        self._dp_force_nnz = int(max_nnz_per_rank)

    def _dp_tail(self, group: FlatGroup, n: int):
        d = self._dp
        buf = d["tails"].get(group.name)
        if buf is None:
            buf = d["tails"][group.name] = self.comm.alloc(f"tail/{group.name}", self.world * _ceil(n, 4) * 4)
        return buf

    def _dp_push_csr(self, crow, col, val, nnz: int, step: int, per: int):
        """all-to-all of this rank's batch by gene shard: piece q of every row is stored straight into slab `rank`
        of rank q's buffer (parity of the consuming step), columns rebased to the shard; then the flags"""
        d, N, r = self._dp, self.world, self.rank
        par = step & 1
        buf = d["csr"][par]
        cnt, start, offs, info = d["scratch"][par]
        base = [p + r * d["slab"] for p in buf.ptr]
        ops.csr_scatter_shards(crow, col, val, d["B"], N, per, d["cap"], base, [b + d["col_off"] for b in base],
                               [b + d["val_off"] for b in base], cnt, start, offs, info)
        ops.peer_signal(self.comm.flag_ptrs(f"csr{par}"), step)

    def prefetch(self, expert_id: str, crow, col, val, ready=None):
        """data parallel: call right AFTER ``train_step`` with the NEXT batch -- its CSR records are exchanged
        underneath the step that was just enqueued, and the next ``train_step`` on the same arrays finds its gathered
        inputs ready (no-op with one process).  ``ready``: a CUDA event after which the arrays are complete (the
        stager's copy event), ``True`` if they already are, ``None`` = whatever is enqueued on the current stream
        (always correct, but the current stream waits for the running step, so nothing overlaps)."""
        if self.comm is None or self._dp is None:
            return None       # buffers exist after the first step
        nnz = int(col.numel())
        key = (crow.data_ptr(), col.data_ptr(), val.data_ptr(), nnz)
        if self._csr_pushed:
            return None       # one batch ahead (two buffers)
        step = self._dp_step + 1
        st = self._dp["stream"]
        if ready is None:
            ready = torch.cuda.Event()
            ready.record()
        with torch.cuda.stream(st), ops.stream_scope(st):
            if ready is not True:
                st.wait_event(ready)               # the batch arrays are complete
            if self._dp.get("safe") is not None:   # every rank is past the step that last read this buffer parity
                st.wait_event(self._dp["safe"])
            g = self.groups[f"experts/{expert_id}"]
            per = g.per if g.sharded else _ceil(self.enc_plan[expert_id][0].K, 128)
            self._dp_push_csr(crow, col, val, nnz, step, per)
        self._csr_pushed[key] = step
        return None

    def _sd(self):
        """graph mode: the device word holding this step's number (flag value of every exchange); else None"""
        return self._dyn["step"] if (self._gmode is not None and self._dyn is not None) else None

    def _dp_host_prepare(self, gexp: FlatGroup, enc0: LayerPlan, crow, col, val, nnz: int, B: int, G: int, Hd: int):
        """host half of the start of a data-parallel step: step number / buffer parity, and -- unless the batch was
        prefetched -- the all-to-all of its records (launched here, i.e. outside a captured graph)"""
        if self._dp is None:
            self._dp_setup(B, getattr(self, "_dp_force_nnz", nnz), enc0.N, Hd)
        d = self._dp
        if B != d["B"]:
            raise RuntimeError(f"data-parallel step was set up for {d['B']} cells per rank, got {B}")
        per = d["per"] = gexp.per if gexp.sharded else _ceil(G, 128)
        self._dp_step += 1
        step = self._dp_step
        key = (crow.data_ptr(), col.data_ptr(), val.data_ptr(), nnz)
        pushed = self._csr_pushed.pop(key, None)
        if pushed != step:
            if pushed is not None or self._csr_pushed:
                # a batch was prefetched and then NOT stepped next (e.g. the loop changed its data source): its
                # records sit in this step's buffer with the flags already raised.  Every rank is in the same
                # situation (symmetric program), so all of them skip this step number -- and with it the buffer
                self._csr_pushed.clear()
                self._dp_step += 1
                step = self._dp_step
            self._dp_push_csr(crow, col, val, nnz, step, per)
        d["host"] = (step, per)
        return step, per

    def _dp_prepare(self, gexp: FlatGroup, enc0: LayerPlan, crow, col, val, nnz: int, B: int, G: int, Hd: int):
        """start of a data-parallel step: the pieces of every rank's batch that fall into this rank's gene shard
        (all-to-all, normally prefetched) -> row ranges, window pointer table and packed records over the received
        slabs, in place"""
        N, r = self.world, self.rank
        if self._gmode is not None and self._gmode.get("dp_host") is not None:
            step, per = self._gmode["dp_host"]       # graph mode: the host half already ran (outside the capture)
        else:
            step, per = self._dp_host_prepare(gexp, enc0, crow, col, val, nnz, B, G, Hd)
        d = self._dp
        par = step & 1
        NB = N * B
        ops.peer_wait(self.comm.local_flags(f"csr{par}"), N, step, step_dev=self._sd())
        # the N received slabs are used in place: row ranges over one array spanning all slabs, window pointers
        # and packed records on top of it -- nothing is copied
        buf = d["csr"][par].local
        rbeg, rend = self.ws("dp.rbeg", (NB,), torch.int32), self.ws("dp.rend", (NB,), torch.int32)
        ops.slab_rows(buf, d["slab"], B, N, rbeg, rend)
        n_rec = (N * d["slab"] - d["val_off"]) // 4
        col_s = buf[d["col_off"]:d["col_off"] + 4 * n_rec].view(torch.int32)
        val_s = buf[d["val_off"]:d["val_off"] + 4 * n_rec].view(torch.float32)
        tp = ops.csr_tile_ptr_rows(rbeg, rend, col_s, val_s, NB, per, n_rec - 8,
                                   self.ws("dp.tp64", (NB * ((per + 63) // 64 + 1),), torch.int32),
                                   self.ws("dp.packed", (n_rec + 8,), torch.int32))
        info = d["scratch"][par][3]
        g0, g1, crow_s = r * per, min(G, (r + 1) * per), None
        return dict(step=step, per=per, g0=g0, g1=g1, NB=NB, crow=crow_s, col=col_s, val=val_s, tp=tp, info=info,
                    rbeg=rbeg, rend=rend, n_rec=n_rec,
                    shard_ssq=self.ws("dp.shard_ssq", (1,), torch.float64), loss_part=self.ws("dp.loss_part", (N,), torch.float64))

    def _dp_first_layer_fwd(self, dpm, gexp: FlatGroup, lp: LayerPlan, B: int):
        """Y[B, H1] of this rank's cells = sum over gene shards (all ranks) + bias"""
        d, N, r, step = self._dp, self.world, self.rank, dpm["step"]
        H1 = lp.N
        W16 = gexp.own_rows(lp.lin.weight, gexp.p16)
        if W16.shape[0] != dpm["per"]:       # single-process test route: pad the view up to the 128-aligned shard
            W16 = self._dp_padded("dp.W1pad", W16, dpm["per"])
        ev = self._t0("csr_linear_fwd")
        S = d["S"]
        ops.csr_linear_fwd_tc_routed(dpm["tp"][1], dpm["tp"][0], dpm["NB"], dpm["per"], W16,
                                     [p + r * S * B * H1 * 4 for p in d["Yin"].ptr], B, S, B * H1)
        self._t1(ev)
        ops.peer_signal(self.comm.flag_ptrs("Y"), step, step_dev=self._sd())
        ev = self._t0("dp_wait_Y")
        ops.peer_wait(self.comm.local_flags("Y"), N, step, step_dev=self._sd())
        self._t1(ev)
        Y = self.ws("enc0.y32", (B, H1))
        ops.slab_sum(d["Yin"].local.view(torch.float32), N * S, B * H1, B * H1, out32=Y, bias=lp.b, H=H1)
        return Y

    def _dp_padded(self, name, t, rows):
        out = self.ws(name, (rows,) + tuple(t.shape[1:]), t.dtype, zero=True)
        out[:t.shape[0]].copy_(t)
        return out

    def _dp_decoder(self, dpm, gexp: FlatGroup, out: LayerPlan, h16, B: int):
        """all-gather h, fused decoder + loss on (all cells) x (this rank's genes), dWout / dbout rows, and the dh
        partial sums routed to the cells' owners.  Returns dh[B, Hd] of this rank's cells."""
        d, N, r, step = self._dp, self.world, self.rank, dpm["step"]
        Hd, per, NB = out.K, dpm["per"], dpm["NB"]
        # every rank has cut its shard out of this step's gathered CSR (the Y partials this rank already received
        # came after that): the CSR buffer of the other parity may be refilled for the next step from here on --
        # i.e. the prefetch of the next batch overlaps the tensor-bound decoder / dWout / dh kernels, not the
        # latency-bound chain of small kernels before them
        if not (self._gmode is not None and self._gmode["graphs"] is not None):   # (captured: recorded by the replay)
            d["safe"] = torch.cuda.Event()
            d["safe"].record()
        ops.peer_push(h16, B * Hd * 2, [p + r * B * Hd * 2 for p in d["hall"].ptr], self.comm.flag_ptrs("h"), step,
                      self.comm.ticket, step_dev=self._sd())
        ev = self._t0("dp_wait_h")
        ops.peer_wait(self.comm.local_flags("h"), N, step, step_dev=self._sd())
        self._t1(ev)
        h_all = d["hall"].local.view(torch.bfloat16).view(NB, Hd)
        W16 = gexp.own_rows(out.lin.weight, gexp.p16)
        bout = gexp.own_rows(out.lin.bias)
        gW = gexp.own_rows(out.lin.weight, gexp.g)
        gb = gexp.own_rows(out.lin.bias, gexp.g)
        if W16.shape[0] != per:
            W16, bout = self._dp_padded("dp.Woutpad", W16, per), self._dp_padded("dp.boutpad", bout, per)
            gW_out, gb_out = self.ws("dp.gWoutpad", (per, Hd)), self.ws("dp.gboutpad", (per,), zero=True)
            gb_out.zero_()
        else:
            gW_out, gb_out = gW, gb
        ldd = _ceil(per, 64)
        dl = self.ws("dp.dlogits16", (NB, ldd), torch.bfloat16, zero=True)
        ev = self._t0("decoder_mse_fused")
        ops.decoder_mse_fused_blocks(h_all, W16, bout, per, dpm["crow"], dpm["col"], dpm["val"], dl, dpm["loss_part"],
                                     B, dpm["tp"][0])
        self._t1(ev)
        dpm["shard_ssq"].zero_()
        ev = self._t0("dWout_gemm")
        ops.gemm(dl, 1, h_all, 1, per, Hd, NB, C32=gW_out, sumsq_out=dpm["shard_ssq"])
        self._t1(ev)
        ops.colsum(dl, gb_out, M=NB, N=per, accumulate=True)
        if gW_out is not gW:
            gW.copy_(gW_out[:gW.shape[0]])
            gb.copy_(gb_out[:gb.shape[0]])
        ops.sumsq(gb_out, dpm["shard_ssq"])
        ev = self._t0("dh_gemm")
        ops.gemm_routed(dl, 0, W16, 1, NB, Hd, per, Hd, [p + r * B * Hd * 4 for p in d["dhin"].ptr], B)
        self._t1(ev)
        ops.peer_signal(self.comm.flag_ptrs("dh"), step, step_dev=self._sd())
        ev = self._t0("dp_wait_dh")
        ops.peer_wait(self.comm.local_flags("dh"), N, step, step_dev=self._sd())
        self._t1(ev)
        dh = self.ws("dh", (B, Hd))
        ops.slab_sum(d["dhin"].local.view(torch.float32), N, B * Hd, B * Hd, out32=dh)
        return dh, dl

    def _dp_first_layer_grad(self, dpm, lp: LayerPlan, dY16, B: int):
        """dW1[genes_r] = X_all[:, genes_r]^T dY_all: this rank's rows of the gradient SUMMED over ranks"""
        d, N, r, step = self._dp, self.world, self.rank, dpm["step"]
        gexp, H1, per = lp.group, lp.N, dpm["per"]
        ops.peer_push(dY16, B * H1 * 2, [p + r * B * H1 * 2 for p in d["dYall"].ptr], self.comm.flag_ptrs("dY"), step,
                      self.comm.ticket, step_dev=self._sd())
        # everything but dW1 is final now: the small (replicated) gradients travel while dW1 is computed
        gvae = self.groups["vae"]
        self._dp_allreduce_start(dpm, gexp, gexp.tail_lo, gexp.n)
        self._dp_allreduce_start(dpm, gvae, 0, gvae.n)
        ops.peer_wait(self.comm.local_flags("dY"), N, step, step_dev=self._sd())
        dY_all = d["dYall"].local.view(torch.bfloat16).view(dpm["NB"], H1)
        gW = gexp.own_rows(lp.lin.weight, gexp.g)
        if gW.shape[0] != per:
            tmp = self.ws("dp.gW1pad", (per, H1))
            ops.csr_linear_bwd_w_tc(dpm["tp"][1], dpm["tp"][0], dpm["NB"], per, dY_all, tmp, sumsq_out=dpm["shard_ssq"])
            gW.copy_(tmp[:gW.shape[0]])
        else:
            ops.csr_linear_bwd_w_tc(dpm["tp"][1], dpm["tp"][0], dpm["NB"], per, dY_all, gW, sumsq_out=dpm["shard_ssq"])

    def _dp_allreduce_start(self, dpm, group: FlatGroup, lo: int, hi: int):
        """push g[lo:hi] of a replicated group into slab `rank` of every rank"""
        n = _ceil(hi - lo, 4)
        buf = self._dp_tail(group, n)
        ops.peer_push(group.g[lo:lo + n], n * 4, [p + self.rank * n * 4 for p in buf.ptr],
                      self.comm.flag_ptrs(f"tail/{group.name}"), dpm["step"], self.comm.ticket, step_dev=self._sd())

    def _dp_allreduce_finish(self, dpm, group: FlatGroup, lo: int, hi: int):
        """g[lo:hi] <- sum over ranks"""
        n = _ceil(hi - lo, 4)
        buf = self._dp_tail(group, n)
        ops.peer_wait(self.comm.local_flags(f"tail/{group.name}"), self.world, dpm["step"], step_dev=self._sd())
        ops.slab_sum(buf.local.view(torch.float32), self.world, n, n, out32=group.g[lo:lo + n])

    def _dp_finish_scalars(self, dpm, sc, s_norm_expert):
        """exchange the per-rank scalars: loss share of every rank's cells and the sum of squares of this rank's
        gradient rows; afterwards sc[0] = recon of THIS rank's cells, s_norm_expert += all shards"""
        d, N, r, step = self._dp, self.world, self.rank, dpm["step"]
        SC = d["SC"]
        mine = self.ws("dp.scal_mine", (SC,), torch.float64, zero=True)
        mine[:N].copy_(dpm["loss_part"])
        mine[N:N + 1].copy_(dpm["shard_ssq"])
        if "od_ssq" in dpm:
            mine[N + 1:N + 2].copy_(dpm["od_ssq"])
        ops.peer_push(mine, SC * 8, [p + r * SC * 8 for p in d["scal"].ptr], self.comm.flag_ptrs("scal"), step,
                      self.comm.ticket, step_dev=self._sd())
        ops.peer_wait(self.comm.local_flags("scal"), N, step, step_dev=self._sd())
        ops.dp_scalars(d["scal"].local.view(torch.float64), N, SC, r, sc[0:1], s_norm_expert)
        if "od_ssq" in dpm:      # (logged only: the discriminator's update is not clipped)
            dpm["od_norm_slot"].add_(d["scal"].local.view(torch.float64).view(N, SC)[:, N + 1].sum())

    def _after_reparameterize(self, z32, z16, B: int):
        """CLVAE.after_reparameterize (clvae.py:89-111): identity, or the conditional layers (whose host plan for
        this batch was made at the start of the step)"""
        if self.cond is None:
            return z32, z16
        return self.cond.forward(z32, self.ws, z16 is not None)

    # ----------------------------------------------------------------------------------------- step
    def train_step(self, expert_id: str, crow, col, val, nnz: int, kl_weight: float, eps=None,
                   labels: Optional[Dict[str, torch.Tensor]] = None, masks=None, nnz_cap: Optional[int] = None,
                   metadata=None):
        """One optimisation step on a CSR batch already resident on the device (no host sync).
        Returns the step record (device scalar block etc.) for ``scalars()``.  With ``use_graph`` (pipelined mode,
        single process) the step is replayed from captured CUDA graphs; ``nnz_cap`` (optional) = the densest batch
        to size the graph's input buffers for."""
        if self.cond is not None:
            if metadata is None:
                raise ValueError("conditional layers need the batch's metadata")
            self.cond.make_plan(metadata, expert_id, crow.numel() - 1)    # host: rows grouped by value, one H2D copy
        if self.pipeline_optimizer:
            if self._hp is None:
                lo_pri, hi_pri = torch.cuda.Stream.priority_range()
                self._hp = torch.cuda.Stream(self.device, priority=hi_pri)
                self._bg = torch.cuda.Stream(self.device, priority=lo_pri)
            cur = torch.cuda.current_stream()
            self._hp.wait_stream(cur)
            with torch.cuda.stream(self._hp), ops.stream_scope(self._hp):
                graphable = (self.use_graph and masks is None and self.timers is None and self.cond is None
                             and self.precision == "bf16" and not L._mask_queue)
                if graphable:
                    rec = self._graph_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, nnz_cap)
                else:
                    rec = self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, masks)
            cur.wait_stream(self._hp)
            return rec
        with ops.stream_scope(torch.cuda.current_stream()):
            return self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, masks)

    # ------------------------------------------------------------------------------------ graph mode
    def _set_gmode(self, mode):
        self._gmode = mode
        for g in self.groups.values():
            g.dyn_active = mode is not None

    def _ensure_dyn(self):
        if self._dyn is None:
            names = list(self.groups)
            nbytes = _ceil(16 + 8 * len(names), 16)
            dev = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self._dyn = dict(dev=dev, seed=dev[0:8].view(torch.int64), klw=dev[8:12].view(torch.float32),
                             step=dev[12:16].view(torch.int32),
                             host=[torch.zeros(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(8)],
                             copied=[None] * 8, slot=0, index={n: i for i, n in enumerate(names)}, labels={},
                             inputs={})
            for n, i in self._dyn["index"].items():
                self.groups[n].bc_dev = dev[16 + 8 * i:24 + 8 * i].view(torch.float32)
        return self._dyn

    def _write_dyn(self, kl_weight: float, dp_step: int = 0):
        """this step's scalars -> pinned block -> device block (one small async copy, stream-ordered before the step)"""
        d = self._dyn
        k = d["slot"]
        d["slot"] = (k + 1) % len(d["host"])
        if d["copied"][k] is not None:
            d["copied"][k].synchronize()
        h = d["host"][k]
        h[0:8].view(torch.int64)[0] = self._next_seed()
        f = h[8:].view(torch.float32)
        f[0] = float(kl_weight)
        h[12:16].view(torch.int32)[0] = dp_step & 0x7FFFFFFF
        for n, i in d["index"].items():
            bc1, bc2 = self.groups[n].bias_corrections()
            f[2 + 2 * i], f[3 + 2 * i] = bc1, bc2
        d["dev"].copy_(h, non_blocking=True)
        d["copied"][k] = torch.cuda.Event()
        d["copied"][k].record()

    def _graph_step(self, expert_id, crow, col, val, nnz, kl_weight, eps, labels, nnz_cap):
        d = self._ensure_dyn()
        # reparameterisation noise at a fixed address, drawn (or injected) OUTSIDE the graph: torch's generator then
        # advances exactly as in a stream-launched step
        Bn, Z = crow.numel() - 1, self.Z
        eps_static = d["inputs"].get(("eps", Bn, Z))
        if eps_static is None:
            eps_static = d["inputs"][("eps", Bn, Z)] = torch.empty(Bn, Z, device=self.device)
        if eps is not None or L._noise_queue:
            eps_static.copy_(eps if eps is not None else L.draw_noise(Bn, Z, self.device), non_blocking=True)
        else:
            eps_static.normal_()
        eps = eps_static
        n_adv = min(len(self.adv), self.n_hidden)
        if n_adv and labels is not None:      # labels at fixed addresses (the captured launches read them there)
            for c, t in labels.items():
                st = d["labels"].get((c, t.numel()))
                if st is None:
                    st = d["labels"][(c, t.numel())] = torch.empty_like(t)
                st.copy_(t, non_blocking=True)
            labels = {c: d["labels"][(c, t.numel())] for c, t in labels.items()}
        stepped = [self.groups["vae"], self.groups[f"experts/{expert_id}"]] + [a.group for a in self.adv[:n_adv]]
        if expert_id in self.odisc:
            stepped.append(self.odisc[expert_id]["group"])
        for g in stepped:
            g.advance()
        Bp1 = crow.numel()
        if self.comm is not None:
            # data parallel: the batch is exchanged by gene shard OUTSIDE the graphs (prefetched under the previous
            # step, or inline here); the captured launches read only the received slabs (fixed addresses, one pair of
            # graphs per buffer parity) and take the step number -- the value every flag is raised to -- from the
            # device block
            enc0, out = self.enc_plan[expert_id][0], self.dec_plan[expert_id][-1]
            gexp = self.groups[f"experts/{expert_id}"]
            step, per = self._dp_host_prepare(gexp, enc0, crow, col, val, nnz, Bp1 - 1, enc0.K, out.K)
            self._write_dyn(kl_weight, step)
            key = (expert_id, Bp1, step & 1)
            dp_host = (step, per)
        else:
            self._write_dyn(kl_weight)
            # the batch is copied to fixed addresses first (25 MB device-to-device, a few microseconds): ONE graph
            # pair per (expert, cell count) serves every batch, wherever it was staged
            gin = d["inputs"].get((expert_id, Bp1))
            if gin is None or gin["cap"] < nnz:
                cap = max(int(nnz * 1.25) + 1024, int(nnz_cap or 0))
                gin = d["inputs"][(expert_id, Bp1)] = dict(
                    cap=cap, crow=torch.empty(Bp1, dtype=torch.int32, device=self.device),
                    col=torch.empty(cap, dtype=torch.int32, device=self.device),
                    val=torch.empty(cap, dtype=torch.float32, device=self.device))
                self._graphs.pop((expert_id, Bp1), None)       # captured against the old addresses
            # (by a kernel: a cudaMemcpyAsync would queue behind the H2D transfers of the batches staged ahead)
            if (crow.data_ptr() | col.data_ptr() | val.data_ptr()) % 16 == 0 and crow.dtype == torch.int32 \
                    and col.is_contiguous() and val.is_contiguous():
                ops.copy_bytes(gin["crow"], crow, 4 * Bp1)
                ops.copy_bytes(gin["col"], col, 4 * nnz)
                ops.copy_bytes(gin["val"], val, 4 * nnz)
            else:
                gin["crow"].copy_(crow, non_blocking=True)
                gin["col"][:nnz].copy_(col, non_blocking=True)
                gin["val"][:nnz].copy_(val, non_blocking=True)
            crow, col, val, nnz_cap = gin["crow"], gin["col"][:nnz], gin["val"][:nnz], gin["cap"]
            key = (expert_id, Bp1)
            dp_host = None
        e = self._graphs.get(key)
        if e is None:
            # first visit: eager, with the device-side scalars (allocates every workspace the capture will need)
            self._set_gmode(dict(cap=nnz_cap, graphs=None, dp_host=dp_host))
            try:
                rec = self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, None)
            finally:
                self._set_gmode(None)
            self._graphs[key] = dict()
            return rec
        if "gA" not in e:
            gA, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            self._set_gmode(dict(cap=nnz_cap, graphs=(gA, gB), events={}, dp_host=dp_host))
            gexp = self.groups[f"experts/{expert_id}"]
            gexp.join_background()
            ops.set_pdl(False)    # plain graph nodes replay faster than nodes with programmatic edges (measured)
            try:
                gA.capture_begin(capture_error_mode="thread_local")   # (packing threads keep issuing copies)
                rec = self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, None)
                gB.capture_end()
                events = self._gmode["events"]
            finally:
                self._set_gmode(None)
                ops.set_pdl(True)
            e.update(gA=gA, gB=gB, rec=rec, bg=rec.pop("launch_bg", None), gexp=gexp, events=events, replays=0)
        e["replays"] += 1
        e["gA"].replay()
        if dp_host is not None:
            # (see _dp_decoder) from here the CSR buffer of the other parity may be refilled for the next step
            self._dp["safe"] = torch.cuda.Event()
            self._dp["safe"].record()
        e["gexp"].join_background()     # the output layer's update of the previous step (background stream)
        e["gB"].replay()
        if e["bg"] is not None:
            e["bg"]()
        # (the record's tensors are the graph's static outputs; per-step host values go into a fresh copy)
        self.last = dict(e["rec"], kl_weight=float(kl_weight))
        return self.last

    def finish(self):
        """join optimizer work still in flight on the background stream (call before reading weights outside
        the engine when ``pipeline_optimizer`` is on; ``state_dict``/evaluation do it themselves)"""
        for g in self.groups.values():
            g.join_background()

    def _train_step(self, expert_id, crow, col, val, nnz, kl_weight, eps, labels, masks):
        dev = self.device
        if masks is None:
            injected = L.draw_dropout_masks()
            if injected is not None:    # keyed by the reference's module paths -> engine layer tags
                masks = {L.dropout_tag(k): v for k, v in injected.items() if k.split(".")[0] != "experts"
                         or k.split(".")[1] == expert_id}
        enc, dec = self.enc_plan[expert_id], self.dec_plan[expert_id]
        B = crow.numel() - 1
        G = enc[0].K
        Z = self.Z
        out = dec[-1]
        gscale = 1.0 / self.world      # DDP semantics: the MEAN of the per-rank gradients is applied
        bf = self.precision == "bf16"
        n_adv = min(len(self.adv), self.n_hidden)   # zip(hidden, adversarials) truncates (cmmvae_model.py:67-70)
        # scalar slots (double): 0 recon | 1..3 kl,sum mu,sum var | then norms | then CE sums | 2 data-parallel info
        n_ce = sum(len(a.conditions) for a in self.adv[:n_adv])
        sc = torch.zeros(4 + 2 + 2 * n_adv + 2 * n_ce + 2 + 2, dtype=torch.float64, device=dev)
        od_slot = 4 + 2 + 2 * n_adv + 2 * n_ce      # [loss, grad norm^2] of the output discriminator; then 2 DP info
        s_norm = lambda k: sc[4 + k:5 + k]  # noqa: E731   0 vae, 1 expert, 2.. disc_i, then gen_i
        ce_base = 4 + 2 + 2 * n_adv

        # ---------------- forward ----------------
        gexp, gvae = self.groups[f"experts/{expert_id}"], self.groups["vae"]
        gexp.zero_vector_grads()
        gvae.zero_vector_grads()
        caches = {}
        x32 = x16 = None
        # tensor-pipe SpMM (tile densified in smem) above the density where it beats the gather kernel
        tc_ok = bf and self.spmm_tc and enc[0].N % 8 == 0 and G <= 65536
        dpm = None
        tp = None
        ev = self._t0("csr_prep")
        if self.comm is not None:
            # data parallel: every rank takes the same route (tensor pipe, gene shards), whatever its batch density
            if not (tc_ok and self._tc(out.K)):
                raise RuntimeError("the data-parallel route needs the bf16 tensor-pipe kernels (hidden sizes % 8 == 0)")
            dpm = self._dp_prepare(gexp, enc[0], crow, col, val, nnz, B, G, out.K)
            use_tc_spmm = True
        elif self._gmode is not None:
            # graph mode: always the tensor-pipe route; nnz is read on the device, `cap` records are packed
            if not tc_ok:
                raise RuntimeError("graph mode needs the bf16 tensor-pipe SpMM (hidden size % 8 == 0, G <= 65536)")
            use_tc_spmm = True
            cap = self._gmode["cap"]
            tp = ops.csr_tile_ptr_dyn(crow, col, val, G, cap,
                                      self.ws("tp64", (B * ((G + 63) // 64 + 1),), torch.int32),
                                      self.ws_cap("packed", (cap + 3) // 4 * 4 + 4, torch.int32))
        else:
            use_tc_spmm = tc_ok and self.spmm_picks_tensor(nnz, B, G)
            if use_tc_spmm:
                n_packed = (nnz + 3) // 4 * 4 + 4
                tp = ops.csr_tile_ptr(crow, col, val, G, nnz,
                                      self.ws("tp64", (B * ((G + 63) // 64 + 1),), torch.int32),
                                      self.ws_cap("packed", n_packed, torch.int32))
        self._t1(ev)
        ev_mid = None
        for j, lp in enumerate(enc):
            Y_pre = self._dp_first_layer_fwd(dpm, gexp, lp, B) if (dpm is not None and j == 0) else None
            x32, x16, caches[("enc", j)] = self._layer_fwd(f"enc{j}", lp, x32, x16, B,
                                                            csr=(crow, col, val, G, tp) if j == 0 else None,
                                                            masks=masks, Y_pre=Y_pre)
            if j == 0:
                ev_mid = self._t0("mid_fwd")     # everything between the two gene-sized layers
        hidden = []
        for j, lp in enumerate(self.vaeenc_plan):
            x32, x16, caches[("venc", j)] = self._layer_fwd(f"venc{j}", lp, x32, x16, B, masks=masks)
            if lp.return_hidden and lp.relu:
                hidden.append(("venc", j, x32, x16))
        q32, q16 = x32, x16
        ML = self.ws("ML", (B, 2 * Z))
        if self._tf(self.Hv, 2 * Z):
            ops.gemm(q32, 0, self.Wmv32, 0, B, 2 * Z, self.Hv, bias=self.bmv, C32=ML, tf32=True)
        elif self._tc(self.Hv, 2 * Z):
            ops.gemm(q16, 0, self.Wmv16, 0, B, 2 * Z, self.Hv, bias=self.bmv, C32=ML)
        else:
            ops.gemm(q32, 0, self.Wmv32, 0, B, 2 * Z, self.Hv, bias=self.bmv, C32=ML, use_tc=False)
        if eps is None:
            eps = L.draw_noise(B, Z, dev)
        z32 = self.ws("z32", (B, Z))
        z16 = self.ws("z16", (B, Z), torch.bfloat16) if bf else None
        ops.reparam_kl_fwd(ML, eps, Z, self.var_eps, z32, z16, sc[1:4])
        if self.hidden_z:
            hidden.append(("z", 0, z32, z16))
        x32, x16 = self._after_reparameterize(z32, z16, B)
        z_out = x32
        for j, lp in enumerate(self.vaedec_plan):
            x32, x16, caches[("vdec", j)] = self._layer_fwd(f"vdec{j}", lp, x32, x16, B, masks=masks)
        for j, lp in enumerate(dec[:-1]):
            x32, x16, caches[("dec", j)] = self._layer_fwd(f"dec{j}", lp, x32, x16, B, masks=masks)
        h32, h16 = x32, x16
        self._t1(ev_mid)
        capturing = self._gmode is not None and self._gmode["graphs"] is not None
        if capturing:       # first graph ends here: the join with the background stream happens between the two
            self._gmode["graphs"][0].capture_end()
        gexp.join_background()      # the output layer's update of the previous step (background stream)
        if capturing:
            self._gmode["graphs"][1].capture_begin(pool=self._gmode["graphs"][0].pool(),
                                                   capture_error_mode="thread_local")
        fused = self._tc(out.K) and bf
        H1 = out.K
        # single process: the two big weight-gradient kernels add their own sum of squares to the clip norm
        fuse_norm = dpm is None and fused and use_tc_spmm
        if dpm is not None:
            # decoder + loss + dWout rows + dh partial sums on (all cells) x (this rank's genes)
            dh, dl = self._dp_decoder(dpm, gexp, out, h16, B)
        elif fused:
            ldd = _ceil(G, 64)
            dl = self.ws("dlogits16", (B, ldd), torch.bfloat16, zero=True)
            wsb = self.ws("tileptr", (ops.decoder_mse_fused_workspace_bytes(B, G),), torch.uint8)
            ev = self._t0("decoder_mse_fused")
            ops.decoder_mse_fused(h16, out.W16, out.b, G, crow, col, val, dl, sc[0:1], wsb,
                                  tile_ptr=tp[0] if tp is not None else None)
            self._t1(ev)
        else:
            logits = self.ws("logits32", (B, G))
            ops.gemm(h32, 0, out.W32, 0, B, G, out.K, bias=out.b, C32=logits, use_tc=False)
            dl = self.ws("dlogits32", (B, G))
            ops.mse_relu_csr(logits, G, crow, col, val, False, dl, None, sc[0:1])

        od = self.odisc.get(expert_id)
        if od is not None:
            if not (fused and use_tc_spmm):
                raise RuntimeError("the output discriminator runs on the fused bf16 decoder route only")
            if dpm is not None:
                dpm["od_norm_slot"] = sc[od_slot + 1:od_slot + 2]
                self._dp_output_disc_step(od, expert_id, dpm, dl, B, sc[od_slot:od_slot + 1], sc[od_slot + 1:od_slot + 2])
            else:
                self._output_disc_step(od, expert_id, crow, col, val, nnz, tp, dl, B, G, sc[od_slot:od_slot + 1],
                                       sc[od_slot + 1:od_slot + 2])

        # ---------------- adversaries: discriminator update, then generator pass ----------------
        d_hidden = {}
        slot = ce_base
        if n_adv:
            assert labels is not None, "adversaries need labels"
            for i in range(n_adv):
                ap = self.adv[i]
                hid, hid16 = hidden[i][2], hidden[i][3]
                code, ac, logits_a = self._adv_fwd(f"adv{i}", ap, hid, hid16, B)
                slots = [sc[slot + k:slot + k + 1] for k in range(len(ap.conditions))]
                slot += len(ap.conditions)
                dla = self._adv_loss(f"adv{i}", ap, logits_a, labels, 1.0, B, slots)
                self._adv_bwd(f"adv{i}", ap, code, ac, dla, B, need_dx=False)
                if dpm is not None:      # sum of the discriminator gradients over ranks (replicated group)
                    self._dp_allreduce_start(dpm, ap.group, 0, ap.group.n)
                    self._dp_allreduce_finish(dpm, ap.group, 0, ap.group.n)
                ap.group.grad_norm_sq(s_norm(2 + i))
                ap.group.clip_adam(s_norm(2 + i), self.clip.get("adversarial"), gscale, advance=self._gmode is None)
            for i in range(n_adv):
                ap = self.adv[i]
                hid, hid16 = hidden[i][2], hidden[i][3]
                code, ac, logits_a = self._adv_fwd(f"adv{i}", ap, hid, hid16, B)
                slots = [sc[slot + k:slot + k + 1] for k in range(len(ap.conditions))]
                slot += len(ap.conditions)
                dla = self._adv_loss(f"adv{i}", ap, logits_a, labels, float(self.adv_weight), B, slots)
                d_hidden[i] = self._adv_bwd(f"adv{i}", ap, code, ac, dla, B, need_dx=True)
                ap.group.grad_norm_sq(s_norm(2 + n_adv + i))   # "generator_i" norm: logged, never applied

        # ---------------- backward ----------------
        if dpm is not None:
            pass                                                           # dWout / dbout / dh came with the decoder
        elif fused:
            dh = self.ws("dh", (B, H1))
            ev = self._t0("dWout_gemm")
            ops.gemm(dl, 1, h16, 1, G, H1, B, C32=out.gW,                 # dWout = dlogits^T h (+ its ||.||^2)
                     sumsq_out=s_norm(1) if fuse_norm else None)
            self._t1(ev)
            ops.colsum(dl, out.gb, M=B, N=G, accumulate=True)
            ev = self._t0("dh_gemm")
            ops.gemm(dl, 0, out.W16, 1, B, H1, G, C32=dh)                 # dh = dlogits Wout
            self._t1(ev)
        else:
            dh = self.ws("dh", (B, H1))
            ops.gemm(dl, 1, h32, 1, G, H1, B, C32=out.gW, use_tc=False)
            ops.colsum(dl, out.gb, accumulate=True)
            ops.gemm(dl, 0, out.W32, 1, B, H1, G, C32=dh, use_tc=False)
        d = dh
        ev_mid = self._t0("mid_bwd")
        for j in reversed(range(len(dec) - 1)):
            d = self._layer_bwd(f"dec{j}", dec[j], caches[("dec", j)], d, B)
        for j in reversed(range(len(self.vaedec_plan))):
            d = self._layer_bwd(f"vdec{j}", self.vaedec_plan[j], caches[("vdec", j)], d, B)
        dz = self.cond.backward(d, self.ws) if self.cond is not None else d
        for i in range(n_adv):
            if hidden[i][0] == "z":
                ops.axpy(dz, d_hidden[i], -1.0)                            # GRL: -alpha * grad, alpha = 1
        dML = self.ws("dML", (B, 2 * Z))
        dML16 = self.ws("dML16", (B, 2 * Z), torch.bfloat16) if bf else None
        if self._gmode is not None:   # graph mode: the KL weight of this step lives in device memory
            ops.reparam_kl_bwd(ML, eps, dz, Z, self.var_eps, 1.0 / B, dML, dML16, kl_weight_dev=self._dyn["klw"])
        else:
            ops.reparam_kl_bwd(ML, eps, dz, Z, self.var_eps, float(kl_weight) / B, dML, dML16)
        dq = self.ws("dq", (B, self.Hv))
        if self._tf(self.Hv, 2 * Z):
            self._on_side(lambda: (ops.colsum(dML, self.gbmv, accumulate=True),
                                   ops.gemm(dML, 1, q32, 1, 2 * Z, self.Hv, B, C32=self.gWmv, tf32=True)))
            ops.gemm(dML, 0, self.Wmv32, 1, B, self.Hv, 2 * Z, C32=dq, tf32=True)
        elif self._tc(self.Hv, 2 * Z):
            self._on_side(lambda: (ops.colsum(dML, self.gbmv, accumulate=True),
                                   ops.gemm(dML16, 1, q16, 1, 2 * Z, self.Hv, B, C32=self.gWmv)))
            ops.gemm(dML16, 0, self.Wmv16, 1, B, self.Hv, 2 * Z, C32=dq)
        else:
            self._on_side(lambda: (ops.colsum(dML, self.gbmv, accumulate=True),
                                   ops.gemm(dML, 1, q32, 1, 2 * Z, self.Hv, B, C32=self.gWmv, use_tc=False)))
            ops.gemm(dML, 0, self.Wmv32, 1, B, self.Hv, 2 * Z, C32=dq, use_tc=False)
        d = dq
        for j in reversed(range(len(self.vaeenc_plan))):
            for i in range(n_adv):
                if hidden[i][0] == "venc" and hidden[i][1] == j:
                    ops.axpy(d, d_hidden[i], -1.0)
            d = self._layer_bwd(f"venc{j}", self.vaeenc_plan[j], caches[("venc", j)], d, B)
        if dpm is not None:
            csc = ("dp", dpm)
        elif use_tc_spmm:
            csc = ("tc", tp, G, s_norm(1) if fuse_norm else None)
        else:
            cptr, ridx, cval = ops.csr_transpose(crow, col, val, G, nnz, self.ws("cptr", (G + 1,), torch.int32),
                                                 self.ws_cap("ridx", max(nnz, 1), torch.int32),
                                                 self.ws_cap("cval", max(nnz, 1)),
                                                 self.ws("cursor", (G + 1,), torch.int32))
            csc = ("gather", cptr, ridx, cval, G)
        for j in reversed(range(len(enc))):
            if j == 0:
                self._t1(ev_mid)
                if dpm is not None:
                    self._join_side()     # the small weight gradients are about to travel: they must be complete
            ev = self._t0("csr_linear_bwd_w+bn") if j == 0 else None
            d = self._layer_bwd(f"enc{j}", enc[j], caches[("enc", j)], d, B, need_dx=(j > 0),
                                csc=csc if j == 0 else None)
            self._t1(ev)

        # ---------------- grad norms, clip, Adam ----------------
        self._join_side()
        ev = self._t0("norm+clip_adam")
        if dpm is not None:
            ev2 = self._t0("dp_wait_grads")
            self._dp_allreduce_finish(dpm, gexp, gexp.tail_lo, gexp.n)
            self._dp_allreduce_finish(dpm, gvae, 0, gvae.n)
            self._t1(ev2)
            gvae.grad_norm_sq(s_norm(0))
            ops.sumsq(gexp.g[gexp.tail_lo:gexp.n], s_norm(1))          # replicated part (summed over ranks)
            self._dp_finish_scalars(dpm, sc, s_norm(1))                 # + every rank's rows; recon of MY cells
            sc[-2:].copy_(dpm["info"])
        else:
            gvae.grad_norm_sq(s_norm(0))
            if self.cond is not None:       # same optimizer as the VAE (cmmvae_model.py:311): one clip norm
                self.cond.add_norm_sq(s_norm(0))
            gexp.grad_norm_sq(s_norm(1), skip=[enc[0].lin.weight, out.lin.weight] if fuse_norm else None)
        bg = self._bg if self.pipeline_optimizer else None
        gvae.clip_adam(s_norm(0), self.clip.get("vae"), gscale, advance=self._gmode is None)
        if self.cond is not None:
            self.cond.clip_adam(s_norm(0), self.clip.get("vae"), gscale)
        launch_bg = gexp.clip_adam(s_norm(1), self.clip.get("expert"), gscale, background=bg,
                                   defer_background=capturing, advance=self._gmode is None)
        self._t1(ev)

        self.last = dict(sc=sc, B=B, Z=Z, kl_weight=float(kl_weight), expert_id=expert_id, n_adv=n_adv,
                         gscale=gscale, ce_base=ce_base, z=z_out, dl=dl, dp=dpm is not None,
                         od_slot=od_slot if od is not None else None)
        if launch_bg is not None:
            self.last["launch_bg"] = launch_bg
        return self.last

    # ---------------------------------------------------------------------------------------- logs
    def scalars_async(self, rec=None):
        """start a non-blocking copy of the step's scalar block into pinned host memory; returns
        (pinned tensor, event) to pass to ``scalars(host=...)`` later"""
        rec = rec or self.last
        host = torch.empty(rec["sc"].shape, dtype=torch.float64, pin_memory=True)
        host.copy_(rec["sc"], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev

    def scalars(self, rec=None, host=None) -> Dict[str, float]:
        """Host copy (one sync) of every value the reference logs for the step, untagged keys."""
        rec = rec or self.last
        if host is not None:
            host[1].synchronize()
            sc = host[0].tolist()
        else:
            sc = rec["sc"].cpu().tolist()
        B, Z, n_adv = rec["B"], rec["Z"], rec["n_adv"]
        if rec.get("dp") and sc[-1] != 0:
            raise RuntimeError(f"data parallel: a gene shard of this rank's batch holds {int(sc[-2])} non-zeros, more "
                               "than the exchange capacity agreed on the first step (CMMVAE_DP_CSR_HEADROOM, default "
                               "1.5 x nnz / world; raise it for gene panels whose shards are unevenly populated)")
        out = {"recon_loss": sc[0], "kl_loss": sc[1] / B, "kl_weight": rec["kl_weight"],
               "Mean": sc[2] / (B * Z), "Variance": sc[3] / (B * Z)}
        total = out["recon_loss"] + rec["kl_weight"] * out["kl_loss"]
        out["grad_norms/vae"] = math.sqrt(sc[4]) * rec["gscale"]
        out[f"grad_norms/expert_{rec['expert_id']}"] = math.sqrt(sc[5]) * rec["gscale"]
        slot = rec["ce_base"]
        for tag_i, tag in enumerate(("discriminator", "generator")):
            for i in range(n_adv):
                ap = self.adv[i]
                summed = 0.0
                for c in ap.conditions:
                    out[f"{tag}_{i + 1}/adversarial_loss/{c}"] = sc[slot]
                    summed += sc[slot]
                    slot += 1
                out[f"{tag}_{i + 1}/adversarial_loss/summed"] = summed
                out[f"grad_norms/{tag}_{i + 1}"] = math.sqrt(sc[6 + tag_i * n_adv + i]) * (
                    rec["gscale"] if tag == "discriminator" else 1.0)
                if tag == "generator":
                    total += self.adv_weight * summed
        out["loss"] = total
        if rec.get("od_slot") is not None:     # reference writer tags: meta_disc/md_<species> (meta_discriminators.py:165-167)
            out[f"meta_disc/md_{rec['expert_id']}"] = sc[rec["od_slot"]]
            out[f"grad_norms/output_discriminator_{rec['expert_id']}"] = math.sqrt(sc[rec["od_slot"] + 1]) * rec["gscale"]
        return out

    # ------------------------------------------------------------------------------- eval forward
    @torch.no_grad()
    def eval_step(self, expert_id: str, crow, col, val, eps=None, kl_weight: float = 1.0, metadata=None):
        """validation_step arithmetic (cmmvae_model.py:219-245): eval-mode forward + ELBO on the fused
        decoder path.  Returns the scalar record (use ``scalars``-like host read via ``eval_scalars``)."""
        enc, dec = self.enc_plan[expert_id], self.dec_plan[expert_id]
        B, G, Z = crow.numel() - 1, enc[0].K, self.Z
        bf = self.precision == "bf16"
        if self.cond is not None:
            if metadata is None:
                raise ValueError("conditional layers need the batch's metadata")
            self.cond.make_plan(metadata, expert_id, B)
        self.groups[f"experts/{expert_id}"].sync_master()   # joins a background update; data parallel: gathers the rows
        sc = torch.zeros(4, dtype=torch.float64, device=self.device)
        x32 = x16 = None
        for j, lp in enumerate(enc):
            x32, x16, _ = self._layer_fwd(f"enc{j}", lp, x32, x16, B,
                                          csr=(crow, col, val, G, None) if j == 0 else None, training=False)
        for j, lp in enumerate(self.vaeenc_plan):
            x32, x16, _ = self._layer_fwd(f"venc{j}", lp, x32, x16, B, training=False)
        ML = self.ws("ML", (B, 2 * Z))
        if self._tc(self.Hv, 2 * Z):
            ops.gemm(x16, 0, self.Wmv16, 0, B, 2 * Z, self.Hv, bias=self.bmv, C32=ML)
        else:
            ops.gemm(x32, 0, self.Wmv32, 0, B, 2 * Z, self.Hv, bias=self.bmv, C32=ML, use_tc=False)
        if eps is None:
            eps = L.draw_noise(B, Z, self.device)
        z32 = self.ws("z32", (B, Z))
        z16 = self.ws("z16", (B, Z), torch.bfloat16) if bf else None
        ops.reparam_kl_fwd(ML, eps, Z, self.var_eps, z32, z16, sc[1:4])
        x32, x16 = self._after_reparameterize(z32, z16, B)
        z_out = x32          # (vae.py:98-102 returns z AFTER after_reparameterize)
        for j, lp in enumerate(self.vaedec_plan):
            x32, x16, _ = self._layer_fwd(f"vdec{j}", lp, x32, x16, B, training=False)
        for j, lp in enumerate(dec[:-1]):
            x32, x16, _ = self._layer_fwd(f"dec{j}", lp, x32, x16, B, training=False)
        out = dec[-1]
        if self._tc(out.K) and bf:
            dl = self.ws("dlogits16", (B, _ceil(G, 64)), torch.bfloat16, zero=True)
            wsb = self.ws("tileptr", (ops.decoder_mse_fused_workspace_bytes(B, G),), torch.uint8)
            ops.decoder_mse_fused(x16, out.W16, out.b, G, crow, col, val, dl, sc[0:1], wsb)
        else:
            logits = self.ws("logits32", (B, G))
            ops.gemm(x32, 0, out.W32, 0, B, G, out.K, bias=out.b, C32=logits, use_tc=False)
            ops.mse_relu_csr(logits, G, crow, col, val, False, None, None, sc[0:1])
        s = sc.cpu().tolist()
        kl = s[1] / B
        return {"loss": s[0] + kl_weight * kl, "recon_loss": s[0], "kl_loss": kl, "kl_weight": kl_weight,
                "z": z_out}
