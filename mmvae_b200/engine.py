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
  DP        one gradient all-reduce per group over NCCL when torch.distributed is initialised

Parameters of each optimizer group live in one flat fp32 buffer (``FlatGroup``); the nn.Parameters of
the modules are views into it, so ``state_dict()`` keeps the reference's names and shapes.  The first
expert-encoder weight is stored physically transposed ([genes, hidden]) for coalesced SpMM row reads
and exposed as a ``[hidden, genes]`` view.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
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

    Layout: ``[segment 0 | segment 1 | ... | tail]``.  With one process the distinction is moot (one range).
    With ``world`` > 1 ranks each *segment* (the big weight matrices) is ZeRO-1 sharded: gradients are
    reduce-scattered, every rank runs clip+Adam on its 1/world shard only (Adam state exists only for the
    shard) and the refreshed bf16 shadow is all-gathered; the fp32 master copy of a segment is current
    only inside the owner's shard until ``sync_master()``.  The *tail* (biases, BatchNorm affine, small
    matrices) is replicated: all-reduced gradients, identical Adam on every rank."""

    ALIGN = 64

    def __init__(self, name: str, chains: List[List[tuple]], device, lr=5e-3, weight_decay=1e-6,
                 betas=(0.9, 0.999), eps=1e-8, segments: Optional[List[List[List[tuple]]]] = None,
                 split_first: int = 1, row_shard_first: bool = False):
        """``chains`` (tail) / ``segments[i]`` (sharded when world > 1): lists of chains; a chain is a list
        of ``(param, transposed)`` stored back to back (e.g. mean/var head weights = one matrix)."""
        self.name, self.lr, self.wd, self.betas, self.eps = name, lr, weight_decay, betas, eps
        self.world, self.rank = dp.world_size(), dp.rank()
        self.params: List[nn.Parameter] = []
        self.offset: Dict[int, int] = {}
        self.transposed: Dict[int, bool] = {}
        segments = segments or []
        seg_align = self.world * 256
        total = 0
        self.seg_bounds: List[tuple] = []

        def place(chain_list):
            nonlocal total
            for chain in chain_list:
                total = _ceil(total, self.ALIGN)
                for p, tr in chain:
                    self.params.append(p)
                    self.offset[id(p)] = total
                    self.transposed[id(p)] = tr
                    total += p.numel()

        self.first_rows: List[tuple] = []      # row ranges of the first parameter covered by segments 0..k-1
        # row_shard_first (world > 1): segment 0 holds ONE row-major matrix and is padded to world * Rs rows
        # (Rs a multiple of 128) so that rank r's ZeRO shard is exactly rows [r*Rs, (r+1)*Rs): the rank can then
        # COMPUTE its shard of the summed gradient from all-gathered inputs instead of reduce-scattering it
        self.row_shard: Optional[tuple] = None   # (rows per rank, width, padded rows)
        for si, seg in enumerate(segments):
            total = _ceil(total, seg_align)
            lo = total
            place(seg)
            if si == 0 and row_shard_first and self.world > 1:
                assert len(seg) == 1 and len(seg[0]) == 1, "row-sharded segment = one matrix"
                p0, tr0 = seg[0][0]
                rows, width = (p0.shape[1], p0.shape[0]) if tr0 else (p0.shape[0], p0.shape[1])
                per = dp.shard_rows(rows, self.world)
                assert (per * width) % 256 == 0
                total = lo + self.world * per * width
                self.row_shard = (per, width, self.world * per)
                self.seg_bounds.append((lo, total))
                self.first_rows = [(0, 0)]
                continue
            total = _ceil(total, seg_align)
            if si == 0 and split_first > 1 and self.world > 1:
                # cut segment 0 inside its first (big, row-major) parameter at 128-row boundaries so that
                # finished row ranges of its gradient can be exchanged while the rest is still computed
                p0, tr0 = seg[0][0]
                rows, width = (p0.shape[1], p0.shape[0]) if tr0 else (p0.shape[0], p0.shape[1])
                per = _ceil((rows + split_first - 1) // split_first, 128)
                cuts = [r for r in range(per, rows, per)]
                if all((r * width) % seg_align == 0 for r in cuts) and self.offset[id(p0)] == lo:
                    edges = [0] + cuts + [rows]
                    starts = [lo + r * width for r in edges[:-1]]
                    ends = starts[1:] + [total]
                    self.seg_bounds += list(zip(starts, ends))
                    self.first_rows = list(zip(edges[:-1], edges[1:]))
                    continue
            self.seg_bounds.append((lo, total))
            if si == 0:
                self.first_rows = [(0, 0)]
        self.n_first = max(len(self.first_rows), 1)   # segments 0..n_first-1 = pieces of the original segment 0
        self.tail_lo = total
        place(chains)
        self.n = _ceil(max(total, 4), 4)
        self.sharded = self.world > 1 and len(self.seg_bounds) > 0
        # ranges the optimizer walks: (flat lo, flat hi, grad buffer, offset into m/v)
        self.p = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.g = torch.zeros(self.n, device=device, dtype=torch.float32)
        self.p16 = torch.zeros(self.n, device=device, dtype=torch.bfloat16)
        self.gs: List[torch.Tensor] = []       # reduce-scatter outputs (one per segment)
        self.ranges: List[tuple] = []
        mv = 0
        if self.sharded:
            for lo, hi in self.seg_bounds:
                ns = (hi - lo) // self.world
                gs = torch.zeros(ns, device=device, dtype=torch.float32)
                self.gs.append(gs)
                own = lo + self.rank * ns
                self.ranges.append((own, own + ns, gs, mv))
                mv += ns
            self.ranges.append((self.tail_lo, self.n, self.g[self.tail_lo:self.n], mv))
            mv += self.n - self.tail_lo
        else:
            self.ranges.append((0, self.n, self.g, 0))
            mv = self.n
        self.m = torch.zeros(mv, device=device, dtype=torch.float32)
        self.v = torch.zeros(mv, device=device, dtype=torch.float32)
        self.step_count = 0
        self.applied = False    # set by the fused step after its own clip+Adam launch; consumed by FlatAdam.step()
        self._vec_range: Optional[tuple] = None
        self._deferred: Optional[torch.cuda.Event] = None   # output-layer update in flight on the background stream
        self.first_by_inputs = False   # set per step by the engine when gs[0] was produced from gathered inputs
        self._ag_pending: List = []
        for p in self.params:
            phys = self.phys(p)
            src = p.data.to(device)
            phys.copy_(src.t() if self.transposed[id(p)] else src)
            p.data = phys.t() if self.transposed[id(p)] else phys
            gphys = self.phys(p, self.g)
            p.grad = gphys.t() if self.transposed[id(p)] else gphys
            L.SHADOWS[id(p)] = self.phys(p, self.p16)
        self.refresh_shadow()

    def zero_vector_grads(self):
        """one fill per step over the gradients of the 1-D parameters (biases, BatchNorm affine): the kernels
        that produce them accumulate, instead of each issuing its own memsets in the middle of the backward
        pass.  Matrix gradients are overwritten by their GEMMs and need no zeroing."""
        if self._vec_range is None:
            vec = [(self.offset[id(p)], self.offset[id(p)] + p.numel()) for p in self.params if p.dim() == 1]
            if not vec:
                self._vec_range = (0, 0)
            elif self.seg_bounds:       # tail = vectors first, then (data parallel) the small matrices
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

    def refresh_shadow(self):
        ops.cast_bf16(self.p, self.p16)

    # ---- gradient exchange (world > 1) ----
    def exchange_segment_async(self, i: int):
        """reduce-scatter (SUM) of segment i's gradient into this rank's shard buffer"""
        if not self.sharded:
            return None
        if i == 0 and self.first_by_inputs:
            return None    # the shard of the summed gradient was computed in place (StepEngine._first_layer_grad_shard)
        lo, hi = self.seg_bounds[i]
        return torch.distributed.reduce_scatter_tensor(self.gs[i], self.g[lo:hi], async_op=True)

    def exchange_rest_async(self):
        """all-reduce (SUM) of everything that is not sharded"""
        if self.world == 1:
            return None
        lo = self.tail_lo if self.sharded else 0
        return dp.allreduce_sum_(self.g[lo:self.n], async_op=True)

    def grad_norm_sq(self, out: torch.Tensor, skip: Optional[List[nn.Parameter]] = None):
        """out (double[1], pre-zeroed) += || sum_r g_r ||^2 over the whole group (padding is zero).
        ``skip``: parameters whose contribution a producer kernel already added to ``out`` (single process)."""
        if skip and not self.sharded and self.world == 1:
            cuts = sorted((self.offset[id(p)], self.offset[id(p)] + p.numel()) for p in skip)
            lo = 0
            for a, b in cuts + [(self.n, self.n)]:
                a4, lo4 = a // 4 * 4, _ceil(lo, 4)   # ranges start/stop on parameter boundaries (64-aligned)
                if a4 > lo4:
                    ops.sumsq(self.g[lo4:a4], out)
                lo = b
            return
        if self.sharded:
            for gs in self.gs:
                ops.sumsq(gs, out)
            torch.distributed.all_reduce(out)
            ops.sumsq(self.g[self.tail_lo:self.n], out)
        else:
            ops.sumsq(self.g, out)

    def clip_adam(self, norm_sq: torch.Tensor, max_norm: Optional[float], grad_scale: float = 1.0,
                  background: Optional[torch.cuda.Stream] = None):
        """fused clip + Adam over the group.  ``background`` (single process only): the LAST segment (the output
        layer, which the next forward pass reads last) is updated on that low-priority stream, after everything
        else, so the update runs underneath the next step's forward; ``wait_shadow("rest")`` joins it."""
        self.step_count += 1
        self.applied = True
        self._ag_pending = []
        hyper = (self.lr, self.betas[0], self.betas[1], self.eps, self.wd, self.step_count)
        if background is not None and not self.sharded and len(self.seg_bounds) >= 2:
            lo1, hi1 = self.seg_bounds[-1]
            for lo, hi in ((0, lo1), (hi1, self.n)):
                if hi > lo:
                    ops.clip_adam(self.p[lo:hi], self.g[lo:hi], self.m[lo:hi], self.v[lo:hi], self.p16[lo:hi],
                                  norm_sq, max_norm or 0.0, grad_scale, *hyper)
            first_done = torch.cuda.Event()
            first_done.record()
            # the clip norm lives in the step's scalar block: keep the allocator from recycling that block for a
            # later step while the background launch may still read it
            norm_sq.record_stream(background)
            with torch.cuda.stream(background), ops.stream_scope(background):
                background.wait_event(first_done)
                ops.clip_adam(self.p[lo1:hi1], self.g[lo1:hi1], self.m[lo1:hi1], self.v[lo1:hi1], self.p16[lo1:hi1],
                              norm_sq, max_norm or 0.0, grad_scale, *hyper, background=True)
                self._deferred = torch.cuda.Event()
                self._deferred.record(background)
            return
        for i, (lo, hi, g, mv) in enumerate(self.ranges):
            n = hi - lo
            ops.clip_adam(self.p[lo:hi], g, self.m[mv:mv + n], self.v[mv:mv + n], self.p16[lo:hi], norm_sq,
                          max_norm or 0.0, grad_scale, *hyper)
            if self.sharded and i < len(self.seg_bounds):
                # publish the refreshed bf16 shard right away (segment 0 = first-layer weight is needed first by
                # the next forward); consumers wait just before they read it
                slo, shi = self.seg_bounds[i]
                self._ag_pending.append(torch.distributed.all_gather_into_tensor(
                    self.p16[slo:shi], self.p16[lo:hi], async_op=True))

    def wait_shadow(self, which: Optional[str] = None):
        """make the current stream wait for the all-gather of bf16 shadows: "first" = the pieces of the original
        segment 0 (first-layer weight), "rest" = the other segments, None = all; "rest"/None also join an
        output-layer update still running on the background stream"""
        for j, w in enumerate(self._ag_pending):
            hit = which is None or (which == "first") == (j < self.n_first)
            if w is not None and hit:
                w.wait()
                self._ag_pending[j] = None
        if which != "first" and self._deferred is not None:
            torch.cuda.current_stream().wait_event(self._deferred)
            self._deferred = None

    def logical(self, p: nn.Parameter, buf: torch.Tensor) -> torch.Tensor:
        """view of ``buf`` (laid out like the value buffer) with parameter ``p``'s logical shape"""
        t = self.phys(p, buf)
        return t.t() if self.transposed[id(p)] else t

    def full_moments(self):
        """Adam moments laid out like the value buffer (``[n]`` each).  Single process: the live buffers.
        ZeRO-sharded: fresh full-size copies, the sharded segments all-gathered from their owners."""
        if not self.sharded:
            return self.m, self.v
        out = []
        for src in (self.m, self.v):
            full = torch.zeros(self.n, device=src.device, dtype=torch.float32)
            for (lo, hi), (own_lo, own_hi, _, mv) in zip(self.seg_bounds, self.ranges):
                torch.distributed.all_gather_into_tensor(full[lo:hi], src[mv:mv + own_hi - own_lo].contiguous())
            _, _, _, mv = self.ranges[-1]
            full[self.tail_lo:self.n].copy_(src[mv:mv + self.n - self.tail_lo])
            out.append(full)
        return out

    def store_moments(self, m_full: torch.Tensor, v_full: torch.Tensor):
        """inverse of ``full_moments`` (no-op single process: the views were written in place)"""
        if not self.sharded:
            return
        for src, dst in ((m_full, self.m), (v_full, self.v)):
            for (own_lo, own_hi, _, mv) in self.ranges:
                dst[mv:mv + own_hi - own_lo].copy_(src[own_lo:own_hi])

    def sync_master(self):
        """all-gather the fp32 master copy of the sharded segments (before state_dict / fp32 evaluation)"""
        if not self.sharded:
            self.wait_shadow()
            return
        self.wait_shadow()
        for (lo, hi), (own_lo, own_hi, _, _) in zip(self.seg_bounds, self.ranges):
            torch.distributed.all_gather_into_tensor(self.p[lo:hi], self.p[own_lo:own_hi])


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

    def __init__(self, group: FlatGroup):
        self.flat = group
        self._max_norm = None
        super().__init__(group.params, dict(lr=group.lr, weight_decay=group.wd, betas=group.betas, eps=group.eps))

    def set_clip(self, max_norm: Optional[float]):
        self._max_norm = max_norm

    @torch.no_grad()
    def step(self, closure=None):
        if self.flat.applied:
            self.flat.applied = False
            return
        if self.flat.world > 1:
            raise NotImplementedError("with torch.distributed the exchange + step run inside training_step")
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
        pg = dict(self.param_groups[0])
        pg["params"] = list(range(len(g.params)))
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
                 precision: Optional[str] = None, device=None):
        self.module = module
        self.device = torch.device(device or "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("StepEngine needs a CUDA device (no CPU fallback)")
        self.precision = precision or L.get_precision()
        # world > 1: the first-layer weight gradient shard is computed from all-gathered inputs (packed CSR, window
        # pointers, dY) instead of reduce-scattering 4*G*H1 bytes of output; CMMVAE_DP_BY_INPUTS=0 -> reduce-scatter
        self.dp_by_inputs = os.environ.get("CMMVAE_DP_BY_INPUTS", "1") != "0" and self.precision == "bf16"
        self._dp_cap = None
        # single process: run the step on a high-priority stream and the output layer's clip+Adam on a
        # low-priority one, underneath the next step's forward pass (37 % of a B=1024 step is optimizer HBM
        # traffic, half of it the output layer whose new value is only needed by the decoder kernel).  Weights
        # read outside the engine need ``finish()`` first, hence opt-in (training_step turns it on in its
        # pipelined mode, sync_logging=False)
        self.pipeline_optimizer = False
        self._hp = self._bg = None
        self.dp_chunks = int(os.environ.get("CMMVAE_DP_CHUNKS", "2"))     # world > 1: pieces the first-layer weight gradient is exchanged in (overlap)
        self.adv_weight = adv_weight
        self.clip = clip or {"vae": 10.0, "expert": 10.0, "adversarial": 10.0}
        vae = module.vae
        if getattr(vae, "conditionals", None):
            raise UnsupportedTopology("conditional layers are outside the fused step (SURVEY.md 8f-1)")
        enc = vae.encoder
        if isinstance(enc.z_transformation, nn.Softmax):
            raise UnsupportedTopology("distribution='ln' is outside the fused step")
        module.to(self.device)   # buffers (BN running stats) and not-yet-flattened params
        dev = self.device
        # ---- optimizer groups as flat buffers (order mirrors configure_optimizers) ----
        self.groups: Dict[str, FlatGroup] = {}
        self.enc_plan: Dict[str, List[LayerPlan]] = {}
        self.dec_plan: Dict[str, List[LayerPlan]] = {}
        for eid, expert in module.experts.items():
            # big matrices: segment 0 = everything but the output layer, segment 1 = the output layer (its
            # gradient is final first, so its exchange overlaps the rest of the backward pass); vectors: tail
            mats = _block_chains(expert.encoder, sparse_first=True, kind="matrix") + \
                _block_chains(expert.decoder, kind="matrix")
            vecs = _block_chains(expert.encoder, kind="vector") + _block_chains(expert.decoder, kind="vector")
            if dp.world_size() > 1 and self.dp_by_inputs:
                # first-layer weight alone in segment 0 with row-aligned shards (its gradient shard is computed
                # from all-gathered inputs); the two small matrices join the replicated tail
                g = FlatGroup(f"experts/{eid}", vecs + mats[1:-1], dev, segments=[mats[:1], mats[-1:]],
                              row_shard_first=True)
            else:
                g = FlatGroup(f"experts/{eid}", vecs, dev, segments=[mats[:-1], mats[-1:]],
                              split_first=self.dp_chunks)
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
        self._ws: Dict[tuple, torch.Tensor] = {}
        # small weight-gradient GEMMs are off the critical path (only the optimizer needs them): they run on a
        # side stream, concurrently with the dX chain of the main stream
        self.side = torch.cuda.Stream(device=self.device)
        self._side_used = False
        self.timers: Optional[Dict[str, list]] = None   # name -> [(start_event, end_event)] when profiling
        self.timer_filter = None   # optional set of section names to time (an event record ends a PDL chain)
        self._seed = 0x5EED
        self.nccl_sms = int(os.environ.get("CMMVAE_NCCL_SMS", "32"))                  # SMs left to communication kernels when world > 1
        self.spmm_tc = True                 # bf16 policy: expert-encoder SpMM on the tensor pipe ...
        self.spmm_tc_min_density = 0.015    # ... when the batch is at least this dense (else gather kernel)
        self.world = 1
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

    def _next_seed(self) -> int:
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) & ((1 << 63) - 1)
        return self._seed

    # ---------------------------------------------------------------------------------- dense layer
    def _linear(self, tag, lp: LayerPlan, x32, x16, B, fuse_relu):
        """Y = x W^T + b (optionally ReLU fused in the GEMM epilogue) -> (y32, y16|None)"""
        y32 = self.ws(tag + ".y32", (B, lp.N))
        if self._tc(lp.K, lp.N):
            y16 = self.ws(tag + ".y16", (B, lp.N), torch.bfloat16) if fuse_relu else None
            ops.gemm(x16, 0, lp.W16, 0, B, lp.N, lp.K, bias=lp.b, relu=fuse_relu, C32=y32, C16=y16)
            return y32, y16
        ops.gemm(x32, 0, lp.W32, 0, B, lp.N, lp.K, bias=lp.b, relu=fuse_relu, C32=y32, use_tc=False)
        return y32, None

    def _layer_fwd(self, tag, lp: LayerPlan, x32, x16, B, csr=None, training=True, masks=None):
        """full layer; returns (out32, out16, cache)"""
        want16 = self.precision == "bf16"
        plain = lp.bn is None and lp.p_drop == 0.0
        if lp.sparse:
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
        cache = dict(x32=x32, x16=x16, Y=Y, mean=None, rstd=None, seed=0, mask=None, p=0.0)
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
            seed = self._next_seed() if (p > 0 and mask is None) else 0
            out32 = self.ws(tag + ".o32", (B, lp.N))
            out16 = self.ws(tag + ".o16", (B, lp.N), torch.bfloat16) if want16 else None
            ops.bn_act_drop_fwd(Y, mean, rstd, lp.gamma if lp.bn is not None else None,
                                lp.beta if lp.bn is not None else None, lp.relu, p, seed, mask, out32, out16)
            cache.update(mean=mean, rstd=rstd, seed=seed, mask=mask, p=p)
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
                                cache["mask"], dY, dY16, lp.ggamma, lp.gbeta, lp.gb, accumulate=True)
        else:
            dY = dOut32
            dY16 = ops.cast_bf16(dY, self.ws(tag + ".dY16", (B, lp.N), torch.bfloat16)) if want16 else None
            ops.colsum(dY, lp.gb, accumulate=True)
        if lp.sparse:
            if csc[0] == "tc":
                _, tp, G, ssq, pending, gathered = csc
                if gathered is not None:
                    self._first_layer_grad_shard(lp.group, gathered, dY16, B)
                    return None
                pieces = lp.group.first_rows if (lp.group.sharded and lp.group.n_first > 1) else [(0, G)]
                for i, (g0, g1) in enumerate(pieces):
                    ops.csr_linear_bwd_w_tc(tp[1], tp[0], B, G, dY16, lp.gW, sumsq_out=ssq, g_begin=g0, g_end=g1)
                    if i + 1 < len(pieces):   # this row range is final: exchange it while the next one is computed
                        pending.append(lp.group.exchange_segment_async(i))
            else:
                _, cptr, ridx, cval, G = csc
                ops.csr_linear_bwd_w(cptr, ridx, cval, B, G, dY, lp.gW)
            return None
        dX = self.ws(tag + ".dX", (B, lp.K)) if need_dx else None
        if self._tc(lp.K, lp.N):
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
            if tc:
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
            if tc:
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

    # ------------------------------------------------------------- data parallel: first layer by inputs
    def _dp_capacity(self, n_packed: int, B: int) -> int:
        """records per rank in the all-gathered packed CSR (identical on every rank; agreed once, with headroom)"""
        if self._dp_cap is None:
            t = torch.tensor([n_packed, B, -B], dtype=torch.int64, device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            mx, bmax, bmin = (int(v) for v in t.tolist())
            if bmax != -bmin:
                raise RuntimeError("data-parallel step needs the same number of cells on every rank")
            self._dp_cap = _ceil(int(mx * 1.15) + 1024, 1024)
        if n_packed > self._dp_cap:
            raise RuntimeError(f"batch with {n_packed} non-zero records exceeds the data-parallel CSR capacity "
                               f"{self._dp_cap} agreed on the first step; set engine._dp_cap (same value on every "
                               "rank) before training on batches of very different density")
        return self._dp_cap

    def _dp_gather_csr(self, gexp: FlatGroup, tp, B: int, cap: int):
        """start the exchange that lets every rank compute ITS gene shard of the summed first-layer weight
        gradient: all-gather of the packed CSR records (4 B per non-zero) and all-to-all of the window pointers
        (rank r receives, from every rank, the pointer rows of r's windows, rebased into the gathered array).
        Runs on the NCCL stream under the forward pass; consumed by ``_first_layer_grad_shard``."""
        per, width, _ = gexp.row_shard
        world, rank, WS = self.world, dp.rank(), gexp.row_shard[0] // 64
        table, packed = tp
        packed_all = self.ws("dp.packed_all", (world * cap,), torch.int32)
        w_pk = torch.distributed.all_gather_into_tensor(packed_all, packed[:cap], async_op=True)
        tp2d = table.view(-1, B)                                # [world * WS + 1, B]
        send = self.ws("dp.tp_send", (world, WS + 1, B), torch.int32)
        send[:, :WS].copy_(tp2d[:world * WS].view(world, WS, B))
        send[:, WS].copy_(tp2d[WS::WS][:world])                 # closing row of every shard
        send += rank * cap
        recv = self.ws("dp.tp_recv", (world, WS + 1, B), torch.int32)
        w_tp = torch.distributed.all_to_all_single(recv.view(-1), send.view(-1), async_op=True)
        return dict(packed_all=packed_all, recv=recv, waits=[w_pk, w_tp], WS=WS)

    def _first_layer_grad_shard(self, gexp: FlatGroup, gathered, dY16, B: int):
        """dW1^T[shard] = X_all^T[shard] . dY_all  -- the shard of the SUM over ranks, written straight into the
        ZeRO gradient shard (no 4*G*H1-byte reduce-scatter; only dY, 2*B*H1 bytes per rank, is exchanged here)"""
        per, width, rows_pad = gexp.row_shard
        world, rank, WS = self.world, dp.rank(), gathered["WS"]
        for w in gathered["waits"]:
            w.wait()
        tp_shard = self.ws("dp.tp_shard", (WS + 1, world * B), torch.int32)
        tp_shard.view(WS + 1, world, B).copy_(gathered["recv"].permute(1, 0, 2))
        dY_all = self.ws("dp.dY_all", (world * B, width), torch.bfloat16)
        torch.distributed.all_gather_into_tensor(dY_all, dY16)
        ops.csr_linear_bwd_w_tc_shard(gathered["packed_all"], tp_shard, world * B, rows_pad, dY_all,
                                      gexp.gs[0].view(per, width), rank * per, (rank + 1) * per)

    # ----------------------------------------------------------------------------------------- step
    def train_step(self, expert_id: str, crow, col, val, nnz: int, kl_weight: float, eps=None,
                   labels: Optional[Dict[str, torch.Tensor]] = None, masks=None):
        """One optimisation step on a CSR batch already resident on the device (no host sync).
        Returns the step record (device scalar block etc.) for ``scalars()``."""
        if self.pipeline_optimizer and dp.world_size() == 1:
            if self._hp is None:
                lo_pri, hi_pri = torch.cuda.Stream.priority_range()
                self._hp = torch.cuda.Stream(self.device, priority=hi_pri)
                self._bg = torch.cuda.Stream(self.device, priority=lo_pri)
            cur = torch.cuda.current_stream()
            self._hp.wait_stream(cur)
            with torch.cuda.stream(self._hp), ops.stream_scope(self._hp):
                rec = self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, masks)
            cur.wait_stream(self._hp)
            return rec
        with ops.stream_scope(torch.cuda.current_stream()):
            return self._train_step(expert_id, crow, col, val, nnz, kl_weight, eps, labels, masks)

    def prefetch(self, expert_id: str, crow, col, val):
        """data parallel: start exchanging the NEXT batch's CSR records while the current step runs (no-op with
        one process).  ``train_step`` on the same arrays then finds its gathered inputs ready."""
        return None

    def finish(self):
        """join optimizer work still in flight on the background stream (call before reading weights outside
        the engine when ``pipeline_optimizer`` is on; ``state_dict``/evaluation do it themselves)"""
        for g in self.groups.values():
            g.wait_shadow()

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
        self.world = dp.world_size()
        # NCCL's all-gather (start of the forward) and reduce-scatter (backward after dWout) kernels hold some
        # SMs while they run: during those windows the persistent kernels plan for the remaining SMs so that
        # no planned CTA has to wait for a free SM; elsewhere they use all 148
        comm_budget = (lambda on: ops.set_sm_budget(148 - self.nccl_sms if on else 148)) if self.world > 1 \
            else (lambda on: None)
        comm_budget(True)
        gscale = 1.0 / self.world
        bf = self.precision == "bf16"
        n_adv = min(len(self.adv), self.n_hidden)   # zip(hidden, adversarials) truncates (cmmvae_model.py:67-70)
        # scalar slots (double): 0 recon | 1..3 kl,sum mu,sum var | then norms | then CE sums
        n_ce = sum(len(a.conditions) for a in self.adv[:n_adv])
        sc = torch.zeros(4 + 2 + 2 * n_adv + 2 * n_ce, dtype=torch.float64, device=dev)
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
        # data parallel: every rank must take the same route (it decides which collectives run), so the
        # density test is dropped there
        by_inputs = tc_ok and self.world > 1 and gexp.row_shard is not None
        use_tc_spmm = tc_ok and (by_inputs or nnz >= self.spmm_tc_min_density * B * G)
        gexp.first_by_inputs = by_inputs
        G_tp = gexp.row_shard[2] if by_inputs else G     # window table padded to world * rows-per-rank genes
        tp = None
        gathered = None
        ev = self._t0("csr_prep")
        if use_tc_spmm:
            n_packed = (nnz + 3) // 4 * 4 + 4
            if by_inputs:
                n_packed = self._dp_capacity(n_packed, B)
            tp = ops.csr_tile_ptr(crow, col, val, G_tp, nnz,
                                  self.ws("tp64", (B * ((G_tp + 63) // 64 + 1),), torch.int32),
                                  self.ws_cap("packed", n_packed, torch.int32))
            if by_inputs:
                gathered = self._dp_gather_csr(gexp, tp, B, n_packed)
        self._t1(ev)
        # index preparation needs no weights: it runs while the previous step's shadow all-gather finishes
        ev = self._t0("dp_wait_shadow_first")
        gexp.wait_shadow("first")   # bf16 shards published by the previous step's optimizer (world > 1)
        self._t1(ev)
        ev_mid = None
        for j, lp in enumerate(enc):
            x32, x16, caches[("enc", j)] = self._layer_fwd(f"enc{j}", lp, x32, x16, B,
                                                            csr=(crow, col, val, G, tp) if j == 0 else None,
                                                            masks=masks)
            if j == 0:
                ev_mid = self._t0("mid_fwd")     # everything between the two gene-sized layers
        hidden = []
        for j, lp in enumerate(self.vaeenc_plan):
            x32, x16, caches[("venc", j)] = self._layer_fwd(f"venc{j}", lp, x32, x16, B, masks=masks)
            if lp.return_hidden and lp.relu:
                hidden.append(("venc", j, x32, x16))
        q32, q16 = x32, x16
        ML = self.ws("ML", (B, 2 * Z))
        if self._tc(self.Hv, 2 * Z):
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
        x32, x16 = z32, z16
        for j, lp in enumerate(self.vaedec_plan):
            x32, x16, caches[("vdec", j)] = self._layer_fwd(f"vdec{j}", lp, x32, x16, B, masks=masks)
        for j, lp in enumerate(dec[:-1]):
            x32, x16, caches[("dec", j)] = self._layer_fwd(f"dec{j}", lp, x32, x16, B, masks=masks)
        h32, h16 = x32, x16
        out = dec[-1]
        self._t1(ev_mid)
        ev = self._t0("dp_wait_shadow_rest")
        gexp.wait_shadow("rest")
        self._t1(ev)
        comm_budget(False)      # all-gathers are done: decoder + dWout run on every SM
        fused = self._tc(out.K) and bf
        if fused:
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
                w = ap.group.exchange_rest_async()
                if w is not None:
                    w.wait()
                ap.group.grad_norm_sq(s_norm(2 + i))
                ap.group.clip_adam(s_norm(2 + i), self.clip.get("adversarial"), gscale)
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
        # single process: the two big weight-gradient kernels add their own sum of squares to the clip norm
        fuse_norm = self.world == 1 and fused and use_tc_spmm
        H1 = out.K
        dh = self.ws("dh", (B, H1))
        if fused:
            ev = self._t0("dWout_gemm")
            ops.gemm(dl, 1, h16, 1, G, H1, B, C32=out.gW,                 # dWout = dlogits^T h (+ its ||.||^2)
                     sumsq_out=s_norm(1) if fuse_norm else None)
            self._t1(ev)
            ops.colsum(dl, out.gb, M=B, N=G, accumulate=True)
            # the output layer's gradient (half of the expert group) is final: start exchanging it now so
            # the transfer overlaps the rest of the backward pass
            pending = [gexp.exchange_segment_async(gexp.n_first)]
            comm_budget(True)
            ev = self._t0("dh_gemm")
            ops.gemm(dl, 0, out.W16, 1, B, H1, G, C32=dh)                 # dh = dlogits Wout
            self._t1(ev)
        else:
            ops.gemm(dl, 1, h32, 1, G, H1, B, C32=out.gW, use_tc=False)
            ops.colsum(dl, out.gb, accumulate=True)
            pending = [gexp.exchange_segment_async(gexp.n_first)]
            ops.gemm(dl, 0, out.W32, 1, B, H1, G, C32=dh, use_tc=False)
        d = dh
        ev_mid = self._t0("mid_bwd")
        for j in reversed(range(len(dec) - 1)):
            d = self._layer_bwd(f"dec{j}", dec[j], caches[("dec", j)], d, B)
        for j in reversed(range(len(self.vaedec_plan))):
            d = self._layer_bwd(f"vdec{j}", self.vaedec_plan[j], caches[("vdec", j)], d, B)
        dz = d
        for i in range(n_adv):
            if hidden[i][0] == "z":
                ops.axpy(dz, d_hidden[i], -1.0)                            # GRL: -alpha * grad, alpha = 1
        dML = self.ws("dML", (B, 2 * Z))
        dML16 = self.ws("dML16", (B, 2 * Z), torch.bfloat16) if bf else None
        ops.reparam_kl_bwd(ML, eps, dz, Z, self.var_eps, float(kl_weight) / B, dML, dML16)
        dq = self.ws("dq", (B, self.Hv))
        if self._tc(self.Hv, 2 * Z):
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
        if use_tc_spmm:
            csc = ("tc", tp, G, s_norm(1) if fuse_norm else None, pending, gathered)
        else:
            cptr, ridx, cval = ops.csr_transpose(crow, col, val, G, nnz, self.ws("cptr", (G + 1,), torch.int32),
                                                 self.ws_cap("ridx", max(nnz, 1), torch.int32),
                                                 self.ws_cap("cval", max(nnz, 1)),
                                                 self.ws("cursor", (G + 1,), torch.int32))
            csc = ("gather", cptr, ridx, cval, G)
        for j in reversed(range(len(enc))):
            if j == 0:
                self._t1(ev_mid)
            ev = self._t0("csr_linear_bwd_w+bn") if j == 0 else None
            d = self._layer_bwd(f"enc{j}", enc[j], caches[("enc", j)], d, B, need_dx=(j > 0),
                                csc=csc if j == 0 else None)
            self._t1(ev)

        # ---------------- grad norms, clip, Adam ----------------
        self._join_side()
        pending.append(gexp.exchange_segment_async(gexp.n_first - 1))   # last piece of W1 + the small matrices
        if not (use_tc_spmm and gexp.n_first > 1):
            for i in range(gexp.n_first - 1):
                pending.append(gexp.exchange_segment_async(i))
        pending.append(gexp.exchange_rest_async())
        pending.append(gvae.exchange_rest_async())
        ev = self._t0("dp_wait_grads")
        for w in pending:
            if w is not None:
                w.wait()
        self._t1(ev)
        comm_budget(False)
        ev = self._t0("norm+clip_adam")
        gvae.grad_norm_sq(s_norm(0))
        gexp.grad_norm_sq(s_norm(1), skip=[enc[0].lin.weight, out.lin.weight] if fuse_norm else None)
        gvae.clip_adam(s_norm(0), self.clip.get("vae"), gscale)
        gexp.clip_adam(s_norm(1), self.clip.get("expert"), gscale,
                       background=self._bg if (self.pipeline_optimizer and self.world == 1) else None)
        self._t1(ev)

        self.last = dict(sc=sc, B=B, Z=Z, kl_weight=float(kl_weight), expert_id=expert_id, n_adv=n_adv,
                         gscale=gscale, ce_base=ce_base, z=z32, dl=dl)
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
        return out

    # ------------------------------------------------------------------------------- eval forward
    @torch.no_grad()
    def eval_step(self, expert_id: str, crow, col, val, eps=None, kl_weight: float = 1.0):
        """validation_step arithmetic (cmmvae_model.py:219-245): eval-mode forward + ELBO on the fused
        decoder path.  Returns the scalar record (use ``scalars``-like host read via ``eval_scalars``)."""
        enc, dec = self.enc_plan[expert_id], self.dec_plan[expert_id]
        B, G, Z = crow.numel() - 1, enc[0].K, self.Z
        bf = self.precision == "bf16"
        if bf:
            self.groups[f"experts/{expert_id}"].wait_shadow(None)
        else:
            self.groups[f"experts/{expert_id}"].sync_master()
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
        x32, x16 = z32, z16
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
                "z": z32}
