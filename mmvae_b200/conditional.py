"""Conditional layers on the latent inside the fused step.

The reference's ``ConditionalLayers`` (modules/base/components.py:467-631) holds, per batch key, one FCBlock per
distinct metadata value (``ConditionalLayer``, components.py:317-413; shared, or one set per species) plus an optional
per-species ``species`` block; ``forward`` routes every cell through the block of ITS value with a Python dict of row
lists and one ``index_select`` / module call / ``index_copy_`` per value, batch key after batch key or -- ``selection
_order == ["parallel"]`` -- side by side with the outputs concatenated in an order drawn per call (components.py:598-631).

``CondBank`` is the B200 side of that module: every block's ``Linear`` lives in one flat parameter bank (slot =
``[W | b]``), a batch is turned ON THE HOST into tiles of <= 32 rows that share a slot (``plan``: one vectorised sort
per batch key, one pinned block, one H2D copy) and ``csrc/conditional.cu`` runs all values of all batch keys in one
launch per direction.  Adam is applied to the slots present in the batch only, each with its own step count, which is
what ``torch.optim.Adam`` does for the reference (parameters whose ``grad is None`` are skipped).
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional

import numpy as np
import pandas as pd
import torch
import torch.nn as nn

from . import ops

ROWS = 32      # rows per tile (kCondRows in csrc/conditional.cu)


class CondUnsupported(NotImplementedError):
    pass


class CondBank:
    def __init__(self, cond: nn.Module, device, lr: float = 5e-3, weight_decay: float = 1e-6, betas=(0.9, 0.999),
                 eps: float = 1e-8):
        self.module = cond
        self.device = torch.device(device)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, tuple(betas), eps
        self.names: List[str] = list(cond.selection_order)
        self.parallel = bool(cond.is_parallel)
        self.shuffle = bool(cond.shuffle_selection_order)
        # ---- slots, in the module's parameter order ----
        blocks = [(path, m) for path, m in cond.named_modules() if hasattr(m, "fc_layers")]
        if not blocks:
            raise CondUnsupported("conditional layers without blocks")
        self.slot_of_path = {path: i for i, (path, _) in enumerate(blocks)}
        first = blocks[0][1]
        for path, b in blocks:
            if len(b.fc_layers) != 1:
                raise CondUnsupported("conditional blocks with more than one layer are outside the fused step")
            parts = dict(b.fc_layers[0].named_children())
            if set(parts) - {"lin", "ln", "af"} or ("af" in parts and type(parts["af"]) is not nn.ReLU):
                raise CondUnsupported("conditional blocks are Linear [+ LayerNorm] [+ ReLU] in the fused step")
            if "ln" in parts and (parts["ln"].elementwise_affine or abs(parts["ln"].eps - 1e-5) > 1e-12):
                raise CondUnsupported("conditional LayerNorm must be the reference's (no affine, eps 1e-5)")
        p0 = dict(first.fc_layers[0].named_children())
        self.layer_norm, self.relu = "ln" in p0, "af" in p0
        self.Zin, self.Zout = p0["lin"].in_features, p0["lin"].out_features
        if self.Zin % 4 or self.Zout % 4:
            raise CondUnsupported("conditional block widths must be multiples of 4 in the fused step")
        if not self.parallel and self.Zin != self.Zout:
            raise CondUnsupported("chained conditional layers need square blocks")
        if (self.Zin + self.Zout) * ROWS * 4 > 200 * 1024:
            raise CondUnsupported("conditional blocks wider than the kernel's shared-memory tile")
        self.n_slots = len(blocks)
        self.S = (self.Zout * self.Zin + self.Zout + 3) // 4 * 4
        n = self.n_slots * self.S
        self.p = torch.zeros(n, device=self.device)
        self.g = torch.zeros(n, device=self.device)
        self.m = torch.zeros(n, device=self.device)
        self.v = torch.zeros(n, device=self.device)
        self.steps = torch.zeros(self.n_slots, dtype=torch.int32, device=self.device)
        self.params: List[nn.Parameter] = []
        nw = self.Zout * self.Zin
        for s, (path, b) in enumerate(blocks):
            lin = b.fc_layers[0].lin
            if (lin.in_features, lin.out_features) != (self.Zin, self.Zout):
                raise CondUnsupported("conditional blocks of different sizes")
            o = s * self.S
            for prm, lo, shape in ((lin.weight, o, (self.Zout, self.Zin)), (lin.bias, o + nw, (self.Zout,))):
                view = self.p[lo:lo + prm.numel()].view(shape)
                view.copy_(prm.data.to(self.device))
                prm.data = view
                prm.grad = self.g[lo:lo + prm.numel()].view(shape)
                self.params.append(prm)
        # ---- value -> slot tables ----
        self.tables: Dict[tuple, tuple] = {}      # (batch key, species | None) -> (keys, slots int32)
        self.block_slot: Dict[tuple, int] = {}    # (batch key, species) -> slot of a per-species plain FCBlock
        self.kind: Dict[str, str] = {}            # batch key -> "shared" | "per_species" | "block"
        for bk, layer in cond.layers.items():
            if hasattr(layer, "conditions"):
                self.kind[bk] = "shared"
                self.tables[(bk, None)] = self._table(f"layers.{bk}", layer)
                continue
            self.kind[bk] = "block"       # (an empty ModuleDict fails at lookup time, as the reference's does)
            for sp, sub in layer.items():
                if hasattr(sub, "conditions"):
                    self.kind[bk] = "per_species"
                    self.tables[(bk, sp)] = self._table(f"layers.{bk}.{sp}", sub)
                else:
                    self.kind[bk] = "block"
                    self.block_slot[(bk, sp)] = self.slot_of_path[f"layers.{bk}.{sp}"]
        self._pinned = [None] * 4
        self._pin_ev = [None] * 4
        self._used_ev = [None] * 4      # recorded behind the last kernel that reads a plan's device arrays
        self._pin_i = 0
        self.plan: Optional[dict] = None

    def _table(self, path: str, layer) -> tuple:
        keys = list(layer.conditions.keys())
        slots = np.array([self.slot_of_path[f"{path}.conditions.{k}"] for k in keys], dtype=np.int32)
        return {k: i for i, k in enumerate(keys)}, slots

    @staticmethod
    def _lookup(tab, values) -> np.ndarray:
        """slot of every metadata value.  The reference formats each value with ``str(v).replace(".", "_")``
        (components.py:355-365, 391-396); values that already are keys are found directly, only the rest is
        formatted"""
        d = tab[0]
        try:          # (a C-level map over a dict: 4x faster than pd.Index.get_indexer on object keys)
            idx = np.fromiter(map(d.__getitem__, values), dtype=np.int64, count=len(values))
        except KeyError:
            idx = np.fromiter((d[str(v).replace(".", "_")] for v in values), dtype=np.int64, count=len(values))
        return tab[1][idx]

    # --------------------------------------------------------------------------------------------- host plan
    def host_plan(self, metadata: pd.DataFrame, species: Optional[str], B: int):
        """(pure host work) rows of the batch grouped by condition value: tiles, row lists, tile range per key"""
        n_c = len(self.names)
        slots = np.empty((n_c, B), dtype=np.int64)
        for c, bk in enumerate(self.names):
            kind = self.kind[bk]
            if kind != "shared" and species is None:
                raise RuntimeError(f"'species' must be set to access non-shared conditional layer for batch_key '{bk}'")
            if kind == "block":
                if (bk, species) not in self.block_slot:
                    raise KeyError(species)
                slots[c] = self.block_slot[(bk, species)]
                continue
            tab = self.tables[(bk, None if kind == "shared" else species)]
            col = metadata[bk]
            if isinstance(col.dtype, pd.CategoricalDtype):      # look the categories up, not the rows
                codes = col.cat.codes.to_numpy()
                if (codes < 0).any():
                    raise KeyError("nan")
                slots[c] = self._lookup(tab, col.cat.categories.to_numpy())[codes]
            else:
                slots[c] = self._lookup(tab, col.to_numpy())
        # slots are unique across batch keys: ONE stable sort of all (key, row) pairs groups everything
        flat = slots.reshape(-1)
        perm = np.argsort(flat, kind="stable")
        ss = flat[perm]
        start = np.flatnonzero(np.concatenate(([True], ss[1:] != ss[:-1])))
        cnt = np.diff(np.concatenate((start, [flat.size])))
        nch = (cnt + ROWS - 1) // ROWS
        n_t = int(nch.sum())
        rep = np.repeat(np.arange(start.size), nch)                  # group of every tile
        within = np.arange(n_t) - np.repeat(np.cumsum(nch) - nch, nch)
        cond_of = (perm[start] // B).astype(np.int32)                # batch key index of every group
        tiles = np.empty((n_t, 4), dtype=np.int32)
        tiles[:, 0] = ss[start][rep]
        tiles[:, 1] = start[rep] + ROWS * within
        tiles[:, 2] = np.minimum(ROWS, cnt[rep] - ROWS * within)
        tiles[:, 3] = cond_of[rep] | ((nch == 1)[rep].astype(np.int32) << 16)     # bit 16: the slot's only tile
        rows = (perm % B).astype(np.int32)
        present = ss[start].astype(np.int32)
        multi = present[nch > 1]            # slots that several tiles ADD into (single tiles overwrite their slot)
        # tile range of every batch key (chained selection launches them one after the other)
        tc = cond_of[rep]
        edge = np.flatnonzero(np.concatenate(([True], tc[1:] != tc[:-1])))
        ranges = [None] * n_c
        for lo, hi in zip(edge, np.concatenate((edge[1:], [n_t]))):
            ranges[int(tc[lo])] = (int(lo), int(hi - lo))
        return tiles, rows, present, ranges, multi

    def make_plan(self, metadata: pd.DataFrame, species: Optional[str], B: int) -> dict:
        """sort the batch by condition value (host, vectorised) and ship tiles / row lists / present slots"""
        order = random.sample(self.names, len(self.names)) if self.shuffle else list(self.names)   # components.py:598-602
        n_c = len(self.names)
        tiles, rows, present, ranges, multi = self.host_plan(metadata, species, B)
        pos = {bk: i for i, bk in enumerate(order)}
        if self.parallel:
            out_col = np.array([pos[bk] * self.Zout for bk in self.names], dtype=np.int32)
            dx_col = np.arange(n_c, dtype=np.int32) * self.Zin
        else:
            out_col = np.zeros(n_c, dtype=np.int32)
            dx_col = np.zeros(n_c, dtype=np.int32)
        # one pinned block, one copy: [tiles | rows | present | out_col | dx_col | multi]
        parts = [tiles.reshape(-1), rows, present, out_col, dx_col, multi]
        sizes = [(p.size + 3) // 4 * 4 for p in parts]
        total = sum(sizes)
        i = self._pin_i
        self._pin_i = (i + 1) % len(self._pinned)
        for ev in (self._pin_ev[i], self._used_ev[i]):      # the block's last copy has left, its last reader is done
            if ev is not None:
                ev.synchronize()
        if self._pinned[i] is None or self._pinned[i][0].numel() < total:
            cap = max(total * 2, 4096)
            self._pinned[i] = (torch.empty(cap, dtype=torch.int32, pin_memory=True),
                               torch.empty(cap, dtype=torch.int32, device=self.device))
        host, dev = self._pinned[i]
        hv = host.numpy()
        offs, o = [], 0
        for p, n in zip(parts, sizes):
            hv[o:o + p.size] = p
            offs.append(o)
            o += n
        dev[:total].copy_(host[:total], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pin_ev[i] = ev
        view = lambda k: dev[offs[k]:offs[k] + parts[k].size]   # noqa: E731
        self.plan = dict(order=order, n_tiles=len(tiles), tiles=view(0), rows=view(1), present=view(2),
                         n_present=int(present.size), out_col=view(3), dx_col=view(4), multi=view(5),
                         n_multi=int(multi.size), ranges=ranges, B=B, ring_slot=i,
                         index={bk: c for c, bk in enumerate(self.names)})
        return self.plan

    # ------------------------------------------------------------------------------------------- device work
    def forward(self, z32: torch.Tensor, ws, want_bf16: bool):
        """CLVAE.after_reparameterize (clvae.py:89-111).  ``ws(name, shape, dtype)`` hands out workspaces.
        Returns (out fp32, out bf16 | None)."""
        pl, B = self.plan, self.plan["B"]
        n_c = len(self.names)
        rstd = ws("cond.rstd", (n_c, B))
        if self.parallel:
            W = n_c * self.Zout
            out, pre = ws("cond.out", (B, W)), ws("cond.pre", (B, W))
            out16 = ws("cond.out16", (B, W), torch.bfloat16) if want_bf16 else None
            ops.cond_fwd(self.p, self.S, self.Zin, self.Zout, pl["tiles"], pl["n_tiles"], pl["rows"], z32,
                         z32.shape[1], out, out16, pre, W, pl["out_col"], rstd, B, self.layer_norm, self.relu)
            pl["x"] = z32
            self._mark_used()
            return out, out16
        cur = z32
        pl["stage_in"] = {}
        last = pl["order"][-1]
        out16 = None
        for bk in pl["order"]:
            c = pl["index"][bk]
            lo, n = pl["ranges"][c]
            out, pre = ws(f"cond.out{c}", (B, self.Zout)), ws(f"cond.pre{c}", (B, self.Zout))
            if bk == last and want_bf16:
                out16 = ws("cond.out16", (B, self.Zout), torch.bfloat16)
            ops.cond_fwd(self.p, self.S, self.Zin, self.Zout, pl["tiles"][4 * lo:], n, pl["rows"], cur, cur.shape[1],
                         out, out16 if bk == last else None, pre, self.Zout, pl["out_col"], rstd, B,
                         self.layer_norm, self.relu)
            pl["stage_in"][bk] = cur
            cur = out
        self._mark_used()
        return cur, out16

    def backward(self, dout: torch.Tensor, ws) -> torch.Tensor:
        """gradients of the present slots (accumulated into the zeroed bank) and dz"""
        pl, B = self.plan, self.plan["B"]
        n_c = len(self.names)
        rstd = ws("cond.rstd", (n_c, B))
        ops.cond_zero_grads(self.g, self.S, pl["multi"], pl["n_multi"])
        if self.parallel:
            W = n_c * self.Zout
            dxc = ws("cond.dx", (B, n_c * self.Zin))
            ops.cond_bwd(self.p, self.g, self.S, self.Zin, self.Zout, pl["tiles"], pl["n_tiles"], pl["rows"], pl["x"],
                         pl["x"].shape[1], dout, ws("cond.pre", (B, W)), W, pl["out_col"], rstd, B, dxc,
                         n_c * self.Zin, pl["dx_col"], self.layer_norm, self.relu)
            return ops.fold_cols(dxc, n_c, ws("cond.dz", (B, self.Zin)))
        d = dout
        for bk in reversed(pl["order"]):
            c = pl["index"][bk]
            lo, n = pl["ranges"][c]
            x = pl["stage_in"][bk]
            dx = ws(f"cond.dx{c}", (B, self.Zin))
            ops.cond_bwd(self.p, self.g, self.S, self.Zin, self.Zout, pl["tiles"][4 * lo:], n, pl["rows"], x,
                         x.shape[1], d, ws(f"cond.pre{c}", (B, self.Zout)), self.Zout, pl["out_col"], rstd, B, dx,
                         self.Zin, pl["dx_col"], self.layer_norm, self.relu)
            d = dx
        return d

    def _mark_used(self):
        ev = torch.cuda.Event()
        ev.record()
        self._used_ev[self.plan["ring_slot"]] = ev

    def add_norm_sq(self, out: torch.Tensor):
        pl = self.plan
        ops.cond_sumsq(self.g, self.S, pl["present"], pl["n_present"], out)

    def clip_adam(self, norm_sq: torch.Tensor, max_norm: Optional[float], grad_scale: float = 1.0):
        pl = self.plan
        ops.cond_adam(self.p, self.g, self.m, self.v, self.S, pl["present"], pl["n_present"], self.steps, norm_sq,
                      max_norm, grad_scale, self.lr, self.betas[0], self.betas[1], self.eps, self.wd)
        self._mark_used()

    # ------------------------------------------------------------------------------- torch.optim.Adam state
    def state_entries(self, first_index: int) -> dict:
        """per-parameter Adam state (torch format) of the slots that have been stepped"""
        steps = self.steps.cpu().tolist()
        out, nw = {}, self.Zout * self.Zin
        for s, t in enumerate(steps):
            if t <= 0:
                continue
            o = s * self.S
            for j, (lo, n, shape) in enumerate(((o, nw, (self.Zout, self.Zin)), (o + nw, self.Zout, (self.Zout,)))):
                out[first_index + 2 * s + j] = {"step": torch.tensor(float(t)),
                                                "exp_avg": self.m[lo:lo + n].view(shape).clone(),
                                                "exp_avg_sq": self.v[lo:lo + n].view(shape).clone()}
        return out

    def load_state_entries(self, state: dict, first_index: int):
        steps = np.zeros(self.n_slots, dtype=np.int32)
        nw = self.Zout * self.Zin
        self.m.zero_()
        self.v.zero_()
        for s in range(self.n_slots):
            o = s * self.S
            for j, (lo, n, shape) in enumerate(((o, nw, (self.Zout, self.Zin)), (o + nw, self.Zout, (self.Zout,)))):
                st = state.get(first_index + 2 * s + j, state.get(str(first_index + 2 * s + j)))
                if st is None:
                    continue
                self.m[lo:lo + n].view(shape).copy_(st["exp_avg"])
                self.v[lo:lo + n].view(shape).copy_(st["exp_avg_sq"])
                steps[s] = int(float(st["step"]))
        self.steps.copy_(torch.from_numpy(steps))
