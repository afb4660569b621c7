"""Host -> HBM feed for CSR minibatches.

The reference hands ``training_step`` a ``torch.sparse_csr_tensor`` built on the host by
``SparseCSRMatrixBatcherDataPipe`` (data/local/cellxgene_datapipe.py:169-193: scipy row slice ->
``torch.sparse_csr_tensor(crow, col, val)``) and lets Lightning move it to the GPU with three separate
pageable copies (sparse tensors cannot be pinned, cellxgene_datamodule.py:95-103).  ``CSRStager`` is the
B200 side of that hand-off: the three arrays of a batch are packed back to back into ONE pinned block
(``crow | col | val``, each 16-byte aligned), shipped with ONE async copy on a dedicated copy stream into a
rotating set of HBM blocks, and handed to the step as views of that block -- indices and values reach the
device bit-identical, int32 / fp32 exactly as scipy emitted them.

    stager = CSRStager(max_cells=B, max_nnz=nnz_cap, device="cuda", depth=3)
    blk = stager.reserve(n_cells, nnz)               # numpy views into the next pinned block
    slice_rows(indptr, indices, data, lo, hi, out=blk)   # the batcher writes the rows in place
    ticket = stager.commit(blk, n_genes)             # async H2D, returns immediately
    x = stager.get(ticket)                           # torch.sparse_csr tensor on the device; the current
                                                     # stream waits for the copy, the host does not
    model.training_step((x, metadata, species), i);  stager.release(ticket)
(``put(crow, col, val, n_genes)`` = reserve + copy in + commit, for producers that already hold arrays.)

A slot is reused ``depth`` puts later; ``put`` first waits (host side) for the event recorded by the
consumer of that slot's previous occupant, so an in-flight step never sees its input overwritten.

``narrow_col=True`` (gene panels of <= 65 536 genes): the pinned block holds the gene ids as uint16 -- 6 instead
of 8 bytes per non-zero over PCIe -- and a widening kernel on the copy stream restores the int32 ``col`` array
in HBM right behind the copy (exact: ids are < 65 536), so consumers still see int32 / fp32.

``StagedCSRBatches(..., pin_chunks=True)``: chunks that serve many batches (or many epochs) are page-locked IN
PLACE once (``cmmvae_host_register``); a batch is then the rebased ``crow`` plus two DMA transfers straight out of
the chunk's ``indices`` / ``data`` arrays -- no host pass over the non-zeros at all.

``StagedCSRBatches(..., workers=k)`` packs batches on ``k`` background threads with the native packer
(``cmmvae_host_slice_rows``: no GIL, no intermediate arrays), several batches ahead of the consumer.
"""
from __future__ import annotations

import collections
import concurrent.futures
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


def block_layout(n_cells: int, nnz: int, col_bytes: int = 4):
    """byte offsets of (crow, col, val) inside one staged block and the block's size"""
    o_crow = 0
    o_col = _align(o_crow + 4 * (n_cells + 1))
    o_val = _align(o_col + col_bytes * nnz)
    return o_crow, o_col, o_val, _align(o_val + 4 * nnz)


def slice_rows(indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, lo: int, hi: int, out=None):
    """rows [lo, hi) of a CSR chunk as (crow int32 rebased to 0, col int32, val fp32) -- what scipy's
    ``chunk[lo:hi]`` yields (cellxgene_datapipe.py:173-183), without building a scipy object.  With
    ``out`` (a ``Block`` from ``CSRStager.reserve(hi - lo, indptr[hi] - indptr[lo])``) the rows are written
    straight into pinned memory."""
    a, b = int(indptr[lo]), int(indptr[hi])
    if out is not None:
        np.subtract(indptr[lo:hi + 1], indptr[lo], out=out.crow, casting="unsafe")
        np.copyto(out.col, indices[a:b], casting="unsafe")     # int32, or uint16 in a narrow block
        out.val[:] = data[a:b]
        return out.crow, out.col, out.val
    crow = (indptr[lo:hi + 1] - indptr[lo]).astype(np.int32, copy=False)
    return crow, indices[a:b].astype(np.int32, copy=False), data[a:b].astype(np.float32, copy=False)


@dataclass
class Ticket:
    slot: int
    n_cells: int
    n_genes: int
    nnz: int
    ready: Optional[torch.cuda.Event]


@dataclass
class Block:
    slot: int
    n_cells: int
    nnz: int
    crow: np.ndarray
    col: np.ndarray
    val: np.ndarray


class CSRStager:
    def __init__(self, max_cells: int, max_nnz: int, device="cuda", depth: int = 3, narrow_col: bool = False):
        if depth < 2:
            raise ValueError("CSRStager needs depth >= 2 (one block in flight, one being filled)")
        self.device = torch.device(device)
        self.max_cells, self.max_nnz, self.depth = int(max_cells), int(max_nnz), int(depth)
        self.narrow = bool(narrow_col)
        self.col_bytes = 2 if self.narrow else 4
        self.nbytes = block_layout(self.max_cells, self.max_nnz, self.col_bytes)[3]
        self.on_gpu = self.device.type == "cuda"
        if self.on_gpu and not torch.cuda.is_available():
            raise RuntimeError("CSRStager(device='cuda') needs a CUDA device")
        if self.on_gpu and self.device.index is None:      # (worker threads select the device by index)
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.host = [torch.empty(self.nbytes, dtype=torch.uint8, pin_memory=self.on_gpu) for _ in range(depth)]
        self.dev = [torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        # narrow blocks: the int32 col array the kernels read, widened on the device behind each copy
        self.col32 = [torch.empty(self.max_nnz + 8, dtype=torch.int32, device=self.device) for _ in range(depth)] \
            if self.narrow else None
        self.stream = torch.cuda.Stream(self.device) if self.on_gpu else None
        self._consumed = [None] * depth     # event: the step that read this slot has been enqueued and finished
        self._copied = [None] * depth       # event: the H2D copy out of the pinned block has finished
        self._n = 0
        self.bytes_staged = 0

    # ------------------------------------------------------------------------------------------ put
    def reserve(self, n_cells: int, nnz: int) -> "Block":
        """next pinned block, as numpy views (crow int32 [n_cells+1], col int32 [nnz], val fp32 [nnz]) the
        batcher fills IN PLACE (``slice_rows(..., out=block)``) -- no intermediate host copy"""
        if n_cells < 0 or n_cells > self.max_cells or nnz < 0 or nnz > self.max_nnz:
            raise ValueError(f"batch of {n_cells} cells / {nnz} nnz exceeds the stager's "
                             f"{self.max_cells} cells / {self.max_nnz} nnz")
        slot = self._n % self.depth
        self._n += 1
        if self._copied[slot] is not None:
            self._copied[slot].synchronize()      # the pinned block is free once its copy has left
        # sections sit at offsets fixed by the ring's CAPACITY, so a slot hands out the same device addresses for
        # every batch (the fused step replays CUDA graphs keyed by them)
        o_crow, o_col, o_val, _ = block_layout(self.max_cells, self.max_nnz, self.col_bytes)
        h = self.host[slot].numpy()
        return Block(slot, n_cells, nnz,
                     h[o_crow:o_crow + 4 * (n_cells + 1)].view(np.int32),
                     h[o_col:o_col + self.col_bytes * nnz].view(np.uint16 if self.narrow else np.int32),
                     h[o_val:o_val + 4 * nnz].view(np.float32))

    def commit(self, b: "Block", n_genes: int) -> Ticket:
        """ship a filled block: one async H2D copy on the copy stream.  A block may be committed again
        (unchanged) as long as it has not been handed out by a later ``reserve``"""
        if b.n_cells >= 0 and (int(b.crow[0]) != 0 or int(b.crow[-1]) != b.nnz):
            raise ValueError("inconsistent CSR arrays (crow[0] must be 0, crow[-1] == len(col) == len(val))")
        o_crow, o_col, o_val, _ = block_layout(self.max_cells, self.max_nnz, self.col_bytes)
        if self.narrow and int(n_genes) > 65536:
            raise ValueError(f"narrow_col stages 16-bit gene ids; the panel has {n_genes} genes")
        # two copies per batch: [crow | col] (contiguous prefix) and val; only the bytes the batch uses travel
        n1 = _align(o_col + self.col_bytes * b.nnz)
        n2 = _align(4 * b.nnz)
        self.bytes_staged = n1 + n2
        slot = b.slot
        if not self.on_gpu:
            self.dev[slot][:n1].copy_(self.host[slot][:n1])
            self.dev[slot][o_val:o_val + n2].copy_(self.host[slot][o_val:o_val + n2])
            if self.narrow:
                self.col32[slot][:b.nnz].copy_(self.dev[slot][o_col:o_col + 2 * b.nnz].view(torch.uint16).to(torch.int32))
            return Ticket(slot, b.n_cells, int(n_genes), b.nnz, None)
        with torch.cuda.stream(self.stream):
            if self._consumed[slot] is not None:
                self.stream.wait_event(self._consumed[slot])    # do not overwrite a block a step still reads
            self.dev[slot][:n1].copy_(self.host[slot][:n1], non_blocking=True)
            self.dev[slot][o_val:o_val + n2].copy_(self.host[slot][o_val:o_val + n2], non_blocking=True)
            if self.narrow:
                from . import ops
                ops.widen_u16_i32(self.dev[slot][o_col:], self.col32[slot], b.nnz, stream=self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        return Ticket(slot, b.n_cells, int(n_genes), b.nnz, ev)

    def commit_from_chunk(self, b: "Block", indices: np.ndarray, data: np.ndarray, a: int, n_genes: int) -> Ticket:
        """ship a batch whose ``crow`` sits in the reserved block and whose ids / values are the page-locked chunk's
        ``indices[a:a+nnz]`` (int32) / ``data[a:a+nnz]`` (fp32): three DMA transfers, no host copy.  Wide stager only"""
        if self.narrow:
            raise ValueError("commit_from_chunk ships int32 gene ids: use a stager with narrow_col=False")
        if indices.dtype != np.int32 or data.dtype != np.float32:
            raise ValueError("commit_from_chunk needs int32 indices and float32 data")
        if int(b.crow[0]) != 0 or int(b.crow[-1]) != b.nnz:
            raise ValueError("inconsistent CSR arrays (crow[0] must be 0, crow[-1] == len(col) == len(val))")
        o_crow, o_col, o_val, _ = block_layout(self.max_cells, self.max_nnz, self.col_bytes)
        n0 = _align(4 * (b.n_cells + 1))
        self.bytes_staged = n0 + 8 * b.nnz
        slot, d = b.slot, self.dev[b.slot]
        if not self.on_gpu:
            d[:n0].copy_(self.host[slot][:n0])
            d[o_col:o_col + 4 * b.nnz].copy_(torch.from_numpy(indices[a:a + b.nnz].view(np.uint8)))
            d[o_val:o_val + 4 * b.nnz].copy_(torch.from_numpy(data[a:a + b.nnz].view(np.uint8)))
            return Ticket(slot, b.n_cells, int(n_genes), b.nnz, None)
        from . import ops
        with torch.cuda.stream(self.stream):
            if self._consumed[slot] is not None:
                self.stream.wait_event(self._consumed[slot])
            d[:n0].copy_(self.host[slot][:n0], non_blocking=True)
            if b.nnz:
                ops.h2d_async(d[o_col:], indices[a:a + b.nnz], 4 * b.nnz, stream=self.stream)
                ops.h2d_async(d[o_val:], data[a:a + b.nnz], 4 * b.nnz, stream=self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        return Ticket(slot, b.n_cells, int(n_genes), b.nnz, ev)

    def put(self, crow, col, val, n_genes: int) -> Ticket:
        """reserve + copy the three arrays in + commit (for producers that already hold numpy arrays)"""
        crow, col, val = (np.asarray(a) for a in (crow, col, val))
        n_cells, nnz = int(crow.shape[0]) - 1, int(col.shape[0])
        if val.shape[0] != nnz or (n_cells >= 0 and (int(crow[-1]) != nnz or int(crow[0]) != 0)):
            raise ValueError("inconsistent CSR arrays (crow[0] must be 0, crow[-1] == len(col) == len(val))")
        if nnz and int(col.max()) >= 2 ** 31:
            raise ValueError("CSR indices do not fit int32")
        if self.narrow and nnz and int(col.max()) >= 65536:
            raise ValueError("narrow_col stages 16-bit gene ids")
        b = self.reserve(n_cells, nnz)
        b.crow[:], b.val[:] = crow, val      # casts int64 -> int32 / f64 -> f32 on the way in
        np.copyto(b.col, col, casting="unsafe")
        return self.commit(b, n_genes)

    # ------------------------------------------------------------------------------------------ get
    def arrays(self, t: Ticket):
        """device views (crow int32 [n_cells+1], col int32 [nnz], val fp32 [nnz]) of a staged batch; the
        current stream waits for the copy"""
        if t.ready is not None:
            torch.cuda.current_stream(self.device).wait_event(t.ready)
        o_crow, o_col, o_val, _ = block_layout(self.max_cells, self.max_nnz, self.col_bytes)
        d = self.dev[t.slot]
        crow = d[o_crow:o_crow + 4 * (t.n_cells + 1)].view(torch.int32)
        col = self.col32[t.slot][:t.nnz] if self.narrow else d[o_col:o_col + 4 * t.nnz].view(torch.int32)
        val = d[o_val:o_val + 4 * t.nnz].view(torch.float32)
        return crow, col, val

    def get(self, t: Ticket) -> torch.Tensor:
        """the staged batch as the ``torch.sparse_csr_tensor`` the reference's ``training_step`` receives"""
        crow, col, val = self.arrays(t)
        x = torch.sparse_csr_tensor(crow, col, val, size=(t.n_cells, t.n_genes))
        x._cmmvae_ready = t.ready if t.ready is not None else True     # for CMMVAEModel.prefetch_batch (data parallel)
        x._cmmvae_cap = self.max_nnz       # elements behind col / val that may be addressed (graph replay)
        return x

    def release(self, t: Ticket):
        """call after the step that consumes ``t`` has been enqueued: its slot may be refilled once that
        step has finished on the device"""
        if self.on_gpu:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._consumed[t.slot] = ev


class StagedCSRBatches:
    """Batches of a chunked CSR dataset, staged ahead of the consumer.

    Iterates like the reference's ``SparseCSRMatrixBatcherDataPipe`` (data/local/cellxgene_datapipe.py:125-193):
    ``source`` yields ``(scipy.sparse.csr_matrix chunk, pandas.DataFrame)`` pairs (what ``LoadCSRMatrixAndDataFrame``
    / ``ShuffleCSRMatrixAndDataFrame`` produce from the ``.npz`` / ``.pkl`` chunk files); every chunk is cut into
    consecutive ``batch_size``-row batches, a trailing short batch is skipped unless ``allow_partials``, metadata
    is ``frame.iloc[i:i+batch_size].reset_index(drop=True)``, and each item is ``(torch.sparse_csr_tensor, metadata)``.

    Difference: rows are never materialised as a scipy slice -- they are written straight from the chunk's
    ``indptr / indices / data`` into a pinned ``CSRStager`` block, shipped with one asynchronous copy while
    earlier batches are being consumed, and the yielded tensor lives on ``device``.  The block of a batch is
    recycled once the consumer asks for the next item, i.e. after it has enqueued its work on the current stream.

    ``workers`` > 0: packing (a 25 MB gather out of pageable memory per 1024-cell batch -- longer than the
    training step itself on one core) runs on that many background threads through the native packer,
    ``ahead`` batches in front of the consumer; batches are still yielded in order.  ``workers`` = 0 packs on
    the consumer's thread with numpy, one batch ahead.  ``narrow_col``: ship gene ids as uint16 (None = whenever
    the chunk has <= 65 536 genes)."""

    def __init__(self, source, batch_size: int, allow_partials: bool = False, device="cuda", depth: int = 3,
                 workers: int = 0, ahead: Optional[int] = None, narrow_col: Optional[bool] = None,
                 pin_chunks: bool = False):
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        self.workers = int(workers)
        self.ahead = int(ahead) if ahead is not None else max(1, 2 * self.workers)
        if depth < 3:
            raise ValueError("StagedCSRBatches needs depth >= 3 (consumed, staged ahead, being filled)")
        depth = max(int(depth), self.ahead + 2)
        self.source, self.batch_size, self.allow_partials = source, int(batch_size), bool(allow_partials)
        self.device, self.depth, self.narrow_col = device, depth, narrow_col
        self.stager: Optional[CSRStager] = None
        self._pool = None
        self.pin_chunks = bool(pin_chunks)
        self._pinned = collections.OrderedDict()     # (addresses) -> (indices, data): page-locked chunks, LRU of 2
        self.chunks_pinned = 0

    def _pin(self, chunk, n_genes: int) -> bool:
        """page-lock the chunk's arrays on first sight (validating the gene ids once); False = pack it instead"""
        ind, dat = chunk.indices, chunk.data
        if not (self.pin_chunks and ind.dtype == np.int32 and dat.dtype == np.float32 and ind.size
                and ind.flags.c_contiguous and dat.flags.c_contiguous and torch.device(self.device).type == "cuda"):
            return False
        key = (ind.ctypes.data, dat.ctypes.data, ind.size)
        if key in self._pinned:
            self._pinned.move_to_end(key)
            return True
        if int(ind.min()) < 0 or int(ind.max()) >= n_genes:
            raise ValueError(f"gene id outside [0, {n_genes})")
        from . import ops
        while len(self._pinned) >= 2:
            self._unpin(next(iter(self._pinned)))
        ops.host_register(ind)
        try:
            ops.host_register(dat)
        except RuntimeError:
            ops.host_unregister(ind)
            raise
        self._pinned[key] = (ind, dat)      # (keeps the arrays alive while they are registered)
        self.chunks_pinned += 1
        return True

    def _unpin(self, key):
        from . import ops
        ind, dat = self._pinned.pop(key)
        torch.cuda.synchronize(self.device)      # transfers out of the chunk may still be in flight
        ops.host_unregister(ind)
        ops.host_unregister(dat)

    def _ensure_capacity(self, nnz: int, n_genes: int, pinned: bool = False):
        narrow = False if pinned else (self.narrow_col if self.narrow_col is not None else n_genes <= 65536)
        if self.stager is None or nnz > self.stager.max_nnz or narrow != self.stager.narrow:
            # a denser batch than any seen so far: new ring with head room (old blocks stay alive with the
            # tickets that reference them until their consumers are done)
            self.stager = CSRStager(self.batch_size, int(nnz * 1.25) + 1024, device=self.device, depth=self.depth,
                                    narrow_col=narrow)

    def _tasks(self):
        for chunk, frame in self.source:
            n_rows, n_genes = chunk.shape
            for lo in range(0, n_rows, self.batch_size):
                hi = min(lo + self.batch_size, n_rows)
                if hi - lo != self.batch_size and not self.allow_partials:
                    continue
                yield chunk, frame, lo, hi, n_genes

    @staticmethod
    def _pack(stager: "CSRStager", blk: "Block", chunk, lo: int, hi: int, n_genes: int, native: bool) -> Ticket:
        indptr, indices, data = chunk.indptr, chunk.indices, chunk.data
        if native and data.dtype == np.float32 and indices.dtype in (np.int32, np.int64) \
                and indptr.dtype in (np.int32, np.int64):
            from . import ops
            if stager.on_gpu:
                torch.cuda.set_device(stager.device)     # worker threads start on device 0
            ops.host_slice_rows(indptr, indices, data, lo, hi, n_genes, blk.crow, blk.col, blk.val)
        else:
            slice_rows(indptr, indices, data, lo, hi, out=blk)
        return stager.commit(blk, n_genes)

    def __iter__(self):
        if self.workers > 0 and self._pool is None:
            self._pool = concurrent.futures.ThreadPoolExecutor(self.workers, thread_name_prefix="csr-pack")
        inflight = collections.deque()    # (stager, ticket or future, metadata), in batch order
        done = None                       # (stager, ticket) handed out on the previous iteration

        def hand_out():
            nonlocal done
            st, tk, md = inflight.popleft()
            if isinstance(tk, concurrent.futures.Future):
                tk = tk.result()
            if done is not None:
                done[0].release(done[1])
            done = (st, tk)
            return st.get(tk), md

        for chunk, frame, lo, hi, n_genes in self._tasks():
            a = int(chunk.indptr[lo])
            nnz = int(chunk.indptr[hi]) - a
            pinned = self._pin(chunk, n_genes)
            self._ensure_capacity(nnz, n_genes, pinned)
            if len(inflight) > self.ahead - 1:     # keep at most ``ahead`` batches staged in front of the consumer
                yield hand_out()
            st = self.stager
            blk = st.reserve(hi - lo, nnz)         # slots are handed out in batch order, on this thread
            meta = frame.iloc[lo:hi].reset_index(drop=True) if frame is not None else None
            if pinned:
                np.subtract(chunk.indptr[lo:hi + 1], a, out=blk.crow, casting="unsafe")
                tk = st.commit_from_chunk(blk, chunk.indices, chunk.data, a, n_genes)
            elif self._pool is not None:
                tk = self._pool.submit(self._pack, st, blk, chunk, lo, hi, n_genes, True)
            else:
                tk = self._pack(st, blk, chunk, lo, hi, n_genes, False)
            inflight.append((st, tk, meta))
        while inflight:
            yield hand_out()
        if done is not None:
            done[0].release(done[1])

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        for key in list(self._pinned):
            self._unpin(key)
