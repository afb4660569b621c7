"""Synthetic CSR minibatches in the reference's batch format: rows sorted ascending and
duplicate-free, int32 indices, values ``log1p(1e4 * c / sum c)`` with ``c = 1 + Poisson(1.5)`` (exactly
``normalize_data``, scripts/data-preprocessing/data_processing_functions.py:24-31).  Used by bench.py, the
tools and the tests for inputs at BASELINE sizes; column popularity is uniform by default or Zipf-skewed
(real scRNA-seq detection rates are heavily skewed; uniform columns are the worst case for the SpMM)."""
from __future__ import annotations

import numpy as np


def synth_csr(n_cells: int, n_genes: int, density: float, seed: int, zipf: float = 0.0, chunk: int = 128):
    """(crow int32 [n_cells+1], col int32 [nnz], val float32 [nnz]) with round(density * n_genes) entries per
    cell.  ``zipf`` > 0: gene g is drawn with probability ~ (rank_g + 1) ** -zipf (ranks shuffled once per
    seed), without replacement inside a cell (Gumbel top-k)."""
    rng = np.random.default_rng(seed)
    per = max(1, int(round(density * n_genes)))
    logw = None
    if zipf > 0:
        ranks = rng.permutation(n_genes).astype(np.float64)
        logw = (-zipf * np.log1p(ranks)).astype(np.float32)
    cols = np.empty((n_cells, per), dtype=np.int32)
    for lo in range(0, n_cells, chunk):
        hi = min(n_cells, lo + chunk)
        keys = rng.random((hi - lo, n_genes), dtype=np.float32)
        if logw is not None:     # Gumbel top-k == sampling without replacement proportional to exp(logw)
            keys = logw - np.log(-np.log(np.clip(keys, 1e-12, 1 - 1e-7)))
            keys = -keys
        idx = np.argpartition(keys, per - 1, axis=1)[:, :per]
        idx.sort(axis=1)
        cols[lo:hi] = idx
    counts = 1.0 + rng.poisson(1.5, size=(n_cells, per))
    vals = np.log1p(1e4 * counts / counts.sum(axis=1, keepdims=True)).astype(np.float32)
    crow = (np.arange(n_cells + 1, dtype=np.int64) * per).astype(np.int32)
    return crow, cols.reshape(-1), vals.reshape(-1)


def synth_chunk(n_cells: int, n_genes: int, density: float, seed: int, zipf: float = 0.0):
    """the same rows as a ``scipy.sparse.csr_matrix`` chunk (what the reference loads from its ``.npz`` files)"""
    import scipy.sparse as sp
    crow, col, val = synth_csr(n_cells, n_genes, density, seed, zipf)
    return sp.csr_matrix((val, col, crow), shape=(n_cells, n_genes))
