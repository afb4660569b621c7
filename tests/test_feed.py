"""CSR batch feed: the staged block must hand the step the arrays scipy emitted, bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from mmvae_b200.feed import CSRStager, block_layout, slice_rows
from oracle.cmmvae_oracle import synth_csr


def _ragged(seed, B=37, G=501):
    rng = np.random.default_rng(seed)
    m = sp.random(B, G, density=0.07, format="csr", dtype=np.float32, random_state=rng)
    m[3] = 0            # an empty cell
    m[B - 1] = 0        # empty last row
    m.eliminate_zeros()
    m.sort_indices()
    return m


def test_block_layout_is_aligned_and_disjoint():
    for B, nnz in [(0, 0), (1, 1), (37, 1234), (1024, 3_099_136)]:
        o_crow, o_col, o_val, size = block_layout(B, nnz)
        assert o_crow == 0 and o_col % 16 == 0 and o_val % 16 == 0 and size % 16 == 0
        assert o_col >= 4 * (B + 1) and o_val >= o_col + 4 * nnz and size >= o_val + 4 * nnz


def test_slice_rows_matches_scipy_row_slicing():
    m = _ragged(0)
    for lo, hi in [(0, 37), (2, 5), (3, 4), (30, 37), (5, 5)]:
        crow, col, val = slice_rows(m.indptr, m.indices, m.data, lo, hi)
        ref = m[lo:hi]
        assert crow.dtype == np.int32 and col.dtype == np.int32 and val.dtype == np.float32
        np.testing.assert_array_equal(crow, ref.indptr)
        np.testing.assert_array_equal(col, ref.indices)
        np.testing.assert_array_equal(val, ref.data)


def test_stager_cpu_round_trip_and_rotation():
    st = CSRStager(max_cells=64, max_nnz=4096, device="cpu", depth=2)
    kept = []
    for seed in range(5):
        m = _ragged(seed)
        t = st.put(m.indptr, m.indices, m.data, m.shape[1])
        crow, col, val = st.arrays(t)
        assert crow.dtype == torch.int32 and col.dtype == torch.int32 and val.dtype == torch.float32
        np.testing.assert_array_equal(crow.numpy(), m.indptr)
        np.testing.assert_array_equal(col.numpy(), m.indices)
        np.testing.assert_array_equal(val.numpy().view(np.uint32), m.data.view(np.uint32))
        x = st.get(t)
        assert x.layout == torch.sparse_csr and tuple(x.shape) == m.shape
        kept.append(t)
    assert kept[0].slot == kept[2].slot == kept[4].slot and kept[1].slot == kept[3].slot


def test_reserve_fill_in_place_commit_and_recommit():
    m = _ragged(11)
    st = CSRStager(max_cells=64, max_nnz=4096, device="cpu", depth=2)
    lo, hi = 4, 29
    blk = st.reserve(hi - lo, int(m.indptr[hi] - m.indptr[lo]))
    slice_rows(m.indptr, m.indices, m.data, lo, hi, out=blk)
    ref = m[lo:hi]
    for _ in range(2):      # an unchanged block may be shipped again
        crow, col, val = st.arrays(st.commit(blk, m.shape[1]))
        np.testing.assert_array_equal(crow.numpy(), ref.indptr)
        np.testing.assert_array_equal(col.numpy(), ref.indices)
        np.testing.assert_array_equal(val.numpy(), ref.data)
    blk.crow[-1] += 1
    with pytest.raises(ValueError, match="inconsistent"):
        st.commit(blk, m.shape[1])


def test_stager_accepts_int64_indices_and_empty_batches():
    st = CSRStager(max_cells=8, max_nnz=16, device="cpu")
    t = st.put(np.array([0, 2, 2, 3], dtype=np.int64), np.array([5, 9, 1], dtype=np.int64),
               np.array([1.5, 2.5, 3.5], dtype=np.float64), 10)
    crow, col, val = st.arrays(t)
    assert crow.tolist() == [0, 2, 2, 3] and col.tolist() == [5, 9, 1] and val.tolist() == [1.5, 2.5, 3.5]
    t = st.put(np.zeros(4, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.float32), 10)
    crow, col, val = st.arrays(t)
    assert crow.tolist() == [0, 0, 0, 0] and col.numel() == 0 and val.numel() == 0


def test_stager_rejects_oversize_and_inconsistent_input():
    st = CSRStager(max_cells=4, max_nnz=8, device="cpu")
    with pytest.raises(ValueError, match="exceeds"):
        st.put(np.arange(7, dtype=np.int32), np.zeros(6, np.int32), np.zeros(6, np.float32), 10)
    with pytest.raises(ValueError, match="exceeds"):
        st.put(np.array([0, 9], np.int32), np.zeros(9, np.int32), np.zeros(9, np.float32), 10)
    with pytest.raises(ValueError, match="inconsistent"):
        st.put(np.array([0, 3], np.int32), np.zeros(2, np.int32), np.zeros(2, np.float32), 10)
    with pytest.raises(ValueError, match="depth"):
        CSRStager(4, 8, device="cpu", depth=1)


@pytest.mark.gpu
def test_stager_gpu_bit_identical_over_many_rotations():
    B, G = 256, 60530
    st = CSRStager(max_cells=B, max_nnz=int(B * G * 0.06), device="cuda", depth=3)
    tickets, srcs = [], []
    for seed in range(7):
        crow, col, val = synth_csr(B, G, 0.05, seed)
        t = st.put(crow, col, val, G)
        dcrow, dcol, dval = st.arrays(t)
        # consume on the current stream (a reduction stands in for the step), then release the slot
        got = (dcrow.clone(), dcol.clone(), dval.clone())
        st.release(t)
        tickets.append(got)
        srcs.append((crow, col, val))
    torch.cuda.synchronize()
    for (dcrow, dcol, dval), (crow, col, val) in zip(tickets, srcs):
        np.testing.assert_array_equal(dcrow.cpu().numpy(), crow)
        np.testing.assert_array_equal(dcol.cpu().numpy(), col)
        np.testing.assert_array_equal(dval.cpu().numpy().view(np.uint32), val.view(np.uint32))
    assert st.host[0].is_pinned()


@pytest.mark.gpu
def test_training_step_through_the_stager_matches_direct_feed(tmp_path):
    import pandas as pd
    from helpers import CONDITIONS, GoldenCase, build_b200_model, csr_batch
    from mmvae_b200 import layers as L
    from mmvae_b200.modules.base import KLAnnealingFn
    gc = GoldenCase("core_human")
    L.set_precision("fp32")
    try:
        logs = []
        for staged in (False, True):
            model = build_b200_model(gc, tmp_path, kl_fn=KLAnnealingFn(0.5))
            model.load_state_dict({f"module.{k}": v for k, v in gc.state("init").items()}, strict=True)
            model.cuda().train()
            model.configure_optimizers()
            st = CSRStager(64, 4096, device="cuda", depth=2)
            rec = {}
            for t in range(gc.n_steps):
                s = gc.step(t)
                L.inject_noise(s["eps"].cuda())
                meta = pd.DataFrame({c: [f"{c}_{int(i)}" for i in s["labels"][c]] for c in CONDITIONS})
                if staged:
                    tk = st.put(s["crow"], s["col"], s["val"], gc.genes["human"])
                    x = st.get(tk)
                else:
                    x = csr_batch(s["crow"], s["col"], s["val"], gc.genes["human"])
                model.training_step((x, meta, "human"), t)
                if staged:
                    st.release(tk)
                rec.update({f"{t}/{k}": float(v) for k, v in model.logged_metrics.items()})
            logs.append(rec)
        assert logs[0].keys() == logs[1].keys() and len(logs[0]) > 0
        for k in logs[0]:
            assert logs[0][k] == pytest.approx(logs[1][k], rel=1e-6), k
    finally:
        L.set_precision("bf16")


def _chunks(n_chunks, rows, G=157, seed=0):
    import pandas as pd
    out = []
    for k in range(n_chunks):
        m = sp.random(rows[k], G, density=0.08 + 0.05 * k, format="csr", dtype=np.float32,
                      random_state=np.random.default_rng(seed + k))
        m.sort_indices()
        out.append((m, pd.DataFrame({"cell": [f"c{k}_{i}" for i in range(rows[k])], "chunk": k})))
    return out


@pytest.mark.parametrize("allow_partials", [False, True])
def test_staged_batches_iterate_like_the_reference_batcher(allow_partials):
    """same batches, in the same order, as SparseCSRMatrixBatcherDataPipe (cellxgene_datapipe.py:169-193):
    consecutive row slices of each chunk, short trailing batch only with allow_partials, metadata re-indexed"""
    from mmvae_b200.feed import StagedCSRBatches
    bs = 16
    chunks = _chunks(3, [50, 16, 33])        # second chunk is denser than the first: the ring must grow
    want = []
    for m, frame in chunks:
        for i in range(0, m.shape[0], bs):
            b = m[i:i + bs]
            if b.shape[0] != bs and not allow_partials:
                continue
            want.append((b, frame.iloc[i:i + bs].reset_index(drop=True)))
    assert len(want) == (8 if allow_partials else 6)
    n = 0
    # an item is valid until the next one is requested (its block is recycled after that): compare as we go
    staged = iter(StagedCSRBatches(chunks, bs, allow_partials=allow_partials, device="cpu"))
    for b, wmeta in want:
        x, meta = next(staged)
        n += 1
        assert x.layout == torch.sparse_csr and tuple(x.shape) == b.shape
        np.testing.assert_array_equal(x.crow_indices().numpy(), b.indptr)
        np.testing.assert_array_equal(x.col_indices().numpy(), b.indices)
        np.testing.assert_array_equal(x.values().numpy().view(np.uint32), b.data.view(np.uint32))
        assert meta.equals(wmeta)
    assert n == len(want) and next(staged, None) is None


def test_staged_batches_keep_a_batch_valid_until_the_next_one_is_requested():
    from mmvae_b200.feed import StagedCSRBatches
    chunks = _chunks(1, [64])
    it = iter(StagedCSRBatches(chunks, 8, device="cpu", depth=3))
    x0, _ = next(it)
    snap = (x0.crow_indices().clone(), x0.col_indices().clone(), x0.values().clone())
    x1, _ = next(it)            # batch 2 has been staged by now; batch 0's block must still be intact
    assert torch.equal(x0.crow_indices(), snap[0]) and torch.equal(x0.col_indices(), snap[1])
    assert torch.equal(x0.values(), snap[2])
    assert len(list(it)) == 6
    with pytest.raises(ValueError):
        StagedCSRBatches(chunks, 8, device="cpu", depth=2)


def test_native_packer_equals_scipy_slicing_narrow_and_wide():
    """cmmvae_host_slice_rows (host C code of libcmmvae_b200.so): same rows as scipy's chunk[lo:hi], bit for bit,
    for int32 / int64 index arrays, into int32 and uint16 column blocks; out-of-range gene ids are refused."""
    from mmvae_b200 import ops
    m = _ragged(5)
    for idx_t in (np.int32, np.int64):
        indptr, indices = m.indptr.astype(idx_t), m.indices.astype(idx_t)
        for lo, hi in [(0, 37), (2, 5), (3, 4), (30, 37), (5, 5)]:
            ref = m[lo:hi]
            n = int(ref.nnz)
            for col_t in (np.int32, np.uint16):
                crow, col, val = np.full(hi - lo + 1, -7, np.int32), np.zeros(n, col_t), np.zeros(n, np.float32)
                assert ops.host_slice_rows(indptr, indices, m.data, lo, hi, m.shape[1], crow, col, val) == n
                np.testing.assert_array_equal(crow, ref.indptr)
                np.testing.assert_array_equal(col.astype(np.int64), ref.indices)
                np.testing.assert_array_equal(val.view(np.uint32), ref.data.view(np.uint32))
    with pytest.raises(ValueError, match="gene id"):
        ops.host_slice_rows(m.indptr, m.indices, m.data, 0, 37, 100, np.zeros(38, np.int32),
                            np.zeros(m.nnz, np.int32), np.zeros(m.nnz, np.float32))


@pytest.mark.parametrize("workers", [0, 3])
def test_staged_batches_narrow_columns_and_worker_threads(workers):
    """uint16 gene ids on the wire + background packing threads: same batches, same order, same bits as the
    reference batcher's scipy slices"""
    from mmvae_b200.feed import StagedCSRBatches
    chunks = _chunks(3, [40, 64, 24])
    batcher = StagedCSRBatches(chunks, 8, device="cpu", depth=3, workers=workers, narrow_col=True)
    got = [(torch.sparse_csr_tensor(x.crow_indices().clone(), x.col_indices().clone(), x.values().clone(),
                                    size=x.shape), meta) for x, meta in batcher]    # blocks are recycled: copy out
    batcher.close()
    want = [(c[lo:lo + 8], f.iloc[lo:lo + 8]) for c, f in chunks for lo in range(0, c.shape[0] - 7, 8)]
    assert len(got) == len(want) and batcher.stager.narrow
    for (x, meta), (ref, fref) in zip(got, want):
        np.testing.assert_array_equal(x.crow_indices().numpy(), ref.indptr)
        np.testing.assert_array_equal(x.col_indices().numpy(), ref.indices)
        assert x.col_indices().dtype == torch.int32
        np.testing.assert_array_equal(x.values().numpy().view(np.uint32), ref.data.view(np.uint32))
        assert list(meta["cell"]) == list(fref["cell"])


@pytest.mark.gpu
def test_pinned_chunks_route_ships_the_same_batches_without_a_host_pass():
    """pin_chunks=True: every chunk is page-locked in place once and its batches are DMA'd straight out of the
    chunk's own indices / data arrays -- same tensors as scipy's chunk[lo:hi], bit for bit, over chunk changes"""
    from mmvae_b200.feed import StagedCSRBatches
    bs = 16
    chunks = _chunks(4, [64, 48, 80, 32], G=911)
    feed = StagedCSRBatches(chunks, bs, device="cuda", pin_chunks=True)
    n = 0
    it = iter(feed)
    for m, frame in chunks:
        for i in range(0, m.shape[0], bs):
            b = m[i:i + bs]
            x, meta = next(it)
            torch.cuda.synchronize()
            n += 1
            assert x.is_cuda and tuple(x.shape) == b.shape
            np.testing.assert_array_equal(x.crow_indices().cpu().numpy(), b.indptr)
            np.testing.assert_array_equal(x.col_indices().cpu().numpy(), b.indices)
            np.testing.assert_array_equal(x.values().cpu().numpy().view(np.uint32), b.data.view(np.uint32))
            assert meta.equals(frame.iloc[i:i + bs].reset_index(drop=True))
    assert next(it, None) is None and n == 14
    assert feed.chunks_pinned == 4 and len(feed._pinned) == 2        # LRU of two registered chunks
    feed.close()
    assert not feed._pinned
    bad = _chunks(1, [32], G=911)
    bad[0][0].indices[5] = 911
    with pytest.raises(ValueError, match="gene id"):
        next(iter(StagedCSRBatches(bad, bs, device="cuda", pin_chunks=True)))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [0, 1, 7, 8, 4099, 1 << 20])
def test_device_feed_kernels_widen_and_copy(n):
    """cmmvae_widen_u16_i32 (narrow gene ids -> int32 in HBM) and cmmvae_copy_bytes (kernel copy to the fixed
    addresses the step's graphs read), all lengths incl. the non-multiple-of-16-byte tails"""
    from mmvae_b200 import ops
    g = torch.Generator().manual_seed(n)
    src = torch.randint(0, 65536, (n + 8,), generator=g, dtype=torch.int32)
    u16 = src.to(torch.uint16).cuda()
    out = torch.full((n + 8,), -1, dtype=torch.int32, device="cuda")
    ops.widen_u16_i32(u16, out, n)
    assert torch.equal(out[:n].cpu(), src[:n]) and bool((out[n:] == -1).all())
    a = torch.randint(0, 256, (4 * n + 19,), generator=g, dtype=torch.uint8).cuda()
    b = torch.zeros_like(a)
    ops.copy_bytes(b, a, 4 * n + 3)
    assert torch.equal(b[:4 * n + 3], a[:4 * n + 3]) and not bool(b[4 * n + 3:].any())
