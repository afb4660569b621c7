"""Thin torch-facing wrappers over the C ABI (include/cmmvae_b200.h).

Every function takes CUDA tensors, passes raw device pointers + sizes + the current CUDA stream to
libcmmvae_b200.so through ctypes and returns tensors.  torch is used for device memory and
streams only.  CPU tensors are rejected: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib

F32, BF16 = 0, 1
_c = ctypes


def lib():
    return _lib.load()


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"cmmvae_b200.{what} failed ({rc}): {lib().cmmvae_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return _c.c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("cmmvae_b200 kernels need CUDA tensors (no CPU fallback)")
    return _c.c_void_p(t.data_ptr())


_stream_override = None   # set by the step engine for the duration of a step (avoids ~1 us per launch)


def _stream():
    if _stream_override is not None:
        return _stream_override
    return _c.c_void_p(torch.cuda.current_stream().cuda_stream)


class stream_scope:
    """``with ops.stream_scope(torch_stream):`` every op launches on that stream without asking torch for the
    current stream on each call.  The caller guarantees it IS torch's current stream inside the scope."""

    def __init__(self, stream):
        self.handle = _c.c_void_p(stream.cuda_stream)

    def __enter__(self):
        global _stream_override
        self.prev, _stream_override = _stream_override, self.handle
        return self

    def __exit__(self, *exc):
        global _stream_override
        _stream_override = self.prev
        return False


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def launch_count() -> int:
    return int(lib().cmmvae_launch_count())


def set_pdl(on: bool) -> None:
    _check(lib().cmmvae_set_pdl(int(bool(on))), "set_pdl")


def set_sm_budget(sms: int) -> None:
    _check(lib().cmmvae_set_sm_budget(int(sms)), "set_sm_budget")


# ------------------------------------------------------------------------------------------ sparse
def csr_linear_fwd(crow, col, val, G: int, Wt, bias, out=None):
    """Y = X_csr @ Wt + bias;  Wt is [G,H] (f32|bf16), contiguous."""
    B = crow.numel() - 1
    H = Wt.shape[1]
    assert Wt.shape[0] == G and Wt.is_contiguous()
    assert crow.dtype == torch.int32 and col.dtype == torch.int32 and val.dtype == torch.float32
    if out is None:
        out = torch.empty(B, H, device=Wt.device, dtype=torch.float32)
    _check(lib().cmmvae_csr_linear_fwd(_ptr(crow), _ptr(col), _ptr(val), B, G, H, _ptr(Wt), _dt(Wt), _ptr(bias),
                                       _ptr(out), _stream()), "csr_linear_fwd")
    return out


def csr_transpose(crow, col, val, G: int, nnz: int, cptr=None, ridx=None, cval=None, cursor=None):
    B = crow.numel() - 1
    dev = crow.device
    if cptr is None:
        cptr = torch.empty(G + 1, device=dev, dtype=torch.int32)
    if cursor is None:
        cursor = torch.empty(G + 1, device=dev, dtype=torch.int32)
    if ridx is None:
        ridx = torch.empty(max(nnz, 1), device=dev, dtype=torch.int32)
    if cval is None:
        cval = torch.empty(max(nnz, 1), device=dev, dtype=torch.float32)
    _check(lib().cmmvae_csr_transpose(_ptr(crow), _ptr(col), _ptr(val), B, G, _c.c_longlong(nnz), _ptr(cptr),
                                      _ptr(ridx), _ptr(cval), _ptr(cursor), _stream()), "csr_transpose")
    return cptr, ridx, cval


def csr_linear_bwd_w(cptr, ridx, cval, B: int, G: int, dY, out):
    H = dY.shape[1]
    assert dY.dtype == torch.float32 and dY.is_contiguous() and out.is_contiguous() and out.shape == (G, H)
    _check(lib().cmmvae_csr_linear_bwd_w(_ptr(cptr), _ptr(ridx), _ptr(cval), B, G, H, _ptr(dY), _ptr(out),
                                         _stream()), "csr_linear_bwd_w")
    return out


def csr_tile_ptr(crow, col, val, G: int, nnz: int, tp=None, packed=None):
    """(tile_ptr int32 [G/64+1, B] window-major, packed uint32-as-int32 [nnz padded]) for the tensor-pipe SpMM"""
    B = crow.numel() - 1
    fn = lib().cmmvae_csr_tile_ptr_bytes
    fn.restype = _c.c_size_t
    n = int(fn(B, G)) // 4
    fp = lib().cmmvae_csr_packed_bytes
    fp.restype = _c.c_size_t
    m = int(fp(_c.c_longlong(nnz))) // 4
    if tp is None:
        tp = torch.empty(n, dtype=torch.int32, device=crow.device)
    if packed is None:
        packed = torch.empty(m, dtype=torch.int32, device=crow.device)
    assert tp.numel() >= n and packed.numel() >= m
    _check(lib().cmmvae_csr_tile_ptr(_ptr(crow), _ptr(col), _ptr(val), B, G, _c.c_longlong(nnz), _ptr(tp),
                                     _ptr(packed), _stream()), "csr_tile_ptr")
    return tp, packed


def csr_tile_ptr_dyn(crow, col, val, G: int, cap: int, tp, packed):
    """csr_tile_ptr whose non-zero count is read on the device (crow[B]); ``cap`` records are written"""
    B = crow.numel() - 1
    assert tp.numel() >= B * ((G + 63) // 64 + 1) and packed.numel() >= (cap + 3) // 4 * 4 + 4
    _check(lib().cmmvae_csr_tile_ptr_dyn(_ptr(crow), _ptr(col), _ptr(val), B, G, _c.c_longlong(cap), _ptr(tp),
                                         _ptr(packed), _stream()), "csr_tile_ptr_dyn")
    return tp, packed


def csr_linear_fwd_tc(packed, tile_ptr, B: int, G: int, Wt16, bias, out=None):
    H = Wt16.shape[1]
    assert Wt16.dtype == torch.bfloat16 and Wt16.is_contiguous() and Wt16.shape[0] == G
    if out is None:
        out = torch.empty(B, H, device=Wt16.device, dtype=torch.float32)
    _check(lib().cmmvae_csr_linear_fwd_tc(_ptr(packed), _ptr(tile_ptr), B, G, H, _ptr(Wt16), _ptr(bias),
                                          _ptr(out), _stream()), "csr_linear_fwd_tc")
    return out


def csr_linear_bwd_w_tc(packed, tile_ptr, B: int, G: int, dY16, out, sumsq_out=None, g_begin=0, g_end=0):
    H = dY16.shape[1]
    assert dY16.dtype == torch.bfloat16 and dY16.is_contiguous() and out.is_contiguous() and out.shape == (G, H)
    _check(lib().cmmvae_csr_linear_bwd_w_tc(_ptr(packed), _ptr(tile_ptr), B, G, H, _ptr(dY16), _ptr(out),
                                            int(g_begin), int(g_end), _ptr(sumsq_out), _stream()),
           "csr_linear_bwd_w_tc")
    return out


def csr_linear_bwd_w_tc_shard(packed_all, tp_shard, B_all: int, G_pad: int, dY16_all, out_shard, g_begin: int,
                              g_end: int, sumsq_out=None):
    """gene shard [g_begin, g_end) of X_all^T . dY_all (all ranks' cells); ``tp_shard``/``out_shard`` hold only
    the shard's windows / rows"""
    H = dY16_all.shape[1]
    assert dY16_all.dtype == torch.bfloat16 and dY16_all.is_contiguous() and out_shard.is_contiguous()
    assert out_shard.shape == (g_end - g_begin, H) and tp_shard.numel() >= ((g_end - g_begin) // 64 + 1) * B_all
    _check(lib().cmmvae_csr_linear_bwd_w_tc_shard(_ptr(packed_all), _ptr(tp_shard), B_all, G_pad, H, _ptr(dY16_all),
                                                  _ptr(out_shard), int(g_begin), int(g_end), _ptr(sumsq_out),
                                                  _stream()), "csr_linear_bwd_w_tc_shard")
    return out_shard


def mse_relu_csr(logits, G: int, crow, col, val, write_xhat: bool, dl32, dl16, loss_sum):
    B = logits.shape[0]
    ldd = (dl32 if dl32 is not None else dl16).stride(0) if (dl32 is not None or dl16 is not None) else 0
    _check(lib().cmmvae_mse_relu_csr(_ptr(logits), logits.stride(0), B, G, _ptr(crow), _ptr(col), _ptr(val),
                                     int(write_xhat), _ptr(dl32), _ptr(dl16), ldd, _ptr(loss_sum), _stream()),
           "mse_relu_csr")


def decoder_mse_fused_workspace_bytes(B: int, G: int) -> int:
    fn = lib().cmmvae_decoder_mse_fused_workspace_bytes
    fn.restype = _c.c_size_t
    return int(fn(B, G))


def decoder_mse_fused(h16, Wout16, bout, G: int, crow, col, val, dl16, loss_sum, workspace=None, tile_ptr=None):
    B, H = h16.shape
    if workspace is None and tile_ptr is None:
        workspace = torch.empty(decoder_mse_fused_workspace_bytes(B, G), dtype=torch.uint8, device=h16.device)
    _check(lib().cmmvae_decoder_mse_fused(_ptr(h16), h16.stride(0), _ptr(Wout16), Wout16.stride(0), _ptr(bout), B, G,
                                          H, _ptr(crow), _ptr(col), _ptr(val), _ptr(tile_ptr), _ptr(dl16),
                                          dl16.stride(0), _ptr(loss_sum), _ptr(workspace), _stream()),
           "decoder_mse_fused")


# -------------------------------------------------------------------------------- BN / act / drop
_BN_SCRATCH = {}


def bn_stats_scratch(H: int, device) -> torch.Tensor:
    """the zero-initialised, self-cleaning scratch ``bn_stats`` needs for width H (cached per device)"""
    key = (int(H), str(device))
    s = _BN_SCRATCH.get(key)
    if s is None:
        fn = lib().cmmvae_bn_stats_scratch_bytes
        fn.restype = _c.c_size_t
        s = _BN_SCRATCH[key] = torch.zeros(int(fn(int(H))) // 8, dtype=torch.float64, device=device)
    return s


def bn_stats(Y, eps, momentum, mean, rstd, running_mean, running_var, scratch=None):
    B, H = Y.shape
    if scratch is None:
        scratch = bn_stats_scratch(H, Y.device)
    _check(lib().cmmvae_bn_stats(_ptr(Y), B, H, _c.c_float(eps), _c.c_float(momentum), _ptr(mean), _ptr(rstd),
                                 _ptr(running_mean), _ptr(running_var), _ptr(scratch), _stream()), "bn_stats")


def rstd_from_var(var, eps, rstd):
    _check(lib().cmmvae_rstd_from_var(_ptr(var), var.numel(), _c.c_float(eps), _ptr(rstd), _stream()),
           "rstd_from_var")


def bn_act_drop_fwd(Y, mean, rstd, gamma, beta, relu, p_drop, seed, mask, out32, out16, seed_base=None):
    """``seed_base`` (device uint64[1], optional): effective seed = *seed_base + seed (graph replay)"""
    B, H = Y.shape
    _check(lib().cmmvae_bn_act_drop_fwd_dyn(_ptr(Y), B, H, _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(beta), int(relu),
                                            _c.c_float(p_drop), _c.c_ulonglong(seed), _ptr(seed_base), _ptr(mask),
                                            _ptr(out32), _ptr(out16), _stream()), "bn_act_drop_fwd")


def bn_act_drop_bwd(dOut, Y, out, mean, rstd, gamma, relu, p_drop, seed, mask, dY, dY16, dgamma, dbeta, dbias,
                    accumulate=False, seed_base=None):
    B, H = dOut.shape
    _check(lib().cmmvae_bn_act_drop_bwd_dyn(_ptr(dOut), _ptr(Y), _ptr(out), B, H, _ptr(mean), _ptr(rstd), _ptr(gamma),
                                            int(relu), _c.c_float(p_drop), _c.c_ulonglong(seed), _ptr(seed_base),
                                            _ptr(mask), _ptr(dY), _ptr(dY16), _ptr(dgamma), _ptr(dbeta), _ptr(dbias),
                                            int(accumulate), _stream()),
           "bn_act_drop_bwd")


# ------------------------------------------------------------------------------------------ GEMMs
def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1
    return t.stride(0)


def gemm(A, transA, Bm, transB, M, N, K, bias=None, relu=False, accumulate=False, C32=None, C16=None,
         use_tc: Optional[bool] = None, sumsq_out=None, tf32: bool = False):
    """C[M,N] = act(opA(A) opB(B) + bias) (+C).  A: [M,K] or (transA) [K,M]; B: [N,K] or (transB) [K,N].
    bf16 operands -> tcgen05 kind::f16; f32 operands with ``tf32`` -> tcgen05 kind::tf32; f32 otherwise -> CUDA cores."""
    C = C32 if C32 is not None else C16
    ldc = _ld(C)
    if C32 is not None and C16 is not None:
        assert _ld(C32) == _ld(C16)
    if tf32:
        assert A.dtype == torch.float32 and Bm.dtype == torch.float32
        _check(lib().cmmvae_gemm_tf32_tc(_ptr(A), _ld(A), int(transA), _ptr(Bm), _ld(Bm), int(transB), M, N, K,
                                         _ptr(bias), int(relu), int(accumulate), _ptr(C32), _ptr(C16), ldc,
                                         _ptr(sumsq_out), _stream()), "gemm_tf32_tc")
        return C
    tc = (A.dtype == torch.bfloat16) if use_tc is None else use_tc
    if tc:
        assert A.dtype == torch.bfloat16 and Bm.dtype == torch.bfloat16
        _check(lib().cmmvae_gemm_bf16_tc(_ptr(A), _ld(A), int(transA), _ptr(Bm), _ld(Bm), int(transB), M, N, K,
                                         _ptr(bias), int(relu), int(accumulate), _ptr(C32), _ptr(C16), ldc,
                                         _ptr(sumsq_out), _stream()), "gemm_bf16_tc")
        return C
    assert A.dtype == torch.float32 and Bm.dtype == torch.float32 and sumsq_out is None
    _check(lib().cmmvae_gemm_f32(_ptr(A), _ld(A), int(transA), _ptr(Bm), _ld(Bm), int(transB), M, N, K, _ptr(bias),
                                 int(relu), int(accumulate), _ptr(C32), _ptr(C16), ldc, _stream()), "gemm_f32")
    return C


def colsum(X, out, accumulate=False, M=None, N=None):
    M = X.shape[0] if M is None else M
    N = X.shape[1] if N is None else N
    _check(lib().cmmvae_colsum(_ptr(X), _dt(X), M, N, _ld(X), _ptr(out), int(accumulate), _stream()), "colsum")
    return out


# --------------------------------------------------------------------------------- latent / losses
def reparam_kl_fwd(ML, eps, Z, var_eps, z32, z16, sums):
    B = ML.shape[0]
    _check(lib().cmmvae_reparam_kl_fwd(_ptr(ML), _ptr(eps), B, Z, _c.c_float(var_eps), _ptr(z32), _ptr(z16),
                                       _ptr(sums), _stream()), "reparam_kl_fwd")


def reparam_kl_bwd(ML, eps, dz, Z, var_eps, kl_scale, dML, dML16, kl_weight_dev=None):
    """``kl_weight_dev`` (device float[1], optional): effective scale = kl_scale * *kl_weight_dev (graph replay)"""
    B = ML.shape[0]
    _check(lib().cmmvae_reparam_kl_bwd_dyn(_ptr(ML), _ptr(eps), _ptr(dz), B, Z, _c.c_float(var_eps),
                                           _c.c_float(kl_scale), _ptr(kl_weight_dev), _ptr(dML), _ptr(dML16),
                                           _stream()), "reparam_kl_bwd")


def softmax_ce_sum(logits, C, labels, scale, dlogits, loss_sum):
    B = logits.shape[0]
    assert labels.dtype == torch.int64
    _check(lib().cmmvae_softmax_ce_sum(_ptr(logits), _ld(logits), B, C, _ptr(labels), _c.c_float(scale),
                                       _ptr(dlogits), _ld(dlogits) if dlogits is not None else 0, _ptr(loss_sum),
                                       _stream()), "softmax_ce_sum")


# --------------------------------------------------------------------------------------- optimiser
def sumsq(g, norm_sq):
    _check(lib().cmmvae_sumsq(_ptr(g), _c.c_longlong(g.numel()), _ptr(norm_sq), _stream()), "sumsq")


def clip_adam(p, g, m, v, p16, norm_sq, max_norm, grad_scale, lr, beta1, beta2, eps, wd, step, background=False,
              bc_dev=None):
    """``bc_dev`` (device float[2], optional): the two bias corrections live in device memory (graph replay)"""
    if bc_dev is not None:
        _check(lib().cmmvae_clip_adam_dyn(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p16), _c.c_longlong(p.numel()),
                                          _ptr(norm_sq), _c.c_float(max_norm if max_norm else 0.0),
                                          _c.c_float(grad_scale), _c.c_float(lr), _c.c_float(beta1), _c.c_float(beta2),
                                          _c.c_float(eps), _c.c_float(wd), _ptr(bc_dev), int(background), _stream()),
               "clip_adam_dyn")
        return
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    fn = lib().cmmvae_clip_adam_bg if background else lib().cmmvae_clip_adam
    _check(fn(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(p16), _c.c_longlong(p.numel()),
                                  _ptr(norm_sq), _c.c_float(max_norm if max_norm else 0.0), _c.c_float(grad_scale),
                                  _c.c_float(lr), _c.c_float(beta1), _c.c_float(beta2), _c.c_float(eps),
                                  _c.c_float(wd), _c.c_float(bc1), _c.c_float(bc2), _stream()), "clip_adam")


# --------------------------------------------------------------------------------------- utilities
def cast_bf16(src, dst):
    _check(lib().cmmvae_cast_f32_bf16(_ptr(src), _ptr(dst), _c.c_longlong(src.numel()), _stream()), "cast_f32_bf16")
    return dst


def transpose(src, dst):
    R, C = src.shape
    _check(lib().cmmvae_transpose(_ptr(src), _ptr(dst), _dt(src), R, C, _ld(src), _ld(dst), _stream()), "transpose")
    return dst


def axpy(a, b, alpha):
    _check(lib().cmmvae_axpy(_ptr(a), _ptr(b), _c.c_float(alpha), _c.c_longlong(a.numel()), _stream()), "axpy")


def fold_cols(dcat, n: int, dz):
    """dz[b, j] = sum_k dcat[b, k*Z + j] (backward of tile_cols)"""
    B, Z = dz.shape
    _check(lib().cmmvae_fold_cols(_ptr(dcat), B, Z, int(n), _ptr(dz), _stream()), "fold_cols")
    return dz


def cond_fwd(params, S, Zin, Zout, tiles, n_tiles, rows, x, ldx, out, out16, pre, ldo, out_col, rstd, B, layer_norm, relu):
    _check(lib().cmmvae_cond_fwd(_ptr(params), _c.c_longlong(S), Zin, Zout, _ptr(tiles), int(n_tiles), _ptr(rows), _ptr(x),
                                 int(ldx), _ptr(out), _ptr(out16), _ptr(pre), int(ldo), _ptr(out_col), _ptr(rstd), int(B),
                                 int(layer_norm), int(relu), _stream()), "cond_fwd")


def cond_bwd(params, grads, S, Zin, Zout, tiles, n_tiles, rows, x, ldx, dout, pre, ldo, out_col, rstd, B, dx, lddx,
             dx_col, layer_norm, relu):
    _check(lib().cmmvae_cond_bwd(_ptr(params), _ptr(grads), _c.c_longlong(S), Zin, Zout, _ptr(tiles), int(n_tiles),
                                 _ptr(rows), _ptr(x), int(ldx), _ptr(dout), _ptr(pre), int(ldo), _ptr(out_col),
                                 _ptr(rstd), int(B), _ptr(dx), int(lddx), _ptr(dx_col), int(layer_norm), int(relu),
                                 _stream()), "cond_bwd")


def cond_zero_grads(grads, S, present, n_present):
    _check(lib().cmmvae_cond_zero_grads(_ptr(grads), _c.c_longlong(S), _ptr(present), int(n_present), _stream()),
           "cond_zero_grads")


def cond_sumsq(grads, S, present, n_present, out):
    _check(lib().cmmvae_cond_sumsq(_ptr(grads), _c.c_longlong(S), _ptr(present), int(n_present), _ptr(out), _stream()),
           "cond_sumsq")


def cond_adam(params, grads, m, v, S, present, n_present, steps, norm_sq, max_norm, grad_scale, lr, b1, b2, eps, wd):
    _check(lib().cmmvae_cond_adam(_ptr(params), _ptr(grads), _ptr(m), _ptr(v), _c.c_longlong(S), _ptr(present),
                                  int(n_present), _ptr(steps), _ptr(norm_sq), _c.c_float(max_norm or 0.0),
                                  _c.c_float(grad_scale), _c.c_float(lr), _c.c_double(b1), _c.c_double(b2),
                                  _c.c_float(eps), _c.c_float(wd), _stream()), "cond_adam")


def widen_u16_i32(src_u16, dst_i32, n: int, stream=None):
    """dst int32[n] = src uint16[n] (gene ids travel narrow over PCIe; mmvae_b200.feed)"""
    st = _c.c_void_p(stream.cuda_stream) if stream is not None else _stream()
    _check(lib().cmmvae_widen_u16_i32(_ptr(src_u16), _ptr(dst_i32), _c.c_longlong(n), st), "widen_u16_i32")
    return dst_i32


def host_slice_rows(indptr, indices, data, lo: int, hi: int, n_genes: int, crow_out, col_out, val_out) -> int:
    """HOST packer (numpy arrays in, numpy views of a pinned block out): rows [lo, hi) of a CSR chunk; see
    cmmvae_host_slice_rows.  Releases the GIL for the duration of the copy."""
    import numpy as np
    fn = lib().cmmvae_host_slice_rows
    fn.restype = _c.c_longlong
    p = lambda a: _c.c_void_p(a.ctypes.data)  # noqa: E731
    n = fn(p(indptr), indptr.dtype.itemsize, p(indices), indices.dtype.itemsize, p(data), _c.c_longlong(lo),
           _c.c_longlong(hi), _c.c_longlong(n_genes), p(crow_out), p(col_out), int(col_out.dtype == np.uint16),
           p(val_out))
    if n < 0:
        raise ValueError(f"host_slice_rows failed ({n}): {lib().cmmvae_last_error().decode()}")
    return int(n)


def copy_bytes(dst: torch.Tensor, src: torch.Tensor, nbytes: int) -> None:
    """device-to-device copy of ``nbytes`` by a kernel on the current stream (see cmmvae_copy_bytes)"""
    _check(lib().cmmvae_copy_bytes(_ptr(dst), _ptr(src), _c.c_longlong(nbytes), _stream()), "copy_bytes")


def host_register(arr) -> None:
    """page-lock a numpy array's memory in place (cmmvae_host_register); raises RuntimeError if the driver refuses"""
    if lib().cmmvae_host_register(_c.c_void_p(arr.ctypes.data), _c.c_longlong(arr.nbytes)) != 0:
        raise RuntimeError(lib().cmmvae_last_error().decode())


def host_unregister(arr) -> None:
    if lib().cmmvae_host_unregister(_c.c_void_p(arr.ctypes.data)) != 0:
        raise RuntimeError(lib().cmmvae_last_error().decode())


def h2d_async(dst: torch.Tensor, src, nbytes: int, stream=None) -> None:
    """cudaMemcpyAsync of ``nbytes`` from a numpy array (view) into a device tensor, on ``stream``"""
    st = _c.c_void_p(stream.cuda_stream) if stream is not None else _stream()
    _check(lib().cmmvae_h2d_async(_ptr(dst), _c.c_void_p(src.ctypes.data), _c.c_longlong(nbytes), st), "h2d_async")


# ------------------------------------------------------------------- data parallel over peer memory
def _ptr_array(ptrs):
    arr = (_c.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    return arr


def csr_linear_fwd_tc_routed(packed, tile_ptr, B: int, G: int, Wt16, route_ptrs, route_rows: int, n_split: int = 1,
                             split_stride: int = 0):
    """partial first-layer product of a gene shard for the cells of all ranks; row r goes to route[r // route_rows]
    (gene piece z of n_split: split_stride elements further)"""
    H = Wt16.shape[1]
    assert Wt16.dtype == torch.bfloat16 and Wt16.is_contiguous() and Wt16.shape[0] == G
    _check(lib().cmmvae_csr_linear_fwd_tc_routed(_ptr(packed), _ptr(tile_ptr), B, G, H, _ptr(Wt16),
                                                 _ptr_array(route_ptrs), len(route_ptrs), int(route_rows),
                                                 int(n_split), _c.c_longlong(split_stride), _stream()),
           "csr_linear_fwd_tc_routed")


def gemm_routed(A, transA, Bm, transB, M, N, K, ldc, route_ptrs, route_rows: int, n_split: int = 1,
                split_stride: int = 0):
    assert A.dtype == torch.bfloat16 and Bm.dtype == torch.bfloat16
    _check(lib().cmmvae_gemm_bf16_tc_routed(_ptr(A), _ld(A), int(transA), _ptr(Bm), _ld(Bm), int(transB), M, N, K,
                                            int(ldc), _ptr_array(route_ptrs), len(route_ptrs), int(route_rows),
                                            int(n_split), _c.c_longlong(split_stride), _stream()),
           "gemm_bf16_tc_routed")


def decoder_mse_fused_blocks(h16, Wout16, bout, G: int, crow, col, val, dl16, loss_sums, loss_rows: int, tile_ptr):
    B, H = h16.shape
    assert loss_sums.dtype == torch.float64 and loss_sums.numel() >= (B + loss_rows - 1) // loss_rows
    _check(lib().cmmvae_decoder_mse_fused_blocks(_ptr(h16), h16.stride(0), _ptr(Wout16), Wout16.stride(0), _ptr(bout),
                                                 B, G, H, _ptr(crow), _ptr(col), _ptr(val), _ptr(tile_ptr), _ptr(dl16),
                                                 dl16.stride(0), _ptr(loss_sums), int(loss_rows), _ptr(None),
                                                 _stream()), "decoder_mse_fused_blocks")


def peer_push(src, nbytes: int, dst_ptrs, flag_ptrs, step: int, ticket, step_dev=None):
    """src -> dst_ptrs[i] on every rank, then (flag_ptrs not None) raise the flags to ``step`` (``step_dev``: device
    uint32 holding the step number instead -- graph replay)"""
    _check(lib().cmmvae_peer_push(_ptr(src), _c.c_longlong(nbytes), _ptr_array(dst_ptrs),
                                  _ptr_array(flag_ptrs) if flag_ptrs is not None else None,
                                  len(dst_ptrs), _c.c_uint(step & 0xFFFFFFFF), _ptr(step_dev), _ptr(ticket), _stream()),
           "peer_push")


def peer_signal(flag_ptrs, step: int, step_dev=None):
    _check(lib().cmmvae_peer_signal(_ptr_array(flag_ptrs), len(flag_ptrs), _c.c_uint(step & 0xFFFFFFFF),
                                    _ptr(step_dev), _stream()), "peer_signal")


def peer_wait(local_flags, n_peers: int, step: int, step_dev=None):
    _check(lib().cmmvae_peer_wait(_ptr(local_flags), int(n_peers), _c.c_uint(step & 0xFFFFFFFF), _ptr(step_dev),
                                  _stream()), "peer_wait")


def slab_sum(slabs, n_slabs: int, slab_stride: int, n: int, out32=None, out16=None, bias=None, H: int = 0):
    _check(lib().cmmvae_slab_sum(_ptr(slabs), int(n_slabs), _c.c_longlong(slab_stride), _c.c_longlong(n), _ptr(bias),
                                 int(H), _ptr(out32), _ptr(out16), _stream()), "slab_sum")


def csr_scatter_shards(crow, col, val, B: int, n_dst: int, per: int, cap: int, dst_crow, dst_col, dst_val, cnt, start,
                       offs, info):
    """all-to-all of the batch by gene shard: piece q of every row -> rank q's slab (peer pointers)"""
    _check(lib().cmmvae_csr_scatter_shards(_ptr(crow), _ptr(col), _ptr(val), B, n_dst, per, cap, _ptr_array(dst_crow),
                                           _ptr_array(dst_col), _ptr_array(dst_val), _ptr(cnt), _ptr(start),
                                           _ptr(offs), _ptr(info), _stream()), "csr_scatter_shards")


def slab_rows(slabs, slab_bytes: int, B: int, n_src: int, row_begin, row_end):
    _check(lib().cmmvae_slab_rows(_ptr(slabs), _c.c_longlong(slab_bytes), B, n_src, _ptr(row_begin), _ptr(row_end),
                                  _stream()), "slab_rows")


def csr_tile_ptr_rows(row_begin, row_end, col, val, B: int, G: int, n_records: int, tp, packed):
    assert tp.numel() >= B * ((G + 63) // 64 + 1) and packed.numel() >= (n_records + 3) // 4 * 4 + 4
    _check(lib().cmmvae_csr_tile_ptr_rows(_ptr(row_begin), _ptr(row_end), _ptr(col), _ptr(val), B, G,
                                          _c.c_longlong(n_records), _ptr(tp), _ptr(packed), _stream()),
           "csr_tile_ptr_rows")
    return tp, packed


def dp_scalars(slabs, n_src: int, stride: int, rank: int, out_recon, out_norm):
    _check(lib().cmmvae_dp_scalars(_ptr(slabs), n_src, stride, rank, _ptr(out_recon), _ptr(out_norm), _stream()),
           "dp_scalars")


# ------------------------------------------------------------------------------ output discriminator
def mask_vals_by_dl(crow, col, val, dl16, val_m):
    B = crow.numel() - 1
    _check(lib().cmmvae_mask_vals_by_dl(_ptr(crow), _ptr(col), _ptr(val), B, _ptr(dl16), dl16.stride(0), _ptr(val_m),
                                        _stream()), "mask_vals_by_dl")
    return val_m


def mask_vals_by_dl_rows(row_begin, row_end, col, val, dl16, val_m):
    _check(lib().cmmvae_mask_vals_by_dl_rows(_ptr(row_begin), _ptr(row_end), _ptr(col), _ptr(val), row_begin.numel(),
                                             _ptr(dl16), dl16.stride(0), _ptr(val_m), _stream()), "mask_vals_by_dl_rows")
    return val_m


def sigmoid_fwd(x, out32=None, out16=None):
    _check(lib().cmmvae_sigmoid_fwd(_ptr(x), _c.c_longlong(x.numel()), _ptr(out32), _ptr(out16), _stream()),
           "sigmoid_fwd")


def sigmoid_bwd(dout, out, dx=None, dx16=None):
    _check(lib().cmmvae_sigmoid_bwd(_ptr(dout), _ptr(out), _c.c_longlong(out.numel()), _ptr(dx), _ptr(dx16), _stream()),
           "sigmoid_bwd")


def bce_sigmoid(a, label: float, p_out, da, loss):
    _check(lib().cmmvae_bce_sigmoid(_ptr(a), a.numel(), _c.c_float(label), _ptr(p_out), _ptr(da), _ptr(loss),
                                    _stream()), "bce_sigmoid")
