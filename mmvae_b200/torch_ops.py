"""``torch.ops.cmmvae.*`` -- the C-ABI entry points of libcmmvae_b200.so registered as torch custom operators
("thin C-ABI torch custom-op layer").  Each operator is a few lines: it hands the tensors' device
pointers, sizes and the current CUDA stream to the C function (through ``mmvae_b200.ops``) and returns; only a CUDA
implementation is registered, so CPU tensors fail in the dispatcher -- there is no CPU kernel to fall back to.

The module route (``mmvae_b200.layers``: the autograd Functions under FCBlock / Encoder, i.e. what
``CMMVAE.forward`` -> ``vae.elbo`` -> ``loss.backward()`` executes) calls these operators.  The fused
``training_step`` (``mmvae_b200.engine``) binds the same C ABI directly: a step is ~70 launches of a few microseconds
each, and the dispatcher's per-call cost would be a fifth of the step.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops

_lib = torch.library.Library("cmmvae", "DEF")


def _def(schema: str, fn):
    name = schema.split("(")[0]
    _lib.define(schema)
    _lib.impl(name, fn, "CUDA")


# ---- expert-encoder first layer on a CSR batch (K1 / K1b) --------------------------------------------------------
def _csr_linear_fwd(crow, col, val, G: int, Wt, bias):
    return ops.csr_linear_fwd(crow, col, val, G, Wt, bias)


def _csr_linear_bwd_w(crow, col, val, G: int, dY):
    B = crow.numel() - 1
    cptr, ridx, cval = ops.csr_transpose(crow, col, val, G, int(col.numel()))
    out = torch.empty(G, dY.shape[1], device=dY.device, dtype=torch.float32)
    return ops.csr_linear_bwd_w(cptr, ridx, cval, B, G, dY.contiguous(), out)


_def("csr_linear_fwd(Tensor crow, Tensor col, Tensor val, int G, Tensor Wt, Tensor bias) -> Tensor", _csr_linear_fwd)
_def("csr_linear_bwd_w(Tensor crow, Tensor col, Tensor val, int G, Tensor dY) -> Tensor", _csr_linear_bwd_w)


# ---- dense GEMM (K4 / K12): tcgen05 bf16, tcgen05 tf32 or CUDA-core fp32 by operand dtype / flag --------------------
def _gemm(A, transA: bool, Bm, transB: bool, M: int, N: int, K: int, bias: Optional[torch.Tensor], relu: bool,
          C32: Optional[torch.Tensor], C16: Optional[torch.Tensor], tf32: bool) -> None:
    if tf32:
        ops.gemm(A, int(transA), Bm, int(transB), M, N, K, bias=bias, relu=relu, C32=C32, C16=C16, tf32=True)
    else:
        ops.gemm(A, int(transA), Bm, int(transB), M, N, K, bias=bias, relu=relu, C32=C32, C16=C16,
                 use_tc=(A.dtype == torch.bfloat16))


_def("gemm(Tensor A, bool transA, Tensor B, bool transB, int M, int N, int K, Tensor? bias, bool relu, "
     "Tensor(a!)? C32, Tensor(b!)? C16, bool tf32) -> ()", _gemm)
_def("colsum(Tensor X, Tensor(a!) out) -> ()", lambda X, out: (ops.colsum(X, out), None)[1])
_def("cast_bf16(Tensor src, Tensor(a!) dst) -> ()", lambda src, dst: (ops.cast_bf16(src, dst), None)[1])


# ---- BatchNorm + ReLU + Dropout (K2 / K3) ---------------------------------------------------------------------------
def _bn_stats(Y, eps: float, momentum: float, mean, rstd, running_mean, running_var) -> None:
    ops.bn_stats(Y, eps, momentum, mean, rstd, running_mean, running_var)


def _bn_act_drop_fwd(Y, mean, rstd, gamma, beta, relu: bool, p_drop: float, seed: int, out) -> None:
    ops.bn_act_drop_fwd(Y, mean, rstd, gamma, beta, relu, p_drop, seed, None, out, None)


def _bn_act_drop_bwd(dOut, Y, out, mean, rstd, gamma, relu: bool, p_drop: float, seed: int, dY, dgamma, dbeta) -> None:
    ops.bn_act_drop_bwd(dOut, Y, out, mean, rstd, gamma, relu, p_drop, seed, None, dY, None, dgamma, dbeta, None)


_def("bn_stats(Tensor Y, float eps, float momentum, Tensor(a!) mean, Tensor(b!) rstd, Tensor(c!)? running_mean, "
     "Tensor(d!)? running_var) -> ()", _bn_stats)
_def("rstd_from_var(Tensor var, float eps, Tensor(a!) rstd) -> ()",
     lambda var, eps, rstd: (ops.rstd_from_var(var, eps, rstd), None)[1])
_def("bn_act_drop_fwd(Tensor Y, Tensor? mean, Tensor? rstd, Tensor? gamma, Tensor? beta, bool relu, float p_drop, "
     "int seed, Tensor(a!) out) -> ()", _bn_act_drop_fwd)
_def("bn_act_drop_bwd(Tensor dOut, Tensor Y, Tensor out, Tensor? mean, Tensor? rstd, Tensor? gamma, bool relu, "
     "float p_drop, int seed, Tensor(a!) dY, Tensor(b!)? dgamma, Tensor(c!)? dbeta) -> ()", _bn_act_drop_bwd)


# ---- reparameterisation + KL (K8 / K9) ---------------------------------------------------------------------------------
def _reparam_kl_fwd(ML, eps, var_eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    B, Z2 = ML.shape
    z = torch.empty(B, Z2 // 2, device=ML.device)
    sums = torch.empty(3, dtype=torch.float64, device=ML.device)
    ops.reparam_kl_fwd(ML, eps, Z2 // 2, var_eps, z, None, sums)
    return z, sums


def _reparam_kl_bwd(ML, eps, dz: Optional[torch.Tensor], var_eps: float, kl_scale: float):
    dML = torch.empty_like(ML)
    ops.reparam_kl_bwd(ML, eps, dz, ML.shape[1] // 2, var_eps, kl_scale, dML, None)
    return dML


_def("reparam_kl_fwd(Tensor ML, Tensor eps, float var_eps) -> (Tensor, Tensor)", _reparam_kl_fwd)
_def("reparam_kl_bwd(Tensor ML, Tensor eps, Tensor? dz, float var_eps, float kl_scale) -> Tensor", _reparam_kl_bwd)


# ---- fused expert-decoder output + ReLU + sum-MSE against the CSR batch (K5-K7) -----------------------------------------
def _decoder_mse_fused(h16, Wout16, bout, G: int, crow, col, val) -> Tuple[torch.Tensor, torch.Tensor]:
    B = h16.shape[0]
    dl = torch.zeros(B, (G + 63) // 64 * 64, dtype=torch.bfloat16, device=h16.device)
    loss = torch.empty(1, dtype=torch.float64, device=h16.device)
    ops.decoder_mse_fused(h16, Wout16, bout, G, crow, col, val, dl, loss)
    return loss, dl


_def("decoder_mse_fused(Tensor h16, Tensor Wout16, Tensor bout, int G, Tensor crow, Tensor col, Tensor val) "
     "-> (Tensor, Tensor)", _decoder_mse_fused)


# ---- grad-norm / clip / Adam (K13-K15) ------------------------------------------------------------------------------------
def _clip_adam(p, g, m, v, p16: Optional[torch.Tensor], norm_sq, max_norm: float, grad_scale: float, lr: float,
               beta1: float, beta2: float, eps: float, wd: float, step: int) -> None:
    ops.clip_adam(p, g, m, v, p16, norm_sq, max_norm, grad_scale, lr, beta1, beta2, eps, wd, step)


_def("sumsq(Tensor g, Tensor(a!) norm_sq) -> ()", lambda g, ns: (ops.sumsq(g, ns), None)[1])
_def("clip_adam_(Tensor(a!) p, Tensor g, Tensor(b!) m, Tensor(c!) v, Tensor(d!)? p16, Tensor norm_sq, float max_norm, "
     "float grad_scale, float lr, float beta1, float beta2, float eps, float wd, int step) -> ()", _clip_adam)

OPS = ("csr_linear_fwd", "csr_linear_bwd_w", "gemm", "colsum", "cast_bf16", "bn_stats", "rstd_from_var",
       "bn_act_drop_fwd", "bn_act_drop_bwd", "reparam_kl_fwd", "reparam_kl_bwd", "decoder_mse_fused", "sumsq",
       "clip_adam_")
