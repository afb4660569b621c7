"""Execution of one FCBlock layer / the latent head on the sm_100a kernels, with hand-written
backward wrapped in ``torch.autograd.Function`` so the nn.Modules of ``mmvae_b200.modules`` compose
under autograd like the reference's.  (The fused ``training_step`` in ``mmvae_b200.engine`` drives the
same kernels directly, without autograd.)

Precision policy: ``"bf16"`` (default) = bf16 GEMM operands on the tcgen05 path, fp32 accumulation,
fp32 everywhere else; ``"fp32"`` = CUDA-core fp32 GEMMs (exact path, also used for shapes the TMA
path cannot address).  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import collections
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from . import torch_ops  # noqa: F401  (registers torch.ops.cmmvae.*)

C = torch.ops.cmmvae     # the module route goes through the custom-op layer (see torch_ops.py)

_PRECISION = "bf16"
BN_SCRATCH = {}


def set_precision(p: str) -> None:
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError(p)
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


# ---- reparameterisation noise ------------------------------------------------------------------
_noise_queue = collections.deque()


def inject_noise(eps: torch.Tensor) -> None:
    """Queue a tensor to be used as the next reparameterisation noise (parity runs inject the
    reference's eps)."""
    _noise_queue.append(eps)


def draw_noise(B: int, Z: int, device) -> torch.Tensor:
    if _noise_queue:
        eps = _noise_queue.popleft()
        assert tuple(eps.shape) == (B, Z), (eps.shape, (B, Z))
        return eps.to(device=device, dtype=torch.float32).contiguous()
    return torch.randn(B, Z, device=device, dtype=torch.float32)


# ---- dropout masks (parity runs) -----------------------------------------------------------------
_mask_queue = collections.deque()


def inject_dropout_masks(masks: dict) -> None:
    """Queue keep-masks (uint8 CUDA tensors, 1 = keep) for the next fused training step, keyed by the
    reference's sub-module path of the dropout layer (``experts.human.encoder.fc_layers.0.dr`` ...).  CPU and
    GPU generators cannot be bit-matched, so parity runs inject the masks the oracle used;
    without injection the kernels draw their own counter-based Bernoulli mask."""
    _mask_queue.append(masks)


def draw_dropout_masks():
    return _mask_queue.popleft() if _mask_queue else None


def dropout_tag(path: str):
    """engine layer tag of a reference dropout sub-module path (None if it is not part of the fused step)"""
    parts = path.split(".")
    try:
        j = int(parts[parts.index("fc_layers") + 1])
    except (ValueError, IndexError):
        return None
    if parts[0] == "experts":
        return f"{'enc' if parts[2] == 'encoder' else 'dec'}{j}"
    if parts[:2] == ["vae", "encoder"]:
        return f"venc{j}"
    if parts[:2] == ["vae", "decoder"]:
        return f"vdec{j}"
    return None


# ---- helpers -----------------------------------------------------------------------------------
def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: mmvae_b200 runs on CUDA (sm_100a) only -- got a {t.device} tensor; "
                           "there is no CPU fallback")


# bf16 shadows kept in sync by the fused optimizer (mmvae_b200.engine): id(param) -> bf16 tensor with
# the parameter's PHYSICAL layout.  Without a shadow, weights are cast on the fly.
SHADOWS = {}


def bf16_of(t: torch.Tensor) -> torch.Tensor:
    """fresh bf16 copy of an fp32 tensor (cast kernel)"""
    src = t.detach().contiguous()
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=t.device)
    C.cast_bf16(src, dst)
    return dst


def tc_ok(*dims) -> bool:
    return _PRECISION == "bf16" and all(d % 8 == 0 for d in dims)


def csr_parts(x: torch.Tensor):
    """(crow int32, col int32, val f32, nnz) of a torch.sparse_csr batch, bit-preserving."""
    parts = getattr(x, "_cmmvae_parts", None)
    if parts is None:
        crow = x.crow_indices().to(torch.int32)
        col = x.col_indices().to(torch.int32)
        val = x.values().to(torch.float32)
        parts = (crow.contiguous(), col.contiguous(), val.contiguous(), int(col.numel()))
        try:      # the same arrays on every call: a batch announced with prefetch_batch is recognised by its addresses
            x._cmmvae_parts = parts
        except Exception:
            pass
    return parts


def _scratch(H: int, device):
    return ops.bn_stats_scratch(H, device)


def dense_linear(x32, x16, W, bias, relu=False, W16=None):
    """y = act(x W^T + b) -> (y32, y16|None)."""
    B, K = x32.shape if x32 is not None else x16.shape
    N = W.shape[0]
    dev = W.device
    y32 = torch.empty(B, N, device=dev, dtype=torch.float32)
    if tc_ok(K, N) and W.is_contiguous():
        if x16 is None:
            x16 = bf16_of(x32)
        y16 = torch.empty(B, N, device=dev, dtype=torch.bfloat16)
        C.gemm(x16, False, W16 if W16 is not None else bf16_of(W), False, B, N, K, bias, relu, y32, y16, False)
        return y32, y16
    if x32 is None:
        x32 = x16.float()
    Wc = W if W.stride(1) == 1 else W.contiguous()
    C.gemm(x32.contiguous(), False, Wc, False, B, N, K, bias, relu, y32, None, False)
    return y32, None


def dense_linear_bwd(dY32, x32, W, need_dx=True, W16=None):
    """(dX, dW, db) of y = x W^T + b."""
    B, N = dY32.shape
    K = W.shape[1]
    dev = W.device
    dW = torch.empty(N, K, device=dev, dtype=torch.float32)
    db = torch.empty(N, device=dev, dtype=torch.float32)
    C.colsum(dY32, db)
    dX = torch.empty(B, K, device=dev, dtype=torch.float32) if need_dx else None
    if tc_ok(K, N) and W.is_contiguous():
        dY16, x16 = bf16_of(dY32), bf16_of(x32)
        C.gemm(dY16, True, x16, True, N, K, B, None, False, dW, None, False)
        if need_dx:
            C.gemm(dY16, False, W16 if W16 is not None else bf16_of(W), True, B, K, N, None, False, dX, None, False)
    else:
        Wc = W if W.stride(1) == 1 else W.contiguous()
        C.gemm(dY32, True, x32.contiguous(), True, N, K, B, None, False, dW, None, False)
        if need_dx:
            C.gemm(dY32, False, Wc, True, B, K, N, None, False, dX, None, False)
    return dX, dW, db


class _Spec:
    """static description of the fused part of a layer"""
    __slots__ = ("has_bn", "relu", "p", "training", "eps", "momentum", "csr", "G", "W16")


class _LayerFn(torch.autograd.Function):
    """lin (+BatchNorm) (+ReLU) (+Dropout) in the CUDA kernels; the input is either a dense fp32
    matrix or (through ``spec.csr``) a CSR batch."""

    @staticmethod
    def forward(ctx, x, W, b, gamma, beta, rm, rv, spec: _Spec):
        dev = W.device
        if spec.csr is not None:
            crow, col, val, nnz = spec.csr
            if _PRECISION == "bf16" and spec.W16 is not None:
                Wt = spec.W16  # physical [G,H] shadow
            else:
                Wt = W.t()
                if not Wt.is_contiguous():
                    Wt = Wt.contiguous()
                if _PRECISION == "bf16":
                    Wt = bf16_of(Wt)
            Y = C.csr_linear_fwd(crow, col, val, spec.G, Wt, b)
        else:
            Y, _ = dense_linear(x, None, W, b, relu=False, W16=spec.W16)
        B, H = Y.shape
        mean = rstd = None
        if spec.has_bn:
            mean = torch.empty(H, device=dev)
            rstd = torch.empty(H, device=dev)
            if spec.training:
                C.bn_stats(Y, spec.eps, spec.momentum, mean, rstd, rm, rv)
            else:
                mean = rm
                C.rstd_from_var(rv, spec.eps, rstd)
        p = spec.p if spec.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p > 0 else 0
        if spec.has_bn or spec.relu or p > 0:
            out = torch.empty_like(Y)
            C.bn_act_drop_fwd(Y, mean, rstd, gamma if spec.has_bn else None, beta if spec.has_bn else None,
                              bool(spec.relu), p, seed, out)
        else:
            out = Y
        ctx.spec, ctx.seed, ctx.p = spec, seed, p
        ctx.save_for_backward(x if spec.csr is None else None, W, gamma, Y, out, mean, rstd)
        return out

    @staticmethod
    def backward(ctx, dOut):
        spec = ctx.spec
        x, W, gamma, Y, out, mean, rstd = ctx.saved_tensors
        if spec.has_bn and not spec.training:
            raise RuntimeError("backward through eval-mode BatchNorm is not implemented")
        dOut = dOut.contiguous()
        B, H = dOut.shape
        dev = W.device
        dgamma = dbeta = None
        if spec.has_bn or spec.relu or ctx.p > 0:
            dY = torch.empty_like(dOut)
            if spec.has_bn:
                dgamma, dbeta = torch.empty(H, device=dev), torch.empty(H, device=dev)
            C.bn_act_drop_bwd(dOut, Y, out, mean, rstd, gamma if spec.has_bn else None, bool(spec.relu), ctx.p,
                              ctx.seed, dY, dgamma, dbeta)
        else:
            dY = dOut
        if spec.csr is not None:
            crow, col, val, nnz = spec.csr
            dWt = C.csr_linear_bwd_w(crow, col, val, spec.G, dY)
            db = torch.empty(H, device=dev)
            C.colsum(dY, db)
            return None, dWt.t(), db, dgamma, dbeta, None, None, None
        dX, dW, db = dense_linear_bwd(dY, x, W, need_dx=ctx.needs_input_grad[0], W16=spec.W16)
        return dX, dW, db, dgamma, dbeta, None, None, None


def run_layer(layer: nn.Sequential, x: torch.Tensor, training: bool, want_hidden: bool = False):
    """Execute one ``lin[/bn][/ln][/af][/dr]`` layer.  Returns (output, post-activation tensor or None).
    ``want_hidden``: the caller collects the post-activation tensor (reference: components.py:309-313,
    taken right after ``af`` and before dropout), so dropout is then not fused."""
    parts = dict(layer.named_children())
    lin: nn.Linear = parts["lin"]
    bn: Optional[nn.BatchNorm1d] = parts.get("bn")
    ln, af, dr = parts.get("ln"), parts.get("af"), parts.get("dr")
    _require_cuda(lin.weight, "FCBlock")
    fuse_tail = ln is None and (af is None or type(af) is nn.ReLU)
    fuse_drop = fuse_tail and dr is not None and not (want_hidden and af is not None)

    spec = _Spec()
    spec.has_bn = bn is not None
    spec.relu = int(fuse_tail and af is not None)
    spec.p = float(dr.p) if fuse_drop else 0.0
    spec.W16 = SHADOWS.get(id(lin.weight))
    spec.training = training
    spec.eps = bn.eps if bn is not None else 0.0
    spec.momentum = bn.momentum if bn is not None else 0.0
    spec.csr, spec.G = None, lin.in_features
    if x.layout == torch.sparse_csr:
        _require_cuda(x, "FCBlock")
        spec.csr = csr_parts(x)
        xin = None
    else:
        _require_cuda(x, "FCBlock")
        xin = x.to(torch.float32).contiguous()
    if bn is not None and training:
        bn.num_batches_tracked += 1
    out = _LayerFn.apply(xin, lin.weight, lin.bias, bn.weight if bn is not None else None,
                         bn.bias if bn is not None else None, bn.running_mean if bn is not None else None,
                         bn.running_var if bn is not None else None, spec)
    post_act = out if (fuse_tail and af is not None and spec.p == 0.0) else None
    if fuse_tail and dr is not None and not fuse_drop:
        out = dr(out)
    if not fuse_tail:
        # sublayers outside the hot path (LayerNorm, non-ReLU activations) run as stock torch modules
        for name in ("ln", "af", "dr"):
            m = parts.get(name)
            if m is not None:
                out = m(out)
                if name == "af":
                    post_act = out
    return out, post_act


class _LatentFn(torch.autograd.Function):
    """[mu | logvar] -> (mu, var, z) with z = mu + eps * sqrt(exp(logvar) + var_eps)."""

    @staticmethod
    def forward(ctx, ML, eps, var_eps):
        B, Z2 = ML.shape
        Z = Z2 // 2
        z, _ = C.reparam_kl_fwd(ML, eps, float(var_eps))
        mu = ML[:, :Z]
        var = torch.exp(ML[:, Z:]) + var_eps
        ctx.save_for_backward(ML, eps)
        ctx.var_eps = var_eps
        return mu, var, z

    @staticmethod
    def backward(ctx, dmu, dvar, dz):
        ML, eps = ctx.saved_tensors
        B, Z2 = ML.shape
        Z = Z2 // 2
        dML = C.reparam_kl_bwd(ML, eps, dz.contiguous() if dz is not None else None, float(ctx.var_eps), 0.0)
        if dmu is not None:
            dML[:, :Z] += dmu
        if dvar is not None:
            dML[:, Z:] += dvar * torch.exp(ML[:, Z:])
        return dML, None, None


def latent_head(q: torch.Tensor, mean_encoder: nn.Linear, var_encoder: nn.Linear, var_eps: float):
    """(mu, var, z): both heads as one GEMM (weights concatenated), then the fused latent kernel."""
    _require_cuda(q, "Encoder")
    W = torch.cat([mean_encoder.weight, var_encoder.weight], 0)
    b = torch.cat([mean_encoder.bias, var_encoder.bias], 0)
    spec = _Spec()
    spec.has_bn, spec.relu, spec.p, spec.training, spec.eps, spec.momentum = False, 0, 0.0, False, 0.0, 0.0
    spec.csr, spec.G, spec.W16 = None, W.shape[1], None
    ML = _LayerFn.apply(q.to(torch.float32).contiguous(), W, b, None, None, None, None, spec)
    eps = draw_noise(q.shape[0], mean_encoder.out_features, q.device)
    return _LatentFn.apply(ML, eps, float(var_eps))
