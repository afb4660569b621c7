"""CPU oracle for the CMMVAE training step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  Nothing under ``mmvae_b200/`` imports it.

It is a functional restatement (plain torch CPU fp32 tensors + explicit formulas; autograd is
used only to differentiate the restated forward, exactly as the reference does through
``manual_backward``) of the one hot path this repo accelerates:

    reference ``CMMVAEModel.training_step``      src/cmmvae/models/cmmvae_model.py:138-217
      expert encode  (FCBlock)                   src/cmmvae/modules/base/components.py:250-314,851-853
      VAE encoder + reparameterisation           src/cmmvae/modules/base/components.py:768-809
      prior / decoder                            src/cmmvae/modules/vae.py:80-102
      expert decode                              src/cmmvae/modules/cmmvae.py:85-113
      ELBO (KL + sum-MSE)                        src/cmmvae/modules/vae.py:104-152
      GRL adversary, discriminator-then-generator src/cmmvae/models/cmmvae_model.py:59-136
      gradient reversal                          src/cmmvae/modules/base/components.py:879-899
      grad-norm logging                          src/cmmvae/models/base_model.py:111-123
      clip-by-norm + Adam(lr 5e-3, wd 1e-6)      src/cmmvae/models/cmmvae_model.py:203-213,306-319
      KL annealing                               src/cmmvae/modules/base/annealing_fn.py:1-42

The arithmetic of the reference lives in un-vendored third-party code: ``torch`` (unpinned in
the reference's setup.cfg:28; 2.11.0+cu128 in this image) and ``lightning`` (unpinned,
setup.cfg:32; absent from this image).  Their published algorithms are restated here:
``nn.Linear`` (y = x W^T + b), ``nn.BatchNorm1d(momentum=0.01, eps=1e-3)`` training statistics
and running-stat update, ``Normal.rsample`` (loc + eps*scale), ``_kl_normal_normal``,
``F.mse_loss(reduction='sum')``, ``CrossEntropyLoss(reduction='sum')``,
``clip_grad_norm_`` (coef = min(1, max_norm/(norm+1e-6))), ``torch.optim.Adam`` (L2 weight
decay, bias-corrected, eps outside the sqrt).

PINNED: the reference's own tests hold no numeric vectors for this path, so the
oracle is pinned against outputs of the *reference itself executed in the build container*
(``tests/golden/make_golden.py`` imports ``/root/reference/src/cmmvae`` unmodified behind a
25-line lightning stand-in and stores inputs/outputs in ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this file against every stored vector.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# model description (mirrors the YAML trees of configs/model/*.yaml, names only)
# --------------------------------------------------------------------------------------------


@dataclass
class BlockSpec:
    """One FCBlock (components.py:193-314): per-layer options already broadcast."""

    layers: List[int]
    use_batch_norm: List[bool]
    use_layer_norm: List[bool]
    activation: List[Optional[str]]  # "relu" | None
    dropout_rate: List[float]
    return_hidden: List[bool]

    @staticmethod
    def make(layers, bn=False, ln=False, act="relu", dropout=0.0, return_hidden=False):
        layers = list(layers)
        if len(layers) == 1:  # components.py:121-122
            layers = layers * 2
        n = len(layers) - 1

        def bc(v):
            return list(v) if isinstance(v, (list, tuple)) else [v] * n

        return BlockSpec(layers, bc(bn), bc(ln), bc(act), bc(dropout), bc(return_hidden))

    @property
    def n_layers(self):
        return len(self.layers) - 1


@dataclass
class AdversarySpec:
    encoder: BlockSpec
    conditions: Dict[str, int]  # condition -> number of classes (rows of the human csv)


@dataclass
class CondSpec:
    """ConditionalLayers (components.py:467-631): per batch key one ConditionalLayer (components.py:317-413) holding
    an FCBlock per metadata value -- here the one-layer block the shipped topology uses (human_only.yaml:61-68:
    ``layers: [latent]`` -> Linear(latent, latent), components.py:115-116, + LayerNorm without affine, :277-278)."""
    names: List[str]                          # conditionals in constructor order, "species" included
    species_specific: List[str] = field(default_factory=list)   # batch keys whose layers sit under ModuleDict[species]
    parallel: bool = True                     # selection_order == ["parallel"]: outputs concatenated (:617-631)
    layer_norm: bool = True
    relu: bool = False


@dataclass
class ModelSpec:
    experts: Dict[str, Dict[str, BlockSpec]]  # id -> {"encoder": BlockSpec, "decoder": BlockSpec}
    vae_encoder: BlockSpec
    vae_decoder: BlockSpec
    latent_dim: int
    hidden_z: bool = False
    var_eps: float = 1e-4
    adversarials: List[AdversarySpec] = field(default_factory=list)
    adv_weight: float = 1.0
    clip: Optional[float] = 10.0  # vae / expert / adversarial clip-by-norm value (None = off)
    lr: float = 5e-3
    weight_decay: float = 1e-6
    betas: Sequence[float] = (0.9, 0.999)
    adam_eps: float = 1e-8
    conditionals: Optional["CondSpec"] = None   # CLVAE.conditionals (clvae.py:42-87)


# --------------------------------------------------------------------------------------------
# forward building blocks
# --------------------------------------------------------------------------------------------


def csr_to_dense(crow, col, val, n_cols):
    """x.to_dense() of cmmvae_model.py:162-163 on raw CSR arrays (int32 crow/col, fp32 val)."""
    crow = np.asarray(crow, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=np.float32)
    n_rows = crow.shape[0] - 1
    out = np.zeros((n_rows, n_cols), dtype=np.float32)
    rows = np.repeat(np.arange(n_rows), np.diff(crow))
    np.add.at(out, (rows, col), val)  # duplicate-free input => plain assignment semantics
    return torch.from_numpy(out)


FAST_CSR = False  # set by bench.py's CPU legs: use torch's sparse-CSR addmm (what the reference executes)


def csr_linear(crow, col, val, weight, bias):
    """First expert-encoder layer on a CSR batch: Y = X W^T + b (components.py:276,306).

    Restated as an explicit per-nonzero accumulation so that CSR indexing is exercised by the
    oracle itself (not by torch's sparse addmm).  With FAST_CSR the same product goes through
    ``F.linear`` on a ``torch.sparse_csr`` tensor, i.e. exactly the ATen call the reference makes
    (used for full-size CPU timing; tests check both forms agree)."""
    if FAST_CSR:
        x = torch.sparse_csr_tensor(torch.as_tensor(np.asarray(crow)), torch.as_tensor(np.asarray(col)),
                                    torch.as_tensor(np.asarray(val, dtype=np.float32)),
                                    size=(len(crow) - 1, weight.shape[1]))
        return F.linear(x, weight, bias)
    crow = torch.as_tensor(np.asarray(crow, dtype=np.int64))
    col = torch.as_tensor(np.asarray(col, dtype=np.int64))
    val = torch.as_tensor(np.asarray(val, dtype=np.float32))
    n_rows = crow.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n_rows), crow[1:] - crow[:-1])
    contrib = weight.t()[col] * val[:, None]  # [nnz, H]
    y = torch.zeros(n_rows, weight.shape[0], dtype=weight.dtype)
    y = y.index_add(0, rows, contrib)
    return y + bias


def batch_norm_train(y, gamma, beta, running_mean, running_var, momentum=0.01, eps=1e-3):
    """nn.BatchNorm1d training mode (components.py:279): biased variance for normalisation,
    unbiased for the running update.  Returns (out, new_running_mean, new_running_var)."""
    n = y.shape[0]
    mean = y.mean(0)
    var = ((y - mean) ** 2).mean(0)
    out = (y - mean) / torch.sqrt(var + eps) * gamma + beta
    with torch.no_grad():
        unbiased = var * (n / max(n - 1, 1))
        new_rm = (1 - momentum) * running_mean + momentum * mean
        new_rv = (1 - momentum) * running_var + momentum * unbiased
    return out, new_rm.detach(), new_rv.detach()


def batch_norm_eval(y, gamma, beta, running_mean, running_var, eps=1e-3):
    return (y - running_mean) / torch.sqrt(running_var + eps) * gamma + beta


def layer_norm(y, eps=1e-5):
    """nn.LayerNorm(n_out, elementwise_affine=False) (components.py:281)."""
    mean = y.mean(-1, keepdim=True)
    var = ((y - mean) ** 2).mean(-1, keepdim=True)
    return (y - mean) / torch.sqrt(var + eps)


def fcblock(x, spec: BlockSpec, prefix: str, P: Dict[str, torch.Tensor], training: bool,
            new_buffers: Dict[str, torch.Tensor], dropout_masks: Optional[Dict[str, torch.Tensor]] = None,
            csr=None):
    """FCBlock.forward (components.py:292-314).  ``csr`` = (crow, col, val) when the block input
    is the sparse batch (first layer of an expert encoder); then ``x`` is ignored for layer 0.
    Returns (out, hidden_list)."""
    hidden = []
    for i in range(spec.n_layers):
        lp = f"{prefix}.fc_layers.{i}"
        w, b = P[f"{lp}.lin.weight"], P[f"{lp}.lin.bias"]
        if i == 0 and csr is not None:
            x = csr_linear(*csr, w, b)
        else:
            x = x @ w.t() + b
        if spec.use_batch_norm[i]:
            g, be = P[f"{lp}.bn.weight"], P[f"{lp}.bn.bias"]
            rm, rv = P[f"{lp}.bn.running_mean"], P[f"{lp}.bn.running_var"]
            if training:
                x, nrm, nrv = batch_norm_train(x, g, be, rm, rv)
                new_buffers[f"{lp}.bn.running_mean"] = nrm
                new_buffers[f"{lp}.bn.running_var"] = nrv
                new_buffers[f"{lp}.bn.num_batches_tracked"] = P[f"{lp}.bn.num_batches_tracked"] + 1
            else:
                x = batch_norm_eval(x, g, be, rm, rv)
        if spec.use_layer_norm[i]:
            x = layer_norm(x)
        if spec.activation[i] == "relu":
            x = torch.relu(x)
            if spec.return_hidden[i]:
                hidden.append(x)
        elif spec.activation[i] is not None:
            raise NotImplementedError(spec.activation[i])
        p = spec.dropout_rate[i]
        if p > 0 and training:
            if dropout_masks is None or f"{lp}.dr" not in dropout_masks:
                raise ValueError(f"oracle needs an injected dropout mask for {lp}.dr")
            x = x * dropout_masks[f"{lp}.dr"] / (1.0 - p)
    return x, hidden


def latent_head(q, P, prefix, eps_noise, var_eps):
    """Encoder.forward after the FC block (components.py:790-801)."""
    mu = q @ P[f"{prefix}.mean_encoder.weight"].t() + P[f"{prefix}.mean_encoder.bias"]
    lv = q @ P[f"{prefix}.var_encoder.weight"].t() + P[f"{prefix}.var_encoder.bias"]
    var = torch.exp(lv) + var_eps
    sigma = torch.sqrt(var)
    z = mu + eps_noise * sigma  # Normal.rsample
    return mu, sigma, z


def kl_normal_std(mu, sigma):
    """kl_divergence(Normal(mu,sigma), Normal(0,1)).sum(-1).mean()  (vae.py:136-138);
    torch's _kl_normal_normal: 0.5*(var_ratio + t1 - 1 - log(var_ratio))."""
    var_ratio = sigma ** 2
    t1 = mu ** 2
    kl = 0.5 * (var_ratio + t1 - 1.0 - torch.log(var_ratio))
    return kl.sum(-1).mean()


def cross_entropy_sum(logits, labels):
    """nn.CrossEntropyLoss(reduction='sum') (cmmvae_model.py:54)."""
    lse = torch.logsumexp(logits, dim=1)
    picked = logits.gather(1, labels.view(-1, 1)).squeeze(1)
    return (lse - picked).sum()


class _GRL(torch.autograd.Function):
    """components.py:879-899: identity forward, -alpha * grad backward."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg() * ctx.alpha, None


# --------------------------------------------------------------------------------------------
# optimiser pieces
# --------------------------------------------------------------------------------------------


def grad_norm(grads: Sequence[torch.Tensor]) -> float:
    """log_gradient_norms (base_model.py:111-123): sqrt(sum_p ||g_p||_2^2), python floats."""
    total = 0.0
    for g in grads:
        total += float(torch.linalg.vector_norm(g.detach(), 2)) ** 2
    return total ** 0.5


def clip_by_norm(grads: Sequence[torch.Tensor], max_norm: float):
    """torch.nn.utils.clip_grad_norm_ as called by Lightning's clip_gradients(..., 'norm')."""
    norms = torch.stack([torch.linalg.vector_norm(g, 2) for g in grads])
    total = torch.linalg.vector_norm(norms, 2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], float(total)


def adam_update(p, g, m, v, step, lr, wd, betas, eps):
    """torch.optim.Adam single-tensor update (non-decoupled weight decay)."""
    b1, b2 = betas
    g = g + wd * p
    m = m + (g - m) * (1 - b1)  # lerp_
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    step_size = lr / bc1
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - step_size * (m / denom)
    return p, m, v


class OptState:
    """Per-optimizer-group Adam state: name -> (m, v, step).  Groups mirror
    configure_optimizers (cmmvae_model.py:299-324): experts/<id>, vae, adversarials/<i>."""

    def __init__(self):
        self.m: Dict[str, torch.Tensor] = {}
        self.v: Dict[str, torch.Tensor] = {}
        self.step: Dict[str, int] = {}


def group_of(name: str) -> str:
    parts = name.split(".")
    if parts[0] == "experts":
        return f"experts/{parts[1]}"
    if parts[0] == "adversarials":
        return f"adversarials/{int(parts[1]) + 1}"  # optimizers are numbered from 1 (cmmvae_model.py:316)
    return "vae"


def is_param(name: str) -> bool:
    return not (name.endswith("running_mean") or name.endswith("running_var") or name.endswith("num_batches_tracked"))


def apply_group_step(P, grads: Dict[str, torch.Tensor], opt: OptState, spec: ModelSpec, clip: Optional[float]):
    """clip (optional) + Adam over one optimizer group; parameters without a gradient are skipped
    entirely, as torch Adam does.  Returns the pre-clip norm."""
    names = [n for n in grads if grads[n] is not None]
    gl = [grads[n] for n in names]
    norm = grad_norm(gl)
    if clip is not None:
        gl, _ = clip_by_norm(gl, clip)
    for n, g in zip(names, gl):
        t = opt.step.get(n, 0) + 1
        m = opt.m.get(n, torch.zeros_like(P[n]))
        v = opt.v.get(n, torch.zeros_like(P[n]))
        p, m, v = adam_update(P[n].detach(), g, m, v, t, spec.lr, spec.weight_decay, spec.betas, spec.adam_eps)
        P[n] = p
        opt.m[n], opt.v[n], opt.step[n] = m, v, t
    return norm


# --------------------------------------------------------------------------------------------
# the step
# --------------------------------------------------------------------------------------------


def conditional_layers(cs: CondSpec, P, x, keys: Dict[str, List[str]], species: str, order: List[str]):
    """ConditionalLayers.forward (components.py:581-631) in the given ``order`` (the reference draws it with
    ``random.sample`` per call when the selection is parallel or unordered, :598-600) with ConditionalLayer.forward
    (components.py:369-413: rows grouped by their metadata value, each group through that value's block, results
    written back in the original row order).  ``keys``: batch key -> formatted condition key of every row."""
    def block(h, prefix):
        h = h @ P[f"{prefix}.fc_layers.0.lin.weight"].t() + P[f"{prefix}.fc_layers.0.lin.bias"]
        if cs.layer_norm:
            h = torch.nn.functional.layer_norm(h, h.shape[-1:])
        return torch.relu(h) if cs.relu else h

    branches = []
    for bk in order:
        if bk == "species":
            y = block(x, f"vae.conditionals.layers.species.{species}")
        else:
            base = f"vae.conditionals.layers.{bk}" + (f".{species}" if bk in cs.species_specific else "")
            groups: Dict[str, List[int]] = {}
            for i, k in enumerate(keys[bk]):
                groups.setdefault(k, []).append(i)
            y = torch.zeros(x.shape[0], P[f"{base}.conditions.{keys[bk][0]}.fc_layers.0.lin.weight"].shape[0])
            for k, rows in groups.items():
                idx = torch.tensor(rows)
                y = y.index_copy(0, idx, block(x.index_select(0, idx), f"{base}.conditions.{k}"))
        if cs.parallel:
            branches.append(y)
        else:
            x = y
    return torch.cat(branches, dim=1) if branches else x


def forward(spec: ModelSpec, P, expert_id, csr, n_genes, eps_noise, training, new_buffers,
            dropout_masks=None, cond=None):
    """CMMVAE.forward (cmmvae.py:85-113) + BaseVAE.forward (vae.py:98-102).  ``cond`` (with ``spec.conditionals``):
    dict(keys=..., order=...) for CLVAE.after_reparameterize (clvae.py:89-111)."""
    ex = spec.experts[expert_id]
    s, _ = fcblock(None, ex["encoder"], f"experts.{expert_id}.encoder", P, training, new_buffers,
                   dropout_masks, csr=csr)
    q, hidden = fcblock(s, spec.vae_encoder, "vae.encoder.fc", P, training, new_buffers, dropout_masks)
    mu, sigma, z = latent_head(q, P, "vae.encoder", eps_noise, spec.var_eps)
    if spec.hidden_z:
        hidden = hidden + [z]
    zc = z
    if spec.conditionals is not None:
        zc = conditional_layers(spec.conditionals, P, z, cond["keys"], expert_id, cond["order"])
    d, _ = fcblock(zc, spec.vae_decoder, "vae.decoder", P, training, new_buffers, dropout_masks)
    xhat, _ = fcblock(d, ex["decoder"], f"experts.{expert_id}.decoder", P, training, new_buffers, dropout_masks)
    return mu, sigma, zc, xhat, hidden      # (vae.py:98-102 returns z AFTER after_reparameterize)


def elbo(mu, sigma, x_dense, xhat, kl_weight):
    """BaseVAE.elbo (vae.py:136-152)."""
    kl = kl_normal_std(mu, sigma)
    recon = ((xhat - x_dense) ** 2).sum()
    return {"loss": recon + kl_weight * kl, "recon_loss": recon, "kl_loss": kl, "kl_weight": kl_weight}


def adversary_losses(spec: ModelSpec, P, hidden, labels, detach, new_buffers):
    """CMMVAEModel.grf (cmmvae_model.py:59-101).  Returns (list of summed losses, per-condition)."""
    out, per_cond = [], []
    for i, (h, adv) in enumerate(zip(hidden, spec.adversarials)):
        h = h.detach() if detach else _GRL.apply(h, 1)
        a, _ = fcblock(h, adv.encoder, f"adversarials.{i}.encoder", P, True, new_buffers)
        head_losses = {}
        for cond in labels:
            lp = f"adversarials.{i}.heads.{cond}.fc_layers.0.lin"
            logits = a @ P[f"{lp}.weight"].t() + P[f"{lp}.bias"]
            head_losses[cond] = cross_entropy_sum(logits, labels[cond])
        out.append(torch.stack(list(head_losses.values())).sum())
        per_cond.append(head_losses)
    return out, per_cond


def train_step(spec: ModelSpec, P: Dict[str, torch.Tensor], opt: Dict[str, OptState], expert_id: str,
               crow, col, val, eps_noise, kl_weight: float, labels: Optional[Dict[str, torch.Tensor]] = None,
               dropout_masks=None, return_grads: bool = True, cond=None):
    """One CMMVAEModel.training_step (cmmvae_model.py:138-217) on state ``P`` (updated in place).

    Returns a dict: logs (exact reference keys, untagged), grads (pre-clip), z.
    """
    n_genes = spec.experts[expert_id]["encoder"].layers[0]
    csr = (crow, col, val)
    x_dense = csr_to_dense(crow, col, val, n_genes)
    logs: Dict[str, float] = {}
    new_buffers: Dict[str, torch.Tensor] = {}

    main_names = [n for n in P if is_param(n) and group_of(n) in ("vae", f"experts/{expert_id}")]
    adv_names = [n for n in P if is_param(n) and n.startswith("adversarials.")]
    for n in main_names + adv_names:
        P[n] = P[n].detach().requires_grad_(True)

    mu, sigma, z, xhat, hidden = forward(spec, P, expert_id, csr, n_genes, eps_noise, True, new_buffers,
                                         dropout_masks, cond=cond)
    ld = elbo(mu, sigma, x_dense, xhat, kl_weight)
    logs["recon_loss"] = float(ld["recon_loss"])
    logs["kl_loss"] = float(ld["kl_loss"])
    logs["kl_weight"] = float(kl_weight)
    logs["Mean"] = float(mu.mean())
    logs["Variance"] = float((sigma ** 2).mean())
    total = ld["loss"]

    grads_out: Dict[str, torch.Tensor] = {}
    if spec.adversarials:
        assert labels is not None
        # (i) discriminator pass on detached hidden reps, one backward/step per adversary
        d_losses, d_per = adversary_losses(spec, P, hidden, labels, True, new_buffers)
        for i, dl in enumerate(d_losses):
            for cond, v_ in d_per[i].items():
                logs[f"discriminator_{i + 1}/adversarial_loss/{cond}"] = float(v_)
            logs[f"discriminator_{i + 1}/adversarial_loss/summed"] = float(dl)
        for i, dl in enumerate(d_losses):
            names = [n for n in adv_names if n.startswith(f"adversarials.{i}.")]
            gs = torch.autograd.grad(dl, [P[n] for n in names], allow_unused=True)
            gd = {n: g for n, g in zip(names, gs)}
            if return_grads:
                for n, g in gd.items():
                    grads_out[f"discriminator/{n}"] = g
            logs[f"grad_norms/discriminator_{i + 1}"] = apply_group_step(
                P, gd, opt.setdefault(f"adversarials/{i + 1}", OptState()), spec, spec.clip)
            for n in names:
                P[n] = P[n].detach().requires_grad_(True)
        # (ii) generator pass through GRL with the *updated* adversaries
        g_losses, g_per = adversary_losses(spec, P, hidden, labels, False, new_buffers)
        for i, gl in enumerate(g_losses):
            for cond, v_ in g_per[i].items():
                logs[f"generator_{i + 1}/adversarial_loss/{cond}"] = float(v_)
            logs[f"generator_{i + 1}/adversarial_loss/summed"] = float(gl)
            total = total + gl * spec.adv_weight
    logs["loss"] = float(total)

    all_names = main_names + (adv_names if spec.adversarials else [])
    gs = torch.autograd.grad(total, [P[n] for n in all_names], allow_unused=True)
    g_all = {n: g for n, g in zip(all_names, gs)}
    if return_grads:
        for n in main_names:
            grads_out[n] = g_all[n]
    vae_g = {n: g_all[n] for n in main_names if group_of(n) == "vae"}
    exp_g = {n: g_all[n] for n in main_names if group_of(n) != "vae"}
    # generator_i norms: adversary grads of the main backward, logged and never applied
    for i in range(len(spec.adversarials)):
        gl = [g_all[n] for n in adv_names if n.startswith(f"adversarials.{i}.") and g_all[n] is not None]
        logs[f"grad_norms/generator_{i + 1}"] = grad_norm(gl)
    logs["grad_norms/vae"] = apply_group_step(P, vae_g, opt.setdefault("vae", OptState()), spec, spec.clip)
    logs[f"grad_norms/expert_{expert_id}"] = apply_group_step(
        P, exp_g, opt.setdefault(f"experts/{expert_id}", OptState()), spec, spec.clip)
    for n, b in new_buffers.items():
        P[n] = b
    for n in list(P):
        P[n] = P[n].detach()
    return {"logs": logs, "grads": grads_out, "z": z.detach(), "mu": mu.detach(), "sigma": sigma.detach()}


def eval_step(spec: ModelSpec, P, expert_id, crow, col, val, eps_noise, kl_weight, cond=None):
    """CMMVAEModel.validation_step (cmmvae_model.py:219-245): eval-mode forward + ELBO."""
    n_genes = spec.experts[expert_id]["encoder"].layers[0]
    x_dense = csr_to_dense(crow, col, val, n_genes)
    with torch.no_grad():
        mu, sigma, z, xhat, hidden = forward(spec, P, expert_id, (crow, col, val), n_genes, eps_noise, False, {},
                                             cond=cond)
        ld = elbo(mu, sigma, x_dense, xhat, kl_weight)
    return {"logs": {k: float(v) for k, v in ld.items()}, "z": z, "xhat": xhat, "mu": mu, "sigma": sigma}


# --------------------------------------------------------------------------------------------
# output discriminator (BASELINE config 4) -- restated from runners/meta_discriminators.py
# --------------------------------------------------------------------------------------------


def output_discriminator_step(Pd: Dict[str, torch.Tensor], opt: OptState, xhat: torch.Tensor, label: float,
                              lr: float = 1e-3):
    """One training step of a species' output discriminator on a (detached) reconstruction:
    nn.Sequential(Linear(G,128), Sigmoid, Linear(128,64), Sigmoid, Linear(64,1), Sigmoid)
    (meta_discriminators.py:33-49), ``binary_cross_entropy(out, label, reduction='mean')`` (131-134),
    ``torch.optim.Adam(lr=1e-3)`` with torch's defaults (103).  ``Pd``: state_dict of the Sequential
    ('0.weight', '0.bias', '2.weight', ...), updated in place.  Returns loss and gradients."""
    names = ["0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias"]
    for n in names:
        Pd[n] = Pd[n].detach().requires_grad_(True)
    a = xhat.detach()
    for i in (0, 2, 4):
        a = torch.sigmoid(a @ Pd[f"{i}.weight"].t() + Pd[f"{i}.bias"])
    truth = torch.full_like(a, float(label))
    loss = F.binary_cross_entropy(a, truth, reduction="mean")
    grads = dict(zip(names, torch.autograd.grad(loss, [Pd[n] for n in names])))
    for n in names:
        t = opt.step.get(n, 0) + 1
        m = opt.m.get(n, torch.zeros_like(Pd[n]))
        v = opt.v.get(n, torch.zeros_like(Pd[n]))
        p, m, v = adam_update(Pd[n].detach(), grads[n], m, v, t, lr, 0.0, (0.9, 0.999), 1e-8)
        Pd[n], opt.m[n], opt.v[n], opt.step[n] = p, m, v, t
    return {"loss": float(loss), "grads": grads}


# --------------------------------------------------------------------------------------------
# synthetic inputs -- shared by tests and bench so both sides see the same bytes
# --------------------------------------------------------------------------------------------


def synth_csr(n_cells: int, n_genes: int, density: float, seed: int):
    """Row-sorted, duplicate-free int32 CSR; values log1p(1e4*c/sum c), c = 1+Poisson(1.5)
    (scripts/data-preprocessing/data_processing_functions.py:24-31)."""
    rng = np.random.default_rng(seed)
    per = max(1, int(round(density * n_genes)))
    crow = np.zeros(n_cells + 1, dtype=np.int32)
    cols, vals = [], []
    for i in range(n_cells):
        c = np.sort(rng.choice(n_genes, size=per, replace=False)).astype(np.int32)
        counts = 1.0 + rng.poisson(1.5, size=per)
        v = np.log1p(1e4 * counts / counts.sum()).astype(np.float32)
        cols.append(c)
        vals.append(v)
        crow[i + 1] = crow[i] + per
    return crow, np.concatenate(cols), np.concatenate(vals)
