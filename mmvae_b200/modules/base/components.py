"""Building blocks of the CMMVAE network, B200-native.

Public surface = the reference's ``cmmvae.modules.base.components`` (same class names, constructor
arguments, forward signatures and ``state_dict`` keys; reference: src/cmmvae/modules/base/components.py):
``FCBlockConfig`` 40-174, ``ConcatBlockConfig`` 177-190, ``FCBlock`` 193-314, ``ConditionalLayer`` 317-413,
``collect_species_files`` 420-464, ``ConditionalLayers`` 467-631, ``Adversarial`` 638-674, ``Encoder``
676-809, ``Expert`` 812-857, ``Experts`` 860-876, ``GradientReversalFunction`` 879-899.

What differs is where the arithmetic runs: every Linear / BatchNorm / ReLU / Dropout of an ``FCBlock``
executes in the hand-written sm_100a kernels of ``libcmmvae_b200.so`` (``mmvae_b200.layers``); a
sparse-CSR batch entering an expert encoder goes through the CSR SpMM kernel without densification.
Construction and validation are pure host logic and work without a GPU; ``forward`` requires CUDA
tensors and raises otherwise (there is no CPU fallback).
"""
from __future__ import annotations

import os
import random
from collections import OrderedDict, defaultdict
from typing import Callable, List, Literal, Optional, Type, Union

import pandas as pd
import torch
import torch.nn as nn
from torch.distributions import Normal

from mmvae_b200 import layers as L


def is_iterable(obj) -> bool:
    """True when ``iter(obj)`` works (strings and dicts count, ints and None do not)."""
    try:
        iter(obj)
        return True
    except TypeError:
        return False


# per-layer options of an FCBlock: name -> (accepted type, compare with issubclass?, None allowed?)
_LAYER_OPTIONS = OrderedDict(
    dropout_rate=(float, False, False),
    use_batch_norm=(bool, False, False),
    use_layer_norm=(bool, False, False),
    activation_fn=(nn.Module, True, True),
    return_hidden=(bool, False, False),
)


class FCBlockConfig:
    """Options for one :class:`FCBlock`.

    ``layers`` lists the widths ``[n_0, n_1, ..., n_L]``; consecutive pairs become Linear layers and a
    single width ``[n]`` means one ``n -> n`` layer.  Every other option may be a scalar (applied to
    all layers) or a list with one entry per layer.
    """

    def __init__(self, layers: List[int], dropout_rate: Union[float, List[float]] = 0.0,
                 use_batch_norm: Union[bool, List[bool]] = False, use_layer_norm: Union[bool, List[bool]] = False,
                 return_hidden: Union[bool, List[bool]] = False,
                 activation_fn: Union[Optional[Type[nn.Module]], List[Optional[Type[nn.Module]]]] = None):
        if not isinstance(layers, list):
            raise ValueError(f"layers must be a list found type: {type(layers)}")
        if not all(isinstance(width, int) and width > 0 for width in layers):
            raise ValueError("layers must be positive integers")
        self.layers = layers * 2 if len(layers) == 1 else layers
        given = dict(dropout_rate=dropout_rate, use_batch_norm=use_batch_norm, use_layer_norm=use_layer_norm,
                     activation_fn=activation_fn, return_hidden=return_hidden)
        for name in _LAYER_OPTIONS:
            value = given[name]
            setattr(self, name, value if is_iterable(value) else [value] * self.n_layers)
        self.validate()

    @property
    def n_layers(self) -> int:
        if not hasattr(self, "layers"):
            raise RuntimeError("n_layers called before layers initialized")
        return max(len(self.layers) - 1, 1)

    def _validate_option(self, name, req_type, comparison_fn=isinstance, optional=False):
        values = getattr(self, name)
        if values is None and not optional:
            raise ValueError(f"{name} is not optional but value is None")
        if len(values) != self.n_layers:
            raise ValueError(f"Length of '{name}' must match the length of 'layers':{len(values)} != {self.n_layers}")
        accepted = (req_type, type(None)) if optional else (req_type,)
        for v in values:
            if v is None and optional:
                continue
            try:
                ok = v is not None and comparison_fn(v, accepted)
            except TypeError:
                ok = False
            if not ok:
                raise ValueError(f"All elements in '{name}' must be a {str(req_type)}")

    def validate(self) -> None:
        for name, (req_type, by_subclass, optional) in _LAYER_OPTIONS.items():
            self._validate_option(name, req_type, issubclass if by_subclass else isinstance, optional)


class ConcatBlockConfig(FCBlockConfig):
    """Options of the single extra layer CLVAE prepends to the decoder in ``parallel`` mode; holds
    scalars only (no ``layers``), exactly like the reference object."""

    def __init__(self, dropout_rate: float = 0.0, use_batch_norm: bool = False, use_layer_norm: bool = False,
                 return_hidden: bool = False, activation_fn: Optional[Type[nn.Module]] = None):
        self.dropout_rate = dropout_rate
        self.use_batch_norm = use_batch_norm
        self.use_layer_norm = use_layer_norm
        self.return_hidden = return_hidden
        self.activation_fn = activation_fn


class FCBlock(nn.Module):
    """Stack of ``lin -> [bn] -> [ln] -> [af] -> [dr]`` layers.

    ``fc_layers`` is an ``nn.Sequential`` of per-layer ``nn.Sequential`` s whose children are named
    ``lin``/``bn``/``ln``/``af``/``dr`` so checkpoints interchange with the reference.  The children are
    parameter holders: ``forward`` hands each layer to ``mmvae_b200.layers.run_layer`` which executes
    it in the CUDA kernels (GEMM or CSR SpMM, fused BN/ReLU/dropout)."""

    def __init__(self, config: FCBlockConfig):
        super().__init__()
        config.validate()
        self.config = config
        widths = config.layers
        self.fc_layers = nn.Sequential(*[
            self._make_layer(n_in, n_out, config.use_batch_norm[i], config.use_layer_norm[i],
                             config.activation_fn[i], config.dropout_rate[i], config.return_hidden[i])
            for i, (n_in, n_out) in enumerate(zip(widths[:-1], widths[1:]))
        ])

    @property
    def input_dim(self) -> int:
        return self.config.layers[0]

    @property
    def output_dim(self) -> int:
        return self.config.layers[-1]

    @property
    def can_bypass(self) -> bool:
        return not any(self.config.return_hidden)

    def _make_layer(self, n_in: int, n_out: int, use_batch_norm: bool, use_layer_norm: bool,
                    activation_fn: Optional[Type[nn.Module]], dropout_rate: float,
                    return_hidden: bool) -> nn.Sequential:
        parts = OrderedDict(lin=nn.Linear(n_in, n_out))
        if use_batch_norm:
            parts["bn"] = nn.BatchNorm1d(n_out, momentum=0.01, eps=0.001)
        if use_layer_norm:
            parts["ln"] = nn.LayerNorm(n_out, elementwise_affine=False)
        if activation_fn is not None:
            parts["af"] = activation_fn(dim=1) if issubclass(activation_fn, nn.Softmax) else activation_fn()
        if dropout_rate > 0:
            parts["dr"] = nn.Dropout(p=dropout_rate)
        return nn.Sequential(parts)

    def forward(self, x: torch.Tensor):
        hidden = []
        for i, layer in enumerate(self.fc_layers):
            x, post_act = L.run_layer(layer, x, self.training, want_hidden=self.config.return_hidden[i])
            if self.config.return_hidden[i] and post_act is not None:
                hidden.append(post_act)
        if self.can_bypass:
            return x
        return x, hidden


class ConditionalLayer(nn.Module):
    """One ``FCBlock`` per distinct value of a metadata column; each cell goes through the block of
    its own value and rows keep their order.  The values are read (one per line) from
    ``conditions_path``; '.' in a value becomes '_' in the module key."""

    def __init__(self, batch_key: str, conditions_path: str, fc_block_config: FCBlockConfig):
        super().__init__()
        self.batch_key = batch_key
        values = pd.read_csv(conditions_path, header=None)[0]
        self.conditions = nn.ModuleDict({self.format_condition_key(v): FCBlock(fc_block_config) for v in values})

    def format_condition_key(self, condition: str) -> str:
        return condition.replace(".", "_")

    def forward(self, x: torch.Tensor, metadata: pd.DataFrame, condition: Optional[str] = None):
        if condition:
            return self.conditions[self.format_condition_key(condition)](x)
        keys = metadata[self.batch_key].astype(str).map(self.format_condition_key).tolist()
        rows_of: "OrderedDict[str, list]" = OrderedDict()
        for r, k in enumerate(keys):
            rows_of.setdefault(k, []).append(r)
        out = torch.empty_like(x)
        for k, rows in rows_of.items():
            idx = torch.tensor(rows, device=x.device)
            out.index_copy_(0, idx, self.conditions[k](x.index_select(0, idx)))
        return out


def _is_valid_file(fname, batch_key):
    return fname == f"unique_expression_{batch_key}.csv"


def collect_species_files(directory, batch_keys, species_files=None,
                          is_valid_file: Optional[Callable[[str, str], bool]] = None):
    """Map ``{"shared": {batch_key: path}, species: {batch_key: path}}`` from a directory laid out as
    ``<directory>/shared/*.csv`` + ``<directory>/<species>/*.csv``.  A batch key present under
    ``shared`` is never repeated under a species."""
    accept = is_valid_file or _is_valid_file
    found = {} if species_files is None else species_files

    def scan(folder, skip=()):
        hits = {}
        for fname in os.listdir(folder):
            path = os.path.join(folder, fname)
            if not os.path.isfile(path):
                continue
            key = next((k for k in batch_keys if accept(fname, k)), None)
            if key is not None and key not in skip:
                hits[key] = path
        return hits

    shared_dir = os.path.join(directory, "shared")
    shared = scan(shared_dir) if os.path.isdir(shared_dir) else {}
    found["shared"] = shared
    for entry in os.listdir(directory):
        sub = os.path.join(directory, entry)
        if entry == "shared" or not os.path.isdir(sub):
            continue
        own = scan(sub, skip=shared)
        if own:
            found[entry] = own
    print(f"Collected species files {found}")
    return found


class ConditionalLayers(nn.Module):
    """All conditional layers of a model: shared ones (one per batch key), species-specific ones
    (``ModuleDict`` keyed by species) and an optional per-species ``species`` block.  Applied in a
    fixed order, in a random order (``selection_order`` empty), or in parallel with the outputs
    concatenated (``selection_order == ["parallel"]``)."""

    def __init__(self, directory: str, conditionals: list, fc_block_config: FCBlockConfig,
                 selection_order: Optional[list] = None):
        super().__init__()
        if not os.path.exists(directory):
            raise FileNotFoundError(
                "Could not intialize the conditional layers either due to the directory not existing yet\n"
                f"{directory}")
        conditionals.remove("species")  # no csv is needed for the species conditional
        paths = collect_species_files(directory, conditionals)
        conditionals.append("species")
        self.shared_conditionals = list(paths["shared"].keys())
        self.is_parallel = selection_order[0] == "parallel"
        self.shuffle_selection_order = (not selection_order) or self.is_parallel
        if self.shuffle_selection_order:
            selection_order = conditionals

        modules = {k: ConditionalLayer(k, p, fc_block_config) for k, p in paths["shared"].items()}
        per_species = defaultdict(dict)
        for species, files in paths.items():
            if species != "shared":
                for k, p in files.items():
                    per_species[k][species] = p
        for k, by_species in per_species.items():
            if k in modules:
                raise RuntimeError(f"batch_key '{k}' is shared but attempted to make species specific")
            modules[k] = nn.ModuleDict({s: ConditionalLayer(k, p, fc_block_config) for s, p in by_species.items()})
        if "species" in conditionals:
            assert "species" not in modules
            modules["species"] = nn.ModuleDict({s: FCBlock(fc_block_config) for s in paths if s != "shared"})
        self.layers = nn.ModuleDict(modules)
        self.selection_order = selection_order

    def forward(self, x: torch.Tensor, metadata: pd.DataFrame, species: Optional[str] = None):
        order = (random.sample(self.selection_order, len(self.selection_order))
                 if self.shuffle_selection_order else self.selection_order)
        branches = []
        for key in order:
            layer = self.layers[key]
            if isinstance(layer, nn.ModuleDict):
                if species is None:
                    raise RuntimeError(
                        f"'species' must be set to access non-shared conditional layer for batch_key '{key}'")
                layer = layer[species]
            y = layer(x, metadata) if isinstance(layer, ConditionalLayer) else layer(x)
            if self.is_parallel:
                branches.append(y)
            else:
                x = y
        return torch.cat(branches, dim=1) if branches else x


class Adversarial(nn.Module):
    """Domain classifier on a hidden representation: an encoder ``FCBlock`` followed by one linear head
    per metadata condition.  Head widths come from ``<labels_dir>/human/unique_expression_<c>.csv``
    (row count); the class-level ``labels`` maps condition -> {value: row index} (int64 targets)."""

    labels = defaultdict(dict)

    def __init__(self, encoder: FCBlockConfig, heads: FCBlockConfig, conditions: list, labels_dir: str):
        super().__init__()
        self.encoder = FCBlock(encoder)
        head_blocks = {}
        for condition in conditions:
            table = pd.read_csv(os.path.join(labels_dir, f"human/unique_expression_{condition}.csv"), header=None)
            if condition not in Adversarial.labels.keys():
                for row, value in enumerate(table[0]):
                    Adversarial.labels[condition][value] = row
            heads.layers = [self.encoder.output_dim, len(table)]
            head_blocks[condition] = FCBlock(heads)
        self.heads = nn.ModuleDict(head_blocks)

    def forward(self, x: torch.Tensor):
        code = self.encoder(x)
        return {condition: head(code) for condition, head in self.heads.items()}


def _identity(x):
    return x


class Encoder(nn.Module):
    """Latent encoder: ``q = fc(x)``; ``mu = mean_encoder(q)``; ``var = exp(var_encoder(q)) + var_eps``;
    ``z = mu + eps * sqrt(var)``.  Returns ``(Normal(mu, sqrt(var)), z, hidden)`` when ``return_dist``
    else ``(mu, var, z, hidden)``.  The two heads run as one ``[2Z, n_hidden]`` GEMM and the
    reparameterisation in the fused latent kernel.  Noise comes from :func:`mmvae_b200.layers.draw_noise`
    (injectable for parity runs)."""

    def __init__(self, latent_dim: int, fc_block_config: FCBlockConfig,
                 distribution: Union[Literal["ln"], Literal["normal"]] = "normal", return_dist: bool = False,
                 hidden_z: bool = False, var_eps: float = 1e-4):
        super().__init__()
        self.fc = FCBlock(fc_block_config)
        n_hidden = fc_block_config.layers[-1]
        self.mean_encoder = nn.Linear(n_hidden, latent_dim)
        self.var_encoder = nn.Linear(n_hidden, latent_dim)
        self.z_transformation = nn.Softmax(dim=-1) if distribution == "ln" else _identity
        self.var_eps = var_eps
        self.return_dist = return_dist
        self.hidden_z = hidden_z

    @property
    def n_layers(self) -> int:
        return self.fc.config.n_layers

    def encode(self, x: torch.Tensor):
        return self.fc(x)

    def forward(self, x: torch.Tensor):
        encoded = self.encode(x)
        q, hidden = encoded if isinstance(encoded, tuple) else (encoded, [])
        mu, var, z = L.latent_head(q, self.mean_encoder, self.var_encoder, self.var_eps)
        z = self.z_transformation(z)
        if self.hidden_z:
            hidden.append(z)
        if self.return_dist:
            return Normal(mu, var.sqrt()), z, hidden
        return mu, var, z, hidden


class Expert(nn.Module):
    """Species expert: an encoder and a decoder ``FCBlock`` sharing an id; only ``encode``/``decode``
    are meaningful."""

    def __init__(self, id: str, encoder_config: FCBlockConfig, decoder_config: FCBlockConfig):
        super().__init__()
        self.id = id
        self.encoder = FCBlock(encoder_config)
        self.decoder = FCBlock(decoder_config)

    def forward(self, *args, **kwargs):
        """An expert has no joint forward: call ``encode`` or ``decode``."""
        raise NotImplementedError(self.forward.__doc__)

    def encode(self, x: torch.Tensor):
        return self.encoder(x)

    def decode(self, x: torch.Tensor):
        return self.decoder(x)


class Experts(nn.ModuleDict):
    """``{expert.id: expert}`` plus ``labels``: id -> position."""

    def __init__(self, experts: list):
        super().__init__({e.id: e for e in experts})
        self.labels = {key: i for i, key in enumerate(self.keys())}


class GradientReversalFunction(torch.autograd.Function):
    """Identity in forward, ``-alpha * grad`` in backward (Ganin et al. 2016).  In the fused training
    step the sign flip is folded into the adversary-input gradient; this Function serves callers that
    build their own autograd graphs."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.neg() * ctx.alpha, None
