"""Drop-in plumbing: make the reference's dotted paths resolve to this package and instantiate the
reference's YAML ``class_path``/``init_args`` trees without jsonargparse (absent from this image).

    import mmvae_b200.compat as compat
    compat.install_as_cmmvae()                 # `import cmmvae.models` -> mmvae_b200.models
    model = compat.instantiate(yaml.safe_load(open("configs/model/human_only.yaml")))
"""
from __future__ import annotations

import importlib
import sys
from typing import Any

_ALIASES = ("", ".config", ".constants", ".models", ".models.base_model", ".models.cmmvae_model", ".modules",
            ".modules.vae", ".modules.clvae", ".modules.cmmvae", ".modules.base", ".modules.base.components",
            ".modules.base.annealing_fn", ".modules.base.init")


def install_as_cmmvae() -> None:
    """Register ``cmmvae[...]`` aliases in ``sys.modules`` so YAML files and user code written against
    the reference (``class_path: cmmvae.models.CMMVAEModel``) load this implementation unchanged."""
    for suffix in _ALIASES:
        sys.modules["cmmvae" + suffix] = importlib.import_module("mmvae_b200" + suffix)


def _resolve(path: str) -> Any:
    if path.startswith("cmmvae.") or path == "cmmvae":
        path = "mmvae_b200" + path[len("cmmvae"):]
    module_name, _, attr = path.rpartition(".")
    return getattr(importlib.import_module(module_name), attr)


def instantiate(node: Any) -> Any:
    """Recursively build ``{class_path, init_args}`` trees (jsonargparse convention used by
    configs/model/*.yaml).  Strings that look like dotted class paths of torch.nn (``torch.nn.ReLU``)
    are resolved to the class itself, as jsonargparse does for ``Type[nn.Module]`` arguments."""
    if isinstance(node, dict):
        if "class_path" in node:
            cls = _resolve(node["class_path"])
            kwargs = {k: instantiate(v) for k, v in (node.get("init_args") or {}).items()}
            return cls(**kwargs)
        return {k: instantiate(v) for k, v in node.items()}
    if isinstance(node, list):
        return [instantiate(v) for v in node]
    if isinstance(node, str) and node.startswith("torch.nn."):
        return _resolve(node)
    return node
