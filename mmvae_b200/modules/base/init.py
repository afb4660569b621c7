"""He initialisation of every nn.Linear under a module (reference: src/cmmvae/modules/base/init.py:4-9):
Kaiming-normal with fan_out / relu gain on weights, zeros on biases."""
import torch.nn as nn


def he_init_weights(module: nn.Module) -> None:
    linears = (m for m in module.modules() if isinstance(m, nn.Linear))
    for lin in linears:
        nn.init.kaiming_normal_(lin.weight, mode="fan_out", nonlinearity="relu")
        if lin.bias is not None:
            nn.init.zeros_(lin.bias)
