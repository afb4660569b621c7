"""Weight initialisation applied by ``BaseModel.init_weights`` to the whole LightningModule
(reference src/cmmvae/modules/base/init.py:4-9, base_model.py:106-109): every ``nn.Linear`` gets
He/Kaiming-normal weights (fan_out, ReLU gain, i.e. std = sqrt(2 / out_features)) and a zero bias;
everything else (BatchNorm affine, ...) keeps torch's defaults."""
import torch.nn as nn


def _he_linear(layer: nn.Module) -> None:
    if not isinstance(layer, nn.Linear):
        return
    nn.init.kaiming_normal_(layer.weight, nonlinearity="relu", mode="fan_out")
    if layer.bias is not None:
        nn.init.zeros_(layer.bias)


def he_init_weights(module: nn.Module) -> None:
    module.apply(_he_linear)
