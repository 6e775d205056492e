"""init_weights / get_lr with the semantics of /root/reference/src/utils.py:18-19,46-63."""
from __future__ import annotations

import math

from torch import nn


def get_lr(base_lr: float, batch_size: int, base_batch_size: int = 4) -> float:
    return base_lr * (batch_size / base_batch_size)


def init_weights(module: nn.Module) -> None:
    """conv: N(0, sqrt(2/fan_out)), fan_out = prod(kernel)*out_channels // groups, bias 0; BN: 1 / 0;
    Linear: U(+-1/sqrt(fan_out)).  Iterates ``module.modules()`` in registration order, so the RNG stream
    matches the reference for an identically constructed parameter tree."""
    for m in module.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Conv3d)):
            fan_out = (math.prod(m.kernel_size) * m.out_channels) // m.groups
            nn.init.normal_(m.weight, 0, math.sqrt(2.0 / fan_out))
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            bound = 1.0 / math.sqrt(m.weight.size(0))
            nn.init.uniform_(m.weight, -bound, bound)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
