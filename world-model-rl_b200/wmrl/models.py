"""The MLP heads next to the path: ``DenseModel`` (reward / done / state enc-dec heads,
rssm_world_model/models.py:3-37) and ``ValueCritic`` (trainers/value/critic.py:4-29).  They consume the
``features[B,H,F]`` the native rollout returns.  Same constructor arguments, module tree and ``state_dict``
keys as the reference; on CUDA fp32 inputs their Linear / LayerNorm+activation blocks run on libwmrl's kernels
(``wmrl.functional.mlp_forward``: 3xTF32 tcgen05 GEMM, fused LayerNorm+SiLU/ReLU, forward and backward) -
the first step of SURVEY.md §8(f) rank 1 (fusing them onto the rollout itself comes next)."""
from __future__ import annotations

import torch
from torch import nn

from .functional import mlp_forward, mlp_forward_bf16


def _mlp(in_dim, hidden, depth, out_dim, act, layernorm=True):
    layers, width = [], in_dim
    for _ in range(depth):
        layers.append(nn.Linear(width, hidden))
        if layernorm:
            layers.append(nn.LayerNorm(hidden))
        layers.append(act())
        width = hidden
    layers.append(nn.Linear(width, out_dim))
    return nn.Sequential(*layers)


_PRECISIONS = ("fp32", "bf16")


def set_precision(module: nn.Module, precision: str) -> nn.Module:
    """Switch every wmrl module under ``module`` (RSSM, DenseModel heads, ValueCritic) to ``precision``:
    "fp32" = the reference-accurate kernels, "bf16" = the tcgen05 paths (fp32 in / out, bf16 MMA operands)."""
    if precision not in _PRECISIONS:
        raise ValueError(f"precision must be one of {_PRECISIONS}")
    for m in module.modules():
        if hasattr(m, "set_precision") and m is not module:
            m.set_precision(precision)
        elif isinstance(m, (DenseModel, ValueCritic)):
            m.precision = precision
    if hasattr(module, "set_precision") and not isinstance(module, (DenseModel, ValueCritic)):
        module.set_precision(precision)
    return module


class DenseModel(nn.Module):
    def __init__(self, depth: int = 3, input_dim: int = 230, hidden_dim: int = 256, output_dim: int = 1,
                 hidden_act=nn.SiLU, output_act=nn.Identity, use_layernorm: bool = True, precision: str = "fp32"):
        super().__init__()
        self.fc = _mlp(input_dim, hidden_dim, depth, output_dim, hidden_act, use_layernorm)
        self.output_act = output_act()
        self.precision = precision   # "bf16": the layer-wise tcgen05 chain (wmrl_mlp_*), fp32 in / out

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        run = mlp_forward_bf16 if self.precision == "bf16" else mlp_forward
        return self.output_act(run(self.fc, z))


class ValueCritic(nn.Module):
    def __init__(self, state_dim, num_layers=3, hidden_dim=512, precision: str = "fp32"):
        super().__init__()
        self.state_dim, self.num_layers, self.hidden_dim = state_dim, num_layers, hidden_dim
        self.layers = _mlp(state_dim, hidden_dim, num_layers, 1, nn.ReLU, True)
        self.precision = precision

    def forward(self, x):
        return (mlp_forward_bf16 if self.precision == "bf16" else mlp_forward)(self.layers, x)
