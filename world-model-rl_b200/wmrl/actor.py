"""Policy MLP with the reference Actor's interface (components/models/actor.py:6-71).

[Linear, LayerNorm, ELU] x num_layers, tanh-squashed mean, optional Normal head.
Inside ``RSSM.imagine_rollout`` the deterministic forward of this module runs in
the fused native path (its weights are read straight from these parameters);
``forward`` below is the eager API used for single observations.
"""
from __future__ import annotations

import torch
import torch.distributions as D
from torch import nn

from .utils import FreezeParameters


class Actor(nn.Module):
    def __init__(self, input_dim, output_dim, bound, num_layers=3, hidden_dim=512, action_scale=5.0):
        super().__init__()
        self.input_dim, self.output_dim = input_dim, output_dim
        self.action_scale = action_scale
        self.bound = torch.tensor(bound, dtype=torch.float32)   # plain attribute, as in actor.py:20
        self.num_layers, self.hidden_dim = num_layers, hidden_dim
        blocks, width = [], input_dim
        for _ in range(num_layers):
            blocks += [nn.Linear(width, hidden_dim), nn.LayerNorm(hidden_dim), nn.ELU()]
            width = hidden_dim
        self.layers = nn.Sequential(*blocks)
        self.mu = nn.Linear(hidden_dim, output_dim)
        self.stddev = nn.Linear(hidden_dim, output_dim)

    def to(self, *args, **kwargs):
        super().to(*args, **kwargs)
        self.bound = self.bound.to(*args, **kwargs)
        return self

    def forward(self, x, deterministic=False):
        hid = self.layers(x)
        mu = torch.tanh(self.mu(hid) / self.action_scale) * self.bound
        if deterministic:
            return mu
        std = 1.0 * torch.sigmoid(self.stddev(hid)) + 0.01      # actor.py:54-57
        return D.Independent(D.Normal(mu, std), 1)

    def sample_action(self, x):
        """rsample + entropy + clamp, the use-site the reference pairs with the
        Normal head (transformer_world_model/world_model.py:311-322)."""
        dist = self(x, deterministic=False)
        bound = self.bound.to(x.device)
        action = torch.max(torch.min(dist.rsample(), bound), -bound)
        return action, dist.entropy()

    def compute_action(self, state, deterministic=False):
        with FreezeParameters([self]):
            device = next(self.parameters()).device
            if not torch.is_tensor(state):
                state = torch.tensor(state, dtype=torch.float32, device=device)
            if state.dim() == 1:
                state = state[None, :]
            return self(state, deterministic=deterministic)
