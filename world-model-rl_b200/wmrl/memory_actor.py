"""Stateful online policy (rssm_world_model/memory_actor.py:8-57): encoder ->
posterior -> actor -> prior at B=1.  On CUDA the three RSSM / actor stages are one native cluster kernel
(``RSSMBase.act_step`` -> wmrl_act_step); CPU modules use the eager single-step API."""
from __future__ import annotations

import copy

import torch

from .utils import FreezeParameters


class WorldModelActor:
    def __init__(self, world_model, actor):
        self.world_model, self.actor = world_model, actor
        self.actor_copy = copy.deepcopy(actor)

    def reset_to_original_actor(self):
        self.actor_copy = copy.deepcopy(self.actor)

    def perturb_actor(self, weight_perturbation_size: float = 0.01):
        self.reset_to_original_actor()
        with torch.no_grad():
            for name, p in self.actor_copy.named_parameters():
                if "weight" in name:
                    p.add_(torch.randn_like(p) * weight_perturbation_size)

    def reset(self):
        device = next(self.actor.parameters()).device
        self.state = self.world_model.dynamic_model.initial_state(1)
        self.state.to(device)
        self.action = torch.zeros(1, self.actor.output_dim, device=device)

    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        if obs.dim() in (2, 4):          # add the missing time axis (memory_actor.py:48-49)
            obs = obs.unsqueeze(1)
        device = next(self.actor.parameters()).device
        rssm = self.world_model.dynamic_model
        with FreezeParameters([self.world_model, self.actor, self.actor_copy]):
            embed = self.world_model.encoder(obs.to(device))
            if embed.is_cuda and embed.dtype == torch.float32 and hasattr(rssm, "act_step"):
                # posterior -> actor -> prior as ONE cluster kernel (wmrl_act_step) instead of ~45 launches
                self.action, self.state, _ = rssm.act_step(embed[:, 0], self.state, self.actor)
            else:
                post = rssm.posterior(embed[:, 0], self.state)
                self.action = self.actor(post.get_features(), deterministic=True)
                self.state = rssm.prior(self.action, post)
        return self.action
