"""Value-gradient actor-critic trainer with the reference's interface
(trainers/value/value_trainer.py:30-206).  The lambda-return targets - a triple
nested Python loop of tiny tensor ops in the reference (:62-134) - are one fused
O(H) scan kernel per call here (``wmrl_lambda_return_fwd``; its transposed scan
is the backward), and the done pre-processing (:148-149) is ``wmrl_done_mask``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Optional

import torch

from .functional import clear_panel_cache, done_mask, lambda_return
from .utils import AdamOptim, FreezeParameters


def update_target_network(target_model, model, tau=5e-3):
    """Polyak average target <- tau*model + (1-tau)*target (value_trainer.py:8-18)."""
    with torch.no_grad():
        for tgt, src in zip(target_model.parameters(), model.parameters()):
            tgt.mul_(1.0 - tau).add_(src, alpha=tau)


@dataclass
class ValueGradTrainerLosses:
    value_loss: float
    value_grad_norm: float
    actor_loss: Optional[float]
    actor_grad_norm: Optional[float]
    entropy_loss: Optional[float]


class ValueGradTrainer:
    def __init__(self, actor, critic, actor_lr: float = 0.001, critic_lr: float = 0.001, grad_clip: float = 100,
                 gamma: float = 0.99, lam: float = 0.95, eta: float = 0.001):
        self.gamma, self.lam, self.eta = gamma, lam, eta
        self.actor_lr, self.critic_lr = actor_lr, critic_lr
        self.actor, self.critic = actor, critic
        self.target_critic = copy.deepcopy(critic)
        self.actor_optim = AdamOptim(actor.parameters(), lr=actor_lr, grad_clip=grad_clip)
        self.critic_optim = AdamOptim(critic.parameters(), lr=critic_lr, grad_clip=grad_clip)

    # -- lambda-return ------------------------------------------------------
    def compute_rollout_value(self, target_state_values, rewards, states, dones, k):
        """k-step return G_k of the first step of the slice, [B,1] (value_trainer.py:62-86): sum_{j<h} gamma^j
        r_j (1-d_j) + gamma^h V_h (1-d_h), h = min(k, H-1).  Public in the reference, kept as a thin torch
        shim for callers that use it directly; the lambda-return itself never goes through it here
        (``wmrl_lambda_return_fwd`` evaluates all G_k of all steps in one O(H) scan)."""
        big_h = states.shape[1]
        h = min(int(k), big_h - 1)
        gam = torch.tensor([self.gamma ** i for i in range(big_h)], device=rewards.device, dtype=rewards.dtype)
        ret = (gam[:h][None, :, None] * rewards[:, :h] * (1 - dones[:, :h])).sum(1)
        return ret + gam[h] * target_state_values[:, h] * (1 - dones[:, h])

    def compute_value_target(self, target_state_values, rewards, states, dones):
        """lambda-return of the FIRST step of the given slice, [B,1]
        (value_trainer.py:88-113; ``states`` only fixes the horizon there)."""
        return lambda_return(target_state_values, rewards, dones, self.gamma, self.lam)[:, 0]

    def value_targets(self, target_state_values, rewards, dones):
        """all H targets at once, [B,H,1] (value_trainer.py:126-134)."""
        return lambda_return(target_state_values, rewards, dones, self.gamma, self.lam)

    def value_loss(self, state_samples, reward_samples, done_samples):
        # the critic reads DETACHED states: the reference lets the value loss back-propagate into the rollout
        # (value_trainer.py:121, 153-156), but the actor gradients that produces are zeroed by actor_optim.backward
        # before they are ever used (utils.py:19-21) - detaching skips a whole BPTT with identical updates
        values = self.critic(state_samples.detach())
        with FreezeParameters([self.target_critic]):
            target_values = self.value_targets(self.target_critic(state_samples), reward_samples, done_samples)
        loss = 0.5 * (target_values.detach() - values) ** 2
        return loss.mean(), target_values

    # -- one update (value_trainer.py:138-182) --------------------------------
    def update(self, state_samples, reward_samples, done_samples, entropy=None, critic_only=False):
        done_samples = done_mask(done_samples)
        value_loss, target_values = self.value_loss(state_samples, reward_samples, done_samples)
        value_gn = self.critic_optim.backward(value_loss, retain_graph=not critic_only)
        self.critic_optim.update_parameters()
        actor_loss = actor_gn = entropy_loss = None
        if not critic_only:
            actor_loss = -target_values.mean()
            entropy_loss = 0.0
            if entropy is not None:
                mean_entropy = entropy.mean()
                actor_loss = actor_loss - self.eta * mean_entropy
                entropy_loss = mean_entropy.item()
            actor_gn = self.actor_optim.backward(actor_loss)
            self.actor_optim.update_parameters()
        update_target_network(self.target_critic, self.critic)
        clear_panel_cache()      # the bf16 heads' feature panel (and the view that keeps its source alive) is per update
        return ValueGradTrainerLosses(
            value_loss=value_loss.item(), value_grad_norm=value_gn.item(),
            actor_loss=None if actor_loss is None else actor_loss.item(),
            actor_grad_norm=None if actor_gn is None else actor_gn.item(),
            entropy_loss=entropy_loss)

    def to(self, device):
        self.actor.to(device)
        self.critic.to(device)
        self.target_critic.to(device)

    def save(self, path):
        torch.save({"actor": self.actor.state_dict(), "actor_optim": self.actor_optim.state_dict(),
                    "critic": self.critic.state_dict(), "critic_optim": self.critic_optim.state_dict()},
                   f"{path}/agent.pth")

    def load(self, path):
        ckpt = torch.load(f"{path}/agent.pth")
        self.actor.load_state_dict(ckpt["actor"])
        self.actor_optim.load_state_dict(ckpt["actor_optim"])
        self.critic.load_state_dict(ckpt["critic"])
        self.critic_optim.load_state_dict(ckpt["critic_optim"])
        self.target_critic = copy.deepcopy(self.critic)
