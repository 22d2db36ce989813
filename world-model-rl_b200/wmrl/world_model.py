"""World-model wrapper with the reference's call surface
(rssm_world_model/world_model.py:73-222): ``observe_rollout`` and
``imagine_rollout`` route through the native RSSM path; the encoder / decoder /
reward / done heads are whatever torch modules the caller supplies.  ``update``
(the loss side, world_model.py:120-163) is ordinary torch glue kept so that the
module remains a drop-in; it is not part of the accelerated path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributions as D
import torch.nn.functional as F
from torch import nn

from .utils import AdamOptim, FreezeParameters


@dataclass
class WorldModelTrainingParams:
    recon_coeff: float = 1.0
    dynamic_coeff: float = 1.0
    reward_coeff: float = 10.0
    done_coeff: float = 1.0
    rho: float = 3.0
    lr: float = 6e-4
    grad_clip: float = 1.0


@dataclass
class WorldModelLosses:
    recon_loss: float
    dynamic_model_loss: float
    dynamic_model_loss_clamped: float
    reward_loss: float
    done_loss: float
    loss: float
    grad_norm: float


@dataclass
class ImaginedRollouts:
    rewards: torch.Tensor
    dones: torch.Tensor
    features: torch.Tensor
    observations: Optional[torch.Tensor] = None

    @property
    def shapes(self):
        return (self.rewards.shape, self.dones.shape, self.features.shape)


def _unit_normal_nll(pred: torch.Tensor, target: torch.Tensor, event_dims: int) -> torch.Tensor:
    """-log N(target; pred, 1) summed over the trailing ``event_dims`` dims, mean over the rest."""
    return -D.Independent(D.Normal(pred, 1.0), event_dims).log_prob(target).mean()


class WorldModel(nn.Module):
    model_list = ["encoder", "decoder", "dynamic_model", "reward_model", "done_model", "opt"]

    def __init__(self, encoder, decoder, done_model, reward_model, dynamic_model,
                 params: Optional[WorldModelTrainingParams] = None):
        super().__init__()
        self.params = params or WorldModelTrainingParams()
        self.encoder, self.decoder = encoder, decoder
        self.dynamic_model = dynamic_model
        self.reward_model, self.done_model = reward_model, done_model
        self.opt = AdamOptim(self.parameters(), lr=self.params.lr, grad_clip=self.params.grad_clip)

    # ---- hot-path callers ---------------------------------------------------
    def observe_rollout(self, obs: torch.Tensor, action: torch.Tensor, **noise_kw):
        """world_model.py:110-118 -> (o_emb, prior_sequence, posterior_sequence)."""
        o_emb = self.encoder(obs)
        prior_seq, post_seq = self.dynamic_model.observe_rollout(o_emb, action, **noise_kw)
        return o_emb, prior_seq, post_seq

    def imagine_rollout(self, initial_states, actor, n_steps: int, with_observations: bool = False, **noise_kw):
        """world_model.py:165-189: the whole world model is frozen for the
        duration, so the native backward produces actor gradients only."""
        with FreezeParameters([self]):
            seq = self.dynamic_model.imagine_rollout(initial_states=initial_states, actor=actor,
                                                     n_steps=n_steps, **noise_kw)
            features = seq.get_features()
            rewards = self.reward_model(features)
            dones = self.done_model(features)
            observations = self.decoder(features) if with_observations else None
        return ImaginedRollouts(rewards=rewards, dones=dones, features=features, observations=observations)

    def _kl_terms(self, prior_seq, post_seq, rho: float):
        """(mean max(KL(prior || posterior), rho), mean KL) over [B, T-1] (world_model.py:131-141): one fused libwmrl
        pass over the distribution parameters on CUDA, the torch.distributions expression otherwise."""
        from . import _lib
        from .functional import PathConfig, kl_loss
        disc = hasattr(prior_seq, "logits")
        a1 = prior_seq.logits if disc else prior_seq.means
        if a1.is_cuda and a1.dtype == torch.float32 and a1.shape[1] > 1:
            if disc:
                C_, S_ = a1.shape[2], a1.shape[3]
                if S_ <= 32:
                    cfg = PathConfig(variant=_lib.DISCRETE, D=1, S=S_, C=C_, A=1, Hd=1, E=1)
                    kc, km, _ = kl_loss(cfg, rho, a1, None, post_seq.logits, None)
                    return kc, km
            else:
                cfg = PathConfig(variant=_lib.CONTINUOUS, D=1, S=a1.shape[2], C=1, A=1, Hd=1, E=1)
                kc, km, _ = kl_loss(cfg, rho, a1, prior_seq.stds, post_seq.means, post_seq.stds)
                return kc, km
        kl = D.kl_divergence(prior_seq.get_dist(), post_seq.get_dist())
        return torch.clamp_min(kl, rho).mean(), kl.mean()

    # ---- loss side (torch glue) ---------------------------------------------
    def update(self, prior_state_sequence, posterior_state_sequence, obs, reward, done) -> WorldModelLosses:
        p = self.params
        feats = posterior_state_sequence.get_features()
        kl_clamped, kl_mean = self._kl_terms(prior_state_sequence, posterior_state_sequence, p.rho)
        pred_r, pred_d = self.reward_model(feats), self.done_model(feats)
        if pred_r.is_cuda and pred_r.dtype == torch.float32 and pred_r.shape == reward[:, 1:].shape \
                and pred_d.shape == done[:, 1:].shape:
            from .functional import head_losses      # both losses and their gradients in one fused pass
            reward_loss, done_loss = head_losses(pred_r, reward[:, 1:], pred_d, done[:, 1:].float())
        else:
            reward_loss = _unit_normal_nll(pred_r, reward[:, 1:], 1)
            done_loss = F.binary_cross_entropy_with_logits(pred_d, done[:, 1:].float())
        decoded = self.decoder(feats)
        tgt = obs[:, 1:]
        # the reference flattens [B,T] and treats the remaining dims beyond the first as the event
        flat_dec = decoded.reshape(-1, *decoded.shape[2:])
        flat_tgt = tgt.reshape(-1, *tgt.shape[2:])
        recon_loss = _unit_normal_nll(flat_tgt, flat_dec, max(flat_dec.dim() - 2, 0)) if flat_dec.dim() >= 2 \
            else _unit_normal_nll(flat_tgt, flat_dec, 0)
        loss = (p.recon_coeff * recon_loss + p.dynamic_coeff * kl_clamped + p.reward_coeff * reward_loss
                + p.done_coeff * done_loss)
        grad_norm = self.opt.backward(loss)
        self.opt.update_parameters()
        from .functional import clear_panel_cache
        clear_panel_cache()
        return WorldModelLosses(recon_loss=recon_loss.item(), dynamic_model_loss=kl_mean.item(),
                                dynamic_model_loss_clamped=kl_clamped.item(), reward_loss=reward_loss.item(),
                                done_loss=done_loss.item(), loss=loss.item(), grad_norm=grad_norm.item())

    # ---- checkpoints (world_model.py:191-222) -------------------------------
    def save(self, path, name="world-model-checkpoint.pth", targets=None):
        targets = targets or self.model_list
        torch.save({t: getattr(self, t).state_dict() for t in targets}, f"{path}/{name}")

    def load(self, path, name="world-model-checkpoint.pth", targets=None):
        device = next(self.parameters()).device
        ckpt = torch.load(f"{path}/{name}", map_location=torch.device(device))
        for t in targets or self.model_list:
            getattr(self, t).load_state_dict(ckpt[t])
