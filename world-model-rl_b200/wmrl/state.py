"""Latent-state containers with the reference's names, fields and methods.

Mirrors ``reflect.components.rssm_world_model.state`` (state/base.py:7-19,
state/continuous.py:8-96, state/discrete.py:8-89) so that code written against
the reference - ``get_features``, ``flatten_batch_time``, ``get_dist``,
``append_``, indexing, ``detach`` / ``to`` - keeps working on the tensors the
CUDA path returns.  The kernels write whole ``[B, T, .]`` sequences at once, so
``append_`` is only used by the eager single-step API.

Field conventions: a *state* holds ``[B, .]`` tensors, a *sequence* holds
``[B, T, .]`` tensors whose index 0 is the initial state; everything derived
for losses / actors drops index 0 (state/continuous.py:73-96).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.distributions as D


def _tensor_items(obj):
    return [(f.name, getattr(obj, f.name)) for f in dataclasses.fields(obj)]


@dataclass
class Base:
    """Common tensor-bag behaviour (state/base.py:7-19): ``detach`` returns a
    new object, ``to`` moves the fields in place, ``shapes`` lists them."""

    def detach(self):
        return type(self)(**{k: v.detach() for k, v in _tensor_items(self)})

    @property
    def shapes(self) -> Tuple[Tuple[int, ...], ...]:
        return tuple(tuple(v.shape) for _, v in _tensor_items(self))

    def to(self, device):
        for k, v in _tensor_items(self):
            setattr(self, k, v.to(device))


class _StateOps:
    """[B, .] state: features are cat(deter, stoch) (state/continuous.py:15-19)."""
    _seq_cls = None
    _plural = {}

    def get_features(self) -> torch.Tensor:
        return torch.cat((self.deter_state, self.stoch_state), dim=-1)

    def to_sequence(self):
        return self._seq_cls.from_init(self)


class _SequenceOps:
    """[B, T, .] sequence; index 0 is the initial state."""
    _state_cls = None
    _singular = {}

    @classmethod
    def from_init(cls, init_state):
        return cls(**{plural: getattr(init_state, single).unsqueeze(1)
                      for plural, single in cls._singular.items()})

    def append_(self, other) -> None:
        for plural, single in self._singular.items():
            grown = torch.cat((getattr(self, plural), getattr(other, single).unsqueeze(1)), dim=1)
            setattr(self, plural, grown)

    def __getitem__(self, index):
        return self._state_cls(**{single: getattr(self, plural)[:, index]
                                  for plural, single in self._singular.items()})

    def get_last(self):
        return self[-1]

    def get_features(self) -> torch.Tensor:
        d, s = self.deter_states, self.stoch_states
        if (d.is_cuda and d.dtype == torch.float32 and s.dtype == torch.float32 and d.dim() == 3 and s.dim() == 3
                and d.device == s.device):
            from .functional import sequence_features       # one streaming pass each way instead of torch.cat
            return sequence_features(d, s)
        return torch.cat((d[:, 1:], s[:, 1:]), dim=-1)

    def flatten_batch_time(self):
        flat = {}
        for plural, single in self._singular.items():
            v = getattr(self, plural)[:, 1:]
            flat[single] = v.reshape(-1, *v.shape[2:])
        return self._state_cls(**flat)


# ------------------------------------------------------------------ continuous
@dataclass
class InternalStateContinuous(_StateOps, Base):
    deter_state: torch.Tensor
    stoch_state: torch.Tensor
    mean: torch.Tensor
    std: torch.Tensor

    @classmethod
    def from_mean_std(cls, deter_state, mean, std):
        """reparameterised draw from N(mean, std) (state/continuous.py:21-36)."""
        stoch = D.Independent(D.Normal(mean, std), 1).rsample()
        return cls(deter_state=deter_state, stoch_state=stoch, mean=mean, std=std)


@dataclass
class InternalStateContinuousSequence(_SequenceOps, Base):
    deter_states: torch.Tensor
    stoch_states: torch.Tensor
    means: torch.Tensor
    stds: torch.Tensor

    def get_dist(self):
        """Independent Normal over steps 1.. (state/continuous.py:79-84)."""
        return D.Independent(D.Normal(self.means[:, 1:], self.stds[:, 1:]), 1)


InternalStateContinuous._seq_cls = InternalStateContinuousSequence
InternalStateContinuousSequence._state_cls = InternalStateContinuous
InternalStateContinuousSequence._singular = {
    "deter_states": "deter_state", "stoch_states": "stoch_state", "means": "mean", "stds": "std"}


# -------------------------------------------------------------------- discrete
@dataclass
class InternalStateDiscrete(_StateOps, Base):
    deter_state: torch.Tensor
    stoch_state: torch.Tensor
    logits: torch.Tensor

    @classmethod
    def from_logits(cls, deter_state, logits, temperature: float = 1.0):
        """straight-through one-hot draw per group; the categorical axis is the
        last one (state/discrete.py:20-37)."""
        dist = D.Independent(D.OneHotCategoricalStraightThrough(logits=logits / temperature), 1)
        sample = dist.rsample()
        return cls(deter_state=deter_state, stoch_state=sample.flatten(1), logits=logits)


@dataclass
class InternalStateDiscreteSequence(_SequenceOps, Base):
    deter_states: torch.Tensor
    stoch_states: torch.Tensor
    logits: torch.Tensor

    def get_dist(self, temperature: float = 1.0):
        """state/discrete.py:77-79"""
        return D.Independent(
            D.OneHotCategoricalStraightThrough(logits=self.logits[:, 1:] / temperature), 1)


InternalStateDiscrete._seq_cls = InternalStateDiscreteSequence
InternalStateDiscreteSequence._state_cls = InternalStateDiscrete
InternalStateDiscreteSequence._singular = {
    "deter_states": "deter_state", "stoch_states": "stoch_state", "logits": "logits"}
