"""Host-side glue with the reference's names (reflect/utils.py): the
``FreezeParameters`` context manager whose semantics the imagine path relies on
(eval + requires_grad=False for the duration, utils.py:171-194) and the
``AdamOptim`` wrapper (utils.py:12-50)."""
from __future__ import annotations

from typing import Iterable, List

import torch
from torch import nn


def _params(modules: Iterable[nn.Module]) -> List[nn.Parameter]:
    out: List[nn.Parameter] = []
    for m in modules:
        if m is not None:
            out.extend(m.parameters())
    return out


class FreezeParameters:
    """``with FreezeParameters([m]):`` - modules go to eval mode and their
    parameters stop requiring grad; both are restored on exit."""

    def __init__(self, modules: Iterable[nn.Module]):
        self.modules = list(modules)
        self._was = [p.requires_grad for p in _params(self.modules)]

    def __enter__(self):
        for m in self.modules:
            if m is not None:
                m.eval()
        for p in _params(self.modules):
            p.requires_grad = False

    def __exit__(self, exc_type, exc, tb):
        for m in self.modules:
            if m is not None:
                m.train()
        for p, flag in zip(_params(self.modules), self._was):
            p.requires_grad = flag


class AdamOptim:
    """Adam with global grad-norm clipping, same call surface as utils.py:12-50:
    ``backward(loss)`` zeroes, back-propagates and returns the (clipped) norm,
    ``update_parameters()`` steps."""

    def __init__(self, parameters, lr, betas=(0.9, 0.999), eps=1e-8, grad_clip=torch.inf, weight_decay=0):
        self.parameters = list(parameters)
        self.lr, self.grad_clip = lr, grad_clip
        self.optimizer = torch.optim.Adam(self.parameters, lr=lr, betas=betas, eps=eps,
                                          weight_decay=weight_decay)

    def zero_grad(self):
        self.optimizer.zero_grad()

    def backward(self, loss, retain_graph=False):
        self.zero_grad()
        loss.backward(retain_graph=retain_graph)
        return torch.nn.utils.clip_grad_norm_(self.parameters, self.grad_clip)

    def update_parameters(self):
        self.optimizer.step()

    def _set_lr(self, lr):
        for group in self.optimizer.param_groups:
            group["lr"] = lr

    def update_lr(self, lr):
        self._set_lr(lr)

    def update_lr_by_factor(self, factor):
        self._set_lr(self.lr * factor)

    def state_dict(self):
        return self.optimizer.state_dict()

    def load_state_dict(self, state_dict):
        self.optimizer.load_state_dict(state_dict)

    def to(self, device):
        for slot in self.optimizer.state.values():
            for name, value in list(slot.items()):
                if torch.is_tensor(value):
                    slot[name] = value.to(device)
