"""Exponential moving average of the trainable parameters (upstream stnf/utils/ema.py:9-105).

Same interface (`update`, `apply_shadow`, `restore`, `state_dict`, `load_state_dict`, `.shadow` dict keyed by
parameter name).  Differences in mechanism only: the update is done in place with one fused foreach call, and
`apply_shadow`/`restore` exchange values by copy instead of re-pointing `param.data`, so parameters that are
views into a trainer-owned flat buffer (st_dadk_b200.trainer) stay views.
"""
import torch
import torch.nn as nn


class ModelEMA:
    def __init__(self, model: nn.Module, decay: float = 0.999):
        self.decay = decay
        self.model = model
        self.backup = {}
        self.shadow = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}

    def _tracked(self, model=None):
        return [(n, p) for n, p in (model or self.model).named_parameters() if p.requires_grad]

    @torch.no_grad()
    def update(self, model: nn.Module):
        names, params = zip(*self._tracked(model))
        missing = [n for n in names if n not in self.shadow]
        assert not missing, f"Parameter {missing[0]} not in shadow"
        shadows = [self.shadow[n] for n in names]
        torch._foreach_mul_(shadows, self.decay)
        torch._foreach_add_(shadows, [p.detach() for p in params], alpha=1.0 - self.decay)

    @torch.no_grad()
    def apply_shadow(self):
        for n, p in self._tracked():
            self.backup[n] = p.detach().clone()
            p.copy_(self.shadow[n])

    @torch.no_grad()
    def restore(self):
        for n, p in self._tracked():
            p.copy_(self.backup[n])
        self.backup = {}

    def state_dict(self):
        return {"decay": self.decay, "shadow": self.shadow}

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.shadow = state_dict["shadow"]
