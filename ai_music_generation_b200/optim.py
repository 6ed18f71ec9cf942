"""FusedAdamW — what GPT.configure_optimizers returns (reference: nanoGPT/model.py:263-287 builds
torch.optim.AdamW(fused=True) over two groups; nanoGPT/train.py:285-287,354-357 drives it).

Duck-types torch.optim.AdamW (param_groups with writable 'lr', step(), zero_grad(set_to_none=True),
state_dict()/load_state_dict() in the torch layout: state[i] = {'step','exp_avg','exp_avg_sq'}), but the update is
two launches of the sm_100a arena kernel (decay region, no-decay region) that also apply the pending
clip_grad_norm_ coefficient from a device scalar and refresh the bf16 weight shadow used by the GEMMs.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *, model):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        object.__setattr__(self, "_model", model)
        a = model._arena
        decay = [p for p in a["params"] if p.dim() >= 2]
        nodecay = [p for p in a["params"] if p.dim() < 2]
        if len(self.param_groups) != 2 or [id(p) for p in self.param_groups[0]["params"]] != [id(p) for p in decay] \
                or [id(p) for p in self.param_groups[1]["params"]] != [id(p) for p in nodecay]:
            raise ValueError("FusedAdamW expects the two parameter groups GPT.configure_optimizers builds "
                             "(all dim>=2 tensors, then all dim<2 tensors)")
        self._m = None
        self._v = None
        self._step_t = torch.tensor(0.0)
        self._steps = 0

    # -- state arenas ------------------------------------------------------------------------------------------
    def _ensure_state(self):
        a = self._model._arena
        flat = a["flat"]
        if self._m is None or self._m.device != flat.device or self._m.numel() != flat.numel():
            old = {id(p): self.state.get(p) for p in a["params"]}
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            for p, o in zip(a["params"], a["offs"]):
                m = self._m[o:o + p.numel()].view(p.shape)
                v = self._v[o:o + p.numel()].view(p.shape)
                st = old.get(id(p))
                if st:  # carry over (load_state_dict before the first step, or a device move)
                    m.copy_(st["exp_avg"])
                    v.copy_(st["exp_avg_sq"])
                self.state[p] = {"step": self._step_t, "exp_avg": m, "exp_avg_sq": v}

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self._model
        a = model._arena
        if a["grad"] is None or a["params"][0].grad is None:
            return loss
        if model._grad_sync is not None:
            model._grad_sync.wait()
        model._ensure_device_state()
        self._ensure_state()
        self._steps += 1
        self._step_t += 1
        ss, max_norm = model._pending_clip if model._pending_clip is not None else (None, 0.0)
        nd, total = a["n_decay"], a["total"]
        for gi, (lo, hi) in enumerate(((0, nd), (nd, total))):
            if hi <= lo:
                continue
            g = self.param_groups[gi]
            b1, b2 = g["betas"]
            ops.adamw(a["flat"][lo:hi], a["grad"][lo:hi], self._m[lo:hi], self._v[lo:hi], a["shadow"][lo:hi],
                      lr=float(g["lr"]), beta1=float(b1), beta2=float(b2), eps=float(g["eps"]),
                      weight_decay=float(g["weight_decay"]), step=self._steps, sumsq=ss, max_norm=max_norm)
        model._shadow_fresh = True
        model._pending_clip = None
        return loss

    def zero_grad(self, set_to_none: bool = True):
        a = self._model._arena
        if set_to_none:
            for p in a["params"]:
                p.grad = None
        elif a["grad"] is not None:
            a["grad"].zero_()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = 0
        for st in self.state.values():
            if "step" in st:
                steps = int(float(st["step"]))
                break
        self._steps = steps
        self._step_t = torch.tensor(float(steps))
        self._m = None  # rebuilt (and filled from the loaded per-tensor state) on the next step
        self._v = None
        if self._model._arena["flat"].is_cuda:
            self._ensure_state()
