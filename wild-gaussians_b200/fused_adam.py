"""``FusedAdam``: ``torch.optim.Adam`` with ``step()`` running as ONE kernel over all parameter tensors.

wild-gaussians builds ``torch.optim.Adam(groups, lr=1.0, eps=1e-15)`` over ten parameter groups (method.py:1033-1049) and
calls ``optimizer.step()`` once per training iteration (method.py:2019).  This class IS a ``torch.optim.Adam`` -- same
constructor, same ``param_groups``, same per-parameter state (``step`` / ``exp_avg`` / ``exp_avg_sq``, so the reference's
optimizer surgery in ``_resize_parameter`` / ``prune`` / ``cat`` (method.py:1088-1110, 1262-1330) and ``state_dict``
round trips keep working) -- only ``step`` is replaced: the fp32 arithmetic of ``torch/optim/adam.py:_multi_tensor_adam``
for every element of every tensor in a single HBM-bound pass (``csrc/adam.cu``, C ABI ``gsr_adam_step``).

    opt = FusedAdam(groups, lr=1.0, eps=1e-15)          # drop-in for the constructor call, or
    adopt(model.optimizer)                               # re-class an existing torch.optim.Adam in place

Unsupported settings (amsgrad, maximize, capturable, differentiable, decoupled weight decay, tensor lr / betas, sparse or
non-fp32 / non-CUDA / non-contiguous parameters) fall through to torch's own ``step`` -- there is no silent numeric change.
"""
from __future__ import annotations

import ctypes

import torch

from diff_gaussian_rasterization import _C

__all__ = ["FusedAdam", "adopt"]


def _fusable_group(group) -> bool:
    if group.get("amsgrad") or group.get("maximize") or group.get("capturable") or group.get("differentiable") or \
            group.get("decoupled_weight_decay"):
        return False
    if isinstance(group["lr"], torch.Tensor) or any(isinstance(b, torch.Tensor) for b in group["betas"]):
        return False
    return True


def _fusable_param(p) -> bool:
    g = p.grad
    return (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and g is not None and not g.is_sparse and
            g.dtype == torch.float32 and g.is_contiguous() and g.device == p.device)


class FusedAdam(torch.optim.Adam):
    """See the module docstring.  ``zero_grads_in_step=True`` additionally leaves every gradient all-zero (the in-place
    counterpart of ``zero_grad(set_to_none=False)``) without a second pass over them."""

    zero_grads_in_step = False

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None or not all(_fusable_group(g) for g in self.param_groups):
            return super().step(closure)
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not _fusable_param(p):
                    return super().step(closure)
        by_key: dict = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:                         # torch/optim/adam.py:_init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous() and m.device == p.device and v.device == p.device and
                        m.dtype == torch.float32 and v.dtype == torch.float32):
                    return super().step(closure)
                by_key.setdefault((p.device, float(beta1), float(beta2), float(group["eps"])), []).append((group, p, state))
        for (dev, beta1, beta2, eps), items in by_key.items():
            stream = torch.cuda.current_stream(dev).cuda_stream
            with torch.cuda.device(dev):
                for i in range(0, len(items), _C.ADAM_MAX_SEGMENTS):
                    part = items[i:i + _C.ADAM_MAX_SEGMENTS]
                    segs = (_C.GsrAdamSegment * len(part))()
                    for s, (group, p, state) in zip(segs, part):
                        s.param, s.grad = p.data_ptr(), p.grad.data_ptr()
                        s.exp_avg, s.exp_avg_sq = state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr()
                        s.n, s.step = p.numel(), int(state["step"].item()) + 1      # the count AFTER this update
                        s.lr, s.weight_decay = float(group["lr"]), float(group["weight_decay"])
                    _C._check(_C._lib.gsr_adam_step(segs, len(part), beta1, beta2, eps, int(bool(self.zero_grads_in_step)),
                                                    ctypes.c_void_p(stream)), "gsr_adam_step")
                    for _, _, state in part:                # only once the launch was accepted
                        state["step"] += 1                  # a CPU scalar tensor, like torch keeps it
        return None


def adopt(optimizer: torch.optim.Adam, zero_grads_in_step: bool = False) -> torch.optim.Adam:
    """Turn an existing ``torch.optim.Adam`` instance (e.g. ``model.optimizer`` built by method.py:1049) into a
    ``FusedAdam`` in place: its param_groups, state and hooks are kept, only ``step`` changes."""
    if type(optimizer) is torch.optim.Adam:
        optimizer.__class__ = FusedAdam
    elif not isinstance(optimizer, FusedAdam):
        raise TypeError(f"adopt() expects a torch.optim.Adam, got {type(optimizer).__name__}")
    optimizer.zero_grads_in_step = bool(zero_grads_in_step)
    if hasattr(optimizer, "_patch_step_function"):          # torch wraps `step` per class (profiler / step hooks)
        optimizer._patch_step_function()
    return optimizer
