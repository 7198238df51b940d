"""TEST INFRASTRUCTURE -- CPU restatement of the optimizer step wild-gaussians runs every iteration.

The reference calls ``self.model.optimizer.step()`` (wildgaussians/method.py:2019) on
``torch.optim.Adam(groups, lr=1.0, eps=1e-15)`` (method.py:1049; per-group ``lr``, ``weight_decay`` only on the
appearance-embedding table, :1040).  The algorithm itself lives in the reference's third-party dependency **torch**
(unpinned ``torch`` requirement of the reference; this image: torch 2.11), function
``torch/optim/adam.py:_multi_tensor_adam`` (the "foreach" path torch selects on CUDA).  This file restates that function's
published sequence of fp32 operations in numpy; it is pinned by ``tests/test_adam.py`` against ``torch.optim.Adam`` itself
(imported here, CPU) and is what the CUDA kernel (``csrc/adam.cu``) is compared with.  Only tests / smoke / bench may import it.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """One Adam update of one tensor (numpy fp32 arrays, returned as new arrays); ``step`` is the 1-based count AFTER the
    increment torch performs first (adam.py: ``torch._foreach_add_(device_state_steps, 1)``)."""
    p, g, m, v = (np.asarray(x, dtype=f32) for x in (param, grad, exp_avg, exp_avg_sq))
    if weight_decay != 0:
        g = (g + f32(weight_decay) * p).astype(f32)                       # _foreach_add(grads, params, alpha=weight_decay)
    m = (m + f32(1 - beta1) * (g - m)).astype(f32)                        # _foreach_lerp_(exp_avgs, grads, 1 - beta1)
    v = (v * f32(beta2)).astype(f32)                                      # _foreach_mul_(exp_avg_sqs, beta2)
    v = (v + f32(1 - beta2) * (g * g).astype(f32)).astype(f32)            # _foreach_addcmul_(exp_avg_sqs, grads, grads, 1 - beta2)
    bc1 = 1 - beta1 ** step                                               # Python floats (doubles), as in adam.py
    bc2 = 1 - beta2 ** step
    step_size = (lr / bc1) * -1
    bc2_sqrt = bc2 ** 0.5
    d = (np.sqrt(v).astype(f32) / f32(bc2_sqrt)).astype(f32)              # _foreach_sqrt, _foreach_div_
    d = (d + f32(eps)).astype(f32)                                        # _foreach_add_(., eps)
    p = (p + f32(step_size) * (m / d).astype(f32)).astype(f32)            # _foreach_addcdiv_(params, exp_avgs, ., step_size)
    return p, m, v
