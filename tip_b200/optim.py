"""Adam for the TIP training loop (tip.py:21-30: `torch.optim.Adam(model.parameters(), lr=settings.lr)`) as ONE
CUDA launch over all parameter tensors (tipb_adam_step, csrc/adam.cu).  Same update rule and defaults as
torch.optim.Adam (betas 0.9/0.999, eps 1e-8, no weight decay, no amsgrad); the step counter lives on the device, so
`step()` can be captured in a CUDA graph.  State layout (`exp_avg`, `exp_avg_sq`, `step`) matches torch's, so
`state_dict()`s are interchangeable with `torch.optim.Adam(capturable=True)`."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    def _init_group(self, group):
        tensors = []
        for p in group["params"]:
            if p.grad is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.is_sparse:
                raise _lib.TipbError("tip_b200.optim.Adam takes dense fp32 CUDA parameters only (there is no CPU path)")
            st = self.state[p]
            if len(st) == 0:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise _lib.TipbError("tip_b200.optim.Adam needs contiguous parameters and gradients")
            tensors.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
        if "step" not in group:       # one device counter per group (all its tensors step together)
            dev = group["params"][0].device
            group["step"] = torch.zeros(1, dtype=torch.float32, device=dev)
        for p in group["params"]:     # torch's per-parameter `step` entry, aliasing the group counter
            if len(self.state[p]) and "step" not in self.state[p]:
                self.state[p]["step"] = group["step"]
        return tensors

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        for group in self.param_groups:
            tensors = self._init_group(group)
            n = len(tensors)
            if n == 0:
                continue
            arr = lambda k: (C.c_void_p * n)(*[t[k].data_ptr() for t in tensors])
            numel = (C.c_int64 * n)(*[t[0].numel() for t in tensors])
            b1, b2 = group["betas"]
            check(L.tipb_adam_step(n, arr(0), arr(1), arr(2), arr(3), numel, float(group["lr"]), float(b1), float(b2),
                                   float(group["eps"]), group["step"].data_ptr(), stream()), "adam_step")
        return loss
