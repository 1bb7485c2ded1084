"""Adam for the TIP training loop (tip.py:21-30: `torch.optim.Adam(model.parameters(), lr=settings.lr)`) as ONE
CUDA launch over all parameter tensors (tipb_adam_step, csrc/adam.cu).  Same update rule and defaults as
torch.optim.Adam (betas 0.9/0.999, eps 1e-8, no weight decay, no amsgrad); the step counter lives on the device, so
`step()` can be captured in a CUDA graph.

State layout matches torch's (`state[p] = {step, exp_avg, exp_avg_sq}`, nothing extra in `param_groups`), so
`state_dict()`s are interchangeable with `torch.optim.Adam`: all tensors of a group step together, the kernel reads
ONE device counter per group, and every `state[p]['step']` of the group aliases it.  After `load_state_dict` (which
deep-copies the state and may leave `step` on the CPU or as a Python number) the counter is re-derived from the
loaded per-parameter values and moved to the parameters' device before the next launch."""
import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, stream


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if lr < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        # the remaining keys are torch.optim.Adam's own (fixed at the values this kernel implements) so that a
        # state_dict saved here loads into torch.optim.Adam and back
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                                      foreach=None, capturable=True, differentiable=False, fused=None,
                                      decoupled_weight_decay=False))
        self._counters = {}      # group index -> device step counter (float32, 0-dim); NOT part of the state_dict

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._counters = {}      # re-derive from the loaded state[p]['step'] at the next step()

    def _group_counter(self, gi, group):
        """the group's device counter; (re)created from the per-parameter `step` entries when missing"""
        ctr = self._counters.get(gi)
        if ctr is None:
            dev = None
            loaded = []
            for p in group["params"]:
                dev = p.device if dev is None else dev
                st = self.state.get(p)
                if st and "step" in st:
                    loaded.append(float(st["step"]))         # CPU / CUDA tensor or a number, as torch may store it
            if dev is None or dev.type != "cuda":
                raise _lib.TipbError("tip_b200.optim.Adam takes dense fp32 CUDA parameters only (there is no CPU path)")
            if loaded and min(loaded) != max(loaded):
                raise _lib.TipbError("tip_b200.optim.Adam: parameters of one group carry different step counts")
            ctr = torch.full((), loaded[0] if loaded else 0.0, dtype=torch.float32, device=dev)   # 0-dim like torch's
            self._counters[gi] = ctr
        return ctr

    def _init_group(self, gi, group):
        tensors = []
        ctr = self._group_counter(gi, group)
        for p in group["params"]:
            if p.grad is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32 or p.grad.is_sparse:
                raise _lib.TipbError("tip_b200.optim.Adam takes dense fp32 CUDA parameters only (there is no CPU path)")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            for k in ("exp_avg", "exp_avg_sq"):               # a state loaded with map_location='cpu'
                if st[k].device != p.device or st[k].dtype != torch.float32:
                    st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()
            st["step"] = ctr                                  # torch's per-parameter entry, aliasing the group counter
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise _lib.TipbError("tip_b200.optim.Adam needs contiguous parameters and gradients")
            tensors.append((p, p.grad, st["exp_avg"], st["exp_avg_sq"]))
        return tensors, ctr

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        for gi, group in enumerate(self.param_groups):
            tensors, ctr = self._init_group(gi, group)
            n = len(tensors)
            if n == 0:
                continue
            arr = lambda k: (C.c_void_p * n)(*[t[k].data_ptr() for t in tensors])
            numel = (C.c_int64 * n)(*[t[0].numel() for t in tensors])
            b1, b2 = group["betas"]
            if group.get("weight_decay", 0) or group.get("amsgrad", False) or group.get("maximize", False):
                raise _lib.TipbError("tip_b200.optim.Adam implements plain Adam (no weight decay / amsgrad / maximize)")
            with torch.cuda.device(ctr.device):
                check(L.tipb_adam_step(n, arr(0), arr(1), arr(2), arr(3), numel, float(group["lr"]), float(b1), float(b2),
                                       float(group["eps"]), ctr.data_ptr(), stream(ctr.device)), "adam_step")
        return loss
