"""Adam as the reference configures it (model/model.py:39-40: torch.optim.Adam(params, lr) with torch's defaults: betas (0.9, 0.999),
eps 1e-8, no weight decay), as ONE native launch over all parameter tensors (csrc/train_kernels.cuh adam_kernel) instead of torch's
per-tensor / foreach kernels."""
import numpy as np
import torch

from . import _native


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if weight_decay != 0:
            raise NotImplementedError("FusedAdam: weight_decay is not used by the reference and not implemented")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        """grad_scale multiplies every gradient first (1 / world_size after a summing all-reduce)."""
        assert closure is None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            st = self._tables.setdefault(id(group), {})       # device table cache: not part of the optimizer's state_dict
            group["step"] = int(group.get("step", 0)) + 1    # saved / restored with the param group (bias correction survives a resume)
            for p in ps:
                s = self.state[p]
                if not s:
                    s["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            # the device table is rebuilt (one synchronous H2D copy, which would also drain the GPU queue) only when a tensor moved: with
            # gradients that live in a fixed arena (parallel.GradientBuckets) that is once
            key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
            if st.get("key") != key:
                rows = np.empty((len(ps), 5), dtype=np.int64)
                for i, p in enumerate(ps):
                    s, g = self.state[p], p.grad
                    if not (p.is_cuda and p.is_contiguous() and g.is_contiguous() and p.dtype == torch.float32 and g.dtype == torch.float32):
                        raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters and gradients")
                    rows[i] = (p.data_ptr(), g.data_ptr(), s["exp_avg"].data_ptr(), s["exp_avg_sq"].data_ptr(), p.numel())
                st["table"] = torch.from_numpy(rows).to(ps[0].device)
                st["key"] = key
            dev = ps[0].device
            table = st["table"]
            with torch.cuda.device(dev):
                _native.adam_step(table, len(ps), group["lr"], group["betas"][0], group["betas"][1], group["eps"], group["step"], grad_scale)
            self._bump(ps)
        return None

    @staticmethod
    def _bump(ps):
        # the native kernel wrote through raw pointers: bump the tensors' version counters so that UNet._weights_version() sees the update
        inc = getattr(torch.autograd.graph, "increment_version", None)
        for p in ps:
            if inc is not None:
                inc(p)
            else:
                p.add_(0)
