"""torch_scatter 2.0.8 `scatter(..., reduce=)` semantics on plain torch (CPU).

sum : zeros(dim_size).scatter_add_(index, src)            -> sequential edge-order adds on CPU
mean: sum / count with count clamped to >= 1              (isolated rows stay 0)
"""
import torch


def _expand(index, src):
    # broadcast a 1-D index along dim 0 of src
    shape = [-1] + [1] * (src.dim() - 1)
    return index.view(shape).expand_as(src)


def scatter_sum(src, index, dim_size):
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return out.scatter_add_(0, _expand(index, src), src)


def scatter_mean(src, index, dim_size):
    total = scatter_sum(src, index, dim_size)
    ones = torch.ones(index.shape[0], dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, dim_size)
    count[count < 1] = 1
    shape = [-1] + [1] * (src.dim() - 1)
    return total.true_divide_(count.view(shape))


def scatter(src, index, dim_size, reduce):
    if reduce in ("add", "sum"):
        return scatter_sum(src, index, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim_size)
    raise NotImplementedError(reduce)
