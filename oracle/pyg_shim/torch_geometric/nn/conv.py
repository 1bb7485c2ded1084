"""TEST INFRASTRUCTURE ONLY -- restatement of the PyG 2.0.1 pieces the reference calls.

MessagePassing.propagate (flow='source_to_target', node_dim=-2, edge_index a
[2,E] LongTensor):
  * x_j  = x.index_select(0, edge_index[0])   ("_j" = source)
  * x_i  = x.index_select(0, edge_index[1])   ("_i" = target)
  * aggregation index = edge_index[1], dim_size = x.size(0)
  * message()/update() receive the remaining keyword arguments BY PARAMETER NAME
  * aggregate = torch_scatter.scatter(msg, index, dim_size, reduce=aggr)

GCNConv (normalize, add_self_loops, bias; improved=False):
  gcn_norm: drop existing self loops, append exactly one loop per node AFTER the
  other edges, deg = in-degree (over edge_index[1]) incl. the loop,
  w = deg^-1/2[row] * deg^-1/2[col] with inf -> 0; cached after the first call;
  x' = x @ lin.weight.T (sparse-COO x allowed); out = scatter_add(w * x'[row] -> col) + bias;
  lin.weight glorot-uniform, bias zeros.
"""
import inspect
import math

import torch
from torch import nn

from ..scatter import scatter


class MessagePassing(nn.Module):
    _special = {"edge_index", "edge_index_i", "edge_index_j", "size", "size_i", "size_j",
                "index", "ptr", "dim_size"}

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        assert flow == "source_to_target" and node_dim == -2
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim

    @staticmethod
    def _param_names(fn, skip):
        names = list(inspect.signature(fn).parameters.keys())
        return names[skip:]

    def propagate(self, edge_index, size=None, **kwargs):
        assert edge_index.dtype == torch.long and edge_index.dim() == 2 and edge_index.size(0) == 2
        src, dst = edge_index[0], edge_index[1]

        n_nodes = None
        msg_kwargs = {}
        for name in self._param_names(self.message, 0):
            if name.endswith("_j") or name.endswith("_i"):
                base = name[:-2]
                if base == "edge_index":
                    msg_kwargs[name] = src if name.endswith("_j") else dst
                    continue
                data = kwargs[base]
                n_nodes = data.size(0) if n_nodes is None else n_nodes
                msg_kwargs[name] = data.index_select(0, src if name.endswith("_j") else dst)
            elif name == "edge_index":
                msg_kwargs[name] = edge_index
            else:
                msg_kwargs[name] = kwargs[name]
        if n_nodes is None:
            n_nodes = int(edge_index.max()) + 1

        out = self.message(**msg_kwargs)
        out = self.aggregate(out, dst, dim_size=n_nodes)

        upd_kwargs = {name: kwargs[name] for name in self._param_names(self.update, 1)}
        return self.update(out, **upd_kwargs)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, dim_size=None):
        return scatter(inputs, index, dim_size, self.aggr)

    def update(self, inputs):
        return inputs


def gcn_norm(edge_index, num_nodes, dtype):
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    loops = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    row = torch.cat([row[keep], loops])
    col = torch.cat([col[keep], loops])
    weight = torch.ones(row.numel(), dtype=dtype, device=row.device)
    deg = torch.zeros(num_nodes, dtype=dtype, device=row.device).scatter_add_(0, col, weight)
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return torch.stack([row, col]), dis[row] * weight * dis[col]


class _Linear(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))
        self.weight.data.uniform_(-bound, bound)

    def forward(self, x):
        if x.is_sparse:
            return torch.sparse.mm(x, self.weight.t())
        return x @ self.weight.t()


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, cached=False, **kwargs):
        super().__init__(aggr="add", **kwargs)
        self.in_channels, self.out_channels, self.cached = in_channels, out_channels, cached
        self._cached_edge_index = None
        self.lin = _Linear(in_channels, out_channels)      # PyG's Linear.__init__ initialises itself ...
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin.reset_parameters()                         # ... and GCNConv.reset_parameters() draws it again

    def forward(self, x, edge_index):
        cache = self._cached_edge_index
        if cache is None:
            cache = gcn_norm(edge_index, x.size(0), self.lin.weight.dtype)
            if self.cached:
                self._cached_edge_index = cache
        edge_index, edge_weight = cache
        x = self.lin(x)
        out = self.propagate(edge_index, x=x, edge_weight=edge_weight)
        out += self.bias
        return out

    def message(self, x_j, edge_weight):
        return edge_weight.view(-1, 1) * x_j
