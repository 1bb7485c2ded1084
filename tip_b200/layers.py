"""Drop-in for the module API of the reference's src/layers.py, backed by libtipb200 (sm_100a).

Same class names, constructor arguments, forward signatures, parameter names/shapes and
initialisation draw order as the reference (so state_dicts and seeds carry over); the
device work of every forward/backward goes through the C ABI in include/tipb200.h.
There is no PyG, no torch_scatter, no Triton and no CPU path: CPU tensors are rejected.

Reference lines (relative to the reference root) are cited per class.
"""
import math
import pickle
import weakref

import numpy as np
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import Parameter as Param

from . import neg_sampling as _ns
from . import ops
from ._lib import TipbError
from .neg_sampling import typed_negative_sampling
from .utils import *  # noqa: F401,F403  (the reference re-exports src.utils the same way)
from .utils import is_sparse_identity, process_edges

torch.manual_seed(1111)   # src/layers.py:13
np.random.seed(1111)      # src/layers.py:14 (the device stream is seeded 1111 too, see neg_sampling._state)
EPS = 1e-13               # src/layers.py:15
SERIAL_STREAMS = False    # measurement aid (bench.py): keep every kernel of a step on the caller's stream
# priority of the sampler's side stream in TIP.forward.  The sampler chain (0.4 ms) is shorter than the encoder chain it runs
# beside (0.5 ms), so it gets the default (lowest) priority; a caller that runs the step on a high-priority stream
# (bench.py does: torch.cuda.Stream(priority=-1)) lets the encoder's kernels take SM slots first and the sampler fill the
# gaps: 1.98 -> 1.86 ms per step (profiles/r03m_*).  TIPB_SIDE_PRIORITY overrides it for measurements.
SIDE_PRIORITY = int(os.environ.get("TIPB_SIDE_PRIORITY", "0"))
ENCODER_FIRST = os.environ.get("TIPB_ENCODER_FIRST", "1") != "0"
# FMEncoder: run the second P-P GCN layer only into the rows the hierarchy conv reads (exact; TIPB_PP_READ_ROWS=0: all rows)
PP_READ_ROWS_ONLY = os.environ.get("TIPB_PP_READ_ROWS", "1") != "0"


def _require_cuda(t, who):
    if not t.is_cuda:
        raise TipbError(f"{who}: CUDA tensors only -- tip_b200 has no CPU fallback")


# =============================================================================== convolutions
class _RGCNBase(nn.Module):
    """shared parameters/initialisation of MyRGCNConv and MyRGCNConv2 (src/layers.py:35-74, 115-155)"""

    def __init__(self, in_channels, out_channels, num_relations, num_bases, after_relu, bias=False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases, self.after_relu = num_relations, num_bases, after_relu
        self.basis = Param(torch.Tensor(num_bases, in_channels, out_channels))
        self.att = Param(torch.Tensor(num_relations, num_bases))
        self.root = Param(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Param(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        # draw order att -> root -> basis is part of the seed contract
        self.att.data.normal_(std=1 / np.sqrt(self.num_bases))
        std = 2 / self.in_channels if self.after_relu else 1 / np.sqrt(self.in_channels)
        self.root.data.normal_(std=std)
        self.basis.data.normal_(std=std)
        if self.bias is not None:
            self.bias.data.zero_()

    def _conv(self, x, edge_index, edge_type, range_list, relu=False):
        _require_cuda(x, type(self).__name__)
        assert edge_index.dtype == torch.long and edge_index.dim() == 2 and edge_index.size(0) == 2
        n = x.size(0)
        kw = dict(edge_type=edge_type, range_list=range_list)
        plan_dst = ops.cached_plan(edge_index, n, self.num_relations, by_src=False, **kw)
        plan_src = ops.cached_plan(edge_index, n, self.num_relations, by_src=True, **kw)
        return ops.rgcn_conv(x, self.basis, self.att, self.root, plan_dst, plan_src, bias=self.bias, relu=relu)

    def __repr__(self):
        return "{}({}, {}, num_relations={})".format(type(self).__name__, self.in_channels, self.out_channels,
                                                     self.num_relations)


class MyRGCNConv(_RGCNBase):
    """src/layers.py:21-99 -- edge_type in any order (the typed CSR is built by a stable GPU sort)."""

    def forward(self, x, edge_index, edge_type):
        return self._conv(x, edge_index, edge_type, None)


class MyRGCNConv2(_RGCNBase):
    """src/layers.py:102-193 -- edges sorted by relation with range_list[r] = (start, end).
    Like the reference, `edge_type` is accepted and not read."""

    def forward(self, x, edge_index, edge_type, range_list, _fused_relu=False):
        return self._conv(x, edge_index, None, range_list.to(torch.long), relu=_fused_relu)


class MyHierarchyConv(nn.Module):
    """directed mean-aggregation protein -> drug (src/layers.py:196-247)"""

    def __init__(self, in_dim, out_dim, unigue_source_num, unique_target_num, is_after_relu=True, is_bias=False,
                 **kwargs):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.unique_source_num, self.unique_target_num = unigue_source_num, unique_target_num
        self.is_after_relu = is_after_relu
        self.weight = Param(torch.Tensor(in_dim, out_dim))
        if is_bias:
            self.bias = Param(torch.Tensor(out_dim))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        std = 1 / np.sqrt(self.in_dim) if self.is_after_relu else 2 / np.sqrt(self.in_dim)
        self.weight.data.normal_(std=std)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, x, edge_index, range_list):
        _require_cuda(x, "MyHierarchyConv")
        if self.bias is not None:
            # the reference's `if self.bias:` raises for a multi-element tensor; same outcome here
            raise RuntimeError("Boolean value of Tensor with more than one value is ambiguous")
        n = self.unique_source_num + self.unique_target_num
        assert x.size(0) == n, "x must hold unique_source_num + unique_target_num rows"
        plan_dst = ops.cached_plan(edge_index, n, 1, by_src=False)
        plan_src = ops.cached_plan(edge_index, n, 1, by_src=True)
        out = ops.hier_conv(x, self.weight, plan_dst, plan_src, self.unique_source_num, self.unique_target_num)
        assert out.shape[0] == self.unique_target_num
        return out

    def __repr__(self):
        return "{}({}, {}".format(type(self).__name__, self.in_dim, self.out_dim)


class _Lin(nn.Module):
    """holds `weight` [out, in] under the name torch_geometric's GCNConv uses (`lin.weight`)"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = Param(torch.Tensor(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self):
        bound = math.sqrt(6.0 / (self.weight.size(0) + self.weight.size(1)))   # glorot
        self.weight.data.uniform_(-bound, bound)


class GCNConv(nn.Module):
    """torch_geometric 2.0.1 GCNConv(in, out, cached=True) as used by PPEncoder (src/layers.py:386-387):
    parameter names `lin.weight`, `bias`; symmetric normalisation with remaining self loops; the
    normalised graph of the FIRST call is cached (here: the typed CSR + deg^-1/2)."""

    def __init__(self, in_channels, out_channels, cached=False, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels, self.cached = in_channels, out_channels, cached
        self.lin = _Lin(in_channels, out_channels)       # PyG's Linear initialises itself in __init__ ...
        self.bias = Param(torch.Tensor(out_channels))
        self._cache = None
        self._identity_ok = {}
        self.reset_parameters()                          # ... and GCNConv.reset_parameters draws it again

    def reset_parameters(self):
        self.lin.reset_parameters()
        self.bias.data.zero_()
        self._cache = None

    def _graph(self, edge_index, n):
        c = self._cache
        if c is not None:
            # PyG's cached=True keeps the normalised graph of the FIRST call whatever is passed later; so does this
            # module -- except when that very tensor was overwritten in place since (its version counter moved):
            # every other index structure of the package re-keys on `_version`, and so does this one.
            if c[3]() is not edge_index or c[4] == edge_index._version:
                return c[:3]
        plan_dst = ops.cached_plan(edge_index, n, 1, by_src=False, drop_self_loops=True)
        plan_src = ops.cached_plan(edge_index, n, 1, by_src=True, drop_self_loops=True)
        graph = (plan_dst, plan_src, ops.gcn_norm(plan_dst))
        if self.cached:
            self._cache = graph + (weakref.ref(edge_index), edge_index._version)
        return graph

    def __getstate__(self):      # torch.save(model) (tip.py:36): parameters travel, device-side caches do not
        state = dict(self.__dict__)
        state["_cache"], state["_identity_ok"] = None, {}
        return state

    def _linear(self, x):
        if x.is_sparse:
            key = (x._indices().data_ptr(), tuple(x.shape))
            if key not in self._identity_ok:
                self._identity_ok[key] = is_sparse_identity(x)     # one-time check (host sync at setup)
            if self._identity_ok[key]:
                return ops.transpose2d(self.lin.weight)            # I @ W^T
            return ops.sparse_matmul(x, ops.transpose2d(self.lin.weight))   # general sparse features
        return ops.matmul(x, self.lin.weight, trans_b=True)

    def forward(self, x, edge_index, _fused_relu=False, _pad_rows=0, _row_plans=None):
        assert edge_index.dtype == torch.long and edge_index.dim() == 2 and edge_index.size(0) == 2
        if not edge_index.is_cuda:
            raise TipbError("GCNConv: CUDA tensors only -- tip_b200 has no CPU fallback")
        plan_dst, plan_src, dis = self._graph(edge_index, x.size(0))
        if _row_plans is not None:
            # the caller reads only some rows of the result (FMEncoder: the proteins with a P->D edge): the plans of the
            # edges INTO those rows, with the normalisation of the whole graph.  The other rows come out as their self
            # term + bias and take no gradient; the rows that are read are bit for bit the same, the gradients equal up to
            # the order of their sums (the skipped addends are exact zeros).
            plan_dst, plan_src = _row_plans
        return ops.gcn_spmm(self._linear(x), self.bias, plan_dst, plan_src, dis, relu=_fused_relu, pad_rows=_pad_rows)


# =============================================================================== encoders
class PPEncoder(nn.Module):
    """2 x GCNConv + ReLU on the protein-protein graph (src/layers.py:380-395)"""

    def __init__(self, in_dim, hid1=32, hid2=16):
        super().__init__()
        self.out_dim = hid2
        self.conv1 = GCNConv(in_dim, hid1, cached=True)
        self.conv2 = GCNConv(hid1, hid2, cached=True)

    def forward(self, x, edge_index, _pad_rows=0, _row_plans=None):
        x = self.conv1(x, edge_index, _fused_relu=True)      # bias + ReLU fused into the SpMM epilogue
        return self.conv2(x, edge_index, _pad_rows=_pad_rows, _row_plans=_row_plans)   # (+ zero rows below: the caller's cat with `hdrug`)


class FMEncoder(nn.Module):
    """P-P GCN -> P->D hierarchy -> drug embedding (cat | add) -> 2 x R-GCN (src/layers.py:471-553)"""

    def __init__(self, device, in_dim_drug, num_dd_et, in_dim_prot, uni_num_prot, uni_num_drug, prot_drug_dim=64,
                 num_base=32, n_embed=64, n_hid1=32, n_hid2=16, mod="cat"):
        super().__init__()
        self.num_et, self.out_dim = num_dd_et, n_hid2
        self.uni_num_drug, self.uni_num_prot = uni_num_drug, uni_num_prot
        self.mod = mod
        assert mod in {"add", "cat"}
        if mod == "add":
            assert n_embed == prot_drug_dim
        self.pp_encoder = PPEncoder(in_dim_prot)
        self.embed = Param(torch.Tensor(in_dim_drug, n_embed))
        self.hgcn = MyHierarchyConv(self.pp_encoder.out_dim, prot_drug_dim, uni_num_prot, uni_num_drug)
        self.hdrug = torch.zeros((self.uni_num_drug, self.pp_encoder.out_dim)).to(device)
        self._hdrug_version = self.hdrug._version     # still all zeros while untouched: the cat below is then fused
        rgcn_in_dim = n_embed + self.hgcn.out_dim if mod == "cat" else n_embed
        self.rgcn1 = MyRGCNConv2(rgcn_in_dim, n_hid1, num_dd_et, num_base, after_relu=False)
        self.rgcn2 = MyRGCNConv2(n_hid1, n_hid2, num_dd_et, num_base, after_relu=True)
        self._identity_ok = {}
        self._row_plan_cache, self._row_plan_seen = None, None
        self.reset_parameters()

    def reset_parameters(self):
        self.embed.data.normal_()

    def __getstate__(self):      # torch.save(model): device-side caches do not travel
        state = dict(self.__dict__)
        state["_row_plan_cache"], state["_row_plan_seen"], state["_identity_ok"] = None, None, {}
        return state

    def _embed(self, x_drug):
        if x_drug.is_sparse:
            key = (x_drug._indices().data_ptr(), tuple(x_drug.shape))
            if key not in self._identity_ok:
                self._identity_ok[key] = is_sparse_identity(x_drug)
            if self._identity_ok[key]:
                return self.embed                                   # I @ embed
            return ops.sparse_matmul(x_drug, self.embed)            # general sparse drug features (data/utils.py:117-132)
        return ops.matmul(x_drug, self.embed)

    def _read_row_plans(self, pp_edge_index, dp_edge_index, n_prot):
        """typed CSRs (by target, by source) of the P-P edges whose TARGET protein is the source of a P->D edge: the
        hierarchy conv reads the second GCN layer's output at those proteins only (19 % of them on the polypharmacy
        shape), so that layer's SpMM and its backward gather skip everything else.  Built once per graph (re-keyed on
        the tensors' versions like every other plan); the filtering itself is set-up work on torch index ops."""
        if not PP_READ_ROWS_ONLY or not pp_edge_index.is_cuda:
            return None
        ident = (pp_edge_index.data_ptr(), tuple(pp_edge_index.shape), dp_edge_index.data_ptr(),
                 tuple(dp_edge_index.shape), int(n_prot))
        versions = (pp_edge_index._version, dp_edge_index._version)
        hit = self._row_plan_cache
        if hit is not None and hit[0] == ident and hit[1] == versions:
            return hit[2]
        # A graph that is rewritten in place before every step (an end-to-end loop that copies its inputs each time) does
        # not amortise this structure: it is (re)built only once the tensors were left alone for a whole step.
        seen, self._row_plan_seen = self._row_plan_seen, (ident, versions)
        if hit is not None and seen != (ident, versions):
            return None
        src = dp_edge_index[0]
        read = torch.zeros(n_prot, dtype=torch.bool, device=pp_edge_index.device)
        read[src[src < n_prot]] = True
        sub = pp_edge_index[:, read[pp_edge_index[1]]].contiguous()
        if sub.shape[1] == 0 or sub.shape[1] == pp_edge_index.shape[1]:
            self._row_plan_cache = (ident, versions, None, None)       # nothing to skip (or nothing to read)
            return None
        plans = tuple(ops.TypedCSR(sub.shape[1], n_prot, 1, sub.device, by_src=by_src, drop_self_loops=True)
                      for by_src in (False, True))
        for p in plans:
            p.build(sub)
            p.check_status()
        self._row_plan_cache = (ident, versions, plans, sub)
        return plans

    def drug_input(self, x_drug, d_norm, x_prot, pp_edge_index, dp_edge_index, dp_range_list):
        """src/layers.py:529-547: everything in front of the two R-GCN layers"""
        rows = self._read_row_plans(pp_edge_index, dp_edge_index, x_prot.size(0))
        if self.hdrug._version == self._hdrug_version and self.hdrug.device == pp_edge_index.device:
            # torch.cat((x_prot, hdrug)) with hdrug == 0: the second GCN layer writes into a buffer with zero rows below
            x_prot = self.pp_encoder(x_prot, pp_edge_index, _pad_rows=self.uni_num_drug, _row_plans=rows)
        else:
            x_prot = self.pp_encoder(x_prot, pp_edge_index, _row_plans=rows)
            x_prot = torch.cat((x_prot, self.hdrug.to(x_prot.device)))
        x_prot = self.hgcn(x_prot, dp_edge_index, dp_range_list)
        return ops.drug_input(self._embed(x_drug), d_norm, x_prot, self.mod)   # / d_norm, then cat | add

    def forward(self, x_drug, dd_edge_index, dd_edge_type, dd_range_list, d_norm, x_prot, pp_edge_index,
                dp_edge_index, dp_range_list):
        x_drug = self.drug_input(x_drug, d_norm, x_prot, pp_edge_index, dp_edge_index, dp_range_list)
        x_drug = self.rgcn1(x_drug, dd_edge_index, dd_edge_type, dd_range_list, _fused_relu=True)
        return self.rgcn2(x_drug, dd_edge_index, dd_edge_type, dd_range_list)


class FMEncoderCat(FMEncoder):
    """the stale 'cat'-only duplicate of FMEncoder (src/layers.py:401-468)"""

    def __init__(self, device, in_dim_drug, num_dd_et, in_dim_prot, uni_num_prot, uni_num_drug, prot_drug_dim=16,
                 num_base=32, n_embed=48, n_hid1=32, n_hid2=16):
        super().__init__(device, in_dim_drug, num_dd_et, in_dim_prot, uni_num_prot, uni_num_drug, prot_drug_dim,
                         num_base, n_embed, n_hid1, n_hid2, mod="cat")


class HierEncoder(nn.Module):
    """source embedding -> hierarchy conv onto the targets (src/layers.py:556-575; the PR-HMP-NN ablation)"""

    def __init__(self, source_dim, embed_dim, target_dim, uni_num_source, uni_num_target):
        super().__init__()
        self.embed = Param(torch.Tensor(source_dim, embed_dim))
        self.hgcn = MyHierarchyConv(embed_dim, target_dim, uni_num_source, uni_num_target)
        self._identity_ok = {}
        self.reset_parameters()

    def reset_parameters(self):
        self.embed.data.normal_()

    def forward(self, source_feat, edge_index, range_list, x_norm):
        if source_feat.is_sparse:
            key = (source_feat._indices().data_ptr(), tuple(source_feat.shape))
            if key not in self._identity_ok:
                self._identity_ok[key] = is_sparse_identity(source_feat)
            x = self.embed if self._identity_ok[key] else ops.sparse_matmul(source_feat, self.embed)
        else:
            x = ops.matmul(source_feat, self.embed)
        x = ops.row_scale(x, x_norm)                      # x / x_norm.view(-1, 1)
        return self.hgcn(x, edge_index, range_list)


# =============================================================================== decoder
class InnerProductDecoder(nn.Module):
    """torch_geometric 2.0.1 InnerProductDecoder (imported at src/layers.py:2, MyGAE's default at :258):
    sigmoid(z_i . z_j) -- the DistMult scorer with a single all-ones relation"""

    def _ones(self, z):
        return torch.ones((1, z.shape[1]), dtype=torch.float32, device=z.device)

    def forward(self, z, edge_index, sigmoid=True):
        _require_cuda(z, "InnerProductDecoder")
        et = torch.zeros(edge_index.shape[1], dtype=torch.long, device=z.device)
        return ops.decoder_score(z, self._ones(z), edge_index, et, sigmoid)

    def forward_all(self, z, sigmoid=True):
        _require_cuda(z, "InnerProductDecoder")
        return ops.decoder_sweep(z, self._ones(z), sigmoid)[0]


class MultiInnerProductDecoder(nn.Module):
    """DistMult scorer z_i^T diag(w_r) z_j (src/layers.py:581-595)"""

    def __init__(self, in_dim, num_et):
        super().__init__()
        self.num_et, self.in_dim = num_et, in_dim
        self.weight = Param(torch.Tensor(num_et, in_dim))
        self.reset_parameters()

    def forward(self, z, edge_index, edge_type, sigmoid=True):
        _require_cuda(z, "MultiInnerProductDecoder")
        return ops.decoder_score(z, self.weight, edge_index, edge_type, sigmoid)

    def sweep(self, z, sigmoid=True):
        """scores of all num_nodes^2 pairs for every relation: [num_et, N, N] (BASELINE.json config 5)"""
        return ops.decoder_sweep(z, self.weight, sigmoid)

    def reset_parameters(self):
        self.weight.data.normal_(std=1 / np.sqrt(self.in_dim))


class NNDecoder(nn.Module):
    """two-layer per-endpoint scorer (src/layers.py:598-637; the DR-NN / PR-HMP-NN ablations):
    sigmoid( relu(z_i w1_l1) . w1_l2[r] + relu(z_j w2_l1) . w2_l2[r] )"""

    def __init__(self, in_dim, num_uni_edge_type, l1_dim=16):
        super().__init__()
        self.l1_dim = l1_dim
        self.w1_l1 = Param(torch.Tensor(in_dim, l1_dim))
        self.w1_l2 = Param(torch.Tensor(num_uni_edge_type, l1_dim))
        self.w2_l1 = Param(torch.Tensor(in_dim, l1_dim))
        self.w2_l2 = Param(torch.Tensor(num_uni_edge_type, l1_dim))
        self.reset_parameters()

    def forward(self, z, edge_index, edge_type):
        _require_cuda(z, "NNDecoder")
        return ops.nn_decoder_score(z, self.w1_l1, self.w1_l2, self.w2_l1, self.w2_l2, edge_index, edge_type)

    def reset_parameters(self):     # draw order w1_l1, w2_l1, w1_l2, w2_l2 (src/layers.py:633-637)
        self.w1_l1.data.normal_()
        self.w2_l1.data.normal_()
        self.w1_l2.data.normal_(std=1 / np.sqrt(self.l1_dim))
        self.w2_l2.data.normal_(std=1 / np.sqrt(self.l1_dim))


# =============================================================================== training wrapper
class MyGAE(nn.Module):
    """src/layers.py:253-258"""

    def __init__(self, encoder, decoder=None):
        super().__init__()
        self.encoder = encoder
        self.decoder = InnerProductDecoder() if decoder is None else decoder


class Setting(object):
    """src/layers.py:260-269"""

    def __init__(self, sp_rate=0.9, lr=0.01, prot_drug_dim=16, n_embed=48, n_hid1=32, n_hid2=16, num_base=32):
        self.sp_rate, self.lr = sp_rate, lr
        self.prot_drug_dim, self.n_embed, self.n_hid1, self.n_hid2, self.num_base = (prot_drug_dim, n_embed, n_hid1,
                                                                                     n_hid2, num_base)


class _Data(object):
    """attribute bag standing in for torch_geometric.data.Data.from_dict(...).to(device)"""

    def __init__(self, d, device):
        for k, v in d.items():
            setattr(self, k, v.to(device) if torch.is_tensor(v) else v)


class TIP(nn.Module):
    """src/layers.py:272-375.  `data_path` is the reference's data_dict.pkl (prepare.py:13-47);
    a ready dict can be passed instead through `data=`."""

    def __init__(self, settings, device, mod="cat", data_path="./data/data_dict.pkl", data=None):
        super().__init__()
        self.mod = mod
        assert mod in {"cat", "add"}
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise TipbError("TIP: CUDA device required -- tip_b200 has no CPU fallback")
        self.settings = settings
        self.data = self._prepare_data(data_path, settings.sp_rate, data)
        self._prepare_model()
        self._neg_plan = None
        self._neg_index_buf = None
        self._neg_packed = None
        self._side = None

    def _prepare_data(self, data_path, sp_rate, data_dict):
        if data_dict is None:
            with open(data_path, "rb") as f:
                data_dict = pickle.load(f)
        data_dict = dict(data_dict)
        if sp_rate != 0.9:
            # src/layers.py:282-286: re-split with the requested rate.  The split runs on the device and draws from the
            # device-side numpy-compatible stream (the reference draws from np.random at this point)
            with torch.cuda.device(self.device):
                raw = [torch.as_tensor(e).to(self.device) for e in data_dict["dd_edge_index"]]
                (data_dict["dd_train_idx"], data_dict["dd_train_et"], data_dict["dd_train_range"],
                 data_dict["dd_test_idx"], data_dict["dd_test_et"], data_dict["dd_test_range"]) = process_edges(raw, p=sp_rate)
        data = _Data(data_dict, self.device)
        data.dd_train_range = data.dd_train_range.to(torch.long)
        data.dd_test_range = data.dd_test_range.to(torch.long)
        self.test_neg_index = self._sample_test_negatives(data)      # src/layers.py:293 (drawn before the model exists)
        return data

    def _sample_test_negatives(self, data):
        return typed_negative_sampling(data.dd_test_idx, data.n_drug, data.dd_test_range)

    def _prepare_model(self):
        d, s = self.data, self.settings
        self.encoder = FMEncoder(self.device, d.n_drug_feat, d.n_dd_et, d.n_prot, d.n_prot, d.n_drug, s.prot_drug_dim,
                                 s.num_base, s.n_embed, s.n_hid1, s.n_hid2, mod=self.mod).to(self.device)
        with torch.no_grad():   # the reference's warm-up encoder call (src/layers.py:319); builds and caches the plans
            self.embeddings = self._encode()
        self.decoder = MultiInnerProductDecoder(s.n_hid2, d.n_dd_et).to(self.device)

    def __getstate__(self):      # torch.save(model) (tip.py:36): streams and per-step index buffers are runtime state
        state = dict(self.__dict__)
        for k in ("_neg_plan", "_neg_index_buf", "_neg_packed", "_side"):
            state[k] = None
        return state

    @property
    def _neg_index(self):
        """the last step's negative pairs as the reference has them: int64 [2, E].  The fused step keeps them packed
        (row << 16 | col); they are unpacked here on demand (inspection / tests, not on the training path)."""
        if self._neg_packed is not None:
            return ops.unpack_pairs(self._neg_packed)
        return self._neg_index_buf

    @_neg_index.setter
    def _neg_index(self, value):
        self._neg_index_buf = value

    def invalidate_graph_caches(self):
        """Kept for callers of the first release: every cached index structure (typed CSRs, GCN normalisation,
        positive-pair bitmaps) re-keys itself on the `_version` of its graph tensors, so nothing needs to be done
        after an in-place overwrite.  Dropping the GCN caches here only forces their rebuild."""
        self.encoder.pp_encoder.conv1._cache = None
        self.encoder.pp_encoder.conv2._cache = None

    def _encode(self):
        d = self.data
        return self.encoder(d.d_feat, d.dd_train_idx, d.dd_train_et, d.dd_train_range, d.d_norm, d.p_feat,
                            d.pp_train_indices, d.dp_edge_index, d.dp_range_list)

    def forward(self, check_status=True):
        d = self.data
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=SIDE_PRIORITY)
        cur = torch.cuda.current_stream(self.device)
        side = cur if SERIAL_STREAMS else self._side
        # ---- fused pair pass (csrc/pair_pass.cu): mirrored edge set + z fits in shared memory (the polypharmacy shape)
        member = _ns._membership(d.dd_train_idx, d.n_drug, d.dd_train_range)     # cached; holds a host copy of the ranges
        plan = ops.pair_plan(d.dd_train_idx, d.n_drug, d.n_dd_et, d.dd_train_range, self.settings.n_hid2,
                             rl_host=member.rl_host, validate=check_status)
        if plan is not None:
            if self._neg_packed is None or self._neg_packed.numel() != d.dd_train_idx.shape[1]:
                self._neg_packed = torch.empty(d.dd_train_idx.shape[1], dtype=torch.int32, device=self.device)
            # the negatives do not depend on the encoder: both chains fork here.  The encoder (the longer chain) is
            # enqueued FIRST: a captured graph starts its root branches in capture order, and the sampler's first
            # kernels are wide enough to keep the encoder's first kernel waiting for SM slots otherwise
            fork = torch.cuda.Event()
            fork.record(cur)
            if ENCODER_FIRST:
                self.embeddings = self._encode()
            side.wait_event(fork)
            with torch.cuda.stream(side):
                typed_negative_sampling(d.dd_train_idx, d.n_drug, d.dd_train_range, check_status=check_status,
                                        packed_out=self._neg_packed)
            if not ENCODER_FIRST:
                self.embeddings = self._encode()
            return ops.pair_bce_loss(self.embeddings, self.decoder.weight, plan, self._neg_packed, neg_stream=side)
        # ---- general path: typed CSR of the fresh negatives + segment decoder (large graphs, unmirrored edge sets)
        self._neg_packed = None
        if self._neg_index_buf is None or self._neg_plan is None:
            self._neg_index_buf = torch.empty_like(d.dd_train_idx)
            self._neg_plan = ops.TypedCSR(d.dd_train_idx.shape[1], d.n_drug, d.n_dd_et, self.device, by_src=False,
                                          doubled=True, rel_major=True)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            neg_index = typed_negative_sampling(d.dd_train_idx, d.n_drug, d.dd_train_range,
                                                check_status=check_status, out=self._neg_index_buf)
            self._neg_plan.build(neg_index, range_list=d.dd_train_range)
        self.embeddings = self._encode()
        pos_plan = ops.positive_decoder_plan(d.dd_train_idx, d.n_drug, d.n_dd_et, d.dd_train_range)
        # decoder(pos) / decoder(neg) / the two log-means of src/layers.py:335-340, fused with their gradient;
        # only the negative pass waits for the side stream
        return ops.bce_loss(self.embeddings, self.decoder.weight, pos_plan, self._neg_plan, neg_stream=side)

    def pred(self, dd_idx, dd_et):
        return self.decoder(self.embeddings, dd_idx, dd_et)

    def test(self, print_output=True):
        self.eval()
        d = self.data
        with torch.no_grad():
            pos_score = self.decoder(self.embeddings, d.dd_test_idx, d.dd_test_et)
            neg_score = self.decoder(self.embeddings, self.test_neg_index, d.dd_test_et)
        return self.compute_auprc_auroc_ap_by_et(pos_score, neg_score, d.dd_test_range, print_output)

    def compute_auprc_auroc_ap_by_et(self, pos_score, neg_score, dd_range, print_out):
        """src/layers.py:353-375; the 861 x 3 scikit-learn calls on host copies become one segmented sort + scan on the
        device (ops.eval_auprc_auroc_ap), one device->host copy of the 3 x n_rel record at the end"""
        record = ops.eval_auprc_auroc_ap(pos_score, neg_score, dd_range).cpu().numpy()
        if print_out:
            auprc, auroc, ap = record.sum(axis=1) / self.data.n_dd_et
            print("On test set: auprc:{:0.4f}   auroc:{:0.4f}   ap@50:{:0.4f}    ".format(auprc, auroc, ap))
        return record
