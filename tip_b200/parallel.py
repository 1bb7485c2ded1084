"""Relation sharding of the TIP hot path over the GPUs of one node.

The reference is single-device (README.md:58).  The path shards naturally by relation:

    out = D^-1 sum_r A_r X W_r + X root         and         loss = sum over typed edges

so every rank owns a contiguous, edge-count-balanced block of relations -- their typed CSRs,
`att` rows, decoder rows and negative samples -- while X, `basis`, `root`, `embed` and the small
P-P / P-D graphs (and their weights) are replicated and recomputed redundantly (identical inputs
=> identical results, no communication).  Exchange steps (NCCL all-reduce, sum, fp32):

    forward    partial un-normalised layer outputs [N_d, F1], [N_d, F2]; the loss scalar
    backward   d(z) [N_d, F2], d(x) of both layers [N_d, F1], [N_d, F0];
               d(basis), d(root) of both layers (one packed buffer, `sync_gradients`)

`att` and `decoder.weight` gradients are relation-local and need no exchange.  Negative
sampling stays bit-exact: every rank runs the (replicated) stream walk over ALL relations --
the MT19937 stream is global -- and keeps the pairs of its own relations.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import neg_sampling as _ns
from . import ops
from .layers import TIP


def partition_relations(range_list, world):
    """contiguous blocks of relations with (nearly) equal edge counts -> [(r_lo, r_hi)] * world"""
    rl = np.asarray(range_list.cpu() if torch.is_tensor(range_list) else range_list, dtype=np.int64)
    n_rel = rl.shape[0]
    ends = rl[:, 1]
    total = int(ends[-1]) if n_rel else 0
    cuts = [0]
    for k in range(1, world):
        target = total * k / world
        r = int(np.searchsorted(ends, target, side="left")) + 1      # first block boundary at/after the target
        r = min(max(r, cuts[-1]), n_rel)
        # pick the nearer of the two candidate boundaries
        if r - 1 > cuts[-1] and abs(int(ends[r - 2]) - target) <= abs(int(ends[r - 1]) - target):
            r -= 1
        cuts.append(r)
    cuts.append(n_rel)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


class _Collective(object):
    """sum-reduction over the ranks; `group=None` + world 1 is the identity (single GPU)."""

    def __init__(self, world, group=None):
        self.world, self.group = world, group

    def all_reduce_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class _ReduceFwd(torch.autograd.Function):
    """all-reduce in forward, identity in backward (partial results -> replicated result)"""

    @staticmethod
    def forward(ctx, x, coll):
        return coll.all_reduce_(x.clone())

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ReduceBwd(torch.autograd.Function):
    """identity in forward, all-reduce in backward (replicated input of a relation-local op)"""

    @staticmethod
    def forward(ctx, x, coll):
        ctx.coll = coll
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.coll.all_reduce_(g.contiguous().clone()), None


class ShardedTIP(TIP):
    """TIP whose D-D relations (R-GCN messages, decoder pairs, negatives) are sharded over `world` ranks.
    Every rank constructs the same parameters (same torch seed) and the same data; `rank` selects the shard.
    `collective` can be injected (tests simulate several logical ranks in one process)."""

    def __init__(self, settings, device, mod="cat", data_path="./data/data_dict.pkl", data=None, rank=0, world=1,
                 collective=None):
        super().__init__(settings, device, mod=mod, data_path=data_path, data=data)
        self.rank, self.world = rank, world
        self.coll = collective if collective is not None else _Collective(world)
        d = self.data
        self.blocks = partition_relations(d.dd_train_range, world)
        self.r_lo, self.r_hi = self.blocks[rank]
        rl = d.dd_train_range
        self.e_lo = int(rl[self.r_lo, 0]) if self.r_hi > self.r_lo else 0
        self.e_hi = int(rl[self.r_hi - 1, 1]) if self.r_hi > self.r_lo else 0
        self.n_local_rel = self.r_hi - self.r_lo
        self.local_idx = d.dd_train_idx[:, self.e_lo:self.e_hi].contiguous()
        self.local_range = (rl[self.r_lo:self.r_hi] - self.e_lo).contiguous()
        n = d.n_drug
        n_rel = max(self.n_local_rel, 1)
        kw = dict(range_list=self.local_range if self.n_local_rel else None)
        self.plan_dst = ops.cached_plan(self.local_idx, n, n_rel, by_src=False, **kw)
        self.plan_src = ops.cached_plan(self.local_idx, n, n_rel, by_src=True, **kw)
        # the mean is over ALL incoming edges of ALL relations: degrees are global
        deg = self.plan_dst.field("deg").to(torch.float32).clone()
        self.coll.all_reduce_(deg)
        self.inv_deg_global = 1.0 / deg.clamp(min=1.0)
        self.plan_dst.inv_deg.copy_(self.inv_deg_global)
        # whole relations per rank: a mirrored edge set stays mirrored, and plan_dst then serves the decoder too
        if self.n_local_rel and ops.edges_mirrored(self.local_idx, self.local_range):
            self.pos_plan = self.plan_dst
        else:
            self.pos_plan = ops.cached_plan(self.local_idx, n, n_rel, by_src=False, doubled=True, rel_major=True, **kw)
        self._neg_local = None
        self._neg_plan_local = None

    def invalidate_graph_caches(self):
        """the graph tensors were overwritten in place: re-derive this rank's shard and rebuild its index structures
        (same buffers), including the global degrees"""
        super().invalidate_graph_caches()
        d = self.data
        self.local_idx.copy_(d.dd_train_idx[:, self.e_lo:self.e_hi])
        self.local_range.copy_(d.dd_train_range[self.r_lo:self.r_hi] - self.e_lo)
        rl = self.local_range if self.n_local_rel else None
        plans = [self.plan_dst, self.plan_src] + ([self.pos_plan] if self.pos_plan is not self.plan_dst else [])
        for plan in plans:
            plan.build(self.local_idx, None, rl)
            plan._versions = ops._versions(self.local_idx, None, rl)
        deg = self.plan_dst.field("deg").to(torch.float32).clone()
        self.coll.all_reduce_(deg)
        self.plan_dst.inv_deg.copy_(1.0 / deg.clamp(min=1.0))

    # ---- one R-GCN layer on this rank's relations
    def _rgcn_local(self, conv, x, relu):
        x = _ReduceBwd.apply(x, self.coll)
        att = conv.att[self.r_lo:self.r_hi] if self.n_local_rel else conv.att[:1] * 0.0
        part = ops.rgcn_conv(x, conv.basis, att, conv.root / self.world, self.plan_dst, self.plan_src)
        out = _ReduceFwd.apply(part, self.coll)
        return torch.relu(out) if relu else out

    def _encode(self):
        if not hasattr(self, "plan_dst"):        # constructor warm-up of the base class (replicated, unsharded)
            return super()._encode()
        d, enc = self.data, self.encoder
        x_prot = enc.pp_encoder(d.p_feat, d.pp_train_indices)
        x_prot = torch.cat((x_prot, enc.hdrug.to(x_prot.device)))
        x_prot = enc.hgcn(x_prot, d.dp_edge_index, d.dp_range_list)
        x_drug = enc._embed(d.d_feat) / d.d_norm.view(-1, 1)
        x_drug = torch.cat((x_drug, x_prot), dim=1) if enc.mod == "cat" else x_drug + x_prot
        x_drug = self._rgcn_local(enc.rgcn1, x_drug, relu=True)
        return self._rgcn_local(enc.rgcn2, x_drug, relu=False)

    def forward(self, check_status=True):
        d = self.data
        if self._neg_index is None:
            self._neg_index = torch.empty_like(d.dd_train_idx)
            self._neg_local = torch.empty_like(self.local_idx)
            self._neg_plan_local = ops.TypedCSR(self.local_idx.shape[1], d.n_drug, max(self.n_local_rel, 1), self.device,
                                                by_src=False, doubled=True, rel_major=True)
            self._side = torch.cuda.Stream(device=self.device, priority=-1)   # sampler -> plan is the critical chain
        cur = torch.cuda.current_stream(self.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            # the MT19937 stream is global: walk all relations (replicated), keep this rank's pairs
            neg_all = _ns.typed_negative_sampling(d.dd_train_idx, d.n_drug, d.dd_train_range, check_status=check_status,
                                                  out=self._neg_index)
            self._neg_local.copy_(neg_all[:, self.e_lo:self.e_hi])
            self._neg_plan_local.build(self._neg_local, range_list=self.local_range if self.n_local_rel else None)
        self.embeddings = self._encode()
        z = _ReduceBwd.apply(self.embeddings, self.coll)
        w = self.decoder.weight[self.r_lo:self.r_hi] if self.n_local_rel else self.decoder.weight[:1] * 0.0
        local = ops.bce_loss(z, w, self.pos_plan, self._neg_plan_local, neg_stream=self._side)
        # local means -> share of the global means
        share = float(self.e_hi - self.e_lo) / float(max(d.dd_train_idx.shape[1], 1))
        return _ReduceFwd.apply(local * share, self.coll)

    def sync_gradients(self):
        """all-reduce the gradients of the replicated R-GCN parameters that received partial contributions
        (basis, root of both layers), packed into one buffer"""
        if self.world == 1:
            return
        enc = self.encoder
        params = [enc.rgcn1.basis, enc.rgcn1.root, enc.rgcn2.basis, enc.rgcn2.root]
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        self.coll.all_reduce_(flat)
        off = 0
        for p in params:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p))
            off += n
