"""Relation sharding of the TIP hot path over the GPUs of one node.

The reference is single-device (README.md:58).  The path shards naturally by relation:

    out = D^-1 sum_r A_r X W_r + X root         and         loss = sum over typed edges

so every rank owns a contiguous, edge-count-balanced block of relations -- their typed CSRs, `att` rows, decoder
rows, positive-pair bitmaps and negative samples -- while X, `basis`, `root`, `embed` and the small P-P / P-D graphs
(and their weights) are replicated and recomputed redundantly (identical inputs => identical results, no
communication).  Exchange steps per training step (NCCL, sum, fp32; 5 + the sampler's all-gather):

    forward    partial un-normalised layer outputs [N_d, F1], [N_d, F2]                              (2 all-reduces)
    backward   {d(z) [N_d, F2], loss share}, {d(x2), d(basis2), d(root2)}, {d(x1), d(basis1), d(root1)}   (3, packed)
    sampler    the ranks' offset tables (a few KB), on the sampler's side stream / its own group      (1 all-gather)

Each backward all-reduce follows from the previous one through a layer's backward pass, each forward one through a
layer's forward pass, so five is the dependency minimum of this formulation.  `att` and `decoder.weight` gradients
are relation-local and need no exchange: a rank only ever updates ITS rows, `gather_parameters()` (called by `test()`
and `state_dict()`) brings all rows to all ranks.

Negative sampling stays bit-exact AND sharded (neg_sampling.ShardedSampler): every rank advances the same MT19937
state, scans only the windows of its relations and learns its start offset in the shared stream from the exchanged
tables.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import neg_sampling as _ns
from . import ops
from ._lib import check, lib, ptr, stream
from .layers import TIP, FMEncoder, MultiInnerProductDecoder, MyRGCNConv2, Param


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
    """the exchange steps over the ranks; world 1 is the identity (single GPU).  `sampler_group`: a second process
    group (its own communicator and stream) so that the sampler's all-gather on the side stream does not queue behind
    the encoder's all-reduces."""

    def __init__(self, world, group=None, sampler_group=None):
        self.world, self.group, self.sampler_group = world, group, sampler_group

    def all_reduce_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_(self, out, inp):
        """out [world, n] <- every rank's inp [n]"""
        if self.world == 1:
            out[0].copy_(inp)
            return out
        g = self.sampler_group if self.sampler_group is not None else self.group
        if dist.get_backend(g) == "nccl":
            dist.all_gather_into_tensor(out.view(-1), inp, group=g)
        else:                                   # gloo (CPU-side tests of the host logic)
            dist.all_gather([out[k] for k in range(self.world)], inp, group=g)
        return out


class _ReduceFwd(torch.autograd.Function):
    """all-reduce in forward (in place: the input is a fresh partial result), identity in backward"""

    @staticmethod
    def forward(ctx, x, coll):
        ctx.mark_dirty(x)
        return coll.all_reduce_(x)

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
        return ctx.coll.all_reduce_(g.contiguous()), None


class _ShardedRGCN(torch.autograd.Function):
    """one R-GCN layer on this rank's relations: local conv + all-reduce of the partial output in forward; in backward
    the partial d(x), d(basis), d(root) travel in ONE packed all-reduce, d(att) stays local (own relations)."""

    @staticmethod
    def forward(ctx, x, basis, att_local, root, plan_dst, plan_src, coll, add_root, zero_root):
        x, basis, att_local, root = (t.contiguous() for t in (x, basis, att_local, root))
        n, f_in = x.shape
        n_bases, _, f_out = basis.shape
        n_rel = att_local.shape[0]
        L = lib()
        out = torch.empty((n, f_out), dtype=torch.float32, device=x.device)
        g_saved = torch.empty((n, n_bases, f_in), dtype=torch.float32, device=x.device)
        # x @ root belongs to the sum once: rank 0 adds it, the other ranks pass a zero matrix (no scaling kernels)
        root_eff = root if add_root else zero_root
        ws = ops.workspace(L.tipb_rgcn_workspace_bytes(plan_dst.n_entries, n, n_rel, f_in, f_out, n_bases), x.device)
        check(L.tipb_rgcn_fwd(ptr(plan_dst.buf), plan_dst.n_entries, n, n_rel, ptr(x), ptr(basis), ptr(att_local),
                              ptr(root_eff), None, f_in, f_out, n_bases, 0, ptr(out), ptr(g_saved), ptr(ws), ws.numel(),
                              stream()), "rgcn_fwd")
        coll.all_reduce_(out)
        ctx.save_for_backward(x, basis, att_local, root_eff, g_saved)
        ctx.plan_dst, ctx.plan_src, ctx.coll = plan_dst, plan_src, coll
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, basis, att_local, root_eff, g_saved = ctx.saved_tensors
        plan_dst, plan_src = ctx.plan_dst, ctx.plan_src
        grad_out = grad_out.contiguous()
        n, f_in = x.shape
        n_bases, _, f_out = basis.shape
        n_rel = att_local.shape[0]
        L = lib()
        # d_x | d_basis in one buffer: one all-reduce.  d_root = X^T g does not depend on the shard (x and g are
        # replicated): every rank computes it in full, it does not travel
        sizes = (x.numel(), basis.numel())
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=x.device)
        d_x, d_basis = (t.view(s) for t, s in zip(flat.split(sizes), (x.shape, basis.shape)))
        d_att, d_root = torch.empty_like(att_local), torch.empty_like(root_eff)
        ws = ops.workspace(L.tipb_rgcn_workspace_bytes(plan_src.n_entries, n, n_rel, f_in, f_out, n_bases), x.device)
        check(L.tipb_rgcn_bwd(ptr(plan_src.buf), plan_src.n_entries, n, n_rel, ptr(plan_dst.inv_deg), ptr(x), ptr(basis),
                              ptr(att_local), ptr(root_eff), ptr(g_saved), ptr(grad_out), None, f_in, f_out, n_bases,
                              ptr(d_x), ptr(d_basis), ptr(d_att), ptr(d_root), None, ptr(ws), ws.numel(), stream()),
              "rgcn_bwd")
        # (the g @ root^T term of d_x is present on rank 0 only, like x @ root in forward: the sum holds it once)
        ctx.coll.all_reduce_(flat)
        return d_x, d_basis, d_att, d_root, None, None, None, None, None


def _layers_side_priority():
    from . import layers as _l
    return _l.SIDE_PRIORITY


class ShardedTIP(TIP):
    """TIP whose D-D relations (R-GCN messages, decoder pairs, positive-pair bitmaps, negatives) are sharded over
    `world` ranks.  Every rank constructs the same parameters (same torch seed) and is given the same data; `rank`
    selects the shard.  `collective` can be injected (tests).  `defer_loss_reduce`: forward() returns this rank's
    share of the loss and the global value is all-reduced together with d(z) in backward (read it from
    `last_loss` afterwards) -- one collective less on the critical path."""

    def __init__(self, settings, device, mod="cat", data_path="./data/data_dict.pkl", data=None, rank=0, world=1,
                 collective=None, defer_loss_reduce=False):
        object.__setattr__(self, "rank", int(rank))
        object.__setattr__(self, "world", int(world))
        object.__setattr__(self, "coll", collective if collective is not None else _Collective(world))
        object.__setattr__(self, "defer_loss_reduce", bool(defer_loss_reduce))
        object.__setattr__(self, "_zero_roots", {})
        super().__init__(settings, device, mod=mod, data_path=data_path, data=data)

    # ---- data: the base class moved everything to the device; derive this rank's shard
    def _sample_test_negatives(self, data):
        """src/layers.py:293 sharded: the test-set negatives are drawn from the same global stream, every rank samples
        the relations it owns (bitmaps of its relations only) and the slices are exchanged once"""
        if self.world == 1:
            return super()._sample_test_negatives(data)
        rl = np.ascontiguousarray(data.dd_test_range.cpu().numpy().astype(np.int64))
        blocks = partition_relations(rl, self.world)
        first = [b[0] for b in blocks] + [int(rl.shape[0])]
        r_lo, r_hi = blocks[self.rank]
        e_lo = int(rl[r_lo, 0]) if r_hi > r_lo else 0
        e_hi = int(rl[r_hi - 1, 1]) if r_hi > r_lo else 0
        local_idx = data.dd_test_idx[:, e_lo:e_hi].contiguous()
        local_rl = torch.from_numpy(np.ascontiguousarray(rl[r_lo:r_hi] - e_lo)).to(self.device)
        sampler = _ns.ShardedSampler(local_idx, data.n_drug, local_rl, rl, first, self.rank, self.world, self.coll)
        mine = sampler.sample()
        sizes = [int(rl[b - 1, 1]) - int(rl[a, 0]) if b > a else 0 for a, b in blocks]
        pad = max(max(sizes), 1)
        send = torch.zeros(2 * pad, dtype=torch.long, device=self.device)
        send[:mine.shape[1]] = mine[0]
        send[pad:pad + mine.shape[1]] = mine[1]
        recv = torch.zeros((self.world, 2 * pad), dtype=torch.long, device=self.device)
        self.coll.all_gather_(recv, send)
        out = torch.cat([torch.stack([recv[k, :n], recv[k, pad:pad + n]]) for k, n in enumerate(sizes)], dim=1)
        del sampler
        return out

    def _prepare_model(self):
        d = self.data
        self.rl_host = np.ascontiguousarray(d.dd_train_range.cpu().numpy().astype(np.int64))
        self.blocks = partition_relations(self.rl_host, self.world)
        self.first_rel = [b[0] for b in self.blocks] + [int(self.rl_host.shape[0])]
        self.r_lo, self.r_hi = self.blocks[self.rank]
        self.n_local_rel = self.r_hi - self.r_lo
        self.e_lo = int(self.rl_host[self.r_lo, 0]) if self.n_local_rel else 0
        self.e_hi = int(self.rl_host[self.r_hi - 1, 1]) if self.n_local_rel else 0
        self.local_idx = d.dd_train_idx[:, self.e_lo:self.e_hi].contiguous()
        self.local_rl_host = np.ascontiguousarray(self.rl_host[self.r_lo:self.r_hi] - self.e_lo)
        self.local_range = torch.from_numpy(self.local_rl_host).to(self.device)
        self._build_modules()
        self.last_loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self._build_shard_plans()
        self.sampler = _ns.ShardedSampler(self.local_idx, d.n_drug, self.local_range, self.rl_host, self.first_rel,
                                          self.rank, self.world, self.coll)
        with torch.no_grad():    # the reference's warm-up encoder call (src/layers.py:319)
            self.embeddings = self._encode()

    def _build_modules(self):
        d, s = self.data, self.settings
        self.encoder = FMEncoder(self.device, d.n_drug_feat, d.n_dd_et, d.n_prot, d.n_prot, d.n_drug, s.prot_drug_dim,
                                 s.num_base, s.n_embed, s.n_hid1, s.n_hid2, mod=self.mod).to(self.device)
        self.decoder = MultiInnerProductDecoder(s.n_hid2, d.n_dd_et).to(self.device)

    def _build_shard_plans(self):
        """typed CSRs of the local edges (both orientations) with the GLOBAL in-degrees (the mean runs over all
        incoming edges of all relations), and the decoder's pair plan / positive plan"""
        d = self.data
        n, n_rel = d.n_drug, max(self.n_local_rel, 1)
        kw = dict(range_list=self.local_range if self.n_local_rel else None)
        self.plan_dst = ops.cached_plan(self.local_idx, n, n_rel, by_src=False, **kw)
        self.plan_src = ops.cached_plan(self.local_idx, n, n_rel, by_src=True, **kw)
        deg = self.plan_dst.field("deg").to(torch.float32)
        self.coll.all_reduce_(deg)
        self.plan_dst.inv_deg.copy_(1.0 / deg.clamp(min=1.0))
        self.pair_plan = None
        self.pos_plan = None
        if self.n_local_rel:
            self.pair_plan = ops.pair_plan(self.local_idx, n, n_rel, self.local_range, self.settings.n_hid2,
                                           rl_host=self.local_rl_host)
            if self.pair_plan is None:
                # a mirrored edge set stays mirrored (whole relations per rank): plan_dst then serves the decoder too
                self.pos_plan = self.plan_dst if ops.edges_mirrored(self.local_idx, self.local_range) else \
                    ops.cached_plan(self.local_idx, n, n_rel, by_src=False, doubled=True, rel_major=True, **kw)

    def refresh_shard(self):
        """`local_idx` was overwritten in place (same ranges): rebuild this rank's index structures into the same
        buffers -- typed CSRs, global degrees, positive pairs, bitmaps.  Device work + one small exchange, no host sync
        besides the popcount exchange of the sampler."""
        rl = self.local_range if self.n_local_rel else None
        for plan in {id(p): p for p in (self.plan_dst, self.plan_src, self.pos_plan) if p is not None}.values():
            plan.build(self.local_idx, None, rl)
            plan._versions = ops._versions(self.local_idx, None, rl)
        deg = self.plan_dst.field("deg").to(torch.float32)
        self.coll.all_reduce_(deg)
        self.plan_dst.inv_deg.copy_(1.0 / deg.clamp(min=1.0))
        if self.pair_plan is not None:
            self.pair_plan.repack(self.local_idx)
        self.sampler.rebuild(self.local_idx, self.local_range)

    def invalidate_graph_caches(self):
        """the full graph tensors were overwritten in place: re-derive the shard from them"""
        if hasattr(self.encoder, "pp_encoder"):
            super().invalidate_graph_caches()
        self.local_idx.copy_(self.data.dd_train_idx[:, self.e_lo:self.e_hi])
        self.refresh_shard()

    # ---- encoder
    def _rgcn_local(self, conv, x, relu):
        att = conv.att[self.r_lo:self.r_hi] if self.n_local_rel else conv.att[:1] * 0.0
        zero = self._zero_roots.get(id(conv))
        if zero is None:
            zero = self._zero_roots[id(conv)] = torch.zeros_like(conv.root)
        out = _ShardedRGCN.apply(x, conv.basis, att, conv.root, self.plan_dst, self.plan_src, self.coll, self.rank == 0, zero)
        return torch.relu(out) if relu else out

    def _drug_input(self):
        d = self.data
        return self.encoder.drug_input(d.d_feat, d.d_norm, d.p_feat, d.pp_train_indices, d.dp_edge_index, d.dp_range_list)

    def _encode(self):
        enc = self.encoder
        x = self._rgcn_local(enc.rgcn1, self._drug_input(), relu=True)
        return self._rgcn_local(enc.rgcn2, x, relu=False)

    # ---- step
    def forward(self, check_status=True):
        d = self.data
        if self._side is None:
            # the main chain (P-P encoder, R-GCN, collectives) is the critical one: the sampler must not take its SM slots
            self._side = torch.cuda.Stream(device=self.device, priority=_layers_side_priority())
        cur = torch.cuda.current_stream(self.device)
        from . import layers as _layers
        side = cur if _layers.SERIAL_STREAMS else self._side
        e_local = self.e_hi - self.e_lo
        use_pairs = self.pair_plan is not None
        if use_pairs:
            if self._neg_packed is None or self._neg_packed.numel() != e_local:
                self._neg_packed = torch.empty(e_local, dtype=torch.int32, device=self.device)
        elif self._neg_index_buf is None:
            self._neg_index_buf = torch.empty((2, e_local), dtype=torch.long, device=self.device)
            if self.n_local_rel:
                self._neg_plan = ops.TypedCSR(e_local, d.n_drug, self.n_local_rel, self.device, by_src=False,
                                              doubled=True, rel_major=True)
        side.wait_stream(cur)
        with torch.cuda.stream(side):       # the negatives do not depend on the encoder: sample them meanwhile
            if use_pairs:
                self.sampler.sample(packed_out=self._neg_packed, check_status=check_status)
            else:
                self.sampler.sample(out=self._neg_index_buf, check_status=check_status)
                if self.n_local_rel:
                    self._neg_plan.build(self._neg_index_buf, range_list=self.local_range)
        self.embeddings = self._encode()
        z = _ReduceBwd.apply(self.embeddings, _LossPacker(self)) if self.defer_loss_reduce else \
            _ReduceBwd.apply(self.embeddings, self.coll)
        w = self.decoder.weight[self.r_lo:self.r_hi] if self.n_local_rel else self.decoder.weight[:1] * 0.0
        if not self.n_local_rel:
            local = (z.sum() + w.sum()) * 0.0
        elif use_pairs:
            local = ops.pair_bce_loss(z, w, self.pair_plan, self._neg_packed, neg_stream=side)
        else:
            local = ops.bce_loss(z, w, self.pos_plan, self._neg_plan, neg_stream=side)
        # local means -> share of the global means
        share = float(e_local) / float(max(int(self.rl_host[-1, 1]), 1))
        part = local * share
        if self.defer_loss_reduce:
            self._loss_share = part.detach()
            return part
        return _ReduceFwd.apply(part.clone(), self.coll)

    @property
    def _neg_index(self):
        """this rank's negatives, int64 [2, E_local] (columns e_lo .. e_hi of the unsharded sample)"""
        if self._neg_packed is not None:
            return ops.unpack_pairs(self._neg_packed)
        return self._neg_index_buf

    @_neg_index.setter
    def _neg_index(self, value):
        self._neg_index_buf = value

    def sync_gradients(self):
        """kept for callers of the first release: the R-GCN weight gradients now travel with d(x) in the backward
        all-reduces (`_ShardedRGCN`), nothing is left to do"""
        return None

    # ---- relation-local parameters: every rank trains only ITS rows
    def _local_params(self):
        enc = self.encoder
        return [enc.rgcn1.att, enc.rgcn2.att, self.decoder.weight]

    def _gather_rows(self, tensors):
        """all ranks' [r_lo:r_hi) rows of each [n_rel, *] tensor -> everywhere (one packed all-reduce)"""
        if self.world == 1:
            return
        with torch.no_grad():
            parts = []
            for t in tensors:
                own = torch.zeros_like(t)
                own[self.r_lo:self.r_hi] = t[self.r_lo:self.r_hi]
                parts.append(own.reshape(-1))
            flat = torch.cat(parts)
            self.coll.all_reduce_(flat)
            off = 0
            for t in tensors:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()

    def gather_parameters(self, optimizer=None):
        """Bring the rows of `att` (both R-GCN layers) and `decoder.weight` that other ranks trained to this rank (and
        vice versa): afterwards every rank holds the same, complete parameters.  With `optimizer`, its per-row state
        (exp_avg, exp_avg_sq) is merged the same way, so that a checkpoint written by any rank resumes anywhere."""
        self._gather_rows([p.data for p in self._local_params()])
        if optimizer is not None:
            for p in self._local_params():
                st = optimizer.state.get(p, {})
                self._gather_rows([st[k] for k in ("exp_avg", "exp_avg_sq") if k in st])

    def test(self, print_output=True):
        self.gather_parameters()
        with torch.no_grad():
            self.embeddings = self._encode()        # complete parameters (identical on every rank)
        return super().test(print_output)

    def state_dict(self, *args, **kwargs):
        self.gather_parameters()
        return super().state_dict(*args, **kwargs)


class _LossPacker(object):
    """collective adapter for the d(z) all-reduce that also carries the loss share (defer_loss_reduce)"""

    def __init__(self, model):
        self.model = model

    def all_reduce_(self, g):
        m = self.model
        flat = torch.cat([g.reshape(-1), m._loss_share.reshape(1)])
        m.coll.all_reduce_(flat)
        m.last_loss.copy_(flat[-1])
        return flat[:-1].view_as(g)


class ShardedDDNet(ShardedTIP):
    """BASELINE.json config 3: the D-D-only R-GCN of the reference's test/dd_net_scalable.py:51-80
    (embed[N, 64] -> MyRGCNConv2(64, 32) -> ReLU -> MyRGCNConv2(32, 16) -> ReLU, n_base = 16, DistMult decoder),
    relation-sharded like ShardedTIP (world 1 = the single-GPU model)."""

    N_EMBED, N_HID1, N_HID2, N_BASE = 64, 32, 16, 16

    def _build_modules(self):
        d = self.data
        enc = torch.nn.Module()
        enc.embed = Param(torch.Tensor(d.n_drug_feat, self.N_EMBED))
        enc.rgcn1 = MyRGCNConv2(self.N_EMBED, self.N_HID1, d.n_dd_et, self.N_BASE, after_relu=False)
        enc.rgcn2 = MyRGCNConv2(self.N_HID1, self.N_HID2, d.n_dd_et, self.N_BASE, after_relu=True)
        enc.embed.data.normal_()                       # test/dd_net_scalable.py:79-80
        self.encoder = enc.to(self.device)
        self.decoder = MultiInnerProductDecoder(self.N_HID2, d.n_dd_et).to(self.device)

    def _drug_input(self):
        # d_feat is the sparse identity (test/dd_net_scalable.py:43): x @ embed == embed; x_norm = ones
        return ops.row_scale(self.encoder.embed, self.data.d_norm)

    def _encode(self):
        enc = self.encoder
        x = self._rgcn_local(enc.rgcn1, self._drug_input(), relu=True)
        return self._rgcn_local(enc.rgcn2, x, relu=True)
