"""torch.autograd.Function wrappers over the C ABI (libtipb200.so).

Everything here is plumbing: tensors are made contiguous, outputs and scratch
are allocated with torch, raw pointers + the current CUDA stream go to the
library.  No computation of the hot path happens in Python, and nothing falls
back to torch ops when the library is missing (tip_b200._lib raises).
"""
import functools
import math
import weakref

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, lib, ptr, stream

# ----------------------------------------------------------------------------- scratch memory
_workspaces = {}


def workspace(nbytes, device, tag="ws"):
    """A reusable scratch buffer (grown geometrically).  Library calls are ordered on the current
    stream, so one buffer per (device, tag, stream) is enough and never shared across streams."""
    key = (device.index, tag, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _f32c(t):
    if t.dtype != torch.float32:
        raise _lib.TipbError(f"tip_b200 computes in fp32; got {t.dtype}")
    return t.contiguous()


def _i64c(t):
    if t.dtype != torch.long:
        raise AssertionError("edge tensors must be torch.long (as torch_geometric asserts)")
    return t.contiguous()


def _guarded(fn):
    """run `fn` with the device of its first tensor argument current (the library launches on the current device)"""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        t = next((a for a in args if torch.is_tensor(a)), None)
        if t is None or not t.is_cuda:
            return fn(*args, **kw)          # the callee raises the "CUDA tensors only" error
        with torch.cuda.device(t.device):
            return fn(*args, **kw)
    return wrapper


def _next_pow2(v, lo=4):
    p = lo
    while p < v:
        p *= 2
    return p


# ----------------------------------------------------------------------------- typed CSR plans
class TypedCSR(object):
    """Device-side index structure built by tipb_typed_csr_build (tip_b200/csrc/typed_csr.cu)."""

    def __init__(self, n_edges, n_nodes, n_rel, device, by_src=False, doubled=False, drop_self_loops=False,
                 n_other=None, rel_major=False):
        self.n_edges, self.n_nodes, self.n_rel = int(n_edges), int(n_nodes), int(n_rel)
        self.n_other = int(n_other if n_other is not None else n_nodes)
        self.by_src, self.doubled, self.drop_self_loops = bool(by_src), bool(doubled), bool(drop_self_loops)
        self.rel_major = bool(rel_major)   # segments ordered (relation, node): decoder plans
        self.n_entries = self.n_edges * (2 if doubled else 1)
        self.device = device
        L = lib()
        self.nbytes = L.tipb_typed_csr_bytes(self.n_entries, self.n_nodes, self.n_rel)
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self._layout, self.seg_cap = _lib.csr_layout(self.n_entries, self.n_nodes, self.n_rel)
        self._ws_bytes = L.tipb_typed_csr_workspace_bytes(self.n_entries, self.n_nodes, self.n_rel)

    def build(self, edge_index, edge_type=None, range_list=None):
        edge_index = _i64c(edge_index)
        assert edge_index.dim() == 2 and edge_index.shape[0] == 2 and edge_index.shape[1] == self.n_edges
        edge_type = None if edge_type is None else _i64c(edge_type)
        range_list = None if range_list is None else _i64c(range_list.to(torch.long))
        with torch.cuda.device(self.device):
            ws = workspace(self._ws_bytes, self.device, "csr")
            check(lib().tipb_typed_csr_build(ptr(edge_index), ptr(edge_type), ptr(range_list), self.n_edges,
                                             self.n_nodes, self.n_other, self.n_rel, int(self.by_src), int(self.doubled),
                                             int(self.drop_self_loops), int(self.rel_major), ptr(self.buf), self.nbytes,
                                             ptr(ws), ws.numel(), stream()), "typed_csr_build")
        return self

    # ---- views (tests, inv_deg for the backward pass)
    def field(self, name):
        sizes = {"counts": 16, "eid": self.n_entries, "other": self.n_entries, "seg_ptr": self.seg_cap + 1,
                 "seg_node": self.seg_cap, "seg_rel": self.seg_cap, "node_ptr": self.n_nodes + 1, "deg": self.n_nodes,
                 "inv_deg": self.n_nodes, "rel_seg_ptr": self.n_rel + 1, "rel_seg": self.seg_cap}
        off, n = self._layout[name], sizes[name]
        raw = self.buf[off:off + 4 * n]
        return raw.view(torch.float32 if name == "inv_deg" else torch.int32)

    @property
    def inv_deg(self):
        return self.field("inv_deg")

    def check_status(self):
        """One host sync: raises if an index was out of range when the plan was built."""
        if int(self.field("counts")[2]) != 0:
            raise IndexError("edge_index / edge_type contains an out-of-range index")
        return self


_plan_cache = {}


def clear_plan_cache():
    _plan_cache.clear()
    _dec_plans.clear()
    _pair_plans.clear()
    _item_tables.clear()
    _mirror_cache.clear()


def _tensor_key(t):
    return None if t is None else (t.data_ptr(), tuple(t.shape), str(t.device))


def _versions(*tensors):
    return tuple(None if t is None else t._version for t in tensors)


def cached_plan(edge_index, n_nodes, n_rel=1, edge_type=None, range_list=None, validate=True, **flags):
    """Plans are cached per (edge tensors identity, flags): the reference's modules are stateless but PyG's
    GCNConv(cached=True) keeps its normalised graph the same way.  If a key tensor was modified in place since
    the plan was built (its version counter moved), the plan is rebuilt into the same buffers."""
    key = (_tensor_key(edge_index), _tensor_key(edge_type), _tensor_key(range_list), int(n_nodes), int(n_rel),
           tuple(sorted(flags.items())))
    versions = _versions(edge_index, edge_type, range_list)
    plan = _plan_cache.get(key)
    if plan is None:
        if len(_plan_cache) > 64:
            _plan_cache.clear()
        plan = TypedCSR(edge_index.shape[1], n_nodes, n_rel, edge_index.device, **flags)
        # keep the key tensors alive so that data_ptr() cannot be recycled under the cache
        plan._keepalive = (edge_index, edge_type, range_list)
        plan._versions = None
        _plan_cache[key] = plan
    if plan._versions != versions:
        plan.build(edge_index, edge_type, range_list)
        if validate:
            plan.check_status()
        plan._versions = versions
    return plan


# ----------------------------------------------------------------------------- R-GCN
class _RGCNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, basis, att, root, plan_dst, plan_src, relu):
        x, basis, att, root = _f32c(x), _f32c(basis), _f32c(att), _f32c(root)
        n, f_in = x.shape
        n_bases, _, f_out = basis.shape
        n_rel = att.shape[0]
        assert n == plan_dst.n_nodes and n_rel == plan_dst.n_rel
        L = lib()
        out = torch.empty((n, f_out), dtype=torch.float32, device=x.device)
        g_saved = torch.empty((n, n_bases, f_in), dtype=torch.float32, device=x.device)
        ws_bytes = L.tipb_rgcn_workspace_bytes(plan_dst.n_entries, n, n_rel, f_in, f_out, n_bases)
        ws = workspace(ws_bytes, x.device)
        check(L.tipb_rgcn_fwd(ptr(plan_dst.buf), plan_dst.n_entries, n, n_rel, ptr(x), ptr(basis), ptr(att), ptr(root),
                              None, f_in, f_out, n_bases, int(relu), ptr(out), ptr(g_saved), ptr(ws), ws.numel(),
                              stream()), "rgcn_fwd")
        ctx.save_for_backward(x, basis, att, root, g_saved, out)
        ctx.plan_dst, ctx.plan_src, ctx.relu = plan_dst, plan_src, relu
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, basis, att, root, g_saved, out = ctx.saved_tensors
        plan_dst, plan_src = ctx.plan_dst, ctx.plan_src
        grad_out = _f32c(grad_out)
        n, f_in = x.shape
        n_bases, _, f_out = basis.shape
        n_rel = att.shape[0]
        L = lib()
        d_x, d_basis, d_att, d_root = (torch.empty_like(t) for t in (x, basis, att, root))
        ws_bytes = L.tipb_rgcn_workspace_bytes(plan_src.n_entries, n, n_rel, f_in, f_out, n_bases)
        ws = workspace(ws_bytes, x.device)
        check(L.tipb_rgcn_bwd(ptr(plan_src.buf), plan_src.n_entries, n, n_rel, ptr(plan_dst.inv_deg), ptr(x), ptr(basis),
                              ptr(att), ptr(root), ptr(g_saved), ptr(grad_out), ptr(out) if ctx.relu else None,
                              f_in, f_out, n_bases, ptr(d_x), ptr(d_basis), ptr(d_att), ptr(d_root), None, ptr(ws),
                              ws.numel(), stream()), "rgcn_bwd")
        return d_x, d_basis, d_att, d_root, None, None, None


@_guarded
def rgcn_conv(x, basis, att, root, plan_dst, plan_src, bias=None, relu=False):
    """out = mean-aggregated basis-decomposed relational conv + x @ root (+ bias) (+ ReLU).
    Feature widths that are not powers of two in [4,128] are zero-padded here (exact)."""
    n_bases, f_in, f_out = basis.shape
    fi, fo = _next_pow2(f_in), _next_pow2(f_out)
    if fi > 128 or fo > 128:
        raise _lib.TipbError("rgcn_conv supports feature widths up to 128")
    if (fi, fo) != (f_in, f_out):
        x = F.pad(x, (0, fi - f_in))
        basis = F.pad(basis, (0, fo - f_out, 0, fi - f_in))
        root = F.pad(root, (0, fo - f_out, 0, fi - f_in))
    fuse_relu = relu and bias is None
    out = _RGCNFunction.apply(x, basis, att, root, plan_dst, plan_src, fuse_relu)
    if fo != f_out:
        out = out[:, :f_out]
    if bias is not None:
        out = out + bias
        if relu:
            out = torch.relu(out)
    return out


# ----------------------------------------------------------------------------- GCN SpMM
class _GCNSpmmFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bias, plan_dst, plan_src, dis, relu):
        x = _f32c(x)
        n, f = x.shape
        out = torch.empty_like(x)
        check(lib().tipb_gcn_spmm(ptr(plan_dst.buf), plan_dst.n_entries, n, ptr(dis), ptr(dis), ptr(x),
                                  None if bias is None else ptr(_f32c(bias)), f, int(relu), ptr(out), stream()),
              "gcn_spmm")
        ctx.plan_src, ctx.dis, ctx.relu, ctx.has_bias = plan_src, dis, relu, bias is not None
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = _f32c(grad_out)
        if ctx.relu:
            (out,) = ctx.saved_tensors
            grad_out = grad_out * (out > 0)
        n, f = grad_out.shape
        d_x = torch.empty_like(grad_out)
        plan = ctx.plan_src
        check(lib().tipb_gcn_spmm(ptr(plan.buf), plan.n_entries, n, ptr(ctx.dis), ptr(ctx.dis), ptr(grad_out), None, f, 0,
                                  ptr(d_x), stream()), "gcn_spmm(bwd)")
        d_bias = grad_out.sum(dim=0) if ctx.has_bias else None
        return d_x, d_bias, None, None, None, None


def gcn_norm(plan_dst):
    dis = torch.empty(plan_dst.n_nodes, dtype=torch.float32, device=plan_dst.device)
    with torch.cuda.device(plan_dst.device):
        check(lib().tipb_gcn_norm(ptr(plan_dst.buf), plan_dst.n_entries, plan_dst.n_nodes, ptr(dis), stream()), "gcn_norm")
    return dis


@_guarded
def gcn_spmm(x, bias, plan_dst, plan_src, dis, relu=False):
    f = x.shape[1]
    fp = _next_pow2(f)
    if fp > 128:
        raise _lib.TipbError("gcn_spmm supports feature widths up to 128")
    if fp != f:
        x = F.pad(x, (0, fp - f))
        bias = None if bias is None else F.pad(bias, (0, fp - f))
    out = _GCNSpmmFunction.apply(x, bias, plan_dst, plan_src, dis, relu)
    return out if fp == f else out[:, :f]


# ----------------------------------------------------------------------------- hierarchy conv
class _HierFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, plan_dst, plan_src, n_source, n_target):
        x, weight = _f32c(x), _f32c(weight)
        f_in, f_out = weight.shape
        mean = torch.empty((n_target, f_in), dtype=torch.float32, device=x.device)
        out = torch.empty((n_target, f_out), dtype=torch.float32, device=x.device)
        check(lib().tipb_hier_fwd(ptr(plan_dst.buf), plan_dst.n_entries, n_source, n_target, ptr(x), ptr(weight), f_in,
                                  f_out, ptr(mean), ptr(out), stream()), "hier_fwd")
        ctx.save_for_backward(mean, weight)
        ctx.plan_dst, ctx.plan_src, ctx.n_source, ctx.n_target = plan_dst, plan_src, n_source, n_target
        return out

    @staticmethod
    def backward(ctx, grad_out):
        mean, weight = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        f_in, f_out = weight.shape
        n_source, n_target = ctx.n_source, ctx.n_target
        L = lib()
        d_x = torch.empty((n_source + n_target, f_in), dtype=torch.float32, device=grad_out.device)
        d_w = torch.empty_like(weight)
        ws = workspace(L.tipb_hier_workspace_bytes(n_source, n_target, f_in, f_out), grad_out.device)
        check(L.tipb_hier_bwd(ptr(ctx.plan_src.buf), ctx.plan_src.n_entries, n_source, n_target, ptr(ctx.plan_dst.inv_deg),
                              ptr(mean), ptr(weight), ptr(grad_out), f_in, f_out, ptr(d_x), ptr(d_w), ptr(ws), ws.numel(),
                              stream()), "hier_bwd")
        return d_x, d_w, None, None, None, None


@_guarded
def hier_conv(x, weight, plan_dst, plan_src, n_source, n_target):
    f_in, f_out = weight.shape
    fi, fo = _next_pow2(f_in), _next_pow2(f_out)
    if (fi, fo) != (f_in, f_out):
        x = F.pad(x, (0, fi - f_in))
        weight = F.pad(weight, (0, fo - f_out, 0, fi - f_in))
    out = _HierFunction.apply(x, weight, plan_dst, plan_src, n_source, n_target)
    return out if fo == f_out else out[:, :f_out]


# ----------------------------------------------------------------------------- decoder
class _DecoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, weight, edge_index, edge_type, sigmoid):
        z, weight = _f32c(z), _f32c(weight)
        edge_index, edge_type = _i64c(edge_index), _i64c(edge_type)
        n_edges = edge_index.shape[1]
        out = torch.empty(n_edges, dtype=torch.float32, device=z.device)
        check(lib().tipb_decoder_fwd(ptr(z), ptr(weight), ptr(edge_index), ptr(edge_type), n_edges, z.shape[0],
                                     weight.shape[0], z.shape[1], int(sigmoid), ptr(out), stream()), "decoder_fwd")
        ctx.save_for_backward(z, weight, edge_index, edge_type)
        ctx.sigmoid = sigmoid
        return out

    @staticmethod
    def backward(ctx, grad_out):
        z, weight, edge_index, edge_type = ctx.saved_tensors
        grad_out = _f32c(grad_out)
        n_edges, n_nodes, n_rel, dim = edge_index.shape[1], z.shape[0], weight.shape[0], z.shape[1]
        plan = _decoder_grad_plan(edge_index, edge_type, n_nodes, n_rel)
        L = lib()
        d_z, d_w = torch.empty_like(z), torch.empty_like(weight)
        ws = workspace(L.tipb_decoder_workspace_bytes(n_edges, n_nodes, n_rel, dim), z.device)
        check(L.tipb_decoder_bwd(ptr(plan.buf), n_edges, n_nodes, n_rel, ptr(z), ptr(weight), ptr(grad_out), dim,
                                 int(ctx.sigmoid), ptr(d_z), ptr(d_w), ptr(ws), ws.numel(), stream()), "decoder_bwd")
        return d_z, d_w, None, None, None


_dec_plans = {}     # (n_edges, n_nodes, n_rel, device) -> [plan, weakref(edge_index), weakref(edge_type), versions]


def _decoder_grad_plan(edge_index, edge_type, n_nodes, n_rel):
    """doubled typed CSR for the decoder gradient.  The decoder is called with FRESH negatives every step
    (src/layers.py:333-336), so these plans do not go through the identity-keyed `cached_plan` (which would pin one
    2E-entry plan plus its edge tensors per call): ONE buffer per shape is rebuilt in place, and reused as is only
    while the very same tensor objects (held weakly) are passed again unmodified -- the positive edges."""
    key = (int(edge_index.shape[1]), int(n_nodes), int(n_rel), str(edge_index.device))
    versions = _versions(edge_index, edge_type)
    hit = _dec_plans.get(key)
    if hit is not None and hit[1]() is edge_index and hit[2]() is edge_type and hit[3] == versions:
        return hit[0]
    if hit is None:
        while len(_dec_plans) >= 4:                       # a handful of shapes (train / test, pos / neg share one)
            _dec_plans.pop(next(iter(_dec_plans)))
        plan = TypedCSR(edge_index.shape[1], n_nodes, n_rel, edge_index.device, by_src=False, doubled=True,
                        rel_major=True)
    else:
        plan = hit[0]
    plan.build(edge_index, edge_type)
    _dec_plans[key] = [plan, weakref.ref(edge_index), weakref.ref(edge_type), versions]
    return plan


@_guarded
def decoder_score(z, weight, edge_index, edge_type, sigmoid=True):
    dim = z.shape[1]
    dp = _next_pow2(dim)
    if dp > 32:
        raise _lib.TipbError("decoder supports embedding widths up to 32")
    if dp != dim:
        z, weight = F.pad(z, (0, dp - dim)), F.pad(weight, (0, dp - dim))
    return _DecoderFunction.apply(z, weight, edge_index, edge_type, sigmoid)


class _BCELossFunction(torch.autograd.Function):
    """loss = -mean(log(sig(pos)+eps)) - mean(log(1-sig(neg)+eps)); gradient computed in the forward pass."""

    @staticmethod
    def forward(ctx, z, weight, plan_pos, plan_neg, neg_stream=None):
        z, weight = _f32c(z), _f32c(weight)
        n_nodes, dim = z.shape
        n_rel = weight.shape[0]
        L = lib()
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        d_z, d_w = torch.empty_like(z), torch.empty_like(weight)
        need = max(L.tipb_decoder_workspace_bytes(plan_pos.n_edges, n_nodes, n_rel, dim),
                   L.tipb_decoder_workspace_bytes(plan_neg.n_edges, n_nodes, n_rel, dim))
        ws = workspace(need, z.device)
        for acc, (plan, sign) in enumerate(((plan_pos, 1), (plan_neg, -1))):
            if sign < 0 and neg_stream is not None:
                # the negative plan is built on a side stream; the positive pass above did not need it
                torch.cuda.current_stream(z.device).wait_stream(neg_stream)
            # a by-target plan (one listing per directed edge) is only ever passed for a mirrored edge set
            fused = L.tipb_decoder_bce_fused if plan.doubled else L.tipb_decoder_bce_fused_mirrored
            check(fused(ptr(plan.buf), plan.n_edges, n_nodes, n_rel, ptr(z), ptr(weight), dim, sign, acc,
                        ptr(loss), ptr(d_z), ptr(d_w), ptr(ws), ws.numel(), stream()), "decoder_bce_fused")
        ctx.save_for_backward(d_z, d_w)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        d_z, d_w = ctx.saved_tensors
        return d_z * grad_loss, d_w * grad_loss, None, None, None


_mirror_cache = {}


@_guarded
def edges_mirrored(edge_index, range_list):
    """True iff every relation range is [pairs..., the same pairs with rows swapped...] (src/utils.py:17-23).
    Checked on the device once per (tensor, version); one host synchronisation at that time."""
    key = (_tensor_key(edge_index), _tensor_key(range_list))
    versions = _versions(edge_index, range_list)
    hit = _mirror_cache.get(key)
    if hit is None or hit[0] != versions:
        if len(_mirror_cache) > 64:
            _mirror_cache.clear()
        flag = torch.zeros(1, dtype=torch.int32, device=edge_index.device)
        rl = _i64c(range_list.to(torch.long))
        check(lib().tipb_edges_mirrored(ptr(_i64c(edge_index)), ptr(rl), edge_index.shape[1], rl.shape[0], ptr(flag),
                                        stream()), "edges_mirrored")
        hit = _mirror_cache[key] = (versions, bool(int(flag.item())), (edge_index, range_list))
    return hit[1]


def positive_decoder_plan(edge_index, n_nodes, n_rel, range_list):
    """index structure for the positive pass of bce_loss: the R-GCN's own by-target plan when the edge set is
    mirrored (always the case for process_edges output), else the doubled relation-major plan"""
    if edges_mirrored(edge_index, range_list):
        return cached_plan(edge_index, n_nodes, n_rel, by_src=False, edge_type=None, range_list=range_list)
    return cached_plan(edge_index, n_nodes, n_rel, range_list=range_list, by_src=False, doubled=True, rel_major=True)


@_guarded
def bce_loss(z, weight, plan_pos, plan_neg, neg_stream=None):
    dim = z.shape[1]
    dp = _next_pow2(dim)
    if dp > 32:
        raise _lib.TipbError("decoder supports embedding widths up to 32")
    if dp != dim:
        z, weight = F.pad(z, (0, dp - dim)), F.pad(weight, (0, dp - dim))
    return _BCELossFunction.apply(z, weight, plan_pos, plan_neg, neg_stream)


# ----------------------------------------------------------------------------- fused pair pass
def _pair_items(ranges, chunk, slot_base):
    """work items (relation, first pair, pair count, slot) of one pass: every relation's pairs [start, end) cut into
    near-equal chunks of at most `chunk` pairs; slots numbered relation-major from `slot_base`; rows ordered largest
    first (the launch order).  -> (items int32 [n,4], rel_slot_ptr int32 [n_rel+1] absolute, n_slots)"""
    ranges = np.asarray(ranges, dtype=np.int64).reshape(-1, 2)
    n = ranges[:, 1] - ranges[:, 0]
    k = -(-n // chunk)                                    # chunks per relation (0 for an empty relation)
    ptr_ = np.concatenate([[0], np.cumsum(k)]) + slot_base
    total = int(k.sum())
    rel = np.repeat(np.arange(ranges.shape[0]), k)
    c = np.arange(total) - np.repeat(ptr_[:-1] - slot_base, k)      # chunk index inside its relation
    base, extra = n[rel] // np.maximum(k[rel], 1), n[rel] % np.maximum(k[rel], 1)
    cnt = base + (c < extra)
    start = ranges[rel, 0] + c * base + np.minimum(c, extra)
    items = np.stack([rel, start, cnt, np.arange(total) + slot_base], axis=1).astype(np.int32).reshape(-1, 4)
    order = np.argsort(-items[:, 2], kind="stable")
    return np.ascontiguousarray(items[order]), ptr_.astype(np.int32), total


_item_tables = {}      # (range bytes, chunk) -> host/device item tables: they depend on the relation sizes only


class PairPlan(object):
    """Static tables of the fused pair pass (csrc/pair_pass.cu) for one mirrored, relation-sorted edge set: the
    positive pairs (first half of every relation range, packed), the work items of the positive and the negative
    pass, and their relation-major slot ranges.  Built once per graph; the item tables depend on the relation sizes
    only and are shared between graphs with the same ranges (`rl_host`: a host copy of range_list if the caller has
    one, otherwise one small device->host copy is made)."""

    def __init__(self, edge_index, n_nodes, n_rel, range_list, rl_host=None):
        L = lib()
        dev = edge_index.device
        self.device, self.n_nodes, self.n_rel = dev, int(n_nodes), int(n_rel)
        self.n_edges = int(edge_index.shape[1])
        rl_dev = _i64c(range_list.to(device=dev, dtype=torch.long))
        rl = np.ascontiguousarray(rl_dev.cpu().numpy() if rl_host is None else np.asarray(rl_host, dtype=np.int64))
        # work-item size: ~3 items per resident CTA slot (2 per SM) so that the largest-first schedule has no long tail
        # (a rank of a sharded run owns few relations), whole tiles of 2048 pairs, at most the library's chunk
        chunk_max = int(L.tipb_pair_chunk())
        slots = 2 * torch.cuda.get_device_properties(dev).multi_processor_count
        total_pairs = int(rl[-1, 1]) * 3 // 2 if rl.shape[0] else 0          # E/2 positive + E negative pairs
        chunk = min(chunk_max, max(2048, -(-total_pairs // (3 * slots * 2048)) * 2048))
        key = (rl.tobytes(), chunk, str(dev))
        tables = _item_tables.get(key)
        if tables is None:
            check_cumulative_ranges(rl, self.n_edges)
            items_pos, ptr_pos, n_pos = _pair_items(rl // 2, chunk, 0)
            items_neg, ptr_neg, n_neg = _pair_items(rl, chunk, n_pos)
            items_neg[:, 3] |= 1 << 30                       # pass flag (csrc/pair_pass.cu: PP_NEG_FLAG)
            items = np.concatenate([items_pos, items_neg])
            items = np.ascontiguousarray(items[np.argsort(-items[:, 2], kind="stable")])    # one launch, largest first
            if len(_item_tables) > 8:
                _item_tables.clear()
            tables = _item_tables[key] = (n_pos + n_neg, int(items.shape[0]), torch.from_numpy(items).to(dev),
                                          torch.from_numpy(np.concatenate([ptr_pos, ptr_neg])).to(dev))
        self.n_slots, self.n_items, self.items, self.rel_slot_ptr = tables
        self.pos_packed = torch.empty(max(self.n_edges // 2, 1), dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self._rl_dev = rl_dev
        self.repack(edge_index)

    def repack(self, edge_index):
        """(re)pack the positive pairs from edge_index (same ranges): device work only, no host synchronisation"""
        with torch.cuda.device(self.device):
            check(lib().tipb_pack_half_pairs(ptr(_i64c(edge_index)), ptr(self._rl_dev), self.n_edges, self.n_rel,
                                             self.n_nodes, ptr(self.pos_packed), ptr(self.status), stream()),
                  "pack_half_pairs")

    def check_status(self):
        """One host sync: raises if an edge endpoint was out of range when the pairs were packed."""
        if int(self.status.item()) != 0:
            raise IndexError("edge_index contains an out-of-range node id")
        return self


_pair_plans = {}


def pair_plan(edge_index, n_nodes, n_rel, range_list, dim, rl_host=None, validate=True):
    """the PairPlan of a graph, or None when the fused pair pass does not apply: the edge set must be mirrored
    (src/utils.py:17-23) and z (n_nodes x dim, dim padded to a power of two) must fit in shared memory"""
    if not lib().tipb_pair_pass_supported(int(n_nodes), _next_pow2(int(dim))):
        return None
    if not edges_mirrored(edge_index, range_list):
        return None
    key = (_tensor_key(edge_index), _tensor_key(range_list), int(n_nodes), int(n_rel))
    versions = _versions(edge_index, range_list)
    hit = _pair_plans.get(key)
    if hit is None or hit[0] != versions:
        if len(_pair_plans) > 8:
            _pair_plans.clear()
        plan = PairPlan(edge_index, n_nodes, n_rel, range_list, rl_host=rl_host)
        if validate:
            plan.check_status()
        hit = _pair_plans[key] = (versions, plan, (edge_index, range_list))
    return hit[1]


class _PairBCEFunction(torch.autograd.Function):
    """loss = -mean(log(sig(pos)+eps)) - mean(log(1-sig(neg)+eps)) over packed pairs; gradient computed in forward"""

    @staticmethod
    def forward(ctx, z, weight, plan, neg_packed, neg_stream=None):
        z, weight = _f32c(z), _f32c(weight)
        n_nodes, dim = z.shape
        n_rel = weight.shape[0]
        assert n_nodes == plan.n_nodes and n_rel == plan.n_rel and neg_packed.numel() == plan.n_edges
        L = lib()
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        d_z, d_w = torch.empty_like(z), torch.empty_like(weight)
        ws = workspace(L.tipb_pair_workspace_bytes(plan.n_slots, n_nodes, dim), z.device, "pair")
        e = float(max(plan.n_edges, 1))
        if neg_stream is not None:      # the negatives were sampled on a side stream while the encoder ran
            torch.cuda.current_stream(z.device).wait_stream(neg_stream)
        check(L.tipb_pair_bce_pass(ptr(plan.pos_packed), ptr(neg_packed), ptr(plan.items), plan.n_items, plan.n_slots,
                                   n_nodes, ptr(z), ptr(weight), dim, 2.0 / e, 1.0 / e, ptr(ws), ws.numel(), stream()),
              "pair_bce_pass")
        check(L.tipb_pair_bce_finish(ptr(plan.rel_slot_ptr), plan.n_slots, n_nodes, n_rel, dim, ptr(loss), ptr(d_z), ptr(d_w),
                                     ptr(ws), ws.numel(), stream()), "pair_bce_finish")
        ctx.save_for_backward(d_z, d_w)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        d_z, d_w = ctx.saved_tensors
        return d_z * grad_loss, d_w * grad_loss, None, None, None


@_guarded
def pair_bce_loss(z, weight, plan, neg_packed, neg_stream=None):
    """TIP.forward's loss (src/layers.py:335-340) and its gradient through the fused pair pass"""
    dim = z.shape[1]
    dp = _next_pow2(dim)
    if dp != dim:
        z, weight = F.pad(z, (0, dp - dim)), F.pad(weight, (0, dp - dim))
    return _PairBCEFunction.apply(z, weight, plan, neg_packed, neg_stream)


def unpack_pairs(packed):
    """packed (row << 16 | col) int32 [n] -> int64 [2, n], the reference's LongTensor layout"""
    out = torch.empty((2, packed.numel()), dtype=torch.long, device=packed.device)
    with torch.cuda.device(packed.device):
        check(lib().tipb_unpack_pairs(ptr(packed), packed.numel(), ptr(out), stream()), "unpack_pairs")
    return out


@_guarded
def decoder_sweep(z, weight, sigmoid=True):
    z, weight = _f32c(z), _f32c(weight)
    n, dim = z.shape
    out = torch.empty((weight.shape[0], n, n), dtype=torch.float32, device=z.device)
    check(lib().tipb_decoder_sweep(ptr(z), ptr(weight), n, weight.shape[0], dim, int(sigmoid), ptr(out), stream()),
          "decoder_sweep")
    return out


def check_cumulative_ranges(range_list, n_edges):
    """range_list must be the cumulative [start, end) table of src/utils.py:26-32 tiling [0, n_edges)"""
    rl = np.asarray(range_list.detach().cpu() if torch.is_tensor(range_list) else range_list).astype(np.int64)
    ok = rl.ndim == 2 and rl.shape[1] == 2 and rl.shape[0] > 0
    ok = ok and rl[0, 0] == 0 and rl[-1, 1] == n_edges and np.all(rl[:, 1] >= rl[:, 0]) and np.all(rl[1:, 0] == rl[:-1, 1])
    if not ok:
        raise ValueError("range_list must be the cumulative [start,end) table of src/utils.py:26-32 covering every edge")


@_guarded
def eval_auprc_auroc_ap(pos_score, neg_score, range_list):
    """record[3, n_rel] (float64, on the device): auprc, auroc, ap per relation -- src/layers.py:353-369 without the
    861 host round trips"""
    pos_score, neg_score = _f32c(pos_score.detach()), _f32c(neg_score.detach())
    if not pos_score.is_cuda:
        raise _lib.TipbError("eval_auprc_auroc_ap takes CUDA tensors only (there is no CPU path)")
    rl = _i64c(range_list.to(device=pos_score.device, dtype=torch.long))
    n_edges, n_rel = pos_score.numel(), rl.shape[0]
    assert neg_score.numel() == n_edges and rl.dim() == 2 and rl.shape[1] == 2
    check_cumulative_ranges(rl, n_edges)          # evaluation is not on the training path: one small host copy
    L = lib()
    record = torch.empty((3, n_rel), dtype=torch.float64, device=pos_score.device)
    ws = workspace(L.tipb_eval_workspace_bytes(n_edges, n_rel), pos_score.device, "eval")
    check(L.tipb_eval_auprc_auroc_ap(ptr(pos_score), ptr(neg_score), ptr(rl), n_edges, n_rel, ptr(record), ptr(ws),
                                     ws.numel(), stream()), "eval_auprc_auroc_ap")
    return record


__all__ = ["eval_auprc_auroc_ap", "TypedCSR", "cached_plan", "rgcn_conv", "gcn_norm", "gcn_spmm", "hier_conv", "decoder_score", "bce_loss",
           "decoder_sweep", "workspace", "math"]
