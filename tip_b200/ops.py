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
    _edge_plans.clear()
    _sparse_feats.clear()


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
    def forward(ctx, x, bias, plan_dst, plan_src, dis, relu, pad_rows):
        x = _f32c(x)
        n, f = x.shape
        L = lib()
        # `pad_rows` zero rows under the result: torch.cat((x_prot, hdrug)) of src/layers.py:533 without the copy
        out = torch.empty((n + pad_rows, f), dtype=torch.float32, device=x.device)
        check(L.tipb_gcn_spmm(ptr(plan_dst.buf), plan_dst.n_entries, n, ptr(dis), ptr(dis), ptr(x),
                              None if bias is None else ptr(_f32c(bias)), f, int(relu), ptr(out), stream()), "gcn_spmm")
        if pad_rows:
            check(L.tipb_fill_zero(out.data_ptr() + n * f * 4, pad_rows * f * 4, stream()), "fill_zero")
        ctx.plan_src, ctx.dis, ctx.relu, ctx.has_bias, ctx.n = plan_src, dis, relu, bias is not None, n
        if relu:
            ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = _f32c(grad_out)          # rows beyond n (the zero rows) take no gradient
        n, f = ctx.n, grad_out.shape[1]
        L = lib()
        d_bias = torch.empty(f, dtype=torch.float32, device=grad_out.device) if ctx.has_bias else None
        g = grad_out
        if ctx.relu or ctx.has_bias:        # ReLU mask and bias gradient in one library call
            (out,) = ctx.saved_tensors if ctx.relu else (None,)
            g = torch.empty((n, f), dtype=torch.float32, device=grad_out.device) if ctx.relu else grad_out
            ws = workspace(L.tipb_relu_grad_colsum_workspace_bytes(f), grad_out.device, "colsum")
            check(L.tipb_relu_grad_colsum(ptr(grad_out), ptr(out), n, f, ptr(g) if ctx.relu else None, ptr(d_bias),
                                          ptr(ws), ws.numel(), stream()), "relu_grad_colsum")
        d_x = torch.empty((n, f), dtype=torch.float32, device=grad_out.device)
        plan = ctx.plan_src
        check(L.tipb_gcn_spmm(ptr(plan.buf), plan.n_entries, n, ptr(ctx.dis), ptr(ctx.dis), ptr(g), None, f, 0,
                              ptr(d_x), stream()), "gcn_spmm(bwd)")
        return d_x, d_bias, None, None, None, None, None


def gcn_norm(plan_dst):
    dis = torch.empty(plan_dst.n_nodes, dtype=torch.float32, device=plan_dst.device)
    with torch.cuda.device(plan_dst.device):
        check(lib().tipb_gcn_norm(ptr(plan_dst.buf), plan_dst.n_entries, plan_dst.n_nodes, ptr(dis), stream()), "gcn_norm")
    return dis


@_guarded
def gcn_spmm(x, bias, plan_dst, plan_src, dis, relu=False, pad_rows=0):
    f = x.shape[1]
    fp = _next_pow2(f)
    if fp > 128:
        raise _lib.TipbError("gcn_spmm supports feature widths up to 128")
    if fp != f:
        x = F.pad(x, (0, fp - f))
        bias = None if bias is None else F.pad(bias, (0, fp - f))
    out = _GCNSpmmFunction.apply(x, bias, plan_dst, plan_src, dis, relu, int(pad_rows))
    return out if fp == f else out[:, :f]


# ----------------------------------------------------------------------------- small dense products
class _MatmulFunction(torch.autograd.Function):
    """y = x @ w (w [k, n]) or x @ w^T (w [n, k], torch.nn.Linear layout) (+ bias) (+ ReLU) through tipb_gemm"""

    @staticmethod
    def forward(ctx, x, w, bias, trans_b, relu):
        x, w = _f32c(x), _f32c(w)
        m, k = x.shape
        n = w.shape[0] if trans_b else w.shape[1]
        assert (w.shape[1] if trans_b else w.shape[0]) == k, "matmul: inner dimensions differ"
        y = torch.empty((m, n), dtype=torch.float32, device=x.device)
        check(lib().tipb_gemm(ptr(x), ptr(w), None if bias is None else ptr(_f32c(bias)), m, n, k, int(trans_b), int(relu),
                              ptr(y), stream()), "gemm")
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.trans_b, ctx.relu, ctx.has_bias = trans_b, relu, bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        gy = _f32c(gy)
        m, k = x.shape
        n = gy.shape[1]
        L = lib()
        g, d_bias = gy, None
        if ctx.relu or ctx.has_bias:
            d_bias = torch.empty(n, dtype=torch.float32, device=gy.device) if ctx.has_bias else None
            g = torch.empty_like(gy) if ctx.relu else gy
            ws = workspace(L.tipb_relu_grad_colsum_workspace_bytes(n), gy.device, "colsum")
            check(L.tipb_relu_grad_colsum(ptr(gy), ptr(y) if ctx.relu else None, m, n, ptr(g) if ctx.relu else None,
                                          ptr(d_bias), ptr(ws), ws.numel(), stream()), "relu_grad_colsum")
        d_x = d_w = None
        if ctx.needs_input_grad[0]:
            d_x = torch.empty_like(x)       # g w^T (w [k,n]) or g w (w [n,k])
            check(L.tipb_gemm(ptr(g), ptr(w), None, m, k, n, int(not ctx.trans_b), 0, ptr(d_x), stream()), "gemm(d_x)")
        if ctx.needs_input_grad[1]:
            d_w = torch.empty_like(w)
            a, b, mm, nn = (g, x, n, k) if ctx.trans_b else (x, g, k, n)     # g^T x [n,k]  |  x^T g [k,n]
            ws = workspace(L.tipb_gemm_tn_workspace_bytes(mm, nn), gy.device, "gemm_tn")
            check(L.tipb_gemm_tn(ptr(a), ptr(b), m, mm, nn, ptr(d_w), ptr(ws), ws.numel(), stream()), "gemm_tn")
        return d_x, d_w, d_bias, None, None


@_guarded
def matmul(x, w, trans_b=False, bias=None, relu=False):
    """x [m,k] times w [k,n] (or w^T with w [n,k] when trans_b), optional bias and ReLU, with gradients -- library kernels"""
    if not x.is_cuda:
        raise _lib.TipbError("matmul: CUDA tensors only -- tip_b200 has no CPU fallback")
    return _MatmulFunction.apply(x, w, bias, bool(trans_b), bool(relu))


class _TransposeFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        out = torch.empty((x.shape[1], x.shape[0]), dtype=torch.float32, device=x.device)
        check(lib().tipb_transpose(ptr(x), x.shape[0], x.shape[1], ptr(out), stream()), "transpose")
        return out

    @staticmethod
    def backward(ctx, g):
        g = _f32c(g)
        out = torch.empty((g.shape[1], g.shape[0]), dtype=torch.float32, device=g.device)
        check(lib().tipb_transpose(ptr(g), g.shape[0], g.shape[1], ptr(out), stream()), "transpose")
        return out


@_guarded
def transpose2d(x):
    """contiguous x^T (and a contiguous gradient in the parameter's own layout)"""
    return _TransposeFunction.apply(x)


class _DrugInputFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, e, d_norm, h, mode):
        e, h, d_norm = _f32c(e), _f32c(h), _f32c(d_norm.reshape(-1))
        n, fe = e.shape
        fh = h.shape[1]
        assert h.shape[0] == n and d_norm.numel() == n
        out = torch.empty((n, fe + fh if mode == 0 else fe), dtype=torch.float32, device=e.device)
        check(lib().tipb_drug_input_fwd(ptr(e), ptr(d_norm), ptr(h), n, fe, fh, mode, ptr(out), stream()), "drug_input_fwd")
        ctx.save_for_backward(d_norm)
        ctx.dims = (n, fe, fh, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        (d_norm,) = ctx.saved_tensors
        n, fe, fh, mode = ctx.dims
        g = _f32c(g)
        d_e = torch.empty((n, fe), dtype=torch.float32, device=g.device)
        d_h = torch.empty((n, fh), dtype=torch.float32, device=g.device)
        check(lib().tipb_drug_input_bwd(ptr(g), ptr(d_norm), n, fe, fh, mode, ptr(d_e), ptr(d_h), stream()), "drug_input_bwd")
        return d_e, None, d_h, None


@_guarded
def drug_input(embed_out, d_norm, hier_out, mod):
    """FMEncoder's glue (src/layers.py:541-547): embed_out / d_norm.view(-1, 1), then cat(dim=1) | + with hier_out"""
    return _DrugInputFunction.apply(embed_out, d_norm, hier_out, 0 if mod == "cat" else 1)


def row_scale(x, norm):
    """x / norm.view(-1, 1) (HierEncoder, src/layers.py:572): the drug-input kernel with an empty second block"""
    empty = torch.empty((x.shape[0], 0), dtype=torch.float32, device=x.device)
    return _DrugInputFunction.apply(x, norm, empty, 0)


def _scale_pair(d_z, d_w, grad_loss):
    """(d_z, d_w) * grad_loss (a device scalar) in one library launch"""
    o_z, o_w = torch.empty_like(d_z), torch.empty_like(d_w)
    with torch.cuda.device(d_z.device):
        check(lib().tipb_scale2(ptr(d_z), d_z.numel(), ptr(d_w), d_w.numel(), ptr(_f32c(grad_loss.reshape(1))), ptr(o_z),
                                ptr(o_w), stream()), "scale2")
    return o_z, o_w


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
        return _scale_pair(d_z, d_w, grad_loss) + (None, None, None)


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
        return _scale_pair(d_z, d_w, grad_loss) + (None, None, None)


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


# ----------------------------------------------------------------------------- ablation operators
_edge_plans = {}     # (n_edges, n_nodes, n_rel, device) -> [plan_by_src, plan_by_dst]: one pair of buffers per shape


def _edge_plan_pair(edge_index, edge_type, n_nodes, n_rel):
    key = (int(edge_index.shape[1]), int(n_nodes), int(n_rel), str(edge_index.device))
    pair = _edge_plans.get(key)
    if pair is None:
        while len(_edge_plans) >= 4:
            _edge_plans.pop(next(iter(_edge_plans)))
        pair = _edge_plans[key] = [TypedCSR(edge_index.shape[1], n_nodes, n_rel, edge_index.device, by_src=b) for b in (True, False)]
    for plan in pair:
        plan.build(edge_index, edge_type)
    return pair


class _NNDecoderGather(torch.autograd.Function):
    """score[e] = act(table_a[i_e, r_e] + table_b[j_e, r_e]); the table gradients are segment sums (no atomics)"""

    @staticmethod
    def forward(ctx, table_a, table_b, edge_index, edge_type, sigmoid):
        table_a, table_b = _f32c(table_a), _f32c(table_b)
        edge_index, edge_type = _i64c(edge_index), _i64c(edge_type)
        n_nodes, n_rel = table_a.shape
        n_edges = edge_index.shape[1]
        out = torch.empty(n_edges, dtype=torch.float32, device=table_a.device)
        check(lib().tipb_nn_decoder_fwd(ptr(table_a), ptr(table_b), ptr(edge_index), ptr(edge_type), n_edges, n_nodes, n_rel,
                                        int(sigmoid), ptr(out), stream()), "nn_decoder_fwd")
        ctx.save_for_backward(out, edge_index, edge_type)
        ctx.dims, ctx.sigmoid = (n_nodes, n_rel), sigmoid
        return out

    @staticmethod
    def backward(ctx, grad_out):
        out, edge_index, edge_type = ctx.saved_tensors
        n_nodes, n_rel = ctx.dims
        n_edges = edge_index.shape[1]
        by_src, by_dst = _edge_plan_pair(edge_index, edge_type, n_nodes, n_rel)
        d_a = torch.empty((n_nodes, n_rel), dtype=torch.float32, device=out.device)
        d_b = torch.empty_like(d_a)
        g_ws = torch.empty(max(n_edges, 1), dtype=torch.float32, device=out.device)
        check(lib().tipb_nn_decoder_bwd(ptr(by_src.buf), ptr(by_dst.buf), n_edges, n_nodes, n_rel, ptr(_f32c(grad_out)),
                                        ptr(out), int(ctx.sigmoid), ptr(d_a), ptr(d_b), ptr(g_ws), stream()), "nn_decoder_bwd")
        return d_a, d_b, None, None, None


@_guarded
def nn_decoder_score(z, w1_l1, w1_l2, w2_l1, w2_l2, edge_index, edge_type, sigmoid=True):
    """NNDecoder.forward (src/layers.py:618-631).  The hidden layers are per NODE (relu(z W) does not depend on the
    edge), and so are the per-(node, relation) dot products: two [N, l1] and two [N, R] dense products replace the
    reference's E x l1 gathers; the per-edge work is the gather of two scalars."""
    if not z.is_cuda:
        raise _lib.TipbError("NNDecoder: CUDA tensors only -- tip_b200 has no CPU fallback")
    h1 = matmul(z, w1_l1, relu=True)
    h2 = matmul(z, w2_l1, relu=True)
    table_a = matmul(h1, w1_l2, trans_b=True)
    table_b = matmul(h2, w2_l2, trans_b=True)
    return _NNDecoderGather.apply(table_a, table_b, edge_index, edge_type, sigmoid)


class SparseFeatures(object):
    """index structures of a sparse COO feature matrix S [n_rows, n_cols] (general drug features, data/utils.py:117-132):
    typed CSRs (n_rel = 1) of the entries by row (forward, S @ dense) and by column (gradient, S^T @ g)"""

    def __init__(self, x):
        x = x.coalesce()
        idx, self.values = x.indices(), _f32c(x.values())
        self.n_rows, self.n_cols = int(x.shape[0]), int(x.shape[1])
        n = max(self.n_rows, self.n_cols)
        # "edge" col -> row: by-target plan groups the entries by output row, by-source plan by column
        edges = torch.stack([idx[1], idx[0]]).contiguous()
        self.nnz = int(edges.shape[1])
        self.by_row = TypedCSR(self.nnz, n, 1, x.device, by_src=False).build(edges).check_status()
        self.by_col = TypedCSR(self.nnz, n, 1, x.device, by_src=True).build(edges).check_status()
        self.n = n


class _SpmmValuesFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, feats):
        dense = _f32c(dense)
        assert dense.shape[0] == feats.n_cols
        f = dense.shape[1]
        out = torch.empty((feats.n, f), dtype=torch.float32, device=dense.device)
        check(lib().tipb_spmm_values(ptr(feats.by_row.buf), feats.nnz, feats.n, ptr(feats.values), ptr(dense), f, ptr(out),
                                     stream()), "spmm_values")
        ctx.feats = feats
        return out[:feats.n_rows]

    @staticmethod
    def backward(ctx, g):
        feats = ctx.feats
        f = g.shape[1]
        gp = g
        if feats.n != feats.n_rows:       # rows of the square index space beyond n_rows have no entries
            gp = torch.zeros((feats.n, f), dtype=torch.float32, device=g.device)
            gp[:feats.n_rows] = g
        out = torch.empty((feats.n, f), dtype=torch.float32, device=g.device)
        check(lib().tipb_spmm_values(ptr(feats.by_col.buf), feats.nnz, feats.n, ptr(feats.values), ptr(_f32c(gp)), f, ptr(out),
                                     stream()), "spmm_values(bwd)")
        return out[:feats.n_cols], None


_sparse_feats = {}


@_guarded
def sparse_matmul(x_sparse, dense):
    """torch.matmul(x_sparse, dense) for a sparse COO feature matrix, through the library (plans cached per tensor)"""
    key = (x_sparse._indices().data_ptr(), x_sparse._values().data_ptr(), tuple(x_sparse.shape))
    hit = _sparse_feats.get(key)
    if hit is None:
        if len(_sparse_feats) > 8:
            _sparse_feats.clear()
        hit = _sparse_feats[key] = (SparseFeatures(x_sparse), x_sparse)
    return _SpmmValuesFunction.apply(dense, hit[0])


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


__all__ = ["eval_auprc_auroc_ap", "matmul", "transpose2d", "drug_input", "nn_decoder_score", "sparse_matmul", "TypedCSR", "cached_plan", "rgcn_conv", "gcn_norm", "gcn_spmm", "hier_conv", "decoder_score", "bce_loss",
           "decoder_sweep", "workspace", "math"]
