"""Drop-in for the reference's src/neg_sampling.py, running on the GPU.

    typed_negative_sampling(pos_edge_index, num_nodes, range_list) -> int64 [2, E]

Same name, arguments and result as src/neg_sampling.py:22-26 -- bit for bit, because
the device keeps a numpy-compatible MT19937 state (the reference draws from the
global `np.random` stream seeded at import, src/layers.py:14).  `seed()`,
`get_state()` and `set_state()` mirror `np.random.seed/get_state/set_state` so the
stream can be handed over from numpy (e.g. after the reference's CPU-side
`process_edges`, which consumes the same stream) and back.

The raw MT19937 words depend on the state only, not on the graph, so after every call
the words for the NEXT call are generated on a side stream while the caller goes on
with the encoder/decoder work (`set_prefetch(False)` turns that off).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .ops import _i64c, workspace

BITMAP_LIMIT_GIB = 64  # per device; 10^4 drugs x 4000 relations = 50 GB unsharded, 6 GB per rank on 8 GPUs
Z_SIGMA = 7.0          # half-width of the offset brackets, in standard deviations of the retry-count model
_member_cache = {}     # positive-pair bitmaps + bracket tables, keyed on the identity of (pos_edge_index, range_list)
_rng = {}              # device index -> _DeviceRng
_prefetch = True
_chunked = True        # generate the MT19937 stream with one CTA per CHUNK words (csrc/mt_jump.cu)
CHUNK = 454 * 368      # words per CTA: 64 chunks for the 10.6 M words of a polypharmacy-shape step


class _DeviceRng(object):
    def __init__(self, device):
        self.device = device
        self.state = torch.empty(625, dtype=torch.int32, device=device)   # 624 key words + position
        self.words = None       # untempered stream generated from `state` (624 + n_new words)
        self.n_new = 0
        self.valid = False      # `words` matches the current state
        self.side = torch.cuda.Stream(device=device)      # next step's words: lowest priority, nobody waits for it
        self.ready = None       # event: the prefetch on `side` has finished
        self.forked = False     # a prefetch is in flight that the caller's stream has not joined yet
        self.polys = None       # jump-ahead polynomials [n, 624] and the chunk start windows they produce
        self.windows = None
        with torch.cuda.device(device):
            check(lib().tipb_mt19937_seed(ptr(self.state), 1111, stream()), "mt19937_seed")   # src/layers.py:14

    def join(self):
        """make the current stream wait for an in-flight prefetch (it reads `state` and writes `words`)"""
        if self.forked:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
            self.forked = False

    def invalidate(self):
        self.join()
        self.valid = False

    def _jump_polys(self, n_polys):
        """jump polynomials x^(k*CHUNK) mod phi, k = 1..n_polys, on the device (host computation, once)"""
        if self.polys is None or self.polys.shape[0] < n_polys:
            n_polys = max(n_polys, 2 * (0 if self.polys is None else self.polys.shape[0]))
            host = np.zeros((n_polys, 624), dtype=np.uint32)
            check(lib().tipb_mt19937_jump_polys(CHUNK, n_polys, host.ctypes.data_as(C.c_void_p)), "mt19937_jump_polys")
            self.join()
            self.polys = torch.from_numpy(host.view(np.int32)).to(self.device)
            self.windows = torch.empty((n_polys, 624), dtype=torch.int32, device=self.device)
        return self.polys

    def generate(self, n_new):
        """624 + n_new untempered words from the current state: CHUNK-word pieces on one CTA each (jump-ahead)"""
        L = lib()
        n_new = int(L.tipb_mt19937_stream_words(int(n_new)))
        if self.words is None or self.n_new < n_new:
            self.join()
            self.words = torch.empty(624 + n_new, dtype=torch.int32, device=self.device)
            self.n_new = n_new
        n_chunks = int(L.tipb_mt19937_chunk_count(self.n_new, CHUNK))
        if not _chunked or n_chunks <= 1:
            check(L.tipb_mt19937_generate(ptr(self.state), ptr(self.words), self.n_new, stream()), "mt19937_generate")
            return
        polys = self._jump_polys(n_chunks - 1)
        check(L.tipb_mt19937_generate_chunked(ptr(self.state), ptr(self.words), self.n_new, CHUNK, ptr(polys),
                                              polys.shape[0], ptr(self.windows), stream()), "mt19937_generate_chunked")


def _device_of(device=None):
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TipbError("tip_b200.neg_sampling runs on CUDA devices only (there is no CPU path)")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def _get_rng(device):
    r = _rng.get(device.index)
    if r is None:
        r = _rng[device.index] = _DeviceRng(device)
    return r


def set_prefetch(flag):
    """generate the next call's MT19937 words on a side stream right after each call (default on)"""
    global _prefetch
    _prefetch = bool(flag)


def join_prefetch(device=None):
    """Order the current stream after the in-flight prefetch.  Needed at the end of a CUDA-graph-captured
    step (forked work must rejoin the capturing stream); harmless otherwise."""
    _get_rng(_device_of(device)).join()


def seed(value, device=None):
    """np.random.seed(value) for the device-side stream."""
    device = _device_of(device)
    r = _get_rng(device)
    r.invalidate()
    with torch.cuda.device(device):
        check(lib().tipb_mt19937_seed(ptr(r.state), int(value) & 0xFFFFFFFF, stream()), "mt19937_seed")


def get_state(device=None):
    """-> ('MT19937', key uint32[624], pos, 0, 0.0), the tuple np.random.set_state accepts."""
    st = _get_rng(_device_of(device)).state.cpu().numpy().view(np.uint32)
    return ("MT19937", st[:624].copy(), int(st[624]), 0, 0.0)


def set_state(state, device=None):
    """accepts np.random.get_state() tuples."""
    device = _device_of(device)
    key = np.asarray(state[1], dtype=np.uint32)
    assert key.shape == (624,)
    host = np.concatenate([key, np.array([int(state[2])], dtype=np.uint32)]).view(np.int32)
    r = _get_rng(device)
    r.invalidate()
    r.state.copy_(torch.from_numpy(host))


class _Membership(object):
    """per-relation bitmaps of the positive pairs, their popcounts, and the bracket table"""

    def __init__(self, pos_edge_index, num_nodes, range_list):
        self.member = None
        self.status = torch.zeros(1, dtype=torch.int32, device=pos_edge_index.device)
        self.build(pos_edge_index, num_nodes, range_list)

    def build(self, pos_edge_index, num_nodes, range_list):
        """(re)build bitmaps, popcounts and the bracket table; device buffers are reused when the sizes match"""
        L = lib()
        dev = pos_edge_index.device
        self.n_edges = int(pos_edge_index.shape[1])
        self.n_rel = int(range_list.shape[0])
        self.num_nodes = int(num_nodes)
        self.range_dev = _i64c(range_list.to(device=dev, dtype=torch.long))
        rl = np.ascontiguousarray(self.range_dev.cpu().numpy())        # one-time host copy (setup, not per step)
        self.rl_host = rl
        sizes = rl[:, 1] - rl[:, 0]
        if self.n_rel and not (np.all(sizes >= 0) and np.all(rl[1:, 0] == rl[:-1, 1]) and rl[0, 0] == 0
                               and rl[-1, 1] == self.n_edges):
            raise ValueError("range_list must be the cumulative [start,end) table of src/utils.py:26-32")
        nbytes = L.tipb_neg_bitmap_bytes(num_nodes, max(self.n_rel, 1))
        if nbytes > (BITMAP_LIMIT_GIB << 30):
            raise _lib.TipbError(f"positive-pair bitmaps would need {nbytes >> 30} GiB (shard the relations over more "
                                 "ranks: tip_b200.parallel)")
        if self.member is None or self.member.numel() != max(nbytes // 4, 1):
            self.member = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=dev)
        popcount = torch.zeros(max(self.n_rel, 1), dtype=torch.int32, device=dev)
        check(L.tipb_neg_bitmap_build(ptr(_i64c(pos_edge_index)), ptr(self.range_dev), self.n_edges, num_nodes,
                                      self.n_rel, ptr(self.member), ptr(popcount), stream()), "neg_bitmap_build")
        pop = np.ascontiguousarray(popcount.cpu().numpy())             # one-time host copy
        table = np.zeros((max(self.n_rel, 1), 8), dtype=np.int64)
        totals = np.zeros(4, dtype=np.int64)
        check(L.tipb_neg_table_build(rl.ctypes.data_as(C.c_void_p), pop.ctypes.data_as(C.c_void_p), self.n_rel,
                                     num_nodes, Z_SIGMA, table.ctypes.data_as(C.c_void_p),
                                     totals.ctypes.data_as(C.c_void_p)), "neg_table_build")
        self.table = torch.from_numpy(table).to(dev)
        self.sum_l, self.sum_w, max_index = int(totals[0]), int(totals[1]), int(totals[2])
        # MT words needed so that `max_index` accepted values exist: acceptance is Bernoulli(p_accept) per word
        bits = max(int(num_nodes * num_nodes - 1).bit_length(), 1)
        p_accept = float(num_nodes) ** 2 / float(1 << bits)
        need = max_index / p_accept
        self.n_new = int(need + Z_SIGMA * np.sqrt(need * (1.0 - p_accept) / p_accept + 1.0) + 4096)
        self._versions = (pos_edge_index._version, range_list._version)


def _membership(pos_edge_index, num_nodes, range_list):
    key = (pos_edge_index.data_ptr(), tuple(pos_edge_index.shape), range_list.data_ptr(), tuple(range_list.shape),
           int(num_nodes), str(pos_edge_index.device))
    m = _member_cache.get(key)
    if m is None:
        if len(_member_cache) > 16:
            _member_cache.clear()
        m = _Membership(pos_edge_index, num_nodes, range_list)
        m._keepalive = (pos_edge_index, range_list)
        _member_cache[key] = m
    elif m._versions != (pos_edge_index._version, range_list._version):
        m.build(pos_edge_index, num_nodes, range_list)      # the positives were overwritten in place
    return m


def typed_negative_sampling(pos_edge_index, num_nodes, range_list, check_status=True, out=None, packed_out=None):
    """src/neg_sampling.py:22-26 on the GPU.  `check_status=False` skips the one host
    synchronisation (use it inside CUDA-graph capture; call `last_status()` later).
    `packed_out` (int32 [E], optional): the same pairs as (row << 16 | col), the form the fused training step
    consumes (ops.pair_bce_loss); when it is given and `out` is not, the int64 [2, E] tensor is not materialised and
    `packed_out` is returned."""
    if not pos_edge_index.is_cuda:
        raise _lib.TipbError("typed_negative_sampling takes CUDA tensors only (there is no CPU path)")
    with torch.cuda.device(pos_edge_index.device):     # the library launches on the current device
        return _typed_negative_sampling(pos_edge_index, num_nodes, range_list, check_status, out, packed_out)


def _typed_negative_sampling(pos_edge_index, num_nodes, range_list, check_status, out, packed_out):
    assert pos_edge_index.dtype == torch.long and pos_edge_index.dim() == 2 and pos_edge_index.shape[0] == 2
    num_nodes = int(num_nodes)
    dev = pos_edge_index.device
    m = _membership(pos_edge_index, num_nodes, range_list)
    if packed_out is not None:
        assert packed_out.dtype == torch.int32 and packed_out.numel() == m.n_edges and packed_out.is_contiguous()
        if num_nodes > 65535:
            raise _lib.TipbError("packed negative pairs hold 16-bit node ids")
    elif out is None:
        out = torch.empty((2, m.n_edges), dtype=torch.long, device=dev)
    if m.n_edges == 0:
        return out if out is not None else packed_out
    L = lib()
    rng = _get_rng(dev)
    exact = 0
    if check_status:
        # the status word is sticky (the library only ORs into it): anything pending here was left by an earlier
        # UNCHECKED call, whose negatives -- and every step since -- were therefore not the reference's
        pending = int(m.status.item())
        if pending:
            m.status.zero_()
            raise _lib.TipbError(f"negative sampling: an earlier unchecked call failed (status {pending}); its "
                                 "negatives differ from the reference's.  Use check_status=True or call last_status()")
    while True:
        rng.join()                                   # an earlier prefetch must be done before its words are read
        if not rng.valid or rng.n_new < m.n_new:
            rng.generate(max(m.n_new, rng.n_new))    # on the caller's stream (first call / state was reset)
            rng.valid = True
        n_words = 624 + rng.n_new
        ws = workspace(L.tipb_neg_sample_workspace_bytes(m.n_edges, m.n_rel, n_words, m.sum_l, m.sum_w), dev, "neg")
        check(L.tipb_neg_sample(ptr(rng.state), ptr(rng.words), n_words, ptr(m.member), ptr(m.range_dev), ptr(m.table),
                                m.sum_l, m.sum_w, m.n_edges, num_nodes, m.n_rel, exact, ptr(out), ptr(packed_out),
                                ptr(m.status), ptr(ws), ws.numel(), stream()), "neg_sample")
        code = int(m.status.item()) if check_status else 0
        if code == 0:
            rng.valid = False                        # the state moved on; `words` no longer starts at it
            break
        # a failed call consumed nothing (the library leaves the MT19937 state alone): rerun it from the same state
        m.status.zero_()
        if code & 2:
            raise _lib.TipbError("negative sampling: the retry-round table overflowed "
                                 "(some relation's positive pairs cover almost every cell)")
        if code & 4:
            exact = 1                                # an offset left its bracket: sequential exact path
        if code & 1:
            m.n_new = m.n_new * 2                    # not enough pre-generated words
    if _prefetch:
        # words for the next call, generated concurrently with whatever the caller does next
        cur = torch.cuda.current_stream(dev)
        rng.side.wait_stream(cur)
        with torch.cuda.stream(rng.side):
            rng.generate(max(m.n_new, rng.n_new))
        rng.valid = True
        rng.forked = True
    return out if out is not None else packed_out


def last_status(device=None, clear=True):
    """OR of the sticky status words of all cached samplers on `device`: 0 = EVERY unchecked call since the last
    read had enough pre-generated words and stayed inside its brackets.  The library only ever ORs failure bits
    into the word and a failed call does not advance the MT19937 state, so a failure inside a CUDA-graph replay
    cannot be overwritten by later steps.  One host synchronisation; `clear` resets the words after reading."""
    device = _device_of(device)
    code = 0
    for m in list(_member_cache.values()) + list(_sharded_samplers):
        if m.status.device == device:
            code |= int(m.status.item())
            if clear:
                m.status.zero_()
    return code


def negative_sampling(pos_edge_index, num_nodes):
    """src/neg_sampling.py:5-19 (single relation)."""
    e = pos_edge_index.shape[1]
    rl = torch.tensor([[0, e]], dtype=torch.long, device=pos_edge_index.device)
    return typed_negative_sampling(pos_edge_index, num_nodes, rl)


# =============================================================================== relation-sharded sampler
import weakref

_sharded_samplers = weakref.WeakSet()


class ShardedSampler(object):
    """typed_negative_sampling for ONE rank of a relation-sharded run (tip_b200/parallel.py).

    The accepted MT19937 stream is one global sequence, so the negatives of a relation depend on how many values
    all earlier relations consumed.  A rank holds the bitmaps of ITS relations only, scans only their windows and
    composes them into one table (its start offset -> the next rank's start offset); the ranks all-gather these
    tables (a few KB: the only exchange step), walk them to their own start offset and materialise their pairs.
    Every rank advances the same MT19937 state, so the result is bit for bit the slice [e_lo, e_hi) of what the
    unsharded sampler -- and the reference, src/neg_sampling.py:22-26 -- produces.

    `local_idx` / `local_range`: the edges and ranges (counted from 0) of the relations [r_lo, r_hi);
    `rl_host`: the FULL cumulative range table on the host; `first_rel`: first relation of every rank (+ n_rel);
    `coll`: object with all_gather_(out [world, n], inp [n]) and all_reduce_max_(t) (see parallel._Collective)."""

    def __init__(self, local_idx, num_nodes, local_range, rl_host, first_rel, rank, world, coll):
        L = lib()
        dev = local_idx.device
        self.device, self.num_nodes, self.rank, self.world, self.coll = dev, int(num_nodes), int(rank), int(world), coll
        self.rl_host = np.ascontiguousarray(np.asarray(rl_host, dtype=np.int64))
        self.first_rel_host = [int(v) for v in first_rel]
        self.n_rel = int(self.rl_host.shape[0])
        self.r_lo, self.r_hi = self.first_rel_host[rank], self.first_rel_host[rank + 1]
        self.n_local = self.r_hi - self.r_lo
        self.n_edges = int(self.rl_host[-1, 1]) if self.n_rel else 0
        self.e_lo = int(self.rl_host[self.r_lo, 0]) if self.n_local else 0
        self.e_hi = int(self.rl_host[self.r_hi - 1, 1]) if self.n_local else 0
        assert int(local_idx.shape[1]) == self.e_hi - self.e_lo
        self.range_dev = torch.from_numpy(self.rl_host).to(dev)
        self.first_rel = torch.tensor(self.first_rel_host, dtype=torch.int32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.z_sigma = Z_SIGMA
        self.member = None
        self._pop_all = None
        _sharded_samplers.add(self)
        with torch.cuda.device(dev):
            self.rebuild(local_idx, local_range)

    def rebuild(self, local_idx, local_range):
        """bitmaps of the own relations (device), their popcounts exchanged once, bracket table on the host"""
        L = lib()
        dev = self.device
        nbytes = L.tipb_neg_bitmap_bytes(self.num_nodes, max(self.n_local, 1))
        if nbytes > (BITMAP_LIMIT_GIB << 30):
            raise _lib.TipbError(f"positive-pair bitmaps of this rank would need {nbytes >> 30} GiB")
        if self.member is None or self.member.numel() != max(nbytes // 4, 1):
            self.member = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=dev)
        n_pad = max(max(b - a for a, b in zip(self.first_rel_host[:-1], self.first_rel_host[1:])), 1)
        pop_local = torch.zeros(n_pad, dtype=torch.int32, device=dev)
        lr = _i64c(local_range.to(device=dev, dtype=torch.long))
        if self.n_local:
            check(L.tipb_neg_bitmap_build(ptr(_i64c(local_idx)), ptr(lr), int(local_idx.shape[1]), self.num_nodes,
                                          self.n_local, ptr(self.member), ptr(pop_local), stream()), "neg_bitmap_build")
        pop_all = torch.zeros((self.world, n_pad), dtype=torch.int32, device=dev)
        self.coll.all_gather_(pop_all, pop_local)
        pop_all = pop_all.cpu().numpy()
        self._pop_all = np.concatenate([pop_all[k, : self.first_rel_host[k + 1] - self.first_rel_host[k]]
                                        for k in range(self.world)]).astype(np.int32) if self.n_rel else np.zeros(1, np.int32)
        self._build_table()

    def _build_table(self):
        L = lib()
        table = np.zeros((max(self.n_rel, 1), 8), dtype=np.int64)
        totals = np.zeros(4, dtype=np.int64)
        pop = np.ascontiguousarray(self._pop_all)
        check(L.tipb_neg_table_build(self.rl_host.ctypes.data_as(C.c_void_p), pop.ctypes.data_as(C.c_void_p), self.n_rel,
                                     self.num_nodes, self.z_sigma, table.ctypes.data_as(C.c_void_p),
                                     totals.ctypes.data_as(C.c_void_p)), "neg_table_build")
        self.table = torch.from_numpy(table).to(self.device)
        self.sum_l, self.sum_w, max_index = int(totals[0]), int(totals[1]), int(totals[2])
        firsts = [a for a, b in zip(self.first_rel_host[:-1], self.first_rel_host[1:]) if b > a]
        self.w_max = int(max([int(table[a, 1]) for a in firsts] + [1]))
        self.rank_table = torch.empty(self.w_max, dtype=torch.int32, device=self.device)
        self.all_tables = torch.empty((self.world, self.w_max), dtype=torch.int32, device=self.device)
        bits = max(int(self.num_nodes * self.num_nodes - 1).bit_length(), 1)
        p_accept = float(self.num_nodes) ** 2 / float(1 << bits)
        need = max_index / p_accept
        self.n_new = int(need + Z_SIGMA * np.sqrt(need * (1.0 - p_accept) / p_accept + 1.0) + 4096)

    def sample(self, out=None, packed_out=None, check_status=True):
        """negatives of this rank's relations: int64 [2, E_local] and / or packed int32 [E_local]"""
        with torch.cuda.device(self.device):
            return self._sample(out, packed_out, check_status)

    def _sample(self, out, packed_out, check_status):
        L = lib()
        dev = self.device
        e_local = self.e_hi - self.e_lo
        if packed_out is None and out is None:
            out = torch.empty((2, e_local), dtype=torch.long, device=dev)
        if self.n_edges == 0:
            return out if out is not None else packed_out
        result = out if out is not None else packed_out
        if e_local == 0:            # a rank without relations still takes part in the exchange and advances the stream
            out, packed_out = torch.empty((2, 1), dtype=torch.long, device=dev), None
        rng = _get_rng(dev)
        if check_status:
            pending = int(self.status.item())
            if pending:
                self.status.zero_()
                raise _lib.TipbError(f"negative sampling: an earlier unchecked call failed (status {pending})")
        while True:
            rng.join()
            if not rng.valid or rng.n_new < self.n_new:
                rng.generate(max(self.n_new, rng.n_new))
                rng.valid = True
            n_words = 624 + rng.n_new
            ws = workspace(L.tipb_neg_sample_workspace_bytes(self.n_edges, self.n_rel, n_words, self.sum_l, self.sum_w), dev, "neg")
            check(L.tipb_neg_sample_shard_begin(ptr(rng.state), ptr(rng.words), n_words, ptr(self.member), ptr(self.table),
                                                self.sum_l, self.sum_w, self.n_edges, self.num_nodes, self.n_rel, self.r_lo,
                                                self.r_hi, ptr(self.rank_table), self.w_max, ptr(ws), ws.numel(), stream()),
                  "neg_sample_shard_begin")
            self.coll.all_gather_(self.all_tables, self.rank_table)          # the sampler's only exchange step
            check(L.tipb_neg_sample_shard_end(ptr(rng.state), ptr(rng.words), n_words, ptr(self.member), ptr(self.range_dev),
                                              ptr(self.table), self.sum_l, self.sum_w, self.n_edges, self.num_nodes,
                                              self.n_rel, ptr(self.all_tables), ptr(self.first_rel), self.world, self.rank,
                                              self.w_max, self.r_lo, self.r_hi, self.e_lo, self.e_hi, ptr(out),
                                              ptr(packed_out), ptr(self.status), ptr(ws), ws.numel(), stream()),
                  "neg_sample_shard_end")
            code = int(self.status.item()) if check_status else 0
            if code == 0:
                rng.valid = False
                break
            # every rank walked the same tables and saw the same failure; the state was left untouched: retry
            self.status.zero_()
            if code & 4:
                self.z_sigma *= 2.0                   # an offset left its bracket: widen all brackets (same on every rank)
                self._build_table()
            if code & 1:
                self.n_new = self.n_new * 2
        if _prefetch:
            cur = torch.cuda.current_stream(dev)
            rng.side.wait_stream(cur)
            with torch.cuda.stream(rng.side):
                rng.generate(max(self.n_new, rng.n_new))
            rng.valid = True
            rng.forked = True
        return result
