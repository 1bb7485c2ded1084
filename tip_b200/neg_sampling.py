"""Drop-in for the reference's src/neg_sampling.py, running on the GPU.

    typed_negative_sampling(pos_edge_index, num_nodes, range_list) -> int64 [2, E]

Same name, arguments and result as src/neg_sampling.py:22-26 -- bit for bit, because
the device keeps a numpy-compatible MT19937 state (the reference draws from the
global `np.random` stream seeded at import, src/layers.py:14).  `seed()`,
`get_state()` and `set_state()` mirror `np.random.seed/get_state/set_state` so the
stream can be handed over from numpy (e.g. after the reference's CPU-side
`process_edges`, which consumes the same stream) and back.
"""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream
from .ops import _i64c, workspace

_states = {}        # device index -> uint32 [625] tensor (624 key words + position)
_member_cache = {}  # positive-pair bitmaps, keyed on the identity of (pos_edge_index, range_list)


def _device_of(device=None):
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TipbError("tip_b200.neg_sampling runs on CUDA devices only (there is no CPU path)")
    return torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())


def _state(device):
    st = _states.get(device.index)
    if st is None:
        st = torch.empty(625, dtype=torch.int32, device=device)
        _states[device.index] = st
        with torch.cuda.device(device):
            check(lib().tipb_mt19937_seed(ptr(st), 1111, stream()), "mt19937_seed")   # src/layers.py:14
    return st


def seed(value, device=None):
    """np.random.seed(value) for the device-side stream."""
    device = _device_of(device)
    st = _state(device)
    with torch.cuda.device(device):
        check(lib().tipb_mt19937_seed(ptr(st), int(value) & 0xFFFFFFFF, stream()), "mt19937_seed")


def get_state(device=None):
    """-> ('MT19937', key uint32[624], pos, 0, 0.0), the tuple np.random.set_state accepts."""
    st = _state(_device_of(device)).cpu().numpy().view(np.uint32)
    return ("MT19937", st[:624].copy(), int(st[624]), 0, 0.0)


def set_state(state, device=None):
    """accepts np.random.get_state() tuples."""
    device = _device_of(device)
    key = np.asarray(state[1], dtype=np.uint32)
    assert key.shape == (624,)
    host = np.concatenate([key, np.array([int(state[2])], dtype=np.uint32)]).view(np.int32)
    _state(device).copy_(torch.from_numpy(host))


class _Membership(object):
    """per-relation bitmap of the positive pairs + the host-side facts needed to size a call"""

    def __init__(self, pos_edge_index, num_nodes, range_list):
        L = lib()
        dev = pos_edge_index.device
        self.n_edges = int(pos_edge_index.shape[1])
        self.n_rel = int(range_list.shape[0])
        self.num_nodes = int(num_nodes)
        self.range_dev = _i64c(range_list.to(device=dev, dtype=torch.long))
        rl = self.range_dev.cpu().numpy()                      # one-time host copy (setup, not per step)
        sizes = (rl[:, 1] - rl[:, 0]).astype(np.float64)
        cells = float(num_nodes) ** 2
        if self.n_rel and not (np.all(sizes >= 0) and np.all(rl[1:, 0] == rl[:-1, 1]) and rl[0, 0] == 0
                               and rl[-1, 1] == self.n_edges):
            raise ValueError("range_list must be the cumulative [start,end) table of src/utils.py:26-32")
        nbytes = L.tipb_neg_bitmap_bytes(num_nodes, max(self.n_rel, 1))
        if nbytes > (24 << 30):
            raise _lib.TipbError(f"positive-pair bitmaps would need {nbytes >> 30} GiB")
        self.member = torch.empty(max(nbytes // 4, 1), dtype=torch.int32, device=dev)
        check(L.tipb_neg_bitmap_build(ptr(_i64c(pos_edge_index)), ptr(self.range_dev), self.n_edges, num_nodes,
                                      self.n_rel, ptr(self.member), stream()), "neg_bitmap_build")
        # expected MT words: every kept draw costs 1/p_accept words, every relation redraws its collisions
        bits = max(int(num_nodes * num_nodes - 1).bit_length(), 1)
        p_accept = cells / float(1 << bits)
        dens = np.minimum(sizes / cells, 0.98)
        expect = float((sizes / (1.0 - dens)).sum()) / p_accept
        self.budget = int(expect * 1.08 + 6.0 * np.sqrt(expect + 1.0) + 65536)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.out = None


def _membership(pos_edge_index, num_nodes, range_list):
    key = (pos_edge_index.data_ptr(), tuple(pos_edge_index.shape), pos_edge_index._version, range_list.data_ptr(),
           tuple(range_list.shape), range_list._version, int(num_nodes), str(pos_edge_index.device))
    m = _member_cache.get(key)
    if m is None:
        if len(_member_cache) > 16:
            _member_cache.clear()
        m = _Membership(pos_edge_index, num_nodes, range_list)
        m._keepalive = (pos_edge_index, range_list)
        _member_cache[key] = m
    return m


def typed_negative_sampling(pos_edge_index, num_nodes, range_list, check_status=True, out=None):
    """src/neg_sampling.py:22-26 on the GPU.  `check_status=False` skips the one host
    synchronisation (use it inside CUDA-graph capture; call `last_status()` later)."""
    if not pos_edge_index.is_cuda:
        raise _lib.TipbError("typed_negative_sampling takes CUDA tensors only (there is no CPU path)")
    assert pos_edge_index.dtype == torch.long and pos_edge_index.dim() == 2 and pos_edge_index.shape[0] == 2
    num_nodes = int(num_nodes)
    dev = pos_edge_index.device
    m = _membership(pos_edge_index, num_nodes, range_list)
    st = _state(dev)
    if out is None:
        out = torch.empty((2, m.n_edges), dtype=torch.long, device=dev)
    if m.n_edges == 0:
        return out
    L = lib()
    budget = m.budget
    while True:
        saved = st.clone() if check_status else None
        ws = workspace(L.tipb_neg_sample_workspace_bytes(m.n_edges, m.n_rel, budget), dev, "neg")
        check(L.tipb_neg_sample(ptr(st), ptr(m.member), ptr(m.range_dev), m.n_edges, num_nodes, m.n_rel, budget,
                                ptr(out), ptr(m.status), ptr(ws), ws.numel(), stream()), "neg_sample")
        if not check_status:
            return out
        code = int(m.status.item())
        if code == 0:
            return out
        if code & 2:
            raise _lib.TipbError("negative sampling: the retry-round table overflowed "
                                 "(some relation's positive pairs cover almost every cell)")
        st.copy_(saved)          # out of pre-generated words: rewind the stream and redo with a larger budget
        budget = budget * 2
        m.budget = budget


def last_status(device=None):
    """OR of the status words of all cached samplers on `device` (0 = every unchecked call had enough
    pre-generated words).  One host synchronisation."""
    device = _device_of(device)
    code = 0
    for m in _member_cache.values():
        if m.status.device == device:
            code |= int(m.status.item())
    return code


def negative_sampling(pos_edge_index, num_nodes):
    """src/neg_sampling.py:5-19 (single relation)."""
    e = pos_edge_index.shape[1]
    rl = torch.tensor([[0, e]], dtype=torch.long, device=pos_edge_index.device)
    return typed_negative_sampling(pos_edge_index, num_nodes, rl)
