"""TEST INFRASTRUCTURE ONLY -- CPU oracle for row L of SURVEY.md section 8(a) (edge layout)
and for the CSR-by-relation index structures the CUDA path builds from it.

PINNED against the reference's own `src/utils.py` (imports in the build
container) through tests/golden/layout_*.npz.
"""
import math

import numpy as np

from .neg_sampling_oracle import MT19937


def bernoulli_keep_mask(rng, p, n):
    """legacy np.random.binomial(1, p, n) as src/utils.py:42 calls it.  For n_trials=1
    and p > 0.5 numpy runs the inversion sampler on q = 1-p: one 53-bit double per
    draw (two MT words: (a>>5)*2^26 + (b>>6)) / 2^53; the draw is 1 iff U <= exp(log(1-q))."""
    assert 0.5 < p < 1.0
    q = 1.0 - p
    qn = math.exp(1 * math.log(1.0 - q))
    w = rng.words(2 * n).astype(np.uint64)
    u = ((w[0::2] >> np.uint64(5)) * 67108864.0 + (w[1::2] >> np.uint64(6))) / 9007199254740992.0
    x = (u > qn).astype(np.int64)       # inversion result X in {0,1}; a second loop pass never fires
    return 1 - x


def to_bidirection(edge_index):
    """src/utils.py:17-23: [pairs..., same pairs with the two rows swapped...]."""
    return np.concatenate([edge_index, edge_index[::-1]], axis=1)


def process_edges(rng, raw_edge_list, p=0.9):
    """src/utils.py:35-65 on numpy arrays; returns train/test (index, type, range)."""
    out = {"train": ([], []), "test": ([], [])}
    for r, idx in enumerate(raw_edge_list):
        keep = bernoulli_keep_mask(rng, p, idx.shape[1])
        for name, sel in (("train", np.nonzero(keep)[0]), ("test", np.nonzero(1 - keep)[0])):
            out[name][0].append(to_bidirection(idx[:, sel]))
            out[name][1].append(np.full(2 * sel.size, r, dtype=np.int64))
    res = []
    for name in ("train", "test"):
        sizes = np.array([e.shape[1] for e in out[name][0]], dtype=np.int64)
        ends = np.cumsum(sizes)
        res += [np.concatenate(out[name][0], axis=1), np.concatenate(out[name][1]),
                np.stack([ends - sizes, ends], axis=1)]
    return tuple(res)


def typed_csr(edge_index, edge_type, n_nodes, n_rel, by="dst", rel_major=False):
    """Index structures of the CUDA path (tip_b200/csrc/typed_csr.cu), by definition:
    STABLE sort of the edges by (node, relation) -- or (relation, node) when rel_major --
    ties keep their input order; node is the target (by='dst') or the source (by='src') endpoint.

    returns dict(eid, other, seg_ptr, seg_node, seg_rel, node_ptr, deg); for rel_major plans
    node_ptr is not defined here (the CUDA plan lists a node's segments through rel_seg)."""
    a, b = (1, 0) if by == "dst" else (0, 1)
    node, other = edge_index[a].astype(np.int64), edge_index[b].astype(np.int64)
    if rel_major:
        order = np.lexsort((node, edge_type))
        key = edge_type[order] * n_nodes + node[order]
        first = np.ones(order.size, dtype=bool)
        first[1:] = key[1:] != key[:-1]
        seg_start = np.nonzero(first)[0]
        seg_key = key[seg_start]
        return dict(eid=order.astype(np.int32), other=other[order].astype(np.int32),
                    seg_ptr=np.concatenate([seg_start, [order.size]]).astype(np.int32),
                    seg_node=(seg_key % n_nodes).astype(np.int32), seg_rel=(seg_key // n_nodes).astype(np.int32),
                    deg=np.bincount(node, minlength=n_nodes).astype(np.int32))
    order = np.lexsort((edge_type, node))                  # last key is primary; stable
    key = node[order] * n_rel + edge_type[order]
    first = np.ones(order.size, dtype=bool)
    first[1:] = key[1:] != key[:-1]
    seg_start = np.nonzero(first)[0]
    seg_ptr = np.concatenate([seg_start, [order.size]]).astype(np.int32)
    seg_key = key[seg_start]
    seg_node, seg_rel = (seg_key // n_rel).astype(np.int32), (seg_key % n_rel).astype(np.int32)
    node_ptr = np.searchsorted(seg_node, np.arange(n_nodes + 1)).astype(np.int32)
    deg = np.bincount(node, minlength=n_nodes).astype(np.int32)
    return dict(eid=order.astype(np.int32), other=other[order].astype(np.int32), seg_ptr=seg_ptr,
                seg_node=seg_node, seg_rel=seg_rel, node_ptr=node_ptr, deg=deg)


__all__ = ["MT19937", "bernoulli_keep_mask", "to_bidirection", "process_edges", "typed_csr"]
